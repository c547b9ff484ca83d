"""Pins oracle/policy_ref.py (the torch fp32 restatement of the reference's network math) to golden vectors produced
by executing the unmodified reference MAPPO.train (oracle/gen_golden_algo.py).  Tolerances: north_star's 1e-5
relative for GAE / statistics and 1e-4 for loss and gradient values (both sides are torch CPU fp32 here, so the
observed differences are ~1e-6)."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, AlgoFixture

FIXTURES = sorted(glob.glob(os.path.join(GOLDEN_DIR, "algo_*.npz")))


def load_algo(path):
    fx = AlgoFixture(path)
    w = {k[2:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("w.")}
    buf = {k[4:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("buf.")}
    return fx, w, buf


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_gae_restatement_and_oracle(path, oracle):
    from oracle import policy_ref
    fx, w, buf = load_algo(path)
    adv, vt = policy_ref.gae(buf)
    torch.testing.assert_close(adv, torch.from_numpy(fx["adv"]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(vt, torch.from_numpy(fx["v_target"]), rtol=1e-6, atol=1e-7)
    # the C oracle of the GAE kernel against the reference's own advantages
    adv_c, vt_c = oracle.gae(buf["r"].numpy(), buf["v_n"].numpy(), buf["active"].numpy(), 0.99, 0.95, use_adv_norm=True)
    np.testing.assert_allclose(adv_c, fx["adv"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(vt_c, fx["v_target"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_train_forward_loss_and_gradients(path):
    from oracle import policy_ref
    fx, w, buf = load_algo(path)
    depth = int(fx["meta"][0])
    for k in w:
        if w[k].dtype.is_floating_point and not k.endswith(("weight_u", "weight_v")):
            w[k] = w[k].clone().requires_grad_(True)
    adv, vt = torch.from_numpy(fx["adv"]), torch.from_numpy(fx["v_target"])
    tot_a = tot_c = 0.0
    n_mb = int(fx["n_mb"])
    params = [v for v in w.values() if v.requires_grad]
    for i in range(n_mb):
        idx = torch.from_numpy(fx[f"mb{i}.index"])
        mb = {k: v[idx] for k, v in buf.items()}
        logp, ent, val = policy_ref.train_forward(w, mb, depth)
        torch.testing.assert_close(logp, torch.from_numpy(fx[f"mb{i}.logp"]), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(ent, torch.from_numpy(fx[f"mb{i}.ent"]), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(val, torch.from_numpy(fx[f"mb{i}.val"]), rtol=1e-5, atol=1e-6)
        la, lc = policy_ref.ppo_losses(logp, ent, val, mb, adv[idx], vt[idx])
        assert abs(float(lc) - fx[f"mb{i}.losses"][0]) <= 1e-4 * max(1.0, abs(fx[f"mb{i}.losses"][0]))
        assert abs(float(la) - fx[f"mb{i}.losses"][1]) <= 1e-4 * max(1.0, abs(fx[f"mb{i}.losses"][1]))
        (la + lc).backward()
        # the shared encoder appears under both prefixes: sum its two gradient copies like one module would
        _merge_shared(w)
        torch.nn.utils.clip_grad_norm_(_unique_params(w), 5.0)      # applied to the running accumulation (:710-711)
        tot_a += float(la)
        tot_c += float(lc)
    assert abs(tot_c / n_mb - float(fx["objC"])) < 1e-5 and abs(tot_a / n_mb - float(fx["objA"])) < 1e-5
    for k in fx.files:
        if k.startswith("grad."):
            name = k[5:]
            g = w[name].grad
            assert g is not None, name
            np.testing.assert_allclose(g.numpy(), fx[k], rtol=1e-4, atol=1e-6, err_msg=name)


def _merge_shared(w):
    for k in list(w):
        if k.startswith("actor.shared_net."):
            twin = "critic." + k[len("actor."):]
            a, c = w[k], w[twin]
            if a is c or not a.requires_grad:
                continue
            ga = a.grad if a.grad is not None else torch.zeros_like(a)
            gc = c.grad if c.grad is not None else torch.zeros_like(c)
            s = ga + gc
            w[twin] = a                     # from now on both names are the same leaf
            a.grad = s


def _unique_params(w):
    seen, out = set(), []
    for v in w.values():
        if v.requires_grad and id(v) not in seen:
            seen.add(id(v))
            out.append(v)
    return out
