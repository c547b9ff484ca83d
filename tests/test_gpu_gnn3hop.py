"""GPU parity of the second network family (mappo_3hop.py: GnnExtractor / SharedActor / SharedCritic / MAPPO.train on
csrc/gnn_kernels.cu + the shared GRU / head / GEMM kernels) against golden vectors recorded by executing the unmodified
reference classes of obstacle_differ_3hop/mappo_parallel.py (oracle/gen_golden_gnn3hop.py).
Tolerances: 1e-5 forward activations, 1e-4 losses and gradients (north_star)."""
import glob
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest

from conftest import GOLDEN_DIR

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

FIXTURES = sorted(glob.glob(os.path.join(GOLDEN_DIR, "gnn3hop_*.npz")))


def _args(n, emb):
    return NS(max_train_steps=int(2e8), lr=5e-4, gamma=0.99, lamda=0.95, epsilon=0.05, K_epochs=1, entropy_coef=0.05,
              use_grad_clip=True, use_lr_decay=True, use_adv_norm=True, use_value_clip=True, state_dim=8, num_layers=2,
              gnn_output_dim=emb, gnn_middle_dim=emb, rnn_hidden_dim=emb, n_hops=3, learner_device="cuda", worker_device="cuda",
              evaluator_device="cuda", use_reward_norm=False, use_spectral_norm=True, action_dim=9, pursuer_num=[n])


class _Big:
    def __init__(self, buf, T):
        self.buf, self.T = buf, T

    def get_training_data(self, num, device):
        return {k: v.to(device) for k, v in self.buf.items()}, self.T


def _build(fx):
    import random
    from distributed_multi_agent_reinforcement_learning_b200 import mappo_3hop as m3
    N, O, B, T, mb, emb, seed = (int(v) for v in fx["meta"])
    m = m3.MAPPO(_args(N, emb), B, mb, "Learner")
    if any(k.startswith("w.") for k in fx.files):
        for net, mod in (("actor", m.actor), ("critic", m.critic)):
            sd = {k[len("w." + net + "."):]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("w." + net + ".")}
            assert list(mod.state_dict().keys()) == list(sd.keys())           # same sub-module names / checkpoint layout
            mod.load_state_dict(sd)
    else:
        # compact fixture: the reference's weights for this seed, rebuilt with the generator's construction order on the CPU
        random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
        actor = m3.SharedActor(m3.GnnExtractor(9, emb, emb, 3, True), emb, 9, 2, emb, True)
        critic = m3.SharedCritic(m3.GnnExtractor(9, emb, emb, 3, True), emb, 1, 2, emb, True)
        with torch.no_grad():
            for p in list(actor.parameters()) + list(critic.parameters()):
                if p.dim() == 1:
                    p.add_(0.05 * torch.randn_like(p))
        for net, mod in (("actor", actor), ("critic", critic)):
            for k, v in mod.state_dict().items():
                want = float(fx[f"wsum.{net}.{k}"])
                assert abs(float(v.double().abs().sum()) - want) <= 1e-9 * max(1.0, want), (net, k)
        m.actor, m.critic = actor, critic
        m.finalize()
    return m, (N, O, B, T, mb, emb)


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_rollout_mode_forward_matches_reference(path):
    fx = np.load(path)
    m, (N, O, B, T, mb, emb) = _build(fx)
    cu = lambda k: torch.from_numpy(fx[k]).cuda()
    st, ad = cu("buf.state")[0, 0], cu("buf.adj")[0, 0]
    z = torch.zeros(2, N, emb, device="cuda")
    with torch.no_grad():
        prob, ha, comm_a = m.actor.forward(st, ad, z, cu("buf.actor_comm_embedding")[0, 0], mode=0)
        val, hc, comm_c = m.critic.forward(st, ad, z, cu("buf.critic_comm_embedding")[0, 0], mode=0)
    tol = dict(rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(comm_a, cu("roll.comm_a"), **tol)
    torch.testing.assert_close(comm_c, cu("roll.comm_c"), **tol)
    torch.testing.assert_close(ha, cu("roll.ha"), **tol)
    torch.testing.assert_close(hc, cu("roll.hc"), **tol)
    torch.testing.assert_close(prob, cu("roll.prob"), **tol)
    torch.testing.assert_close(val, cu("roll.val"), **tol)


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_train_matches_reference(path):
    fx = np.load(path)
    m, (N, O, B, T, mb, emb) = _build(fx)
    with torch.no_grad():      # train() starts from the spectral-norm buffers the reference had after its rollout-mode forward
        z = torch.zeros(2, N, emb, device="cuda")
        m.critic.forward(torch.from_numpy(fx["buf.state"])[0, 0].cuda(), torch.from_numpy(fx["buf.adj"])[0, 0].cuda(), z,
                         torch.from_numpy(fx["buf.critic_comm_embedding"])[0, 0].cuda(), mode=0)
    buf = {k[4:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("buf.")}
    objC, objA, ag, cg = m.train(_Big(buf, T), total_steps=B * T)
    assert abs(objC - float(fx["objC"])) <= 1e-4 * max(1.0, abs(float(fx["objC"])))
    assert abs(objA - float(fx["objA"])) <= 1e-4 * max(1.0, abs(float(fx["objA"])))
    checked = 0
    for net, mod, grads in (("actor", m.actor, ag), ("critic", m.critic, cg)):
        for (name, _), g in zip(mod.named_parameters(), grads):
            key = f"{net}.{name}"
            if f"grad.{key}" in fx.files:
                gold = fx[f"grad.{key}"]
                scale = max(1e-6, float(np.abs(gold).max()))
                assert g is not None and np.abs(g - gold).max() <= 1e-4 * scale + 1e-7, (key, np.abs(g - gold).max(), scale)
                checked += 1
            elif f"gnorm.{key}" in fx.files:
                gold = fx[f"gsample.{key}"]
                mine = g.reshape(-1)[::max(1, g.size // 2048)]
                scale = max(1e-6, float(np.abs(gold).max()))
                assert np.abs(mine - gold).max() <= 1e-4 * scale + 1e-7, (key, np.abs(mine - gold).max(), scale)
                assert abs(float(np.linalg.norm(g.astype(np.float64))) - float(fx[f"gnorm.{key}"])) <= 1e-4 * float(fx[f"gnorm.{key}"]) + 1e-9
                checked += 1
            else:
                assert g is None or not np.any(g), key
    assert checked >= 20


def test_entity_agg_kernels_vs_torch():
    from distributed_multi_agent_reinforcement_learning_b200 import mappo_3hop as m3
    torch.manual_seed(0)
    S, N, O, E = 37, 5, 19, 128
    J = N + O
    x = torch.randn(S, N, J, E, device="cuda", requires_grad=True)
    adj = (torch.rand(S, N, J, device="cuda") < 0.3).float()
    adj[0, 0] = 0                                             # an all-zero row -> zeros, no NaN
    for all_ones in (False, True):
        a = torch.ones_like(adj) if all_ones else adj
        w = torch.nn.functional.normalize(a, p=1, dim=-1)
        ref = torch.matmul(w.unsqueeze(-2), x).squeeze(-2)
        out = m3._EntityAgg.apply(x, adj, all_ones)
        torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-6)
        g = torch.randn_like(ref)
        (gx_ref,) = torch.autograd.grad(ref, x, g, retain_graph=True)
        (gx,) = torch.autograd.grad(out, x, g)
        torch.testing.assert_close(gx, gx_ref, rtol=1e-5, atol=1e-6)
        comm = torch.randn(S, N, 2 * E, device="cuda")
        torch.testing.assert_close(m3.comm_agg(comm, adj, all_ones), torch.matmul(w[..., :N], comm), rtol=1e-5, atol=1e-6)
