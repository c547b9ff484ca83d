"""TEST INFRASTRUCTURE — golden vectors of the rollout-mode forward at the PRODUCTION width (embedding_dim 128), produced
by EXECUTING THE UNMODIFIED REFERENCE (`MAPPO.explore_env`, DHGN/mappo_parallel.py:731-827) in the build container.
They pin the fused rollout-step kernel (csrc/policy_fused.cu) directly against the reference's own replay buffer.

To keep the fixtures small the weights are NOT stored: they are the reference's initial weights for torch seed `seed`
(the product's MAPPO reproduces them bit for bit — tests/test_gpu_policy.py::test_initial_weights_equal_reference); a
checksum per tensor is stored instead.  Re-run:  python -m oracle.gen_golden_rollout128
"""
import os
import sys
from copy import deepcopy

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_bootstrap import load_reference, make_cfg, seed_all  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def gen(R, depth, n_def, T, episodes, seed):
    import torch
    cfg = make_cfg(num_defender=n_def, depth=depth, max_steps=T, embedding_dim=128)
    seed_all(seed)
    torch.set_grad_enabled(False)
    worker = R.mappo.MAPPO(cfg, None, None, "Worker")
    sums = {}
    for net, mod in (("actor", worker.actor), ("critic", worker.critic)):
        for k, v in mod.state_dict().items():
            sums[f"wsum.{net}.{k}"] = np.float64(v.double().abs().sum().item())
    env = R.pe.Pursuit_Env(cfg)
    big = R.replay_buffer.BigBuffer()
    o_counts = []
    for _ in range(episodes):
        r, buf, steps = worker.explore_env(env, 1)
        o_counts.append(len(env.boundary_map.obstacle_agent))
        big.concat_buffer(deepcopy(buf))
    b = big.buffer
    fx = dict(p_state=b["p_state"].numpy(), e_state=b["e_state"].numpy(), o_xy=b["o_state"][:, 0, :, :2].numpy().astype(np.int32),
              p_adj=np.packbits(b["p_adj"].numpy().astype(np.uint8), axis=-1, bitorder="little"),
              e_adj=b["e_adj"].numpy().astype(np.uint8)[..., 0],
              o_adj=np.packbits(b["o_adj"].numpy().astype(np.uint8), axis=-1, bitorder="little"),
              hist_a=b["actor_historical_embedding"].numpy(), hist_c=b["critic_historical_embedding"].numpy(),
              v_n=b["v_n"].numpy(), a_n=b["a_n"].numpy().astype(np.int32), a_logprob_n=b["a_logprob_n"].numpy(),
              o_counts=np.array(o_counts, np.int32), meta=np.array([depth, n_def, T, episodes, seed, 128], np.int64))
    fx.update(sums)
    return fx


# the N = 16 case runs the 8-worker-warp variant of the fused kernel (a whole 16-agent env per warp, BASELINE configs[2])
CASES = ((1, 8, 6, 3, 31), (3, 5, 7, 2, 33), (3, 16, 6, 2, 35))


def main(only_n=None):
    R = load_reference()
    for depth, n_def, T, episodes, seed in CASES:
        if only_n is not None and n_def != only_n:
            continue
        fx = gen(R, depth, n_def, T, episodes, seed)
        path = os.path.join(GOLDEN_DIR, f"rollout128_d{depth}_n{n_def}.npz")
        np.savez_compressed(path, **fx)
        print(path, f"{os.path.getsize(path) / 1e6:.2f} MB", "o_counts", fx["o_counts"])


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else None)
