"""Secondary measurement: the 2-D N-vs-E particle env (csrc/envn2n_kernels.cu) — one fused launch per K steps with actions from the
device counter RNG.  Prints one JSON line.   usage: python tools/bench_envn2n.py [B] [N] [E] [K]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_multi_agent_reinforcement_learning_b200.particle_env_n2n import BatchedParticleEnvN2N, EnvN2nArena  # noqa: E402

B, N, E, K = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((1, 65536), (2, 8), (3, 3), (4, 100)))
env = BatchedParticleEnvN2N(B, N, E)
arena = EnvN2nArena(N, E, B, K, env.device)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ms = []
for it in range(6):
    env.reset(seed=it)
    flush.fill_(1)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    env.rollout(arena, K, 0, None, None, seed=it)
    e.record()
    e.synchronize()
    if it >= 2:
        ms.append(s.elapsed_time(e))
t = sum(ms) / len(ms) * 1e-3
stored = sum(getattr(arena, f).numel() * getattr(arena, f).element_size() for f in ("p_state_f32", "e_state_f32", "p_active", "e_active",
                                                                                   "pp_adj_bits", "pe_adj_bits", "assign", "action", "reward", "done"))
print(json.dumps({"workload": f"env_n2n {B} envs x {N} pursuers x {E} evaders x {K} steps (one launch)", "ms_per_launch": t * 1e3,
                  "agent_env_steps_per_sec": B * N * K / t, "record_GBps": stored / t / 1e9,
                  "alive_pursuers_at_end": float(env.p_active.float().mean().item())}))
