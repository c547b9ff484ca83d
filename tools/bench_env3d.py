"""Throughput of the 3-D particle env kernel at BASELINE config 4's per-GPU shape (8192 envs x 32 pursuers, T=200):
one fused launch per K steps with device-generated actions.  Prints one JSON line (secondary measurement; the driver
bench is bench.py).  Also times the oracle on the host for the same shape (bounded sample)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_multi_agent_reinforcement_learning_b200.particle_env import BatchedParticleEnv, Env3dArena  # noqa: E402

B, N, T = int(os.environ.get("B", 8192)), int(os.environ.get("N", 32)), 200
K = int(os.environ.get("K", 20))
eng = BatchedParticleEnv(B, N)
eng.reset(seed=4)
snap = eng.snapshot()
arena = Env3dArena(N, B, T, eng.device)


def episode():
    for t0 in range(0, T, K):
        eng.rollout(arena, K, t0, None, None, seed=0xB200)


for _ in range(3):
    eng.restore(snap)
    episode()
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)
ms = []
for _ in range(10):
    eng.restore(snap)
    flush.fill_(1)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    episode()
    e.record()
    e.synchronize()
    ms.append(s.elapsed_time(e))
ms_ep = float(np.median(ms))
alive = float(eng.p_active.float().mean().item())
# algorithmic bytes per agent-step: state r/w 96 B + active 2 B + records (f32 state 24 + adj 4 + pe 1 + reward 4 + active 4)
bytes_per = 96 + 2 + 24 + 4 + 1 + 4 + 4
out = {"metric": "agent_env_steps_per_sec", "workload": f"env_3d particle env, {N} pursuers, {B} envs, T={T}, K={K} steps/launch",
       "value": B * N * T / (ms_ep * 1e-3), "ms_per_episode": ms_ep, "launches_per_episode": T // K,
       "alive_fraction_at_end": alive,
       "roofline": {"bound": "hbm", "achieved": bytes_per * B * N * T / (ms_ep * 1e-3) / 1e9, "peak": 6547.8, "unit": "GB/s",
                    "algorithmic_bytes_per_agent_step": bytes_per}}
out["roofline"]["frac"] = out["roofline"]["achieved"] / out["roofline"]["peak"]
try:
    from oracle import oracle as orc
    orc.build()
    p = orc.Env3dParams.from_dict({n: getattr(eng.params, n) for n, _ in eng.params._fields_})
    eng.restore(snap)
    Bc = min(B, 2048)
    st = dict(p_state=eng.p_state.cpu().numpy()[:Bc].copy(), p_active=eng.p_active.cpu().numpy()[:Bc].copy(),
              e_state=eng.e_state.cpu().numpy()[:Bc].copy(), e_active=eng.e_active.cpu().numpy()[:Bc].copy(),
              target=eng.target.cpu().numpy()[:Bc].copy(), time_step=np.zeros(Bc, np.int32), reward=np.zeros((Bc, N), np.int32),
              done=np.zeros(Bc, np.uint8), pp_adj=np.zeros((Bc, N, N), np.uint8), pe_adj=np.zeros((Bc, N), np.uint8))
    rng = np.random.default_rng(0)
    t0 = time.perf_counter()
    steps = 40
    for k in range(steps):
        st["action"] = rng.uniform(-1, 1, (Bc, N, 3))
        st["e_action"] = rng.uniform(-1, 1, (Bc, 3))
        orc.env3d_iteration(p, st)
    dt = time.perf_counter() - t0
    out["cpu_baseline"] = {"value": Bc * N * steps / dt, "cores": orc.num_threads(), "kind": "port",
                           "sample": f"{Bc} envs x {steps} steps, oracle C (OpenMP)"}
except Exception as ex:   # the oracle is optional for this tool
    out["cpu_baseline"] = {"error": str(ex)}
print(json.dumps(out))
