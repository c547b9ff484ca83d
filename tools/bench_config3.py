"""Secondary measurement at BASELINE configs[2]'s shape: 16 pursuers, DHGN depth 3, 16 384 envs on one GPU — a slice of T steps of the
network-in-the-loop rollout (observe -> fused policy step of actor and critic -> A* replan when due -> fused env step), timed with
CUDA events.  Prints one JSON line.   usage: python tools/bench_config3.py [B] [N] [depth] [T]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_multi_agent_reinforcement_learning_b200 import default_config  # noqa: E402
from distributed_multi_agent_reinforcement_learning_b200.mappo_parallel import MAPPO  # noqa: E402
from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import BatchedPursuitEnv, RolloutArena  # noqa: E402

B, N, D, T = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((1, 16384), (2, 16), (3, 3), (4, 20)))
cfg = default_config(env__num_defender=N, env__max_steps=T, algo__depth=D)
env = BatchedPursuitEnv(cfg, B, num_maps=256)
env.reset_device(seed=3, tape_len=32)
arena = RolloutArena(env.params, B, T, env.device)
torch.manual_seed(0)
m = MAPPO(cfg, B, max(1, B // 10), "Learner")
snap = env.snapshot()
m.rollout_batched(env, arena, T, seed=1)          # warm-up
torch.cuda.synchronize()
ms = []
for it in range(3):
    env.restore(snap)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    m.rollout_batched(env, arena, T, seed=2 + it)
    e.record()
    e.synchronize()
    ms.append(s.elapsed_time(e))
t = min(ms) * 1e-3
print(json.dumps({"workload": f"pursuit-evasion, {N} pursuers, DHGN depth {D}, {B} envs, {T}-step slice, actor + critic in the loop (eager launches)",
                  "ms_per_env_step": t * 1e3 / T, "agent_env_steps_per_sec": B * N * T / t,
                  "hbm_GB_allocated": torch.cuda.max_memory_allocated() / 1e9}))
