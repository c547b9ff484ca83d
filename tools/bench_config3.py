"""BASELINE configs[2]'s shape on the same footing as the headline bench: 16 pursuers, DHGN depth 3, 16 384 envs on one GPU, the
WHOLE network-in-the-loop episode (observe -> fused policy step of actor and critic -> A* replan when due -> fused env step) as one
CUDA graph of env-group pipelines (`MAPPO.explore_batched`, the public per-iteration call), timed with CUDA events over replays.
Prints one JSON line.   usage: python tools/bench_config3.py [B] [N] [depth] [T] [replays]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_multi_agent_reinforcement_learning_b200 import default_config  # noqa: E402
from distributed_multi_agent_reinforcement_learning_b200.mappo_parallel import MAPPO  # noqa: E402
from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import BatchedPursuitEnv, RolloutArena  # noqa: E402

B, N, D, T, REPS = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((1, 16384), (2, 16), (3, 3), (4, 150), (5, 3)))
cfg = default_config(env__num_defender=N, env__max_steps=T, algo__depth=D)
env = BatchedPursuitEnv(cfg, B, num_maps=256)
env.reset_device(seed=3, tape_len=32)
arena = RolloutArena(env.params, B, T, env.device)
torch.manual_seed(0)
m = MAPPO(cfg, B, max(1, B // 10), "Learner")
snap = env.snapshot()
m.explore_batched(env, arena, T, seed=1)          # captures the episode graph (after an eager warm-up episode)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=env.device)
ms = []
for it in range(REPS):
    env.restore(snap)
    flush.fill_(1)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    m.explore_batched(env, arena, T, seed=1)
    e.record()
    e.synchronize()
    ms.append(s.elapsed_time(e))
assert int(env.evader_status.max()) == 0, "evader search overflow / target tape exhausted"
t = min(ms) * 1e-3
print(json.dumps({"workload": f"pursuit-evasion, {N} pursuers, DHGN depth {D}, {B} envs, whole {T}-step episode, actor + critic in the loop, "
                              f"one CUDA graph of {MAPPO.default_pipelines(B, N)} env-group pipelines",
                  "ms_per_episode": t * 1e3, "ms_per_env_step": t * 1e3 / T, "agent_env_steps_per_sec": B * N * T / t,
                  "hbm_GB_allocated": torch.cuda.max_memory_allocated() / 1e9}))
