"""Profiling aid: kernel-time breakdown of one MAPPO.train epoch at the bench shape (torch.profiler, CUDA activities)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ab  # noqa: E402,F401
import bench  # noqa: E402
from distributed_multi_agent_reinforcement_learning_b200.mappo_parallel import MAPPO  # noqa: E402
from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import BatchedPursuitEnv, RolloutArena  # noqa: E402

cfg = bench.make_cfg()
B, M, T = int(os.environ.get("B", bench.B_PER_GPU)), 32, bench.T_STEPS
wl = bench.host_workload(cfg, B, M, seed=1)
env = BatchedPursuitEnv(cfg, B, num_maps=M)
env.set_maps(wl["grids"], wl["inflated"])
env.set_state(wl["p_state"], wl["e_state"], wl["target"], wl["map_id"], time_step=0)
env.set_target_tape(wl["tape"])
env.start_episode()
arena = RolloutArena(env.params, B, T, env.device)
torch.manual_seed(0)
m = MAPPO(cfg, B, max(1, round(B / 10)), "Learner")
tb = m.rollout_batched(env, arena, T, seed=1)
torch.cuda.synchronize()
m.train(tb, total_steps=B * T, return_numpy=False)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    m.train(tb, total_steps=B * T, return_numpy=False)
    m.update(B * T)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
