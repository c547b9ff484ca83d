"""Profiling aid: a few steps of the network-in-the-loop rollout at the bench shape (4096 envs x 8 pursuers), meant to be
run under `ncu --metrics gpu__time_duration.sum` to get the per-kernel time split of one env step."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from distributed_multi_agent_reinforcement_learning_b200.mappo_parallel import MAPPO  # noqa: E402
from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import BatchedPursuitEnv, RolloutArena  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = bench.make_cfg()
B, M = bench.B_PER_GPU, 32
wl = bench.host_workload(cfg, B, M, seed=1)
env = BatchedPursuitEnv(cfg, B, num_maps=M)
env.set_maps(wl["grids"], wl["inflated"])
env.set_state(wl["p_state"], wl["e_state"], wl["target"], wl["map_id"], time_step=0)
env.set_target_tape(wl["tape"])
env.start_episode()
arena = RolloutArena(env.params, B, T, env.device)
torch.manual_seed(0)
m = MAPPO(cfg, B, B // 10, "Learner")
snap = env.snapshot()
m.rollout_batched(env, arena, T, seed=1)      # warm-up
torch.cuda.synchronize()
env.restore(snap)
torch.cuda.nvtx.range_push("steps")
m.rollout_batched(env, arena, T, seed=1)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("done")
