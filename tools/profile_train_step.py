"""Profiling aid: one training minibatch (forward + backward) at the bench shape, for ncu captures of the training kernels."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from distributed_multi_agent_reinforcement_learning_b200.mappo_parallel import MAPPO  # noqa: E402
from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import BatchedPursuitEnv, RolloutArena  # noqa: E402

cfg = bench.make_cfg()
B, M, T = 512, 32, bench.T_STEPS
wl = bench.host_workload(cfg, B, M, seed=1)
env = BatchedPursuitEnv(cfg, B, num_maps=M)
env.set_maps(wl["grids"], wl["inflated"])
env.set_state(wl["p_state"], wl["e_state"], wl["target"], wl["map_id"], time_step=0)
env.set_target_tape(wl["tape"])
env.start_episode()
arena = RolloutArena(env.params, B, T, env.device)
torch.manual_seed(0)
m = MAPPO(cfg, B, 410, "Learner")          # first minibatch = 410 envs, as in the bench (4096 / 10)
tb = m.rollout_batched(env, arena, T, seed=1)
torch.cuda.synchronize()
m.train(tb, total_steps=B * T, return_numpy=False)
torch.cuda.synchronize()
print("done")
