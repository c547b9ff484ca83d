"""Is a PPO epoch bound by the host issuing it or by the device running it?  Prints, for a few epochs at the bench shape, the
host time until every launch of MAPPO.train is issued and the wall time until the device is done."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ab  # noqa: E402,F401
import bench  # noqa: E402
from distributed_multi_agent_reinforcement_learning_b200.mappo_parallel import MAPPO  # noqa: E402
from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import BatchedPursuitEnv, RolloutArena  # noqa: E402

cfg = bench.make_cfg()
B, M, T = bench.B_PER_GPU, bench.N_MAPS, bench.T_STEPS
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
for i in range(5):
    t0 = time.perf_counter()
    m.train(tb, total_steps=B * T, return_numpy=False)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"epoch {i}: issued after {1e3 * m.last_train_issue_s:.1f} ms, done after {1e3 * (t1 - t0):.1f} ms", flush=True)
