"""Build-container calibration of bench.py's CPU arm (`kind: "port"`): the UNMODIFIED reference's own rollout
(`MAPPO.explore_env`, DHGN/mappo_parallel.py:731-827, through oracle/ref_bootstrap.py) timed next to the oracle port
(`bench.cpu_policy_rollout`: C env + torch-CPU restatement of both networks) on the same shape - 8 pursuers, depth 1, E = 128,
T = 150 - with ONE thread each.  Needs /root/reference, so it cannot run on the GPU box; the result is committed under profiles/."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle.ref_bootstrap import load_reference, make_cfg, seed_all  # noqa: E402

torch.set_num_threads(1)
os.environ["OMP_NUM_THREADS"] = "1"
N, T = 8, 150
R = load_reference()
cfg = make_cfg(num_defender=N, depth=1, max_steps=T, embedding_dim=128)
seed_all(3)
torch.set_grad_enabled(False)
worker = R.mappo.MAPPO(cfg, None, None, "Worker")
env = R.pe.Pursuit_Env(cfg)
t0 = time.perf_counter()
env.reset()
t_reset = time.perf_counter() - t0
t0 = time.perf_counter()
worker.explore_env(env, 1)
t_ep = time.perf_counter() - t0          # run_episode resets the env itself: subtract one reset for the step loop
ref_incl = N * T / t_ep
ref_excl = N * T / max(t_ep - t_reset, 1e-9)

bcfg = bench.make_cfg()
wl = bench.host_workload(bcfg, 16, 16, seed=5)
from oracle import oracle as orc  # noqa: E402
orc.build()
import ctypes  # noqa: E402
ctypes.CDLL("libgomp.so.1").omp_set_num_threads(1)
bench.cpu_policy_rollout(bcfg, wl, 4, 5)
dt, n = bench.cpu_policy_rollout(bcfg, wl, 16, T)
out = {"shape": {"pursuers": N, "depth": 1, "embedding_dim": 128, "steps": T}, "threads": 1,
       "reference_agent_env_steps_per_sec_incl_reset": ref_incl, "reference_agent_env_steps_per_sec_excl_reset": ref_excl,
       "reference_reset_seconds": t_reset, "port_agent_env_steps_per_sec": n / dt,
       "port_over_reference": (n / dt) / ref_excl,
       "note": "one core of the build container; the port is the stronger baseline, so the GPU / CPU ratio bench.py reports is conservative"}
print(json.dumps(out))
with open(os.path.join(ROOT, "profiles", "r2_cpu_arm_calibration.json"), "w") as f:
    json.dump(out, f, indent=1)
