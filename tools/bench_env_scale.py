"""Env-kernel scaling sweep: the fused open-loop rollout kernel (observe -> evader tape -> step -> reward-norm -> store,
K env steps per launch; pursuit_env.py:104-209 of the reference) at growing env counts, reported as agent-env-steps/s and
as HBM GB/s of the ALGORITHMIC bytes (118 B per agent-step at N=8, O=176) against the measured HBM peak.
At BASELINE config 2 (4096 envs) the grid is 0.4 waves and nothing can be bandwidth-bound; this shows where the kernel goes
as the batch grows towards configs 3-5.   usage: python tools/bench_env_scale.py [N] [K] [B ...]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ab  # noqa: E402,F401
from distributed_multi_agent_reinforcement_learning_b200 import default_config  # noqa: E402
from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import BatchedPursuitEnv, RolloutArena  # noqa: E402


def hbm_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        for k in ("hbm_gbs", "hbm_gbps"):
            if k in d:
                return float(d[k]), k
    except Exception:
        pass
    return 6547.8, "fallback"


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    Bs = [int(x) for x in sys.argv[3:]] or [4096, 32768, 262144, 1048576]
    peak, src = hbm_peak()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for B in Bs:
        cfg = default_config(env__num_defender=N, env__max_steps=K)
        env = BatchedPursuitEnv(cfg, B, device="cuda:0", num_maps=256)
        env.reset_device(seed=5)
        arena = RolloutArena(env.params, B, K, env.device)
        e_tape = env.e_state.unsqueeze(0).repeat(K + 1, 1, 1).contiguous()
        p0 = env.p_state.clone()
        O = env.O
        per_agent_step = 72 + 2 * ((O + 7) // 8) + (N + 7) // 8 + 1
        ms = []
        for it in range(6):
            env.p_state.copy_(p0)
            env.start_episode()
            flush.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            env.rollout(arena, K, 0, e_tape=e_tape, seed=it, sync_evader=False)
            e.record()
            e.synchronize()
            if it >= 2:
                ms.append(s.elapsed_time(e))
        t = sum(ms) / len(ms) * 1e-3
        steps = B * N * K
        stored = arena.nbytes()
        print(json.dumps({"B": B, "N": N, "K": K, "ms_per_launch": t * 1e3, "agent_env_steps_per_sec": steps / t,
                          "algorithmic_GBps": steps * per_agent_step / t / 1e9, "frac_of_hbm_peak": steps * per_agent_step / t / 1e9 / peak,
                          "bytes_per_agent_step": per_agent_step, "record_bytes_stored_GBps": stored / t / 1e9,
                          "hbm_peak_GBps": peak, "peak_source": src}), flush=True)
        del env, arena, e_tape, p0
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
