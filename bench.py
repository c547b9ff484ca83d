#!/usr/bin/env python
"""bench.py — headline benchmark of the rollout hot path (driver contract: see the task statement).

Workload (BASELINE.json configs[1]): pursuit-evasion with obstacles, 8 pursuers + 1 A*-driven evader, 60x55 map,
O=176, T=150 steps, 4096 batched envs per B200.  One bench "step" = one whole closed-loop episode of all envs:
per env step  evader A* / waypoint step -> observe (comm adjacency, LoS, obstacle visibility) -> pursuer step ->
reward-norm -> store into the time-major rollout arena.  Actions come from the device-side counter generator
(scripted policy stand-in; the actor/critic network is not in this loop yet — stated in `config.policy`).

metric = agent-env-steps/s = B*N*T / time, whole job (all ranks).
  value : episode replayed from HBM-resident initial state (CUDA graph of 2*T/10 kernel launches: one A* replanning
          launch + one fused 10-step rollout launch per replanning period).
  e2e   : same episode through the public API with HOST buffers: pinned initial states/targets H2D + episode
          reward sums D2H inside the timed region.
--impl reference : the CPU oracle port of the same loop (oracle/marl_oracle.c, all host threads) on a bounded
          sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "agent_env_steps_per_sec"
UNIT = "agent-env-steps/s"
N_AGENTS, B_PER_GPU, T_STEPS, N_MAPS = 8, 4096, 150, 256
STREAM_GROUPS = 16   # independent env sub-batches on separate streams inside the episode graph
SURVEY_BYTES_PER_AGENT_STEP = 118.0   # 72 state r/w + 22 raser row + 1 p_adj + 1 e_adj + 22 o_adj (N=8, O=176)


def workload_config(n_gpus):
    return {"workload": "pursuit_evasion_obstacles_8p_gru_c2", "num_pursuers": N_AGENTS, "envs_per_gpu": B_PER_GPU,
            "global_envs": B_PER_GPU * n_gpus, "max_steps": T_STEPS, "map": "60x55", "num_max_obstacle": 176,
            "map_pool_per_gpu": N_MAPS, "evader": "gpu A* (replan every 10 steps)",
            "policy": "uniform random actions from a device counter RNG (network not in the loop this round)",
            "parallelism": f"dp{n_gpus} (independent env shards, no data-path collective)",
            "stream_groups": STREAM_GROUPS, "l2": "flushed between timed iterations (256 MiB write)"}


def make_cfg():
    from distributed_multi_agent_reinforcement_learning_b200 import default_config
    return default_config(env__num_defender=N_AGENTS, env__max_steps=T_STEPS)


def host_workload(cfg, B, M, seed):
    """Synthetic initial conditions generated on the host with the reference's placement rules (untimed set-up)."""
    import numpy as np
    from distributed_multi_agent_reinforcement_learning_b200 import maps
    rng = maps.GenRng(seed)
    W, H = cfg.map.map_size
    grids = np.zeros((M, W, H), np.uint8)
    infl = np.zeros_like(grids)
    for m in range(M):
        grids[m] = maps.make_obstacle_grid(cfg.map, rng)
    infl = maps.dilate(grids, 2)
    N = cfg.env.num_defender
    ps = np.zeros((B, N, 4))
    es = np.zeros((B, 4))
    tg = np.zeros((B, 2), np.int32)
    mid = (np.arange(B) % M).astype(np.int32)
    for b in range(B):
        work = infl[mid[b]].copy()
        tg[b] = maps.draw_target(infl[mid[b]], rng)
        pxy, cells = maps.place_pursuers(work, N, float(cfg.defender.comm_range), rng)
        ps[b, :, :2] = pxy
        es[b, :2] = maps.place_evader(work, cells, float(cfg.defender.sen_range), rng)
    tape = np.stack([rng.g.integers(0, W, (B, 16)), rng.g.integers(0, H, (B, 16))], axis=-1).astype(np.int32)
    return dict(grids=grids, inflated=infl, p_state=ps, e_state=es, target=tg, map_id=mid, tape=tape)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].startswith("Active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU reference
def cpu_rollout(cfg, wl, n_envs, steps_T, seed):
    """The oracle port of the same closed loop on `n_envs` envs of the workload, all host threads.
    Returns (seconds, agent_env_steps)."""
    import numpy as np
    from oracle import oracle as orc
    from distributed_multi_agent_reinforcement_learning_b200 import env_params_dict, maps
    p = orc.EnvParams.from_dict(env_params_dict(cfg))
    B, N, O, M = n_envs, p.N, p.O, wl["grids"].shape[0]
    # per-map sensor tables with the oracle itself (untimed set-up, like reset in the reference)
    beam = maps.beam_directions(p.sensor_beams)
    used = sorted(set(int(m) for m in wl["map_id"][:B]))
    raser = np.zeros((M, p.W * p.H, O), np.uint8)
    ob_count = np.zeros(M, np.int32)
    for m in used:
        b_, xy, n = orc.boundary_map(p, wl["grids"][m])
        ob_count[m] = min(n, O)
        raser[m, :, :min(n, O)] = orc.raser_map(p, b_, xy, beam).reshape(p.W * p.H, -1)[:, :O]
    st = dict(p_state=wl["p_state"][:B].copy(), e_state=wl["e_state"][:B].copy(), target=wl["target"][:B].copy(),
              path=np.zeros((B, 512, 2), np.int16), path_len=np.zeros(B, np.int32),
              grid=np.ascontiguousarray(wl["grids"]), inflated=np.ascontiguousarray(wl["inflated"]), raser=raser,
              ob_count=ob_count, map_id=wl["map_id"][:B].copy(), action_table=maps.action_table(cfg.defender.vmax),
              tape=np.ascontiguousarray(wl["tape"][:B]), tape_pos=np.zeros(B, np.int32),
              p_adj=np.zeros((B, N, N), np.uint8), o_adj=np.zeros((B, N, O), np.uint8), e_adj=np.zeros((B, N), np.uint8),
              reward=np.zeros((B, N), np.int32), can_apply=np.zeros((B, N), np.uint8), collision=np.zeros(B, np.uint8),
              time_step=np.zeros(B, np.int32), done=np.zeros(B, np.uint8), wf_n=np.zeros(B, np.int64),
              wf_mean=np.zeros((B, N)), wf_S=np.zeros((B, N)), wf_std=np.zeros((B, N)),
              r_norm=np.zeros((B, N), np.float32), status=np.zeros(B, np.int32))
    rng = np.random.default_rng(seed)
    actions = rng.integers(0, 9, (steps_T, B, N)).astype(np.int32)
    t0 = time.perf_counter()
    for t in range(steps_T):
        st["action"] = actions[t]
        orc.rollout_iteration_closed(p, st)
    dt = time.perf_counter() - t0
    assert not st["status"].any(), "oracle evader reported an error"
    return dt, B * N * steps_T


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    cfg = make_cfg()
    cores = orc.num_threads()
    sample = B_PER_GPU    # the C port finishes the whole per-GPU workload in seconds, so the sample is all of it
    wl = host_workload(cfg, sample, min(N_MAPS, sample), seed=0xB200 + 1)
    for _ in range(args.warmup):
        cpu_rollout(cfg, wl, sample, 30, seed=1)
    tot_t, tot_n = 0.0, 0
    for k in range(args.steps):
        dt, n = cpu_rollout(cfg, wl, sample, T_STEPS, seed=2 + k)
        tot_t += dt
        tot_n += n
    value = tot_n / tot_t
    desc = f"{sample} of {B_PER_GPU} envs x {T_STEPS} steps per step (same maps/placement rules, closed loop with A* evader)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from distributed_multi_agent_reinforcement_learning_b200 import _lib
    from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import BatchedPursuitEnv, EpisodeGraph, RolloutArena

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this arm has no CPU fallback")
    _lib.lib()   # fail loudly if the CUDA library is missing
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cfg = make_cfg()
    B, N, T, M = B_PER_GPU, N_AGENTS, T_STEPS, N_MAPS
    wl = host_workload(cfg, B, M, seed=0xB200 + 1 + 7919 * rank)
    env = BatchedPursuitEnv(cfg, B, device=dev, num_maps=M)
    env.set_maps(wl["grids"], wl["inflated"])
    env.set_state(wl["p_state"], wl["e_state"], wl["target"], wl["map_id"], time_step=0)
    env.set_target_tape(wl["tape"])
    env.start_episode()
    arena = RolloutArena(env.params, B, T, dev)
    snap = env.snapshot()
    graph = EpisodeGraph(env, arena, T, seed=0xB200 + rank, groups=STREAM_GROUPS)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: HBM-resident episode --------------------------------------------------------------------
    def timed_resident(n_iter):
        total = 0.0
        for _ in range(n_iter):
            env.restore(snap)
            flush.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            graph.replay()
            e.record()
            e.synchronize()
            total += s.elapsed_time(e)
        return total

    timed_resident(args.warmup)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    ms_total = timed_resident(args.steps)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    status = int(env.evader_status.max().item())
    assert status == 0, f"evader status {status}: search overflow or target tape exhausted"
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * N * T * args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers in, host metrics out ------------------------------------------------------------
    pin = {k: torch.from_numpy(np.ascontiguousarray(wl[k])).pin_memory() for k in ("p_state", "e_state", "target")}
    ep_reward_host = torch.zeros(B, dtype=torch.int64).pin_memory()
    coll_host = torch.zeros(B, dtype=torch.uint8).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in pin.values())
    d2h = ep_reward_host.numel() * 8 + coll_host.numel()

    def e2e_episode():
        env.p_state.copy_(pin["p_state"], non_blocking=True)
        env.e_state.copy_(pin["e_state"], non_blocking=True)
        env.target.copy_(pin["target"], non_blocking=True)
        for n_ in ("path_len", "time_step", "collision", "done", "tape_pos", "evader_status", "wf_n", "wf_mean", "wf_S", "wf_std"):
            getattr(env, n_).zero_()
        graph.replay()
        ep_reward_host.copy_(arena.raw_reward.sum(dim=(0, 2), dtype=torch.int64), non_blocking=True)
        coll_host.copy_(env.collision, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(args.warmup):
        e2e_episode()
    barrier()
    e2e_ms = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        e2e_episode()
        e.record()
        e.synchronize()
        e2e_ms += s.elapsed_time(e)
    barrier()
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * N * T * args.steps / (float(t.item()) * 1e-3)

    # ---- per-kernel durations: CUDA events around every launch of one eager episode (same stream) ----------
    env.restore(snap)
    timers = {}
    env.rollout_closed(arena, T, 0, seed=0xB200 + rank, timers=timers)
    torch.cuda.synchronize()
    kernel_table = {k: {"launches": len(v), "total_ms": sum(a.elapsed_time(b) for a, b in v)} for k, v in timers.items()}
    for v in kernel_table.values():
        v["avg_us"] = 1e3 * v["total_ms"] / v["launches"]
    env.restore(snap)
    ms_episode = ms_total / args.steps
    kernel_table["episode_graph_ms"] = ms_episode
    peak, peak_src = measured_peaks()
    rk = kernel_table["rollout_kernel(closed)"]
    steps_per_launch = T / rk["launches"]
    alg_bytes = SURVEY_BYTES_PER_AGENT_STEP * B * N * steps_per_launch
    achieved = alg_bytes / (rk["avg_us"] * 1e-6) / 1e9
    roofline = {"kernel": "rollout_kernel<8,1,closed> (observe + evader move + step + reward-norm + store, %d env steps per launch)" % steps_per_launch,
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "avg_launch_us": rk["avg_us"], "share_of_step": rk["total_ms"] / sum(v["total_ms"] for k, v in kernel_table.items() if isinstance(v, dict)),
                "note": "118 B/agent-step (SURVEY 8d) x 32768 agents x 10 steps per launch; working set of one launch fits L2, "
                        "so this kernel is latency-bound (fp64 RK4 division chains), not HBM-bound, at 4096 envs"}

    if rank == 0:
        # ---- CPU baseline: oracle port on the box's host cores, bounded sample ------------------------------
        from oracle import oracle as orc
        orc.build()
        cores = orc.num_threads()
        sample = B       # whole per-GPU workload: a few seconds of CPU work on all host threads
        wl_cpu = {k: (v[:sample] if k in ("p_state", "e_state", "target", "map_id", "tape") else v) for k, v in wl.items()}
        cpu_rollout(cfg, wl_cpu, sample, 10, seed=1)
        dt, n = cpu_rollout(cfg, wl_cpu, sample, T, seed=2)
        cpu = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {sample} of {B} envs x {T} steps, oracle/marl_oracle.c with OpenMP on {cores} threads ({dt:.1f} s)"}
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": graph.launches_per_replay * args.steps, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu, "kernel_ms_per_episode": kernel_table}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
