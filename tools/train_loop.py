"""End-to-end driver on one GPU (or one per rank under torchrun): what main.py + runner.py do with Ray actors
(main.py:79-129: broadcast weights -> workers explore -> learners train -> gradient sum -> Adam), here as
reset -> batched rollout (all envs, networks in the loop) -> MAPPO.train -> all-reduce + Adam, per iteration.

    python tools/train_loop.py --envs 512 --iters 10
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_loop.py --envs 512

Prints one line per iteration: mean episode reward per env (sum over pursuers and steps, the quantity the reference's
evaluator tracks), fraction of envs with a collision, losses, rollout / train wall times."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_multi_agent_reinforcement_learning_b200 import default_config, parallel  # noqa: E402
from distributed_multi_agent_reinforcement_learning_b200.mappo_parallel import MAPPO  # noqa: E402
from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import BatchedPursuitEnv, RolloutArena  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=512)
    ap.add_argument("--maps", type=int, default=32)
    ap.add_argument("--pursuers", type=int, default=8)
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--depth", type=int, default=1)
    ap.add_argument("--allreduce", default=None, choices=[None, "update", "minibatch"],
                    help="gradient exchange: once per update (reference, default) or once per PPO minibatch")
    ap.add_argument("--host-reset", action="store_true", help="generate episodes on the host with the reference's generators (slow)")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
    cfg = default_config(env__num_defender=args.pursuers, env__max_steps=args.steps, algo__depth=args.depth,
                         algo__learner_device=str(dev), algo__worker_device=str(dev))
    B, T, N = args.envs, args.steps, args.pursuers
    env = BatchedPursuitEnv(cfg, B, device=dev, num_maps=args.maps)
    arena = RolloutArena(env.params, B, T, dev)
    torch.manual_seed(0)
    mappo = MAPPO(cfg, B, max(1, round(B / 10)), "Learner")
    mappo.sync_weights(0)
    total_steps = 0
    W, H = cfg.map.map_size
    for it in range(args.iters):
        t0 = time.perf_counter()
        if args.host_reset:
            env.reset(seed=1000 * rank + it)                   # host generation with the reference's placement rules
            g = np.random.default_rng(7 + 1000 * rank + it)
            env.set_target_tape(np.stack([g.integers(0, W, (B, 16)), g.integers(0, H, (B, 16))], -1).astype(np.int32))
        else:
            env.reset_device(seed=1000 * rank + it)            # same rules, generated on the GPU (csrc/reset_kernels.cu)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if not args.host_reset:
            assert int(env.reset_fail.sum()) == 0, "device reset could not place every pursuer / evader"
        tb = mappo.rollout_batched(env, arena, T, seed=it * world + rank)     # distinct sampling streams per replica and iteration
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        assert int(env.evader_status.max()) == 0, "evader search overflow / target tape exhausted"
        ep_reward = float(arena.raw_reward.sum(dim=(0, 2)).float().mean())
        collided = float(env.collision.float().mean())
        total_steps += world * B * T
        for _ in range(int(cfg.algo.epochs)):                       # main.py:99: K epochs over the same buffers, one Adam step each
            obj_c, obj_a, _, _ = mappo.train(tb, total_steps, return_numpy=False, allreduce=args.allreduce)
            mappo.update(total_steps)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        stats = torch.tensor([ep_reward, collided, obj_c, obj_a], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(stats)
            stats /= world
        if rank == 0:
            print(f"iter {it:3d}  steps {total_steps:9d}  ep_reward {stats[0]:9.2f}  collided {stats[1]:5.2f}  objC {stats[2]:8.4f}  "
                  f"objA {stats[3]:8.4f}  reset {t1 - t0:5.2f}s  rollout {1e3 * (t2 - t1):7.1f}ms  train {1e3 * (t3 - t2):7.1f}ms", flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
