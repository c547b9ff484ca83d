"""TEST INFRASTRUCTURE — golden vectors of the evaluator path, produced by EXECUTING THE UNMODIFIED REFERENCE
(`evaluate(env, actor, cfg)`, evaluator.py:106-201: one deterministic / argmax episode of the actor in `Pursuit_Env`) in
the build container.  `ray.remote` is stubbed to the identity (it only schedules; no arithmetic).  Records what the function
returns ([episode_reward, step]) and, by wrapping `env.step`, every action vector and reward vector of the episode.

The weights are the reference's initial weights for torch seed `seed` (the product reproduces them bit for bit), the env is
reset by `evaluate` itself from the seeded global RNGs.  Re-run:  python -m oracle.gen_golden_eval
"""
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_bootstrap import load_reference, make_cfg, seed_all  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference_evaluator():
    R = load_reference()
    if "ray" not in sys.modules:
        ray = types.ModuleType("ray")

        def remote(*args, **kw):
            if len(args) == 1 and callable(args[0]) and not kw:
                return args[0]
            return lambda f: f
        ray.remote = remote
        sys.modules["ray"] = ray
    if "environment.pursuit_evasion_game.gif_plotting" not in sys.modules:
        m = types.ModuleType("environment.pursuit_evasion_game.gif_plotting")
        m.sim_moving = None
        sys.modules["environment.pursuit_evasion_game.gif_plotting"] = m
    import evaluator as ref_eval
    return R, ref_eval


def gen(R, ref_eval, n_def, depth, T, seed, embedding_dim):
    import torch
    cfg = make_cfg(num_defender=n_def, depth=depth, max_steps=T, embedding_dim=embedding_dim)
    seed_all(seed)
    torch.set_grad_enabled(False)
    agent = R.mappo.MAPPO(cfg, None, None, "Evaluator")
    env = R.pe.Pursuit_Env(cfg)
    actions, rewards = [], []
    orig_step = env.step

    def step(a):
        actions.append(np.asarray(a).astype(np.int32).copy())
        out = orig_step(a)
        rewards.append(np.asarray(out[0], dtype=np.int32).copy())
        return out
    env.step = step
    episode_reward, last_step = ref_eval.evaluate(env, agent.actor, cfg)
    return dict(ret=np.array([episode_reward, last_step], np.int64), actions=np.stack(actions), rewards=np.stack(rewards),
                collision=np.array([bool(env.collision)]), p_final=np.asarray(env.get_state("defender"), np.float64),
                e_final=np.asarray(env.get_state("attacker"), np.float64),
                meta=np.array([n_def, depth, T, seed, embedding_dim], np.int64))


def main():
    R, ref_eval = load_reference_evaluator()
    for n_def, depth, T, seed, E in ((4, 1, 40, 41, 128), (6, 3, 30, 43, 128)):
        fx = gen(R, ref_eval, n_def, depth, T, seed, E)
        path = os.path.join(GOLDEN_DIR, f"eval_n{n_def}_d{depth}.npz")
        np.savez_compressed(path, **fx)
        print(path, "ret", fx["ret"], "actions", fx["actions"].shape, "collision", fx["collision"])


if __name__ == "__main__":
    main()
