"""TEST INFRASTRUCTURE — golden vectors of the 3-D particle env, produced by EXECUTING THE UNMODIFIED
REFERENCE `environment/env_3d/particle_env.py` in the build container.  Re-run:  python -m oracle.gen_golden_env3d

Recorded per step (black-box observations only): full pursuer / evader state before and after, actions, the evader's
commanded action (returned by the reference's own SLSQP evader `eva.e_f`, scipy — supplied to the port as an input
tape because scipy's SLSQP is third-party arithmetic), rewards, active flags, done, and the adjacency
matrices `get_adj_mat` gives for the pursuer-pursuer (comm range) and pursuer-evader (sensor range) relations.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_bootstrap import load_reference, seed_all  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_env3d():
    load_reference()        # installs the import stubs (sko, matplotlib, ...)
    import environment.env_3d.particle_env as pe3
    return pe3


def _full_state(env, pursuer):
    return np.array(env.get_team_state(pursuer, rules=False), dtype=np.float64)


def _active(env, pursuer):
    lst, idx = (env.p_list, env.p_idx) if pursuer else (env.e_list, env.e_idx)
    return np.array([1 if lst[f"{i}"].active else 0 for i in idx], dtype=np.uint8)


def run_episode(pe3, n, seed, steps, evader="slsqp", crowd=False, action_scale=1.0):
    seed_all(seed)
    env = pe3.ParticleEnv()
    env.initialize(n)
    if n <= 8 and not crowd:
        env.reset()
    else:
        # reference reset() cannot place many pursuers (min_dist 4 in a 10^3 box): same object construction with
        # relaxed spacing, everything else as particle_env.py:135-199
        env.reset.__func__  # noqa: B018  (keep the attribute access honest: we do not call it)
        env.target = [np.random.rand() * 20, np.random.rand() * 20, np.random.rand() * 20]
        env.p_list, env.p_idx, env.e_list, env.e_idx = dict(), list(), dict(), list()
        env.time_step = 0
        centre = np.array([10.0, 10.0, 10.0])
        for i in range(n):
            pos = (centre + np.random.normal(0, 0.6 if crowd else 2.0, 3)).clip(5, 15)
            env.p_list[f"{i}"] = pe3.Pursuer(i, pos[0], pos[1], pos[2], (2 * np.random.rand() - 1) * np.pi,
                                             (2 * np.random.rand() - 1) * np.pi / 2, 0, env.p_vmax, env.p_sen_range,
                                             env.p_comm_range, env.ang_lmt, env.v_lmt)
            env.p_idx.append(i)
        e_pos = centre + np.random.normal(0, 1.0, 3) if crowd else 20 - np.array(env.target)
        env.e_list["0"] = pe3.Evader(0, e_pos[0], e_pos[1], e_pos[2], (2 * np.random.rand() - 1) * np.pi,
                                     (2 * np.random.rand() - 1) * np.pi / 2, 0, env.e_vmax, env.e_sen_range,
                                     env.e_comm_range, env.ang_lmt, env.v_lmt)
        env.e_idx.append(0)
    rng = np.random.RandomState(seed + 77)
    rec = dict(p_before=[], e_before=[], p_active_before=[], e_active_before=[], action=[], e_action=[], e_moved=[],
               p_after=[], e_after=[], p_active=[], e_active=[], reward=[], done=[], pp_adj=[], pe_adj=[])
    captured = {}
    orig_e_f = pe3.eva.e_f

    def spy(**kw):
        out = orig_e_f(**kw)
        captured["a"] = np.array(out, dtype=np.float64)
        return out

    pe3.eva.e_f = spy
    try:
        for t in range(steps):
            p_state = _full_state(env, True)
            rec["p_before"].append(p_state)
            rec["e_before"].append(_full_state(env, False)[0])
            rec["p_active_before"].append(_active(env, True))
            rec["e_active_before"].append(_active(env, False)[0])
            rec["pp_adj"].append(env.get_adj_mat(p_state, p_state, env.p_comm_range, True).astype(np.uint8))
            rec["pe_adj"].append(env.get_adj_mat(p_state, _full_state(env, False), env.p_sen_range, True).astype(np.uint8))
            # evader move (particle_env.py:348-373): its action comes from the reference's own SLSQP evader
            captured.pop("a", None)
            if evader == "slsqp":
                alive = env.get_team_state(True)            # rules=True: only active pursuers are visible to it
                if alive:
                    env.evader_step(alive)
            if "a" not in captured:                         # scripted evader (random commanded action)
                captured["a"] = rng.uniform(-1, 1, 3)
                env.e_list["0"].step(env.step_size, captured["a"])
            rec["e_action"].append(captured["a"].copy())
            rec["e_moved"].append(_full_state(env, False)[0])
            a = rng.uniform(-1, 1, (n, 3)) * action_scale
            rec["action"].append(a)
            reward, done, active = env.step(a)
            rec["reward"].append(np.array(reward, dtype=np.int32))
            rec["done"].append(np.uint8(done))
            rec["p_after"].append(_full_state(env, True))
            rec["e_after"].append(_full_state(env, False)[0])
            rec["p_active"].append(np.array(active, dtype=np.uint8))
            rec["e_active"].append(_active(env, False)[0])
    finally:
        pe3.eva.e_f = orig_e_f
    out = {k: np.stack(v) for k, v in rec.items()}
    out.update(target=np.array(env.target, dtype=np.float64), n=np.int32(n), max_step=np.int32(env.max_step),
               p_vmax=np.float64(env.p_vmax), e_vmax=np.float64(env.e_vmax), kill_radius=np.float64(env.kill_radius),
               ang_lmt=np.float64(env.ang_lmt), v_lmt=np.float64(env.v_lmt), step_size=np.float64(env.step_size),
               p_comm_range=np.float64(env.p_comm_range), p_sen_range=np.float64(env.p_sen_range))
    return out


def reset_record(pe3, n, seed):
    """ParticleEnv.reset() (particle_env.py:135-199) for a fixed numpy seed: initial state for the reset parity test."""
    seed_all(seed)
    env = pe3.ParticleEnv()
    env.initialize(n)
    env.reset()
    return dict(p_state=_full_state(env, True), e_state=_full_state(env, False)[0], target=np.array(env.target), n=np.int32(n),
                seed=np.int32(seed))


def main():
    pe3 = load_env3d()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    jobs = [("env3d_n3_s1", dict(n=3, seed=1, steps=60)),
            ("env3d_n5_s2", dict(n=5, seed=2, steps=60)),
            ("env3d_n8_s3_scripted", dict(n=8, seed=3, steps=80, evader="scripted")),
            ("env3d_n32_s4_crowd", dict(n=32, seed=4, steps=40, evader="scripted", crowd=True)),
            ("env3d_n12_s5_crowd", dict(n=12, seed=5, steps=60, crowd=True))]
    for name, kw in jobs:
        fx = run_episode(pe3, **kw)
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **fx)
        print(name, "steps", len(fx["done"]), "reward sum", int(fx["reward"].sum()), "alive at end",
              int(fx["p_active"][-1].sum()), "evader alive", int(fx["e_active"][-1]), "done at",
              int(np.argmax(fx["done"])) if fx["done"].any() else None)
    rs = reset_record(pe3, 4, 9)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "env3d_reset_n4_s9.npz"), **rs)


if __name__ == "__main__":
    main()
