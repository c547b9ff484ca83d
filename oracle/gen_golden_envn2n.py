"""TEST INFRASTRUCTURE — golden vectors of the 2-D N-vs-E particle env (environment/env_n2n/particle_env.py),
produced by EXECUTING THE UNMODIFIED REFERENCE in the build container.  Re-run:  python -m oracle.gen_golden_envn2n

Per step (black-box observations only): full pursuer / evader state and active flags before the step, the adjacency matrices
`get_adj_mat` gives (pursuer-pursuer at comm range, pursuer-evader at sensor range), `choose_evader('actor')`, the evaders'
commanded actions (from the reference's own SLSQP evader `eva.e_f` when `evader="slsqp"` — scipy is third-party arithmetic, so the
action is an input tape for the port, SURVEY §8c — or scripted uniform actions), the state after `evader_step`, the pursuers'
discrete actions, and reward / done / active / state after `step`.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_bootstrap import load_reference, seed_all  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_envn2n():
    load_reference()        # installs the import stubs (sko, matplotlib, ...)
    import environment.env_n2n.particle_env as pe2
    return pe2


def _state(env, pursuer):
    return np.array(env.get_team_state(pursuer, rules=False), dtype=np.float64)


def _full(env, pursuer):
    lst, idx = (env.p_list, env.p_idx) if pursuer else (env.e_list, env.e_idx)
    return np.array([[lst[f"{i}"].x, lst[f"{i}"].y, lst[f"{i}"].phi, lst[f"{i}"].v] for i in idx], dtype=np.float64)


def _active(env, pursuer):
    lst, idx = (env.p_list, env.p_idx) if pursuer else (env.e_list, env.e_idx)
    return np.array([1 if lst[f"{i}"].active else 0 for i in idx], dtype=np.uint8)


def run_episode(pe2, n, e, seed, steps, evader="slsqp", crowd=False):
    seed_all(seed)
    env = pe2.ParticleEnv()
    env.initialize(n, e)
    env.reset()
    if crowd:        # pull everybody together so that kills / team collisions happen within the episode
        for i in env.p_idx:
            a = env.p_list[f"{i}"]
            a.x, a.y = 10 + 0.35 * (a.x - 10), 10 + 0.35 * (a.y - 10)
            a.phi = (2 * np.random.rand() - 1) * np.pi
        for i in env.e_idx:
            a = env.e_list[f"{i}"]
            a.x, a.y = 10 + np.random.normal(0, 1.2), 10 + np.random.normal(0, 1.2)
            a.phi = (2 * np.random.rand() - 1) * np.pi
    rng = np.random.RandomState(seed + 77)
    keys = ("p_before", "e_before", "p_active_before", "e_active_before", "pp_adj", "pe_adj", "assign", "e_action", "e_moved",
            "action", "reward", "done", "p_after", "e_after", "p_active", "e_active")
    rec = {k: [] for k in keys}
    captured = []
    orig_e_f = pe2.eva.e_f

    def spy(**kw):
        out = orig_e_f(**kw)
        captured.append(float(out))
        return out

    pe2.eva.e_f = spy
    try:
        for t in range(steps):
            p_state, e_state = _state(env, True), _state(env, False)
            rec["p_before"].append(_full(env, True))
            rec["e_before"].append(_full(env, False))
            rec["p_active_before"].append(_active(env, True))
            rec["e_active_before"].append(_active(env, False))
            rec["pp_adj"].append(env.get_adj_mat(p_state, p_state, env.p_comm_range, True).astype(np.uint8))
            rec["pe_adj"].append(env.get_adj_mat(p_state, e_state, env.p_sen_range, True).astype(np.uint8))
            rec["assign"].append(np.asarray(env.choose_evader("actor"), dtype=np.uint8))
            e_act = np.zeros(e, np.float64)
            del captured[:]
            if evader == "slsqp":
                env.evader_step(env.get_team_state(True, rules=False))       # e_f is called once per ACTIVE evader, in index order
                k = 0
                for j in env.e_idx:
                    if rec["e_active_before"][-1][j]:
                        e_act[j] = captured[k]
                        k += 1
            else:
                for j in env.e_idx:
                    ev = env.e_list[f"{j}"]
                    e_act[j] = rng.uniform(-1, 1)
                    if ev.active:
                        ev.step(env.step_size, e_act[j])
            rec["e_action"].append(e_act)
            rec["e_moved"].append(_full(env, False))
            a = rng.randint(0, 9, n).astype(np.int32)
            rec["action"].append(a)
            reward, done, active = env.step(a)
            rec["reward"].append(np.array(reward, dtype=np.int32))
            rec["done"].append(np.uint8(done))
            rec["p_after"].append(_full(env, True))
            rec["e_after"].append(_full(env, False))
            rec["p_active"].append(np.array(active, dtype=np.uint8))
            rec["e_active"].append(_active(env, False))
    finally:
        pe2.eva.e_f = orig_e_f
    out = {k: np.stack(v) for k, v in rec.items()}
    out.update(target=np.array(env.target, dtype=np.float64), n=np.int32(n), e=np.int32(e), episode_limit=np.int32(env.episode_limit),
               p_vmax=np.float64(env.p_vmax), e_vmax=np.float64(env.e_vmax), kill_radius=np.float64(env.kill_radius),
               ang_lmt=np.float64(env.ang_lmt), step_size=np.float64(env.step_size), p_comm_range=np.float64(env.p_comm_range),
               p_sen_range=np.float64(env.p_sen_range))
    return out


def reset_record(pe2, n, e, seed):
    """ParticleEnv.reset() (particle_env.py:195-246) for a fixed numpy seed: initial state for the reset parity test."""
    seed_all(seed)
    env = pe2.ParticleEnv()
    env.initialize(n, e)
    env.reset()
    return dict(p_state=_full(env, True), e_state=_full(env, False), target=np.array(env.target), n=np.int32(n), e=np.int32(e),
                seed=np.int32(seed))


def main():
    pe2 = load_envn2n()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    jobs = [("envn2n_n4_e1_s1", dict(n=4, e=1, seed=1, steps=60)),
            ("envn2n_n6_e2_s2", dict(n=6, e=2, seed=2, steps=60)),
            ("envn2n_n8_e3_s3_scripted", dict(n=8, e=3, seed=3, steps=80, evader="scripted")),
            ("envn2n_n15_e4_s4_crowd", dict(n=15, e=4, seed=4, steps=50, evader="scripted", crowd=True)),
            ("envn2n_n10_e3_s5_crowd", dict(n=10, e=3, seed=5, steps=50, crowd=True))]
    for name, kw in jobs:
        fx = run_episode(pe2, **kw)
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **fx)
        print(name, "steps", len(fx["done"]), "reward sum", int(fx["reward"].sum()), "pursuers alive", int(fx["p_active"][-1].sum()),
              "evaders alive", int(fx["e_active"][-1].sum()), "done at", int(np.argmax(fx["done"])) if fx["done"].any() else None)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "envn2n_reset_n5_e2_s9.npz"), **reset_record(pe2, 5, 2, 9))


if __name__ == "__main__":
    main()
