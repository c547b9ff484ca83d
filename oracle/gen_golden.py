"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE in the build
container (see oracle/ref_bootstrap.py).  The fixtures are committed because /root/reference does not exist
on the GPU box.  Re-run:  python -m oracle.gen_golden [env|edge|algo|all]

Everything stored is an observation of the reference treated as a black box: states, adjacency lists,
rewards, paths, A* call arguments/results, Welford outputs.  Nothing here restates reference arithmetic.
"""
import os
import sys
import random

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_bootstrap import load_reference, make_cfg, seed_all  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _params_of(cfg):
    return dict(
        W=cfg.map.map_size[0], H=cfg.map.map_size[1], N=cfg.env.num_defender, O=cfg.map.num_max_obstacle,
        max_steps=cfg.env.max_steps, difficulty=cfg.env.difficulty, sensor_beams=cfg.sensor.num_beams,
        sensor_radius=cfg.sensor.radius, e_extend_dis=cfg.attacker.extend_dis, e_sen_range=cfg.attacker.sen_range,
        d_step=cfg.defender.step_size, d_tau=cfg.defender.tau, d_vmax=cfg.defender.vmax,
        d_collision_radius=cfg.defender.collision_radius, d_comm_range=cfg.defender.comm_range,
        d_sen_range=cfg.defender.sen_range, e_step=cfg.attacker.step_size, e_tau=cfg.attacker.tau,
        e_vmax=cfg.attacker.vmax, e_collision_radius=cfg.attacker.collision_radius, resolution=cfg.map.resolution)


def _map_record(env, cfg):
    W, H = cfg.map.map_size
    grid = np.asarray(env.occupied_map.grid_map, dtype=np.uint8)
    inflated = np.asarray(env.inflated_map.grid_map != 0, dtype=np.uint8)
    boundary = np.asarray(env.boundary_map.grid_map != 0, dtype=np.uint8)
    bxy = np.asarray(env.boundary_map.obstacles, dtype=np.int32).reshape(-1, 2)
    raser = np.asarray(env.raser_map, dtype=np.uint8)  # [W,H,Ob]
    nb = cfg.sensor.num_beams
    beam_dir = np.array([[np.cos(b * 2 * np.pi / nb), np.sin(b * 2 * np.pi / nb)] for b in range(nb)], dtype=np.float64)
    action_table = np.asarray(env.defender_list[0].actions_mat, dtype=np.float64)
    return dict(grid=grid, inflated=inflated, boundary=boundary, boundary_xy=bxy,
                raser_packed=np.packbits(raser, axis=-1, bitorder="little"), raser_ob=np.int32(raser.shape[-1]),
                beam_dir=beam_dir, action_table=action_table)


class _Recorder:
    """Wraps reference methods to observe can_apply and every A* call."""

    def __init__(self, R, env):
        self.can_apply = []
        self.astar_calls = []
        self._orig_reward = env.defender_reward
        self.targets_drawn = []
        env.defender_reward = self._reward
        ast = env.attacker_list[0].astar
        self._orig_search = ast.searching
        ast.searching = self._search
        self.env = env
        self._orig_init_target = env.init_target
        env.init_target = self._init_target

    def _reward(self, state, next_state):
        r, ok = self._orig_reward(state, next_state)
        self.can_apply.append(bool(ok))
        return r, ok

    def _search(self, s_start, s_goal, obs):
        path, closed = self._orig_search(s_start=s_start, s_goal=s_goal, obs=obs)
        W, H = self.env.map_config.map_size
        blocked = np.zeros((W + 1, H + 1), dtype=np.uint8)
        for (x, y) in obs:
            blocked[int(x), int(y)] = 1
        self.astar_calls.append(dict(start=np.array(s_start, dtype=np.int32), goal=np.array(s_goal, dtype=np.int32),
                                     blocked=blocked, path=np.array(path, dtype=np.int32).reshape(-1, 2),
                                     n_closed=np.int32(len(closed))))
        return path, closed

    def _init_target(self, inflated_map):
        self._orig_init_target(inflated_map=inflated_map)
        self.targets_drawn.append(tuple(self.env.target[0]))


def run_episode(R, cfg, seed, policy, steps=None, tamper=None):
    """policy(env, t, rng) -> list[int] actions.  Returns the fixture dict."""
    seed_all(seed)
    env = R.pe.Pursuit_Env(cfg)
    env.reset()
    if tamper is not None:
        tamper(env)
    rec = _Recorder(R, env)
    norm = R.normalization.Normalization(shape=cfg.env.num_defender)
    N = cfg.env.num_defender
    T = steps or cfg.env.max_steps
    rng = np.random.RandomState(seed + 1000)
    out = dict(p_state=[], e_before=[], e_after=[], action=[], reward=[], can_apply=[], p_adj=[], o_adj=[], e_adj=[],
               collision=[], target=[], path_len=[], path_flat=[], r_norm=[], done=[], e_target_attr=[])
    for t in range(T):
        out["p_state"].append(env.get_state("defender"))
        out["e_before"].append(env.get_state("attacker")[0])
        out["p_adj"].append(env.communicate())
        o_adj, e_adj = env.sensor()
        out["o_adj"].append(o_adj)
        out["e_adj"].append(e_adj)
        out["target"].append(list(env.target[0]))
        out["e_target_attr"].append(list(env.attacker_list[0].target))
        paths = env.attacker_step()
        out["e_after"].append(env.get_state("attacker")[0])
        out["path_len"].append(len(paths[0]))
        out["path_flat"].extend(paths[0])
        a = [int(v) for v in policy(env, t, rng)]
        out["action"].append(a)
        rec.can_apply = []
        r, done, _ = env.step(a)
        out["reward"].append([int(v) for v in r])
        out["can_apply"].append(list(rec.can_apply))
        out["collision"].append(bool(env.collision))
        out["done"].append(bool(done))
        out["r_norm"].append(np.asarray(norm(r), dtype=np.float64))
    out["p_state"].append(env.get_state("defender"))
    out["p_adj"].append(env.communicate())
    o_adj, e_adj = env.sensor()
    out["o_adj"].append(o_adj)
    out["e_adj"].append(e_adj)
    out["target"].append(list(env.target[0]))
    fx = dict(_map_record(env, cfg))
    fx.update({f"param_{k}": np.asarray(v) for k, v in _params_of(cfg).items()})
    fx.update(
        p_state=np.asarray(out["p_state"], dtype=np.float64), e_before=np.asarray(out["e_before"], dtype=np.float64),
        e_after=np.asarray(out["e_after"], dtype=np.float64), action=np.asarray(out["action"], dtype=np.int32),
        reward=np.asarray(out["reward"], dtype=np.int32), can_apply=np.asarray(out["can_apply"], dtype=np.uint8),
        p_adj=np.asarray(out["p_adj"], dtype=np.uint8),
        o_adj_packed=np.packbits(np.asarray(out["o_adj"], dtype=np.uint8), axis=-1, bitorder="little"),
        e_adj=np.asarray(out["e_adj"], dtype=np.uint8).reshape(len(out["e_adj"]), N),
        collision=np.asarray(out["collision"], dtype=np.uint8), done=np.asarray(out["done"], dtype=np.uint8),
        target=np.asarray(out["target"], dtype=np.int32), e_target_attr=np.asarray(out["e_target_attr"], dtype=np.int32),
        path_len=np.asarray(out["path_len"], dtype=np.int32),
        path_flat=np.asarray(out["path_flat"], dtype=np.int32).reshape(-1, 2),
        r_norm=np.asarray(out["r_norm"], dtype=np.float64),
        targets_drawn=np.asarray(rec.targets_drawn, dtype=np.int32).reshape(-1, 2),
        seed=np.int32(seed))
    # A* calls: ragged -> flat
    fx["astar_n"] = np.int32(len(rec.astar_calls))
    if rec.astar_calls:
        fx["astar_start"] = np.stack([c["start"] for c in rec.astar_calls])
        fx["astar_goal"] = np.stack([c["goal"] for c in rec.astar_calls])
        fx["astar_blocked_packed"] = np.packbits(np.stack([c["blocked"] for c in rec.astar_calls]).reshape(len(rec.astar_calls), -1),
                                                 axis=-1, bitorder="little")
        fx["astar_path_len"] = np.array([len(c["path"]) for c in rec.astar_calls], dtype=np.int32)
        fx["astar_path_flat"] = np.concatenate([c["path"] for c in rec.astar_calls], axis=0)
        fx["astar_n_closed"] = np.array([c["n_closed"] for c in rec.astar_calls], dtype=np.int32)
    return fx


def mixed_policy(env, t, rng):
    """Half scripted chaser (pursuit_env.py:211-229 demon), half uniform random — random moves produce the
    pursuer/pursuer and pursuer/obstacle collisions the scripted policy avoids."""
    demon = env.demon()
    return [int(rng.randint(0, 9)) if rng.rand() < 0.5 else demon[i] for i in range(env.num_defender)]


def gen_env():
    R = load_reference()
    cases = [
        ("env_n4_s1", make_cfg(num_defender=4), 1, None),
        ("env_n8_s2", make_cfg(num_defender=8), 2, None),
        ("env_n15_s0", make_cfg(num_defender=15), 0, None),   # SURVEY §8(c) KAT seed
        ("env_n10_conf_s3", make_cfg(num_defender=10, max_steps=250, map_size=(60, 60), center=(30, 30),
                                     num_max_obstacle=110, extend_dis=3, depth=3), 3, 120),
    ]
    for name, cfg, seed, steps in cases:
        fx = run_episode(R, cfg, seed, mixed_policy, steps=steps)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **fx)
        print(name, "total reward", int(fx["reward"].sum()), "collision", int(fx["collision"][-1]),
              "astar calls", int(fx["astar_n"]), "targets drawn", len(fx["targets_drawn"]), os.path.getsize(path), "B")


def gen_kat():
    """The survey's known-answer episode: seed 0, N=15, demon() policy for 150 steps -> total reward -748."""
    R = load_reference()
    fx = run_episode(R, make_cfg(num_defender=15), 0, lambda env, t, rng: env.demon())
    assert int(fx["reward"].sum()) == -748 and int(fx["collision"][-1]) == 1, int(fx["reward"].sum())
    np.savez_compressed(os.path.join(GOLDEN_DIR, "env_n15_s0_demon.npz"), **fx)
    print("kat ok: -748")


def gen_edge():
    """Crafted states at the map edge / in corners / on top of each other, exercising the in-place clip that is
    visible to later agents only (pursuit_env.py:143-145) and evader capture."""
    R = load_reference()
    cfg = make_cfg(num_defender=8, max_steps=40)

    def tamper(env):
        # pairs (A moving out of the map at full speed, B resting 0.35 inside): A's raw proposal is 0.55 from B,
        # A's clipped proposal 0.45 -> B is rejected iff it is evaluated AFTER A (index order).
        spots = [(0.1, 10.0, -2.0, 0.0), (0.45, 10.0, 0.0, 0.0),       # A idx0 < B idx1  -> rewards [0,-1]
                 (0.45, 20.0, 0.0, 0.0), (0.1, 20.0, -2.0, 0.0),       # B idx2 < A idx3  -> rewards [0, 0]
                 (30.0, 53.9, 0.0, 2.0), (30.0, 53.55, 0.0, 0.0),      # top edge
                 (58.55, 40.0, 0.0, 0.0), (58.9, 40.0, 2.0, 0.0)]      # right edge, B first
        for d, (x, y, vx, vy) in zip(env.defender_list, spots):
            assert env.occupied_map.is_unoccupied((x, y))
            d.x, d.y, d.vx, d.vy = float(x), float(y), float(vx), float(vy)
        a = env.attacker_list[0]
        a.x, a.y = 0.3, 10.2

    script = {0: [4, 8, 8, 4, 2, 8, 8, 0]}

    def policy(env, t, rng):
        if t in script:
            return script[t]
        base = [4, 0, 0, 4, 2, 6, 4, 0]
        return [b if rng.rand() < 0.6 else int(rng.randint(0, 9)) for b in base]

    fx = run_episode(R, cfg, 7, policy, steps=40, tamper=tamper)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "env_edge_n8_s7.npz"), **fx)
    print("edge: first-step rewards", fx["reward"][0], "total", fx["reward"].sum(axis=0), "rejected", int((fx["can_apply"] == 0).sum()))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    if what in ("env", "all"):
        gen_env()
    if what in ("kat", "all"):
        gen_kat()
    if what in ("edge", "all"):
        gen_edge()
    if what in ("algo", "all"):
        from oracle.gen_golden_algo import gen_algo
        gen_algo()
