"""TEST INFRASTRUCTURE — numpy-facing wrappers around oracle/libmarl_oracle.so (the plain-C restatement of the
reference hot path, oracle/marl_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmarl_oracle.so")


class EnvParams(C.Structure):
    """Mirror of marl_env_params (include/marl_b200.h)."""
    _fields_ = [(n, C.c_int32) for n in ("W", "H", "N", "O", "max_steps", "difficulty", "sensor_beams",
                                         "sensor_radius", "e_extend_dis", "e_sen_range")] + \
               [(n, C.c_double) for n in ("d_step", "d_tau", "d_vmax", "d_collision_radius", "d_comm_range",
                                          "d_sen_range", "e_step", "e_tau", "e_vmax", "e_collision_radius",
                                          "resolution")]

    @classmethod
    def from_dict(cls, d):
        p = cls()
        for name, ctype in cls._fields_:
            v = d[name]
            setattr(p, name, int(v) if ctype is C.c_int32 else float(v))
        return p

    @classmethod
    def from_fixture(cls, fx):
        return cls.from_dict({k[len("param_"):]: fx[k].item() for k in fx.files if k.startswith("param_")})


def build(force=False):
    src = os.path.join(_HERE, "marl_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "libmarl_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_boundary_map.restype = C.c_int32
        _lib.orc_astar.restype = C.c_int32
        _lib.orc_replan.restype = C.c_int32
        _lib.orc_evader_step.restype = C.c_int32
        _lib.orc_num_threads.restype = C.c_int32
    return _lib


def _p(a, t=None):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def dynamic(state, ux, uy, tau, h):
    s = _c(state, np.float64)
    out = np.empty(4, np.float64)
    lib().orc_dynamic(_p(s), C.c_double(ux), C.c_double(uy), C.c_double(tau), C.c_double(h), _p(out))
    return out


def env_step(p, p_state, e_state, action, grid, action_table, time_step=0, collision=0):
    """One Pursuit_Env.step for one env. Returns dict(p_state, reward, can_apply, collision, time_step, done)."""
    ps = _c(p_state, np.float64).copy()
    es = _c(e_state, np.float64)
    act = _c(action, np.int32)
    g = _c(grid, np.uint8)
    at = _c(action_table, np.float64)
    reward = np.zeros(p.N, np.int32)
    can = np.zeros(p.N, np.uint8)
    col = C.c_uint8(collision)
    ts = C.c_int32(time_step)
    done = C.c_uint8(0)
    lib().orc_env_step(C.byref(p), _p(ps), _p(es), _p(act), _p(g), _p(at), _p(reward), _p(can), C.byref(col),
                       C.byref(ts), C.byref(done))
    return dict(p_state=ps, reward=reward, can_apply=can, collision=col.value, time_step=ts.value, done=done.value)


def communicate(p, p_state):
    ps = _c(p_state, np.float64)
    adj = np.zeros((p.N, p.N), np.uint8)
    lib().orc_communicate(C.byref(p), _p(ps), _p(adj))
    return adj


def sensor(p, p_state, e_state, grid, raser):
    """raser u8 [W,H,Ob] dense. Returns (o_adj u8 [N,O], e_adj u8 [N])."""
    ps = _c(p_state, np.float64)
    es = _c(e_state, np.float64)
    g = _c(grid, np.uint8)
    rs = _c(raser, np.uint8)
    o_adj = np.zeros((p.N, p.O), np.uint8)
    e_adj = np.zeros(p.N, np.uint8)
    lib().orc_sensor(C.byref(p), _p(ps), _p(es), _p(g), _p(rs), C.c_int32(rs.shape[-1]), _p(o_adj), _p(e_adj))
    return o_adj, e_adj


def boundary_map(p, grid, cap=None):
    g = _c(grid, np.uint8)
    cap = cap or p.W * p.H
    b = np.zeros((p.W, p.H), np.uint8)
    xy = np.zeros((cap, 2), np.int32)
    n = lib().orc_boundary_map(C.byref(p), _p(g), _p(b), _p(xy), C.c_int32(cap))
    return b, xy[:min(n, cap)].copy(), n


def raser_map(p, boundary, xy, beam_dir):
    b = _c(boundary, np.uint8)
    xy = _c(xy, np.int32)
    bd = _c(beam_dir, np.float64)
    ob = xy.shape[0]
    out = np.zeros((p.W, p.H, ob), np.uint8)
    lib().orc_raser_map(C.byref(p), _p(b), _p(xy), C.c_int32(ob), _p(bd), _p(out))
    return out


def dilate(p, grid, e):
    g = _c(grid, np.uint8)
    out = np.zeros((p.W, p.H), np.uint8)
    lib().orc_dilate(C.byref(p), _p(g), C.c_int32(e), _p(out))
    return out


def astar(p, blocked, start, goal, cap=4096):
    bl = _c(blocked, np.uint8)
    assert bl.shape == (p.W + 1, p.H + 1)
    path = np.zeros((cap, 2), np.int16)
    nclosed = C.c_int32(0)
    n = lib().orc_astar(C.byref(p), _p(bl), C.c_int(int(start[0])), C.c_int(int(start[1])), C.c_int(int(goal[0])),
                        C.c_int(int(goal[1])), _p(path), C.c_int32(cap), C.byref(nclosed))
    assert n > 0, "path overflow"
    return path[:n].astype(np.int32), nclosed.value


def rescan(p, grid, e, p_state, px, py):
    g = _c(grid, np.uint8)
    ps = _c(p_state, np.float64)
    out = np.zeros((p.W + 1, p.H + 1), np.uint8)
    lib().orc_rescan(C.byref(p), _p(g), C.c_int(e), _p(ps), C.c_int(px), C.c_int(py), _p(out))
    return out


class EvaderState:
    """Mutable per-env evader state used by evader_step."""

    def __init__(self, e_state, target, path=None, cap=4096):
        self.e_state = _c(e_state, np.float64).copy()
        self.target = _c(target, np.int32).copy()
        self.path = np.zeros((cap, 2), np.int16)
        self.path_len = C.c_int32(0)
        if path is not None:
            path = np.asarray(path, np.int16).reshape(-1, 2)
            self.path[:len(path)] = path
            self.path_len = C.c_int32(len(path))
        self.tape_pos = C.c_int32(0)
        self.cap = cap


def evader_step(p, ev, p_state, time_step, grid, inflated, tape):
    ps = _c(p_state, np.float64)
    g = _c(grid, np.uint8)
    inf = _c(inflated, np.uint8)
    tp = _c(tape, np.int32).reshape(-1, 2)
    rc = lib().orc_evader_step(C.byref(p), _p(ev.e_state), _p(ps), _p(ev.target), _p(ev.path), C.byref(ev.path_len),
                               C.c_int32(ev.cap), C.c_int32(time_step), _p(g), _p(inf), _p(tp), C.c_int32(len(tp)),
                               C.byref(ev.tape_pos))
    return rc


class Welford:
    def __init__(self, N):
        self.N = N
        self.n = C.c_int64(0)
        self.mean = np.zeros(N, np.float64)
        self.S = np.zeros(N, np.float64)
        self.std = np.zeros(N, np.float64)

    def __call__(self, x, update=True):
        x = _c(x, np.int32)
        out = np.zeros(self.N, np.float32)
        lib().orc_welford(C.c_int32(self.N), _p(x), C.byref(self.n), _p(self.mean), _p(self.S), _p(self.std), _p(out),
                          C.c_int(1 if update else 0))
        return out


def gae(r, v, active, gamma, lamda, use_adv_norm=True):
    r = _c(r, np.float32)
    v = _c(v, np.float32)
    active = _c(active, np.float32)
    B, T, N = r.shape
    adv = np.zeros_like(r)
    vt = np.zeros_like(r)
    lib().orc_gae(C.c_int32(B), C.c_int32(T), C.c_int32(N), _p(r), _p(v), _p(active), C.c_float(gamma),
                  C.c_float(gamma * lamda), C.c_int(1 if use_adv_norm else 0), _p(adv), _p(vt))
    return adv, vt


def rollout_iteration(p, st):
    """st: dict of contiguous numpy arrays (see orc_rollout_iteration). Mutates in place."""
    B = st["p_state"].shape[0]
    mid = st.get("map_id")
    lib().orc_rollout_iteration(
        C.byref(p), C.c_int32(B), _p(st["p_state"]), _p(st["e_before"]), _p(st["e_after"]), _p(st["action"]),
        _p(st["grid"]), _p(st["raser"]), _p(st["ob_count"]), C.c_int32(st["raser"].shape[-1]),
        _p(mid) if mid is not None else None, _p(st["action_table"]), _p(st["p_adj"]), _p(st["o_adj"]),
        _p(st["e_adj"]), _p(st["reward"]), _p(st["can_apply"]), _p(st["collision"]), _p(st["time_step"]),
        _p(st["done"]), _p(st["wf_n"]), _p(st["wf_mean"]), _p(st["wf_S"]), _p(st["wf_std"]), _p(st["r_norm"]))


def rollout_iteration_closed(p, st):
    """As rollout_iteration but with the A* evader in the loop (st carries e_state/target/path/... arrays)."""
    B = st["p_state"].shape[0]
    lib().orc_rollout_iteration_closed(
        C.byref(p), C.c_int32(B), _p(st["p_state"]), _p(st["e_state"]), _p(st["target"]), _p(st["path"]),
        _p(st["path_len"]), C.c_int32(st["path"].shape[1]), _p(st["action"]), _p(st["grid"]), _p(st["inflated"]),
        _p(st["raser"]), _p(st["ob_count"]), C.c_int32(st["raser"].shape[-1]), _p(st["map_id"]),
        _p(st["action_table"]), _p(st["tape"]), C.c_int32(st["tape"].shape[1]), _p(st["tape_pos"]), _p(st["p_adj"]),
        _p(st["o_adj"]), _p(st["e_adj"]), _p(st["reward"]), _p(st["can_apply"]), _p(st["collision"]),
        _p(st["time_step"]), _p(st["done"]), _p(st["wf_n"]), _p(st["wf_mean"]), _p(st["wf_S"]), _p(st["wf_std"]),
        _p(st["r_norm"]), _p(st["status"]))


def observe_batch(p, st):
    """communicate + sensor for all envs of `st` (fills st['p_adj'], st['o_adj'], st['e_adj'])."""
    B = st["p_state"].shape[0]
    lib().orc_observe_batch(C.byref(p), C.c_int32(B), _p(st["p_state"]), _p(st["e_state"]), _p(st["grid"]),
                            _p(st["raser"]), _p(st["ob_count"]), C.c_int32(st["raser"].shape[-1]), _p(st["map_id"]),
                            _p(st["p_adj"]), _p(st["o_adj"]), _p(st["e_adj"]))


def num_threads():
    return lib().orc_num_threads()


# ---------------------------------------------------------------------------------------------- 3-D particle env
class Env3dParams(C.Structure):
    """Mirror of marl_env3d_params (include/marl_b200.h)."""
    _fields_ = [("N", C.c_int32), ("max_step", C.c_int32)] + \
               [(n, C.c_double) for n in ("p_vmax", "e_vmax", "kill_radius", "ang_lmt", "v_lmt", "step_size",
                                          "comm_range", "sen_range")]

    @classmethod
    def from_dict(cls, d):
        p = cls()
        for name, ctype in cls._fields_:
            setattr(p, name, int(d[name]) if ctype is C.c_int32 else float(d[name]))
        return p

    @classmethod
    def from_fixture(cls, fx):
        return cls.from_dict(dict(N=fx["n"], max_step=fx["max_step"], p_vmax=fx["p_vmax"], e_vmax=fx["e_vmax"],
                                  kill_radius=fx["kill_radius"], ang_lmt=fx["ang_lmt"], v_lmt=fx["v_lmt"],
                                  step_size=fx["step_size"], comm_range=fx["p_comm_range"], sen_range=fx["p_sen_range"]))


def point_step(state, action, v_max, ang_lmt, v_lmt, step_size):
    s = _c(state, np.float64).copy()
    a = _c(action, np.float64)
    lib().orc_point_step(_p(s), _p(a), C.c_double(v_max), C.c_double(ang_lmt), C.c_double(v_lmt), C.c_double(step_size))
    return s


def env3d_step(p, p_state, p_active, e_state, e_active, target, action, time_step):
    """One ParticleEnv.step for one env -> dict(p_state, p_active, e_state, e_active, reward, done, time_step)."""
    ps, pa = _c(p_state, np.float64).copy(), _c(p_active, np.uint8).copy()
    es, ea = _c(e_state, np.float64).copy(), C.c_uint8(int(e_active))
    tg, act = _c(target, np.float64), _c(action, np.float64)
    ts, done = C.c_int32(int(time_step)), C.c_uint8(0)
    reward = np.zeros(p.N, np.int32)
    lib().orc_env3d_step(C.byref(p), _p(ps), _p(pa), _p(es), C.byref(ea), _p(tg), _p(act), C.byref(ts), _p(reward),
                         C.byref(done))
    return dict(p_state=ps, p_active=pa, e_state=es, e_active=ea.value, reward=reward, done=done.value, time_step=ts.value)


def env3d_adjacency(p, p_state, p_active, e_state):
    ps, pa, es = _c(p_state, np.float64), _c(p_active, np.uint8), _c(e_state, np.float64)
    pp = np.zeros((p.N, p.N), np.uint8)
    pe = np.zeros(p.N, np.uint8)
    lib().orc_env3d_adjacency(C.byref(p), _p(ps), _p(pa), _p(es), _p(pp), _p(pe))
    return pp, pe


def env3d_iteration(p, st):
    """Batched adjacency -> evader move -> step over st's contiguous arrays (mutated in place)."""
    B = st["p_state"].shape[0]
    lib().orc_env3d_iteration(C.byref(p), C.c_int32(B), _p(st["p_state"]), _p(st["p_active"]), _p(st["e_state"]),
                              _p(st["e_active"]), _p(st["target"]), _p(st["action"]), _p(st["e_action"]),
                              _p(st["time_step"]), _p(st["reward"]), _p(st["done"]), _p(st["pp_adj"]), _p(st["pe_adj"]))
