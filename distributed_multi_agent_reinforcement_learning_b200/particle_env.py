"""3-D particle pursuit env: the batched B200 engine and the reference-compatible facade.

* `BatchedParticleEnv` — B independent envs resident in HBM (`p_state f64 [B,N,6]` = x,y,z,phi,gamma,v; active flags;
  evader; target), stepped by the sm_100a kernels behind the C-ABI (`marl_env3d_*`, include/marl_b200.h).
* `ParticleEnv` — drop-in for `environment/env_3d/particle_env.py:76-404` (same constructor / `initialize` / `reset` /
  `step` / `get_done` / `get_active` / `get_agent_state` / `get_team_state` / `reward` / `get_adj_mat` /
  `collision_detection` / `evader_step`, same return types).  It is a B=1 view of the engine.  `reset()` draws from the
  global numpy RNG in the reference's order, so equal seeds give equal initial states.

The evader's commanded action is an input (`evader_step(p_state, action=...)`): the reference computes it with scipy
SLSQP (`eva.e_f`), which is third-party arithmetic outside this path.  No CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


class Env3dParams(C.Structure):
    """marl_env3d_params (include/marl_b200.h); defaults = particle_env.py:78-92,119-120."""
    _fields_ = [("N", C.c_int32), ("max_step", C.c_int32)] + \
               [(n, C.c_double) for n in ("p_vmax", "e_vmax", "kill_radius", "ang_lmt", "v_lmt", "step_size",
                                          "comm_range", "sen_range")]

    @classmethod
    def make(cls, N, max_step=200, p_vmax=0.7, e_vmax=1.0, kill_radius=0.5, ang_lmt=np.pi / 4, v_lmt=0.4, step_size=0.5,
             comm_range=6.0, sen_range=3.0):
        p = cls()
        p.N, p.max_step = int(N), int(max_step)
        p.p_vmax, p.e_vmax, p.kill_radius, p.ang_lmt, p.v_lmt = float(p_vmax), float(e_vmax), float(kill_radius), float(ang_lmt), float(v_lmt)
        p.step_size, p.comm_range, p.sen_range = float(step_size), float(comm_range), float(sen_range)
        return p


class Env3dRecords(C.Structure):
    FIELDS = ("p_state_f32", "e_state_f32", "pp_adj_bits", "pe_adj", "reward", "active_f32", "done")
    _fields_ = [(n, C.c_void_p) for n in FIELDS]


class Env3dArena:
    """Time-major records of a fused rollout: one env step of all envs is one contiguous slab."""

    def __init__(self, N, B, T, device, adjacency=True):
        NW = (N + 31) // 32
        f32 = dict(dtype=torch.float32, device=device)
        self.T, self.B, self.N = T, B, N
        self.p_state_f32 = torch.zeros(T, B, N, 6, **f32)
        self.e_state_f32 = torch.zeros(T, B, 6, **f32)
        self.pp_adj_bits = torch.zeros(T, B, N, NW, dtype=torch.int32, device=device) if adjacency else None
        self.pe_adj = torch.zeros(T, B, N, dtype=torch.uint8, device=device) if adjacency else None
        self.reward = torch.zeros(T, B, N, dtype=torch.int32, device=device)
        self.active_f32 = torch.zeros(T, B, N, **f32)
        self.done = torch.zeros(T, B, dtype=torch.uint8, device=device)

    def records(self):
        r = Env3dRecords()
        for n in Env3dRecords.FIELDS:
            t = getattr(self, n)
            setattr(r, n, t.data_ptr() if t is not None else None)
        return r


def counter_uniform_pm1(seed, agent_linear, t, comp):
    """numpy restatement of the device counter RNG (env3d_kernels.cu: rand_pm1) for tests: uniform in [-1,1)."""
    M = (1 << 64) - 1

    def sm(z):
        z = (z + 0x9E3779B97F4A7C15) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)

    h = sm((sm(seed ^ sm(((agent_linear & M) * 0x100000001B3 + t) & M)) + comp) & M)
    return float(h >> 11) * (2.0 / 9007199254740992.0) - 1.0


class BatchedParticleEnv:
    def __init__(self, num_envs, num_pursuers, device="cuda:0", **params):
        if not torch.cuda.is_available():
            raise _lib.MarlError("BatchedParticleEnv needs a CUDA device (no CPU fallback)")
        self.lib = _lib.lib()
        self.device = torch.device(device)
        self.params = Env3dParams.make(num_pursuers, **params)
        self.B, self.N = int(num_envs), int(num_pursuers)
        B, N, dev = self.B, self.N, self.device
        self.p_state = torch.zeros(B, N, 6, dtype=torch.float64, device=dev)
        self.p_active = torch.ones(B, N, dtype=torch.uint8, device=dev)
        self.e_state = torch.zeros(B, 6, dtype=torch.float64, device=dev)
        self.e_active = torch.ones(B, dtype=torch.uint8, device=dev)
        self.target = torch.zeros(B, 3, dtype=torch.float64, device=dev)
        self.time_step = torch.zeros(B, dtype=torch.int32, device=dev)
        self.reward = torch.zeros(B, N, dtype=torch.int32, device=dev)
        self.done = torch.zeros(B, dtype=torch.uint8, device=dev)
        self.pp_adj_bits = torch.zeros(B, N, (N + 31) // 32, dtype=torch.int32, device=dev)
        self.pe_adj = torch.zeros(B, N, dtype=torch.uint8, device=dev)
        self.launches = 0

    def _pp(self):
        return C.byref(self.params)

    def set_state(self, p_state, e_state, target, p_active=None, e_active=None, time_step=0):
        f = lambda a, shape: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).reshape(shape)
        self.p_state.copy_(f(p_state, (self.B, self.N, 6)))
        self.e_state.copy_(f(e_state, (self.B, 6)))
        self.target.copy_(f(target, (self.B, 3)))
        if p_active is None:
            self.p_active.fill_(1)
        else:
            self.p_active.copy_(torch.as_tensor(np.ascontiguousarray(p_active, dtype=np.uint8)).reshape(self.B, self.N))
        if e_active is None:
            self.e_active.fill_(1)
        else:
            self.e_active.copy_(torch.as_tensor(np.ascontiguousarray(e_active, dtype=np.uint8)).reshape(self.B))
        self.time_step.fill_(int(time_step))

    def reset(self, seed=0):
        """Synthetic initial states for throughput runs: the reference's draws (target ~ U[0,20)^3, pursuers ~
        N(10,2) clipped to [5,15], evader = 20 - target, headings uniform) with the min-distance rule RELAXED to the
        kill radius for large N (min_dist 4 in a 10^3 box cannot place 32 pursuers — SURVEY §8d C4)."""
        g = np.random.default_rng(seed)
        B, N = self.B, self.N
        tgt = g.random((B, 3)) * 20
        pos = np.clip(g.normal(10, 2, (B, N, 3)), 5, 15)
        for _ in range(8):       # push apart anything closer than 2 kill radii (keeps step 1 from being a massacre)
            d = np.linalg.norm(pos[:, :, None] - pos[:, None], axis=-1) + np.eye(N)[None] * 1e9
            bad = (d < 2 * self.params.kill_radius).any(-1)
            if not bad.any():
                break
            pos[bad] = np.clip(g.normal(10, 2, (int(bad.sum()), 3)), 5, 15)
        ps = np.zeros((B, N, 6))
        ps[..., :3] = pos
        ps[..., 3] = (2 * g.random((B, N)) - 1) * np.pi
        ps[..., 4] = (2 * g.random((B, N)) - 1) * np.pi / 2
        es = np.zeros((B, 6))
        es[:, :3] = 20 - tgt
        es[:, 3] = (2 * g.random(B) - 1) * np.pi
        es[:, 4] = (2 * g.random(B) - 1) * np.pi / 2
        self.set_state(ps, es, tgt)

    def snapshot(self):
        return {k: getattr(self, k).clone() for k in ("p_state", "p_active", "e_state", "e_active", "target", "time_step")}

    def restore(self, snap):
        for k, v in snap.items():
            getattr(self, k).copy_(v)

    # ---------------------------------------------------------------------------------------------- per-call API
    def step(self, action):
        """action: f64 [B,N,3] device tensor.  Fills self.reward / self.done."""
        P = _lib.ptr
        _lib.check(self.lib.marl_env3d_step(self._pp(), self.B, P(self.p_state), P(self.p_active), P(self.e_state),
                                            P(self.e_active), P(self.target), P(action), P(self.time_step), P(self.reward),
                                            P(self.done), _lib.stream_ptr()), "marl_env3d_step")
        self.launches += 1

    def evader_step(self, e_action):
        P = _lib.ptr
        _lib.check(self.lib.marl_env3d_evader_step(self._pp(), self.B, P(self.e_state), P(self.e_active), P(e_action),
                                                   _lib.stream_ptr()), "marl_env3d_evader_step")
        self.launches += 1

    def adjacency(self, dense=False):
        P = _lib.ptr
        pp_f = torch.zeros(self.B, self.N, self.N, dtype=torch.float32, device=self.device) if dense else None
        pe_f = torch.zeros(self.B, self.N, 1, dtype=torch.float32, device=self.device) if dense else None
        _lib.check(self.lib.marl_env3d_adjacency(self._pp(), self.B, P(self.p_state), P(self.p_active), P(self.e_state),
                                                 P(self.pp_adj_bits), P(self.pe_adj), P(pp_f), P(pe_f), _lib.stream_ptr()),
                   "marl_env3d_adjacency")
        self.launches += 1
        return (pp_f, pe_f) if dense else (self.pp_adj_bits, self.pe_adj)

    # ---------------------------------------------------------------------------------------------- fused rollout
    def rollout(self, arena, K, t0=0, action_tape=None, e_action_tape=None, seed=0):
        P = _lib.ptr
        rec = arena.records()
        _lib.check(self.lib.marl_env3d_rollout(self._pp(), self.B, arena.T, t0, K, P(self.p_state), P(self.p_active),
                                               P(self.e_state), P(self.e_active), P(self.target), P(self.time_step),
                                               P(action_tape), P(e_action_tape), C.c_uint64(seed), C.byref(rec),
                                               _lib.stream_ptr()), "marl_env3d_rollout")
        self.launches += 1


class _PointView:
    """Read-only attribute view of one agent (the reference exposes Pursuer/Evader objects in p_list / e_list)."""

    def __init__(self, env, pursuer, idx):
        self._env, self._pursuer, self.idx = env, pursuer, idx
        self.is_pursuer = pursuer

    def _row(self):
        e = self._env.engine
        return (e.p_state[0, self.idx] if self._pursuer else e.e_state[0]).cpu().numpy()

    x = property(lambda s: float(s._row()[0]))
    y = property(lambda s: float(s._row()[1]))
    z = property(lambda s: float(s._row()[2]))
    phi = property(lambda s: float(s._row()[3]))
    gamma = property(lambda s: float(s._row()[4]))
    v = property(lambda s: float(s._row()[5]))

    @property
    def active(self):
        e = self._env.engine
        return bool((e.p_active[0, self.idx] if self._pursuer else e.e_active[0]).item())


class ParticleEnv:
    """Reference-compatible single env (environment/env_3d/particle_env.py:76-404) on top of the engine."""

    def __init__(self, device="cuda:0"):
        self.p_obs_dim = self.e_obs_dim = 6
        self.env_name = "ParticleEnvBoundGra"
        self.p_vmax, self.e_vmax = 0.7, 1
        self.x_bound = self.y_bound = [5, 15]
        self.p_sen_range, self.p_comm_range, self.e_sen_range, self.e_comm_range = 3, 6, 3, 6
        self.target = [17.5, 17.5, 17.5]
        self.kill_radius, self.ang_lmt, self.v_lmt = 0.5, np.pi / 4, 0.4
        self.random = np.random
        self.state_dim, self.action_dim = 12, 3
        self.n_episode, self.max_step, self.shadow_epi, self.target_return = 0, 200, 1000, 1000
        self.step_size, self.time_step, self.curriculum = 0.5, 0, False
        self.p_num = self.e_num = None
        self.p_list = self.e_list = self.p_idx = self.e_idx = None
        self.device, self.engine = device, None

    def initialize(self, p_num):
        self.p_num, self.e_num = p_num, 1
        self.engine = BatchedParticleEnv(1, p_num, device=self.device, max_step=self.max_step, p_vmax=self.p_vmax,
                                         e_vmax=self.e_vmax, kill_radius=self.kill_radius, ang_lmt=self.ang_lmt,
                                         v_lmt=self.v_lmt, step_size=self.step_size, comm_range=self.p_comm_range,
                                         sen_range=self.p_sen_range)

    def reset(self):
        """particle_env.py:135-199 with the same global-RNG draw order."""
        self.target = [np.random.rand() * 20, np.random.rand() * 20, np.random.rand() * 20]
        self.time_step = 0
        self.n_episode += 1
        if self.n_episode >= self.shadow_epi / 2:
            self.curriculum = False
        sample = []
        while len(sample) < self.p_num:
            newp = np.random.normal(loc=10, scale=2, size=(3,)).clip(5, 15)
            if all(np.linalg.norm(newp - p) >= 4 for p in sample):
                sample.append(newp)
        ps = np.zeros((self.p_num, 6))
        for i in range(self.p_num):
            ps[i, :3] = sample[i]
            ps[i, 3] = (2 * np.random.rand() - 1) * np.pi
            ps[i, 4] = (2 * np.random.rand() - 1) * np.pi / 2
        es = np.zeros(6)
        es[:3] = [20 - self.target[0], 20 - self.target[1], 20 - self.target[2]]
        es[3] = (2 * np.random.rand() - 1) * np.pi
        es[4] = (2 * np.random.rand() - 1) * np.pi / 2
        self.engine.set_state(ps[None], es[None], np.array(self.target)[None])
        self.p_idx, self.e_idx = list(range(self.p_num)), [0]
        self.p_list = {f"{i}": _PointView(self, True, i) for i in self.p_idx}
        self.e_list = {"0": _PointView(self, False, 0)}

    def step(self, action):
        a = torch.as_tensor(np.ascontiguousarray(action, dtype=np.float64)).reshape(1, self.p_num, 3).to(self.engine.device)
        self.engine.step(a)
        self.time_step += 1
        reward = [int(v) for v in self.engine.reward[0].cpu().numpy()]
        return reward, bool(self.engine.done[0].item()), self.get_active()

    def get_done(self):
        e = self.engine
        es = e.e_state[0, :3].cpu().numpy()
        if np.linalg.norm([es[0] - self.target[0], es[1] - self.target[1], es[2] - self.target[2]]) <= self.kill_radius:
            return True
        return int(e.p_active.sum().item()) == 0 or int(e.e_active.sum().item()) == 0

    def get_active(self):
        return [int(v) for v in self.engine.p_active[0].cpu().numpy()]

    def get_agent_state(self, is_pursuer, idx):
        e = self.engine
        return [float(v) for v in (e.p_state[0, idx] if is_pursuer else e.e_state[0]).cpu().numpy()]

    def get_team_state(self, is_pursuer, rules=True):
        e = self.engine
        if is_pursuer:
            st, act = e.p_state[0].cpu().numpy(), e.p_active[0].cpu().numpy()
        else:
            st, act = e.e_state.cpu().numpy(), e.e_active.cpu().numpy()
        return [[float(v) for v in st[i]] for i in range(len(st)) if (act[i] or not rules)]

    def collision_detection(self, agent_idx, is_pursuer, is_inner):
        me = np.array(self.get_agent_state(is_pursuer, agent_idx))
        other = np.array(self.get_team_state(is_pursuer if is_inner else (not is_pursuer)))
        return [1 if np.linalg.norm(me[:3] - o[:3]) <= self.kill_radius else 0 for o in other]

    def agent_reward(self, agent_idx, is_pursuer=True):
        return sum(self.collision_detection(agent_idx, is_pursuer, False)) - (sum(self.collision_detection(agent_idx, is_pursuer, True)) - 1)

    def reward(self, is_pursuer):
        act = self.get_active()
        return [self.agent_reward(i, is_pursuer) if act[i] else 0 for i in self.p_idx]

    def get_adj_mat(self, obs=None, be_obs=None, rag=None, is_pursuer=True):
        """Device path for the two relations the engine knows (pursuer-pursuer at comm range, pursuer-evader at sensor
        range, selected by `rag`); returns np.ndarray [n_obs, n_be_obs] of 0/1 like the reference."""
        pp, pe = self.engine.adjacency(dense=True)
        if be_obs is not None and len(be_obs) == 1 and (rag is None or rag == self.p_sen_range):
            return pe[0].cpu().numpy().astype(np.float64)
        return pp[0].cpu().numpy().astype(np.float64)

    def evader_step(self, p_state=None, action=None):
        """The evader's Point.step for a commanded action in [-1,1]^3 (the reference obtains it from eva.e_f / SLSQP)."""
        if action is None:
            raise _lib.MarlError("ParticleEnv.evader_step: pass action=(phi, gamma, v) commands; the SLSQP evader "
                                 "(environment/env_3d/eva.py) is outside this path")
        a = torch.as_tensor(np.ascontiguousarray(action, dtype=np.float64)).reshape(1, 3).to(self.engine.device)
        self.engine.evader_step(a)
