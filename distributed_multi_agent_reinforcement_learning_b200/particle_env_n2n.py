"""2-D N-pursuers-vs-E-evaders particle env: the batched B200 engine and the reference-compatible facade.

* `BatchedParticleEnvN2N` — B independent envs resident in HBM (`p_state f64 [B,N,4]` = x, y, phi, v; evaders `[B,E,4]`; active
  flags; target), stepped by the sm_100a kernels behind the C-ABI (`marl_envn2n_*`, include/marl_b200.h).
* `ParticleEnv` — drop-in for `environment/env_n2n/particle_env.py:104-462` (same constructor / `initialize` / `reset` / `step` /
  `evader_step` / `get_done` / `get_active` / `get_agent_state` / `get_team_state` / `reward` / `get_adj_mat` /
  `collision_detection` / `choose_evader`, same return types).  It is a B=1 view of the engine.  `reset()` draws from the global
  numpy RNG in the reference's order, so equal seeds give equal initial states.

The evaders' commanded headings are an input (`evader_step(p_state, action=...)`): the reference computes them with scipy SLSQP
(`eva.e_f`), third-party arithmetic outside this path.  No CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


class EnvN2nParams(C.Structure):
    """marl_envn2n_params (include/marl_b200.h); defaults = particle_env.py:106-149."""
    _fields_ = [("N", C.c_int32), ("E", C.c_int32), ("episode_limit", C.c_int32), ("reserved", C.c_int32)] + \
               [(n, C.c_double) for n in ("p_vmax", "kill_radius", "ang_lmt", "step_size", "comm_range", "sen_range")]

    @classmethod
    def make(cls, N, E, episode_limit=100, p_vmax=0.3, kill_radius=0.5, ang_lmt=np.pi / 4, step_size=0.5, comm_range=6.0, sen_range=3.0):
        p = cls()
        p.N, p.E, p.episode_limit = int(N), int(E), int(episode_limit)
        p.p_vmax, p.kill_radius, p.ang_lmt = float(p_vmax), float(kill_radius), float(ang_lmt)
        p.step_size, p.comm_range, p.sen_range = float(step_size), float(comm_range), float(sen_range)
        return p


class EnvN2nRecords(C.Structure):
    FIELDS = ("p_state_f32", "e_state_f32", "p_active", "e_active", "pp_adj_bits", "pe_adj_bits", "assign", "action", "reward", "done")
    _fields_ = [(n, C.c_void_p) for n in FIELDS]


class EnvN2nArena:
    """Time-major records of a fused rollout: one env step of all envs is one contiguous slab."""

    def __init__(self, N, E, B, T, device, observations=True):
        z = lambda *s, dtype: torch.zeros(*s, dtype=dtype, device=device)
        self.T, self.B, self.N, self.E = T, B, N, E
        self.p_state_f32, self.e_state_f32 = z(T, B, N, 4, dtype=torch.float32), z(T, B, E, 4, dtype=torch.float32)
        self.p_active, self.e_active = z(T, B, N, dtype=torch.uint8), z(T, B, E, dtype=torch.uint8)
        self.pp_adj_bits = z(T, B, N, dtype=torch.int32) if observations else None
        self.pe_adj_bits = z(T, B, N, dtype=torch.int32) if observations else None
        self.assign = z(T, B, N, dtype=torch.int8) if observations else None
        self.action, self.reward = z(T, B, N, dtype=torch.int32), z(T, B, N, dtype=torch.int32)
        self.done = z(T, B, dtype=torch.uint8)

    def records(self):
        r = EnvN2nRecords()
        for n in EnvN2nRecords.FIELDS:
            t = getattr(self, n)
            setattr(r, n, t.data_ptr() if t is not None else None)
        return r


def counter_uniform_pm1(seed, agent_linear, t):
    """numpy restatement of the device counter RNG (envn2n_kernels.cu: n2n_rand_pm1) for tests: uniform in [-1,1)."""
    M = (1 << 64) - 1

    def sm(z):
        z = (z + 0x9E3779B97F4A7C15) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)

    h = sm((sm(seed ^ sm(((agent_linear & M) * 0x100000001B3 + t) & M)) + 0x51) & M)
    return float(h >> 11) * (2.0 / 9007199254740992.0) - 1.0


class BatchedParticleEnvN2N:
    def __init__(self, num_envs, num_pursuers, num_evaders, device="cuda:0", e_vmax=1.0, **params):
        if not torch.cuda.is_available():
            raise _lib.MarlError("BatchedParticleEnvN2N needs a CUDA device (no CPU fallback)")
        self.lib = _lib.lib()
        self.device = torch.device(device)
        self.params = EnvN2nParams.make(num_pursuers, num_evaders, **params)
        self.e_vmax = float(e_vmax)
        self.B, self.N, self.E = int(num_envs), int(num_pursuers), int(num_evaders)
        B, N, E, dev = self.B, self.N, self.E, self.device
        self.p_state = torch.zeros(B, N, 4, dtype=torch.float64, device=dev)
        self.p_active = torch.ones(B, N, dtype=torch.uint8, device=dev)
        self.e_state = torch.zeros(B, E, 4, dtype=torch.float64, device=dev)
        self.e_active = torch.ones(B, E, dtype=torch.uint8, device=dev)
        self.target = torch.zeros(B, 2, dtype=torch.float64, device=dev)
        self.time_step = torch.zeros(B, dtype=torch.int32, device=dev)
        self.reward = torch.zeros(B, N, dtype=torch.int32, device=dev)
        self.done = torch.zeros(B, dtype=torch.uint8, device=dev)
        self.pp_adj_bits = torch.zeros(B, N, dtype=torch.int32, device=dev)
        self.pe_adj_bits = torch.zeros(B, N, dtype=torch.int32, device=dev)
        self.assign = torch.zeros(B, N, dtype=torch.int8, device=dev)
        self.launches = 0

    def _pp(self):
        return C.byref(self.params)

    def set_state(self, p_state, e_state, target, p_active=None, e_active=None, time_step=0):
        f = lambda a, shape: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).reshape(shape)
        self.p_state.copy_(f(p_state, (self.B, self.N, 4)))
        self.e_state.copy_(f(e_state, (self.B, self.E, 4)))
        self.target.copy_(f(target, (self.B, 2)))
        for dst, src, shape in ((self.p_active, p_active, (self.B, self.N)), (self.e_active, e_active, (self.B, self.E))):
            if src is None:
                dst.fill_(1)
            else:
                dst.copy_(torch.as_tensor(np.ascontiguousarray(src, dtype=np.uint8)).reshape(shape))
        self.time_step.fill_(int(time_step))

    def reset(self, seed=0):
        """Synthetic initial states for throughput runs: the reference's draws (target ~ U[0,20)^2, pursuers ~ N(10,2) clipped,
        evaders ~ N(20 - target, 2) clipped to [0,20], headings pi/4) WITHOUT the pairwise min-distance rejection (which is a host
        loop in the reference; the facade's `reset` keeps it)."""
        g = torch.Generator(device=self.device).manual_seed(int(seed))
        B, N, E, dev = self.B, self.N, self.E, self.device
        self.target.copy_(torch.rand(B, 2, generator=g, device=dev, dtype=torch.float64) * 20)
        self.p_state.zero_()
        self.p_state[..., :2] = (torch.randn(B, N, 2, generator=g, device=dev, dtype=torch.float64) * 2).clamp(-8, 8) + 10
        self.p_state[..., 2] = np.pi / 4
        self.e_state.zero_()
        self.e_state[..., :2] = ((20 - self.target).unsqueeze(1) + torch.randn(B, E, 2, generator=g, device=dev, dtype=torch.float64) * 2).clamp(0, 20)
        self.e_state[..., 2] = np.pi / 4
        self.e_state[..., 3] = self.e_vmax
        self.p_active.fill_(1)
        self.e_active.fill_(1)
        self.time_step.zero_()

    # ---------------------------------------------------------------------------------------------- kernels
    def step(self, action):
        """ParticleEnv.step for every env; action i32 [B,N] in 0..8.  Returns (reward i32 [B,N], done u8 [B])."""
        assert tuple(action.shape) == (self.B, self.N) and action.dtype == torch.int32
        P = _lib.ptr
        _lib.check(self.lib.marl_envn2n_step(self._pp(), self.B, P(self.p_state), P(self.p_active), P(self.e_state), P(self.e_active),
                                             P(self.target), P(action), P(self.time_step), P(self.reward), P(self.done),
                                             _lib.stream_ptr()), "marl_envn2n_step")
        self.launches += 1
        return self.reward, self.done

    def evader_step(self, e_action):
        """Evader.step of every active evader for commanded headings f64 [B,E] in [-1,1]."""
        assert tuple(e_action.shape) == (self.B, self.E) and e_action.dtype == torch.float64
        P = _lib.ptr
        _lib.check(self.lib.marl_envn2n_evader_step(self._pp(), self.B, P(self.e_state), P(self.e_active), P(e_action), _lib.stream_ptr()),
                   "marl_envn2n_evader_step")
        self.launches += 1

    def observe(self):
        """(pp_adj_bits, pe_adj_bits) i32 [B,N] (bit j = column j) and the nearest-evader assignment i8 [B,N] (-1 = none)."""
        P = _lib.ptr
        _lib.check(self.lib.marl_envn2n_observe(self._pp(), self.B, P(self.p_state), P(self.p_active), P(self.e_state), P(self.e_active),
                                                P(self.pp_adj_bits), P(self.pe_adj_bits), P(self.assign), _lib.stream_ptr()),
                   "marl_envn2n_observe")
        self.launches += 1
        return self.pp_adj_bits, self.pe_adj_bits, self.assign

    def rollout(self, arena, K, t0=0, action_tape=None, e_action_tape=None, seed=0):
        """K fused iterations (observe -> evader move -> step -> store) in ONE launch; tapes i32 [K,B,N] / f64 [K,B,E] or None
        (counter RNG)."""
        if action_tape is not None:
            assert tuple(action_tape.shape) == (K, self.B, self.N) and action_tape.dtype == torch.int32
        if e_action_tape is not None:
            assert tuple(e_action_tape.shape) == (K, self.B, self.E) and e_action_tape.dtype == torch.float64
        P = _lib.ptr
        rec = arena.records()
        _lib.check(self.lib.marl_envn2n_rollout(self._pp(), self.B, arena.T, t0, K, P(self.p_state), P(self.p_active), P(self.e_state),
                                                P(self.e_active), P(self.target), P(self.time_step), P(action_tape), P(e_action_tape),
                                                C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), C.byref(rec), _lib.stream_ptr()),
                   "marl_envn2n_rollout")
        self.launches += 1


# ---------------------------------------------------------------------------------------------------- facade
class _Agent:
    """View of one agent of the B=1 engine with the attributes the reference's callers touch."""

    def __init__(self, env, pursuer, idx):
        self._env, self._p, self.idx, self.is_pursuer = env, pursuer, idx, pursuer

    def _row(self):
        return (self._env._p if self._p else self._env._e)[self.idx]

    x = property(lambda s: float(s._row()[0]))
    y = property(lambda s: float(s._row()[1]))
    phi = property(lambda s: float(s._row()[2]))
    v = property(lambda s: float(s._row()[3]))
    active = property(lambda s: bool((s._env._pa if s._p else s._env._ea)[s.idx]))


class ParticleEnv:
    """Reference-compatible single environment (environment/env_n2n/particle_env.py:104-462)."""

    def __init__(self, device="cuda:0"):
        self.p_obs_dim = self.e_obs_dim = 3
        self.env_name = "ParticleEnvBoundGra"
        self.p_vmax, self.e_vmax = 0.3, 1
        self.x_bound, self.y_bound = [5, 15], [5, 15]
        self.p_sen_range, self.p_comm_range, self.e_sen_range, self.e_comm_range = 3, 6, 3, 6
        self.target, self.kill_radius, self.ang_lmt = [17.5, 17.5], 0.5, np.pi / 4
        self.random = np.random
        self.action_dim, self.n_episode, self.episode_limit, self.shadow_epi, self.target_return = 1, 0, 100, 1000, 1000
        self.step_size, self.time_step, self.curriculum = 0.5, 0, False
        self.p_num = self.e_num = None
        self.p_list, self.p_idx, self.e_list, self.e_idx = {}, [], {}, []
        self.state = None
        self._device = device
        self.engine = None

    def initialize(self, p_num, e_num):
        self.p_num, self.e_num = p_num, e_num
        self.engine = BatchedParticleEnvN2N(1, p_num, e_num, device=self._device, e_vmax=self.e_vmax, episode_limit=self.episode_limit,
                                            p_vmax=self.p_vmax, kill_radius=self.kill_radius, ang_lmt=self.ang_lmt,
                                            step_size=self.step_size, comm_range=self.p_comm_range, sen_range=self.p_sen_range)

    def _pull(self):
        eng = self.engine
        self._p, self._e = eng.p_state[0].cpu().numpy(), eng.e_state[0].cpu().numpy()
        self._pa, self._ea = eng.p_active[0].cpu().numpy(), eng.e_active[0].cpu().numpy()

    def reset(self):
        """particle_env.py:195-246 on the global numpy stream (target, pursuers, evaders, in that order)."""
        self.target = [np.random.rand() * 20, np.random.rand() * 20]
        self.time_step = 0
        self.n_episode += 1
        if self.n_episode >= self.shadow_epi / 2:
            self.curriculum = False
        p_pos = self.gen_init_p_pos()
        e_pos = self.gen_init_e_pos([20 - self.target[0], 20 - self.target[1]])
        p = np.array([[q[0] + 10, q[1] + 10, np.pi / 4, 0.0] for q in p_pos], np.float64)
        e = np.array([[q[0], q[1], np.pi / 4, self.e_vmax] for q in e_pos], np.float64)
        self.engine.set_state(p[None], e[None], np.array(self.target, np.float64)[None], time_step=0)
        self.p_idx, self.e_idx = list(range(self.p_num)), list(range(self.e_num))
        self.p_list = {f"{i}": _Agent(self, True, i) for i in self.p_idx}
        self.e_list = {f"{i}": _Agent(self, False, i) for i in self.e_idx}
        self._pull()

    def _scatter(self, count, centre, lo, hi):
        pts = []
        while len(pts) < count:
            q = np.random.normal(loc=centre, scale=2, size=(2,)).clip(lo, hi)
            if all(np.linalg.norm(q - o) >= 2 for o in pts):
                pts.append(q)
        return pts

    def gen_init_p_pos(self):
        return self._scatter(self.p_num, 0, -8, 8)

    def gen_init_e_pos(self, center):
        return self._scatter(self.e_num, np.array(center), 0, 20)

    def step(self, action):
        act = torch.as_tensor(np.asarray(action, dtype=np.int64).reshape(1, -1).astype(np.int32), device=self.engine.device)
        reward, done = self.engine.step(act)
        self.time_step += 1
        self._pull()
        return [int(v) for v in reward[0].cpu().numpy()], bool(done[0].item()), self.get_active()

    def evader_step(self, p_state=None, action=None):
        """particle_env.py:179-193 with the SLSQP output supplied by the caller: `action` = one commanded heading in [-1,1] per
        evader (entries of inactive evaders are ignored)."""
        if action is None:
            raise _lib.MarlError("evader_step needs `action`: eva.e_f (scipy SLSQP) is outside this path")
        a = torch.as_tensor(np.asarray(action, dtype=np.float64).reshape(1, self.e_num), device=self.engine.device)
        self.engine.evader_step(a)
        self._pull()

    def get_done(self):
        r = self.kill_radius
        if any(np.linalg.norm([e[0] - self.target[0], e[1] - self.target[1]]) <= r for e in self._e):
            return True
        return int(self._pa.sum()) == 0 or int(self._ea.sum()) == 0

    def get_active(self):
        return [int(v) for v in self._pa]

    def get_agent_state(self, is_pursuer, idx):
        row = (self._p if is_pursuer else self._e)[idx]
        return [float(row[0]), float(row[1]), float(row[2])]

    def get_team_state(self, is_pursuer, rules=True):
        act = self._pa if is_pursuer else self._ea
        return [self.get_agent_state(is_pursuer, i) for i in range(len(act)) if (act[i] or not rules)]

    def collision_detection(self, agent_idx, is_pursuer, is_inner):
        me = np.array(self.get_agent_state(is_pursuer, agent_idx))
        other = np.array(self.get_team_state(is_pursuer if is_inner else (not is_pursuer), rules=True))
        return [1 if np.linalg.norm(me[:2] - o[:2]) <= self.kill_radius else 0 for o in other]

    def agent_reward(self, agent_idx, is_pursuer=True):
        return sum(self.collision_detection(agent_idx, is_pursuer, False)) - (sum(self.collision_detection(agent_idx, is_pursuer, True)) - 1)

    def reward(self, is_pursuer):
        return [self.agent_reward(i, is_pursuer) if self._pa[i] else 0 for i in self.p_idx]

    def get_adj_mat(self, obs, be_obs, rag, is_pursuer=True):
        """get_adj_mat (particle_env.py:338-350).  The two relations a policy consumes (pursuers against pursuers at
        p_comm_range, pursuers against evaders at p_sen_range, both on the current full team states) come from the device
        kernel; any other combination is evaluated on the host with the same rule."""
        eng = self.engine
        full_p, full_e = self.get_team_state(True, rules=False), self.get_team_state(False, rules=False)
        if is_pursuer and obs == full_p and ((be_obs == full_p and rag == self.p_comm_range) or (be_obs == full_e and rag == self.p_sen_range)):
            pp, pe, _ = eng.observe()
            words = (pp if be_obs == full_p else pe)[0].cpu().numpy().astype(np.uint32)
            n = len(be_obs)
            return ((words[:, None] >> np.arange(n, dtype=np.uint32)[None, :]) & 1).astype(np.float64)
        act = self._pa if is_pursuer else self._ea
        out = np.zeros((len(obs), len(be_obs)))
        for i, a in enumerate(obs):
            if act[i]:
                for j, b in enumerate(be_obs):
                    if np.linalg.norm([a[0] - b[0], a[1] - b[1]]) <= rag:
                        out[i, j] = 1
        return out

    def choose_evader(self, networks="actor"):
        out = np.zeros((self.p_num, self.e_num))
        if networks == "actor":
            _, _, assign = self.engine.observe()
            for i, j in enumerate(assign[0].cpu().numpy()):
                if j >= 0:
                    out[i, j] = 1
            return out
        for i in self.p_idx:
            if self._pa[i]:
                d = [(j, np.linalg.norm([self._p[i][0] - self._e[j][0], self._p[i][1] - self._e[j][1]])) for j in self.e_idx if self._ea[j]]
                if d:
                    out[i, min(d, key=lambda t: t[1])[0]] = 1
        return out
