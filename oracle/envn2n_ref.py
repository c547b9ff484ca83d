"""TEST INFRASTRUCTURE — CPU restatement (numpy float64 scalars, plain loops: small cases only) of the 2-D N-vs-E particle env,
`environment/env_n2n/particle_env.py` of the reference.  Pinned by tests/golden/envn2n_*.npz (oracle/gen_golden_envn2n.py executes
the unmodified reference); checked in tests/test_oracle_envn2n.py.  Only tests/ may import this module.

State layout used by the product (and here): pursuers f64 [N,4] = (x, y, phi, v), evaders f64 [E,4], active u8 [N] / [E].
"""
import numpy as np

PARK = 1000.0


def default_params(**over):
    """ParticleEnv.__init__ (particle_env.py:106-157)."""
    p = dict(p_vmax=0.3, e_vmax=1.0, p_sen_range=3.0, p_comm_range=6.0, e_sen_range=3.0, e_comm_range=6.0, kill_radius=0.5,
             ang_lmt=np.pi / 4, episode_limit=100, step_size=0.5)
    p.update(over)
    return p


def _turn(a, phi, ang_lmt):
    """Shared heading rule (particle_env.py:41-55 / 73-85): signed, rate-limited turn from phi towards the commanded angle a."""
    if np.sign(a * phi) >= 0:
        delta, sign = abs(a - phi), np.sign(a - phi)
    elif abs(a - phi) < 2 * np.pi - abs(a - phi):
        delta, sign = abs(a - phi), np.sign(a - phi)
    else:
        delta, sign = 2 * np.pi - abs(a - phi), -np.sign(a - phi)
    return sign * np.clip(delta, 0, ang_lmt)


def _wrap(phi):
    if phi > np.pi:
        phi -= 2 * np.pi
    elif phi < -np.pi:
        phi += 2 * np.pi
    return phi


def pursuer_step(s, active, a, prm):
    """Pursuer.step (particle_env.py:34-63): the heading turns even when the pursuer is inactive; only an active one moves."""
    x, y, phi, v_old = (np.float64(t) for t in s)
    if a == 0:
        v = 0
    else:
        v = prm["p_vmax"]
        ang = a * np.pi / 4
        if ang > np.pi:
            ang -= 2 * np.pi
        phi = _wrap(phi + _turn(ang, phi, prm["ang_lmt"]))
    if active:
        x = x + v * np.cos(phi) * prm["step_size"]
        y = y + v * np.sin(phi) * prm["step_size"]
        v_old = v
    return np.array([x, y, phi, v_old], np.float64)


def evader_step(s, a, prm):
    """Evader.step (particle_env.py:70-93) of an ACTIVE evader: moves along the OLD heading, then turns."""
    x, y, phi, v = (np.float64(t) for t in s)
    ang = a * np.pi
    d = _turn(ang, phi, prm["ang_lmt"])
    x = x + v * np.cos(phi) * prm["step_size"]
    y = y + v * np.sin(phi) * prm["step_size"]
    return np.array([x, y, _wrap(phi + d), v], np.float64)


def evaders_move(e_state, e_active, e_action, prm):
    out = e_state.copy()
    for j in range(len(e_state)):
        if e_active[j]:
            out[j] = evader_step(e_state[j], np.float64(e_action[j]), prm)
    return out


def _hits(pos, others, others_active, r):
    return sum(1 for k in range(len(others)) if others_active[k] and np.linalg.norm(pos[:2] - others[k][:2]) <= r)


def step(p_state, p_active, e_state, e_active, action, time_step, target, prm):
    """ParticleEnv.step (particle_env.py:164-177): returns (p_state, p_active, e_state, e_active, reward, done, time_step)."""
    N, E, r = len(p_state), len(e_state), prm["kill_radius"]
    p = np.stack([pursuer_step(p_state[i], p_active[i], int(action[i]), prm) for i in range(N)])
    # reward (:263-285) and the verdicts of update_agent_active (:287-321), all evaluated before anybody is removed
    reward = np.zeros(N, np.int32)
    p_dead, e_dead = np.zeros(N, bool), np.zeros(E, bool)
    for i in range(N):
        if p_active[i]:
            hit, inner = _hits(p[i], e_state, e_active, r), _hits(p[i], p, p_active, r)
            reward[i] = hit - (inner - 1)
            p_dead[i] = bool(inner + hit - 1)
    for j in range(E):
        if e_active[j]:
            e_dead[j] = bool(_hits(e_state[j], p, p_active, r))
    p_act, e_act, e = np.array(p_active, np.uint8).copy(), np.array(e_active, np.uint8).copy(), e_state.copy()
    for i in range(N):
        if p_dead[i]:
            p_act[i] = 0
            p[i, :3] = (PARK, PARK, 0.0)
    for j in range(E):
        if e_dead[j]:
            e_act[j] = 0
            e[j, :3] = (PARK, PARK, 0.0)
    time_step += 1
    done = any(np.linalg.norm([e[j, 0] - target[0], e[j, 1] - target[1]]) <= r for j in range(E)) or p_act.sum() == 0 or e_act.sum() == 0
    return p, p_act, e, e_act, reward, bool(done or time_step >= prm["episode_limit"]), time_step


def adjacency(a_state, a_active, b_state, rng):
    """get_adj_mat (particle_env.py:338-350): row i is zero when agent i is inactive."""
    out = np.zeros((len(a_state), len(b_state)), np.uint8)
    for i in range(len(a_state)):
        if a_active[i]:
            for j in range(len(b_state)):
                if np.linalg.norm([a_state[i][0] - b_state[j][0], a_state[i][1] - b_state[j][1]]) <= rng:
                    out[i, j] = 1
    return out


def choose_evader(p_state, p_active, e_state, e_active, sen_range, networks="actor"):
    """choose_evader (particle_env.py:392-420): the nearest active evader of every active pursuer (first index on ties; for the
    actor only evaders strictly inside the sensor range)."""
    out = np.zeros((len(p_state), len(e_state)), np.uint8)
    for i in range(len(p_state)):
        if not p_active[i]:
            continue
        best, best_d = -1, None
        for j in range(len(e_state)):
            if e_active[j]:
                d = np.linalg.norm([p_state[i][0] - e_state[j][0], p_state[i][1] - e_state[j][1]])
                if (networks != "actor" or d < sen_range) and (best_d is None or d < best_d):
                    best, best_d = j, d
        if best >= 0:
            out[i, best] = 1
    return out


def reset(n, e, prm=None):
    """ParticleEnv.reset (particle_env.py:195-262) on the GLOBAL numpy stream: (p_state [n,4], e_state [e,4], target [2])."""
    target = [np.random.rand() * 20, np.random.rand() * 20]

    def scatter(count, centre, lo, hi):
        pts = []
        while len(pts) < count:
            q = np.random.normal(loc=centre, scale=2, size=(2,)).clip(lo, hi)
            if all(np.linalg.norm(q - o) >= 2 for o in pts):
                pts.append(q)
        return pts
    prm = prm or default_params()
    pp = scatter(n, 0, -8, 8)
    ee = scatter(e, np.array([20 - target[0], 20 - target[1]]), 0, 20)
    p_state = np.array([[q[0] + 10, q[1] + 10, np.pi / 4, 0.0] for q in pp], np.float64)
    e_state = np.array([[q[0], q[1], np.pi / 4, prm["e_vmax"]] for q in ee], np.float64)
    return p_state, e_state, np.array(target, np.float64)
