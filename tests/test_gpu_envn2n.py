"""GPU parity of the 2-D N-vs-E particle env kernels (csrc/envn2n_kernels.cu, through the C-ABI) — SURVEY §8(f) rank 4.

Oracles: tests/golden/envn2n_*.npz (the unmodified reference, oracle/gen_golden_envn2n.py) and oracle/envn2n_ref.py (its numpy
restatement, bit-exact against those fixtures).  Discrete outputs (rewards, done, active flags, adjacency, assignment) must be
identical; continuous state agrees to 1e-12 (CUDA's cos / sin vs numpy's differ in the last ulp)."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

FIXTURES = sorted(p for p in glob.glob(os.path.join(GOLDEN_DIR, "envn2n_*.npz")) if "reset" not in p)
IDS = [os.path.basename(p)[:-4] for p in FIXTURES]
TOL = dict(rtol=0, atol=1e-12)


def _engine(fx, B=1):
    from distributed_multi_agent_reinforcement_learning_b200.particle_env_n2n import BatchedParticleEnvN2N
    return BatchedParticleEnvN2N(B, int(fx["n"]), int(fx["e"]), e_vmax=float(fx["e_vmax"]), episode_limit=int(fx["episode_limit"]),
                                 p_vmax=float(fx["p_vmax"]), kill_radius=float(fx["kill_radius"]), ang_lmt=float(fx["ang_lmt"]),
                                 step_size=float(fx["step_size"]), comm_range=float(fx["p_comm_range"]), sen_range=float(fx["p_sen_range"]))


def _unpack(words, n):
    w = words.cpu().numpy().astype(np.uint32)
    return ((w[..., None] >> np.arange(n, dtype=np.uint32)) & 1).astype(np.uint8)


def _assign_matrix(assign, e):
    a = assign.cpu().numpy()
    out = np.zeros(a.shape + (e,), np.uint8)
    for idx in np.argwhere(a >= 0):
        out[tuple(idx) + (a[tuple(idx)],)] = 1
    return out


@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_stepwise_episode_matches_reference(path):
    fx = np.load(path)
    n, e, T = int(fx["n"]), int(fx["e"]), len(fx["done"])
    env = _engine(fx)
    env.set_state(fx["p_before"][0][None], fx["e_before"][0][None], fx["target"][None], fx["p_active_before"][0][None],
                  fx["e_active_before"][0][None])
    for t in range(T):
        np.testing.assert_allclose(env.p_state[0].cpu().numpy(), fx["p_before"][t], **TOL)
        pp, pe, assign = env.observe()
        assert np.array_equal(_unpack(pp[0], n), fx["pp_adj"][t]), t
        assert np.array_equal(_unpack(pe[0], e), fx["pe_adj"][t]), t
        assert np.array_equal(_assign_matrix(assign[0], e), fx["assign"][t]), t
        env.evader_step(torch.from_numpy(fx["e_action"][t][None].copy()).cuda())
        np.testing.assert_allclose(env.e_state[0].cpu().numpy(), fx["e_moved"][t], **TOL)
        reward, done = env.step(torch.from_numpy(fx["action"][t][None].copy()).cuda())
        assert np.array_equal(reward[0].cpu().numpy(), fx["reward"][t]), t
        assert bool(done[0].item()) == bool(fx["done"][t]), t
        assert np.array_equal(env.p_active[0].cpu().numpy(), fx["p_active"][t]) and np.array_equal(env.e_active[0].cpu().numpy(), fx["e_active"][t]), t
        np.testing.assert_allclose(env.p_state[0].cpu().numpy(), fx["p_after"][t], **TOL)
        np.testing.assert_allclose(env.e_state[0].cpu().numpy(), fx["e_after"][t], **TOL)
    assert int(env.time_step[0].item()) == T


@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_fused_rollout_matches_reference(path):
    from distributed_multi_agent_reinforcement_learning_b200.particle_env_n2n import EnvN2nArena
    fx = np.load(path)
    n, e, T = int(fx["n"]), int(fx["e"]), len(fx["done"])
    B = 3                                                           # three copies of the episode: batch indexing
    env = _engine(fx, B)
    rep = lambda a: np.repeat(a[None], B, axis=0)
    env.set_state(rep(fx["p_before"][0]), rep(fx["e_before"][0]), rep(fx["target"]), rep(fx["p_active_before"][0]), rep(fx["e_active_before"][0]))
    arena = EnvN2nArena(n, e, B, T, env.device)
    a_tape = torch.from_numpy(np.repeat(fx["action"][:, None], B, axis=1).astype(np.int32)).cuda().contiguous()
    e_tape = torch.from_numpy(np.repeat(fx["e_action"][:, None], B, axis=1)).cuda().contiguous()
    K1 = T // 3
    env.rollout(arena, K1, 0, a_tape[:K1].contiguous(), e_tape[:K1].contiguous())          # two launches: t0 handling
    env.rollout(arena, T - K1, K1, a_tape[K1:].contiguous(), e_tape[K1:].contiguous())
    for b in range(B):
        assert np.array_equal(arena.reward[:, b].cpu().numpy(), fx["reward"])
        assert np.array_equal(arena.done[:, b].cpu().numpy(), fx["done"])
        assert np.array_equal(arena.p_active[:, b].cpu().numpy(), fx["p_active_before"]) and np.array_equal(arena.e_active[:, b].cpu().numpy(), fx["e_active_before"])
        assert np.array_equal(_unpack(arena.pp_adj_bits[:, b], n), fx["pp_adj"]) and np.array_equal(_unpack(arena.pe_adj_bits[:, b], e), fx["pe_adj"])
        assert np.array_equal(_assign_matrix(arena.assign[:, b], e), fx["assign"])
        assert np.array_equal(arena.action[:, b].cpu().numpy(), fx["action"])
        np.testing.assert_allclose(arena.p_state_f32[:, b].cpu().numpy(), fx["p_before"].astype(np.float32), rtol=0, atol=1e-4)
        np.testing.assert_allclose(env.p_state[b].cpu().numpy(), fx["p_after"][-1], **TOL)
        np.testing.assert_allclose(env.e_state[b].cpu().numpy(), fx["e_after"][-1], **TOL)


def test_large_batch_rng_rollout_equals_stepwise_api_and_oracle():
    """4096 envs x 8 pursuers x 3 evaders, pursuer actions from the device counter RNG: the fused rollout must equal the per-call
    API replayed with the recorded actions bit for bit, the first envs must agree with the CPU oracle, and the domain invariants
    must hold (removed agents are parked at (1000, 1000) with heading 0, active flags never come back, rewards are bounded)."""
    from distributed_multi_agent_reinforcement_learning_b200.particle_env_n2n import BatchedParticleEnvN2N, EnvN2nArena
    from oracle import envn2n_ref as ref
    B, N, E, K = 4096, 8, 3, 24
    env = BatchedParticleEnvN2N(B, N, E)
    env.reset(seed=7)
    env.p_state[..., :2] = 10 + 0.3 * (env.p_state[..., :2] - 10)              # crowd them: kills and team collisions happen
    env.e_state[..., :2] = 10 + 0.15 * (env.e_state[..., :2] - 10)
    p0, e0, tgt = env.p_state.clone(), env.e_state.clone(), env.target.clone()
    g = torch.Generator(device="cuda").manual_seed(3)
    e_tape = (torch.rand(K, B, E, generator=g, device="cuda", dtype=torch.float64) * 2 - 1).contiguous()
    arena = EnvN2nArena(N, E, B, K, env.device)
    env.rollout(arena, K, 0, None, e_tape, seed=99)
    acts = arena.action.clone()
    assert int(acts.min()) >= 0 and int(acts.max()) <= 8 and len(torch.unique(acts)) == 9
    # (1) per-call API replay
    env2 = BatchedParticleEnvN2N(B, N, E)
    env2.set_state(p0.cpu().numpy(), e0.cpu().numpy(), tgt.cpu().numpy())
    for t in range(K):
        pp, pe, asg = env2.observe()
        assert torch.equal(pp, arena.pp_adj_bits[t]) and torch.equal(pe, arena.pe_adj_bits[t]) and torch.equal(asg, arena.assign[t])
        env2.evader_step(e_tape[t].contiguous())
        reward, done = env2.step(acts[t].contiguous())
        assert torch.equal(reward, arena.reward[t]) and torch.equal(done, arena.done[t])
    for name in ("p_state", "e_state", "p_active", "e_active", "time_step"):
        assert torch.equal(getattr(env, name), getattr(env2, name)), name
    # (2) oracle on the first envs
    prm = ref.default_params()
    for b in range(6):
        p, e = p0[b].cpu().numpy(), e0[b].cpu().numpy()
        pa, ea, ts = np.ones(N, np.uint8), np.ones(E, np.uint8), 0
        for t in range(K):
            e = ref.evaders_move(e, ea, e_tape[t, b].cpu().numpy(), prm)
            p, pa, e, ea, reward, done, ts = ref.step(p, pa, e, ea, acts[t, b].cpu().numpy(), ts, tgt[b].cpu().numpy(), prm)
            assert np.array_equal(reward, arena.reward[t, b].cpu().numpy()) and done == bool(arena.done[t, b].item()), (b, t)
        assert np.array_equal(pa, env.p_active[b].cpu().numpy()) and np.array_equal(ea, env.e_active[b].cpu().numpy())
        np.testing.assert_allclose(env.p_state[b].cpu().numpy(), p, **TOL)
        np.testing.assert_allclose(env.e_state[b].cpu().numpy(), e, **TOL)
    # (3) invariants
    dead_p = env.p_active == 0
    assert dead_p.any() and (env.e_active == 0).any()
    assert (env.p_state[dead_p][:, :2] == 1000.0).all() and (env.e_state[env.e_active == 0][:, :3] == torch.tensor([1000.0, 1000.0, 0.0], device="cuda", dtype=torch.float64)).all()
    pa_t = arena.p_active.to(torch.int16)
    assert (pa_t[1:] <= pa_t[:-1]).all()
    assert int(arena.reward.max()) <= E and int(arena.reward.min()) >= -(N - 1)
    assert (arena.reward[arena.p_active == 0] == 0).all()


def test_facade_reset_and_api_follow_the_reference():
    from distributed_multi_agent_reinforcement_learning_b200.particle_env_n2n import ParticleEnv
    fx = np.load(os.path.join(GOLDEN_DIR, "envn2n_reset_n5_e2_s9.npz"))
    n, e = int(fx["n"]), int(fx["e"])
    np.random.seed(int(fx["seed"]))
    env = ParticleEnv()
    env.initialize(n, e)
    env.reset()
    got_p = np.array([[a.x, a.y, a.phi, a.v] for a in (env.p_list[f"{i}"] for i in env.p_idx)])
    got_e = np.array([[a.x, a.y, a.phi, a.v] for a in (env.e_list[f"{i}"] for i in env.e_idx)])
    assert np.array_equal(got_p.view(np.int64), fx["p_state"].view(np.int64)) and np.array_equal(got_e.view(np.int64), fx["e_state"].view(np.int64))
    assert np.array_equal(np.array(env.target).view(np.int64), fx["target"].view(np.int64))
    p_state, e_state = env.get_team_state(True, rules=False), env.get_team_state(False, rules=False)
    pp = env.get_adj_mat(p_state, p_state, env.p_comm_range, True)
    pe = env.get_adj_mat(p_state, e_state, env.p_sen_range, True)
    assert pp.shape == (n, n) and pe.shape == (n, e) and (np.diag(pp) == 1).all()
    assert env.choose_evader("actor").shape == (n, e)
    env.evader_step(p_state, action=[0.25] * e)
    reward, done, active = env.step([1, 2, 3, 0, 8])
    assert len(reward) == n and isinstance(done, bool) and active == [1] * n and env.time_step == 1
    assert env.reward(True) == [env.agent_reward(i) for i in range(n)] and env.get_done() is False
    assert env.collision_detection(0, True, True)[0] == 1
