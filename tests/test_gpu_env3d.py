"""GPU parity of the 3-D particle env kernels (csrc/env3d_kernels.cu, through the C-ABI) against
 (a) golden vectors recorded from the unmodified reference (tests/golden/env3d_*.npz) and
 (b) the CPU oracle at batch sizes up to BASELINE config 4's per-GPU share (8192 envs x 32 pursuers).
Discrete outputs (reward, active, done, adjacency) bit-exact; fp64 state to 1e-12 (device vs host cos/sin)."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "env3d_n*.npz")))


def _engine(fx, B=1):
    from distributed_multi_agent_reinforcement_learning_b200.particle_env import BatchedParticleEnv
    return BatchedParticleEnv(B, int(fx["n"]), max_step=int(fx["max_step"]), p_vmax=float(fx["p_vmax"]),
                              e_vmax=float(fx["e_vmax"]), kill_radius=float(fx["kill_radius"]), ang_lmt=float(fx["ang_lmt"]),
                              v_lmt=float(fx["v_lmt"]), step_size=float(fx["step_size"]), comm_range=float(fx["p_comm_range"]),
                              sen_range=float(fx["p_sen_range"]))


def _unpack(bits, n):
    w = bits.cpu().numpy().astype(np.uint32)
    return ((w[..., None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(*w.shape[:-1], -1)[..., :n].astype(np.uint8)


@pytest.mark.parametrize("name", NAMES)
def test_per_call_api_vs_reference_golden(name):
    fx = golden(name)
    N = int(fx["n"])
    eng = _engine(fx)
    dev = eng.device
    for t in range(fx["done"].shape[0]):
        eng.set_state(fx["p_before"][t][None], fx["e_before"][t][None], fx["target"][None],
                      fx["p_active_before"][t][None], fx["e_active_before"][t][None], time_step=t)
        pp, pe = eng.adjacency()
        assert np.array_equal(_unpack(pp, N)[0], fx["pp_adj"][t]), (name, t)
        assert np.array_equal(pe.cpu().numpy()[0], fx["pe_adj"][t][:, 0]), (name, t)
        eng.evader_step(torch.from_numpy(fx["e_action"][t][None].copy()).to(dev))
        np.testing.assert_allclose(eng.e_state.cpu().numpy()[0], fx["e_moved"][t], rtol=1e-12, atol=1e-13)
        eng.e_state.copy_(torch.from_numpy(fx["e_moved"][t][None].copy()))     # re-synchronise the continuous state
        eng.step(torch.from_numpy(fx["action"][t][None].copy()).to(dev))
        assert np.array_equal(eng.reward.cpu().numpy()[0], fx["reward"][t]), (name, t)
        assert np.array_equal(eng.p_active.cpu().numpy()[0], fx["p_active"][t]), (name, t)
        assert int(eng.e_active[0]) == int(fx["e_active"][t]) and int(eng.done[0]) == int(fx["done"][t]), (name, t)
        assert int(eng.time_step[0]) == t + 1
        np.testing.assert_allclose(eng.p_state.cpu().numpy()[0], fx["p_after"][t], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(eng.e_state.cpu().numpy()[0], fx["e_after"][t], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("name", NAMES)
def test_fused_rollout_vs_reference_golden(name):
    """Whole episode in ONE launch from the initial state with the action tapes (no re-synchronisation)."""
    from distributed_multi_agent_reinforcement_learning_b200.particle_env import Env3dArena
    fx = golden(name)
    N, T = int(fx["n"]), fx["done"].shape[0]
    eng = _engine(fx)
    eng.set_state(fx["p_before"][0][None], fx["e_before"][0][None], fx["target"][None], fx["p_active_before"][0][None],
                  fx["e_active_before"][0][None])
    arena = Env3dArena(N, 1, T, eng.device)
    eng.rollout(arena, T, 0, torch.from_numpy(fx["action"][:, None].copy()).to(eng.device),
                torch.from_numpy(fx["e_action"][:, None].copy()).to(eng.device))
    assert np.array_equal(arena.reward.cpu().numpy()[:, 0], fx["reward"])
    assert np.array_equal(arena.done.cpu().numpy()[:, 0], fx["done"])
    assert np.array_equal(arena.active_f32.cpu().numpy()[:, 0], fx["p_active_before"].astype(np.float32))
    assert np.array_equal(_unpack(arena.pp_adj_bits, N)[:, 0], fx["pp_adj"])
    assert np.array_equal(arena.pe_adj.cpu().numpy()[:, 0], fx["pe_adj"][..., 0])
    np.testing.assert_allclose(arena.p_state_f32.cpu().numpy()[:, 0], fx["p_before"].astype(np.float32), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(eng.p_state.cpu().numpy()[0], fx["p_after"][-1], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("B,N,K", [(257, 3, 12), (64, 16, 10), (33, 32, 8), (8192, 32, 6), (9, 64, 5), (5, 100, 4)])
def test_fused_rollout_vs_oracle_batched(oracle, B, N, K):
    """Config-4 shape (8192 x 32) and ragged sizes, device counter RNG for actions, against the OpenMP oracle."""
    from distributed_multi_agent_reinforcement_learning_b200.particle_env import (BatchedParticleEnv, Env3dArena,
                                                                                 counter_uniform_pm1)
    eng = BatchedParticleEnv(B, N)
    eng.reset(seed=B + N)
    p = oracle.Env3dParams.from_dict({n: getattr(eng.params, n) for n, _ in eng.params._fields_})
    st = dict(p_state=eng.p_state.cpu().numpy().copy(), p_active=eng.p_active.cpu().numpy().copy(),
              e_state=eng.e_state.cpu().numpy().copy(), e_active=eng.e_active.cpu().numpy().copy(),
              target=eng.target.cpu().numpy().copy(), time_step=np.zeros(B, np.int32), reward=np.zeros((B, N), np.int32),
              done=np.zeros(B, np.uint8), pp_adj=np.zeros((B, N, N), np.uint8), pe_adj=np.zeros((B, N), np.uint8))
    rng = np.random.default_rng(5)
    act = rng.uniform(-1, 1, (K, B, N, 3))
    eact = rng.uniform(-1, 1, (K, B, 3))
    # spot-check the device RNG against its numpy restatement through one generated step
    arena = Env3dArena(N, B, K, eng.device)
    eng.rollout(arena, K, 0, torch.from_numpy(act).to(eng.device), torch.from_numpy(eact).to(eng.device))
    rewards, dones, pps = [], [], []
    for k in range(K):
        st["action"], st["e_action"] = np.ascontiguousarray(act[k]), np.ascontiguousarray(eact[k])
        oracle.env3d_iteration(p, st)
        rewards.append(st["reward"].copy())
        dones.append(st["done"].copy())
        pps.append(st["pp_adj"].copy())
    assert np.array_equal(arena.reward.cpu().numpy(), np.stack(rewards))
    assert np.array_equal(arena.done.cpu().numpy(), np.stack(dones))
    assert np.array_equal(_unpack(arena.pp_adj_bits, N), np.stack(pps))
    assert np.array_equal(eng.p_active.cpu().numpy(), st["p_active"])
    np.testing.assert_allclose(eng.p_state.cpu().numpy(), st["p_state"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(eng.e_state.cpu().numpy(), st["e_state"], rtol=1e-11, atol=1e-12)
    assert int(np.stack(rewards).min()) < 0 or N <= 3      # the batch really exercises collisions
    # generated actions: device counter RNG == numpy restatement (one env, one step)
    eng2 = BatchedParticleEnv(2, N)
    eng2.reset(seed=1)
    s0 = eng2.p_state.cpu().numpy().copy()
    a2 = Env3dArena(N, 2, 1, eng2.device, adjacency=False)
    eng2.rollout(a2, 1, 0, None, None, seed=1234)
    cmd = np.array([[counter_uniform_pm1(1234, 1 * N + i, 0, c) for c in range(3)] for i in range(N)])
    exp = np.stack([oracle.point_step(s0[1, i], cmd[i], eng2.params.p_vmax, eng2.params.ang_lmt, eng2.params.v_lmt,
                                      eng2.params.step_size) for i in range(N)])
    alive = eng2.p_active.cpu().numpy()[1].astype(bool)
    np.testing.assert_allclose(eng2.p_state.cpu().numpy()[1][alive], exp[alive], rtol=1e-11, atol=1e-12)


def test_facade_matches_reference_reset_and_types():
    from distributed_multi_agent_reinforcement_learning_b200.particle_env import ParticleEnv
    rs = golden("env3d_reset_n4_s9")
    import random
    random.seed(int(rs["seed"]))
    np.random.seed(int(rs["seed"]))
    torch.manual_seed(int(rs["seed"]))
    env = ParticleEnv()
    env.initialize(int(rs["n"]))
    env.reset()
    assert np.array_equal(np.array(env.get_team_state(True, False)), rs["p_state"])      # same RNG draw order
    assert np.array_equal(np.array(env.get_team_state(False, False))[0], rs["e_state"])
    assert np.array_equal(np.array(env.target), rs["target"])
    reward, done, active = env.step(np.zeros((4, 3)))
    assert isinstance(reward, list) and isinstance(reward[0], int) and isinstance(done, bool) and active == [1, 1, 1, 1]
    assert env.get_adj_mat(env.get_team_state(True, False), env.get_team_state(True, False), env.p_comm_range).shape == (4, 4)
    env.evader_step(action=[0.1, 0.0, 1.0])
    assert env.time_step == 1 and env.p_list["0"].active
