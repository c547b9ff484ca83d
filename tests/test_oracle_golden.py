"""Pins the CPU oracle (oracle/marl_oracle.c) to golden vectors produced by executing the unmodified reference
(oracle/gen_golden.py).  Bit-exact for every discrete quantity AND for fp64 pursuer states; the evader's fp64
state goes through libm cos/sin/acos vs numpy's and is held to 1e-12 relative."""
import numpy as np
import pytest

from conftest import env_fixture_names, golden, unpack_bits

NAMES = env_fixture_names()


def _knife_edge(fx, t, thr, what):
    """True if some pairwise distance at step t is within 1e-12 of thr (ulp-level ambiguity)."""
    ps = fx[what][t][:, :2]
    d = np.sqrt(((ps[:, None] - ps[None]) ** 2).sum(-1))
    return bool((np.abs(d - thr) < 1e-12).any())


@pytest.mark.parametrize("name", NAMES)
def test_step_bit_exact(oracle, name):
    fx = golden(name)
    p = oracle.EnvParams.from_fixture(fx)
    T = fx["action"].shape[0]
    col = 0
    for t in range(T):
        out = oracle.env_step(p, fx["p_state"][t], fx["e_after"][t], fx["action"][t], fx["grid"], fx["action_table"],
                              time_step=t, collision=col)
        col = out["collision"]
        assert np.array_equal(out["reward"], fx["reward"][t]), (name, t)
        assert np.array_equal(out["can_apply"], fx["can_apply"][t]), (name, t)
        # fp64 state: bit-exact (view as integers so -0.0/NaN could not hide)
        assert np.array_equal(out["p_state"].view(np.int64), fx["p_state"][t + 1].view(np.int64)), (name, t)
        assert out["collision"] == fx["collision"][t]
        assert out["done"] == fx["done"][t]
        assert out["time_step"] == t + 1


@pytest.mark.parametrize("name", NAMES)
def test_closed_loop_rollout_bit_exact(oracle, name):
    """Oracle steps from the initial state only (no re-synchronisation) with the evader tape."""
    fx = golden(name)
    p = oracle.EnvParams.from_fixture(fx)
    ps = fx["p_state"][0].copy()
    col = 0
    for t in range(fx["action"].shape[0]):
        out = oracle.env_step(p, ps, fx["e_after"][t], fx["action"][t], fx["grid"], fx["action_table"], t, col)
        ps, col = out["p_state"], out["collision"]
    assert np.array_equal(ps.view(np.int64), fx["p_state"][-1].view(np.int64))
    assert col == fx["collision"][-1]


@pytest.mark.parametrize("name", NAMES)
def test_observe_bit_exact(oracle, name):
    fx = golden(name)
    p = oracle.EnvParams.from_fixture(fx)
    ob = int(fx["raser_ob"])
    raser = unpack_bits(fx["raser_packed"], ob)
    o_gold = unpack_bits(fx["o_adj_packed"], ob)
    T1 = fx["p_state"].shape[0]
    for t in range(T1):
        adj = oracle.communicate(p, fx["p_state"][t])
        assert np.array_equal(adj, fx["p_adj"][t]), (name, t)
        e = fx["e_before"][t] if t < T1 - 1 else fx["e_after"][t - 1]
        o_adj, e_adj = oracle.sensor(p, fx["p_state"][t], e, fx["grid"], raser)
        assert np.array_equal(o_adj[:, :ob], o_gold[t]), (name, t)
        assert not o_adj[:, ob:].any()
        assert np.array_equal(e_adj, fx["e_adj"][t]), (name, t)
    # the adj[j,1]=1 quirk is part of the contract
    assert (fx["p_adj"][:, :, 1] == 1).all()


@pytest.mark.parametrize("name", NAMES)
def test_maps_bit_exact(oracle, name):
    fx = golden(name)
    p = oracle.EnvParams.from_fixture(fx)
    b, xy, n = oracle.boundary_map(p, fx["grid"])
    assert n == len(fx["boundary_xy"])
    assert np.array_equal(b, fx["boundary"])
    assert np.array_equal(xy, fx["boundary_xy"])
    raser = oracle.raser_map(p, b, xy, fx["beam_dir"])
    assert np.array_equal(raser, unpack_bits(fx["raser_packed"], n))
    assert np.array_equal(oracle.dilate(p, fx["grid"], 2), fx["inflated"])
    # find_boundaries restatement properties (parity for skimage itself is unpinned: SURVEY §8(c))
    assert not (b & ~fx["grid"].astype(bool)).any()


@pytest.mark.parametrize("name", NAMES)
def test_astar_calls(oracle, name):
    fx = golden(name)
    p = oracle.EnvParams.from_fixture(fx)
    n = int(fx["astar_n"])
    assert n > 0
    blocked = unpack_bits(fx["astar_blocked_packed"], (p.W + 1) * (p.H + 1)).reshape(n, p.W + 1, p.H + 1)
    off = 0
    for i in range(n):
        L = int(fx["astar_path_len"][i])
        gold = fx["astar_path_flat"][off:off + L]
        off += L
        path, nclosed = oracle.astar(p, blocked[i], fx["astar_start"][i], fx["astar_goal"][i])
        assert np.array_equal(path, gold), (name, i)
        assert nclosed == fx["astar_n_closed"][i], (name, i)


@pytest.mark.parametrize("name", NAMES)
def test_evader_teacher_forced(oracle, name):
    """One attacker_step at a time from the golden state: path (discrete) exact, fp64 state to 1e-12."""
    fx = golden(name)
    p = oracle.EnvParams.from_fixture(fx)
    T = fx["action"].shape[0]
    tape = fx["targets_drawn"]   # mid-episode init_target draws, in order
    off = 0
    prev_path = None
    tape_pos = 0
    for t in range(T):
        ev = oracle.EvaderState(fx["e_before"][t], fx["e_target_attr"][t], path=prev_path)
        ev.tape_pos.value = tape_pos
        rc = oracle.evader_step(p, ev, fx["p_state"][t], t, fx["grid"], fx["inflated"], tape)
        assert rc == 0
        tape_pos = ev.tape_pos.value
        L = int(fx["path_len"][t])
        gold_path = fx["path_flat"][off:off + L]
        off += L
        assert ev.path_len.value == L, (name, t)
        assert np.array_equal(ev.path[:L].astype(np.int32), gold_path), (name, t)
        np.testing.assert_allclose(ev.e_state, fx["e_after"][t], rtol=1e-12, atol=1e-12)
        assert np.array_equal(ev.target, fx["target"][t + 1]), (name, t)
        prev_path = gold_path


@pytest.mark.parametrize("name", NAMES)
def test_evader_closed_loop(oracle, name):
    """Evader free-running for the whole episode against the golden pursuer states."""
    fx = golden(name)
    p = oracle.EnvParams.from_fixture(fx)
    T = fx["action"].shape[0]
    tape = fx["targets_drawn"]   # mid-episode init_target draws, in order
    ev = oracle.EvaderState(fx["e_before"][0], fx["e_target_attr"][0])
    for t in range(T):
        assert oracle.evader_step(p, ev, fx["p_state"][t], t, fx["grid"], fx["inflated"], tape) == 0
        np.testing.assert_allclose(ev.e_state, fx["e_after"][t], rtol=1e-9, atol=1e-9, err_msg=f"{name} t={t}")
    assert np.array_equal(ev.target, fx["target"][-1])


@pytest.mark.parametrize("name", NAMES)
def test_welford(oracle, name):
    fx = golden(name)
    p = oracle.EnvParams.from_fixture(fx)
    wf = oracle.Welford(p.N)
    for t in range(fx["reward"].shape[0]):
        out = wf(fx["reward"][t])
        np.testing.assert_array_equal(out, fx["r_norm"][t].astype(np.float32))
    assert (wf(fx["reward"][0], update=False) == wf(fx["reward"][0], update=False)).all()


def test_welford_first_sample_is_zero(oracle):
    wf = oracle.Welford(4)
    assert (wf(np.array([-1, 0, 1, -3])) == 0).all()   # std = x on the first sample (normalization.py:14-17)


def test_kat_total_reward(oracle):
    """seed 0, N=15, demon() policy, 150 steps -> -748 and collision=True."""
    fx = golden("env_n15_s0_demon")
    assert int(fx["reward"].sum()) == -748 and int(fx["collision"][-1]) == 1
    np.testing.assert_allclose(fx["e_before"][0][:2], [33.42948579818793, 14.331032510730052], rtol=0, atol=0)
    assert tuple(fx["target"][0]) == (25, 50)
    assert int(fx["grid"].sum()) == 180 and len(fx["boundary_xy"]) == 100


def test_edge_clip_order_quirk():
    """pursuit_env.py:143-145: the in-place clip is visible to later agents only."""
    fx = golden("env_edge_n8_s7")
    assert fx["reward"][0].tolist() == [1, -1, 0, 0, 0, -1, 0, 0]
