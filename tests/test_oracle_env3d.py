"""Pins the CPU oracle of the 3-D particle env (oracle/marl_oracle.c: orc_point_step / orc_env3d_step /
orc_env3d_adjacency) to golden vectors recorded from the unmodified reference environment/env_3d/particle_env.py
(oracle/gen_golden_env3d.py).  Discrete outputs (rewards, active flags, done, adjacency) bit-exact; fp64 states to
1e-12 relative (glibc cos/sin here vs numpy's there)."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, golden

NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "env3d_n*.npz")))
RTOL = 1e-12


def test_fixtures_present():
    assert len(NAMES) >= 5


@pytest.mark.parametrize("name", NAMES)
def test_evader_point_step(oracle, name):
    fx = golden(name)
    p = oracle.Env3dParams.from_fixture(fx)
    for t in range(fx["done"].shape[0]):
        if fx["e_active_before"][t]:
            s = oracle.point_step(fx["e_before"][t], fx["e_action"][t], p.e_vmax, p.ang_lmt, p.v_lmt, p.step_size)
        else:
            s = fx["e_before"][t]
        np.testing.assert_allclose(s, fx["e_moved"][t], rtol=RTOL, atol=1e-13, err_msg=f"{name} t={t}")


@pytest.mark.parametrize("name", NAMES)
def test_step_per_step(oracle, name):
    """Every step restarted from the reference's own pre-step state."""
    fx = golden(name)
    p = oracle.Env3dParams.from_fixture(fx)
    for t in range(fx["done"].shape[0]):
        out = oracle.env3d_step(p, fx["p_before"][t], fx["p_active_before"][t], fx["e_moved"][t], fx["e_active_before"][t],
                                fx["target"], fx["action"][t], t)
        assert np.array_equal(out["reward"], fx["reward"][t]), (name, t)
        assert np.array_equal(out["p_active"], fx["p_active"][t]), (name, t)
        assert out["e_active"] == fx["e_active"][t], (name, t)
        assert out["done"] == fx["done"][t], (name, t)
        np.testing.assert_allclose(out["p_state"], fx["p_after"][t], rtol=RTOL, atol=1e-13)
        np.testing.assert_allclose(out["e_state"], fx["e_after"][t], rtol=RTOL, atol=1e-13)


@pytest.mark.parametrize("name", NAMES)
def test_closed_loop(oracle, name):
    """From the initial state only, with the action tapes: no re-synchronisation with the reference."""
    fx = golden(name)
    p = oracle.Env3dParams.from_fixture(fx)
    ps, pa = fx["p_before"][0].copy(), fx["p_active_before"][0].copy()
    es, ea = fx["e_before"][0].copy(), int(fx["e_active_before"][0])
    total = 0
    for t in range(fx["done"].shape[0]):
        pp, pe = oracle.env3d_adjacency(p, ps, pa, es)
        assert np.array_equal(pp, fx["pp_adj"][t]), (name, t)
        assert np.array_equal(pe, fx["pe_adj"][t][:, 0]), (name, t)
        if ea:
            es = oracle.point_step(es, fx["e_action"][t], p.e_vmax, p.ang_lmt, p.v_lmt, p.step_size)
        out = oracle.env3d_step(p, ps, pa, es, ea, fx["target"], fx["action"][t], t)
        ps, pa, es, ea = out["p_state"], out["p_active"], out["e_state"], out["e_active"]
        assert np.array_equal(out["reward"], fx["reward"][t]), (name, t)
        assert np.array_equal(pa, fx["p_active"][t]), (name, t)
        assert out["done"] == fx["done"][t], (name, t)
        total += int(out["reward"].sum())
    np.testing.assert_allclose(ps, fx["p_after"][-1], rtol=1e-10, atol=1e-12)
    assert total == int(fx["reward"].sum())


def test_batched_iteration_matches_single(oracle):
    fx = golden("env3d_n12_s5_crowd")
    p = oracle.Env3dParams.from_fixture(fx)
    B, N = 3, p.N
    st = dict(p_state=np.ascontiguousarray(np.repeat(fx["p_before"][:1], B, 0)), p_active=np.ones((B, N), np.uint8),
              e_state=np.ascontiguousarray(np.repeat(fx["e_before"][:1], B, 0)), e_active=np.ones(B, np.uint8),
              target=np.ascontiguousarray(np.repeat(fx["target"][None], B, 0)), time_step=np.zeros(B, np.int32),
              reward=np.zeros((B, N), np.int32), done=np.zeros(B, np.uint8), pp_adj=np.zeros((B, N, N), np.uint8),
              pe_adj=np.zeros((B, N), np.uint8))
    for t in range(20):
        st["action"] = np.ascontiguousarray(np.repeat(fx["action"][t][None], B, 0))
        st["e_action"] = np.ascontiguousarray(np.repeat(fx["e_action"][t][None], B, 0))
        oracle.env3d_iteration(p, st)
        for b in range(B):
            assert np.array_equal(st["reward"][b], fx["reward"][t])
            assert np.array_equal(st["pp_adj"][b], fx["pp_adj"][t])
            assert np.array_equal(st["p_active"][b], fx["p_active"][t])
