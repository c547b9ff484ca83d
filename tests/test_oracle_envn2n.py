"""The CPU restatement of the 2-D N-vs-E particle env (oracle/envn2n_ref.py) against golden vectors produced by executing the
unmodified reference (environment/env_n2n/particle_env.py; oracle/gen_golden_envn2n.py).  Same numpy arithmetic => bit-exact."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import envn2n_ref as ref

FIXTURES = sorted(p for p in glob.glob(os.path.join(GOLDEN_DIR, "envn2n_*.npz")) if "reset" not in p)
IDS = [os.path.basename(p)[:-4] for p in FIXTURES]


def params(fx):
    return ref.default_params(p_vmax=float(fx["p_vmax"]), e_vmax=float(fx["e_vmax"]), kill_radius=float(fx["kill_radius"]),
                              ang_lmt=float(fx["ang_lmt"]), step_size=float(fx["step_size"]), episode_limit=int(fx["episode_limit"]),
                              p_comm_range=float(fx["p_comm_range"]), p_sen_range=float(fx["p_sen_range"]))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_episode_bit_exact(path):
    fx = np.load(path)
    prm = params(fx)
    T = len(fx["done"])
    p, e = fx["p_before"][0].copy(), fx["e_before"][0].copy()
    pa, ea = fx["p_active_before"][0].copy(), fx["e_active_before"][0].copy()
    ts = 0
    kills = 0
    for t in range(T):
        assert np.array_equal(bits(p), bits(fx["p_before"][t])) and np.array_equal(bits(e), bits(fx["e_before"][t])), t
        assert np.array_equal(ref.adjacency(p, pa, p, prm["p_comm_range"]), fx["pp_adj"][t]), t
        assert np.array_equal(ref.adjacency(p, pa, e, prm["p_sen_range"]), fx["pe_adj"][t]), t
        assert np.array_equal(ref.choose_evader(p, pa, e, ea, prm["p_sen_range"]), fx["assign"][t]), t
        e = ref.evaders_move(e, ea, fx["e_action"][t], prm)
        assert np.array_equal(bits(e), bits(fx["e_moved"][t])), t
        p, pa, e, ea, reward, done, ts = ref.step(p, pa, e, ea, fx["action"][t], ts, fx["target"], prm)
        assert np.array_equal(reward, fx["reward"][t]) and done == bool(fx["done"][t]), t
        assert np.array_equal(pa, fx["p_active"][t]) and np.array_equal(ea, fx["e_active"][t]), t
        assert np.array_equal(bits(p), bits(fx["p_after"][t])) and np.array_equal(bits(e), bits(fx["e_after"][t])), t
        kills += int((fx["e_active_before"][t] > fx["e_active"][t]).sum())
    assert T >= 50


def test_fixtures_cover_kills_and_team_collisions():
    seen_kill = seen_team = seen_inactive_turn = False
    for path in FIXTURES:
        fx = np.load(path)
        seen_kill |= bool((fx["e_active_before"] > fx["e_active"]).any())
        seen_team |= bool((fx["reward"] < 0).any())
        # an inactive pursuer keeps turning (Pursuer.step updates phi outside `if self.active`)
        dead = (fx["p_active_before"] == 0)
        seen_inactive_turn |= bool((dead & (fx["p_after"][..., 2] != fx["p_before"][..., 2])).any())
    assert seen_kill and seen_team and seen_inactive_turn


def test_reset_matches_reference_stream():
    fx = np.load(os.path.join(GOLDEN_DIR, "envn2n_reset_n5_e2_s9.npz"))
    np.random.seed(int(fx["seed"]))
    p, e, target = ref.reset(int(fx["n"]), int(fx["e"]))
    assert np.array_equal(bits(p), bits(fx["p_state"])) and np.array_equal(bits(e), bits(fx["e_state"]))
    assert np.array_equal(bits(target), bits(fx["target"]))
