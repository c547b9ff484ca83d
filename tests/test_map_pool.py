"""Map pool on-disk format (MAPPO_parallel_main.py:65-75, :109-121 of the reference) — SURVEY §8(f) rank 2.

The pools used here are assembled from tests/golden/env_*.npz, i.e. from grids, boundary grids, boundary lists and "raser" tables
that the unmodified reference produced (oracle/gen_golden.py), so `install(verify=True)` on the GPU checks our table-building
kernels against the reference through the pool path, entry by entry."""
import os

import numpy as np
import pytest

from conftest import env_fixture_names, golden, unpack_bits


def _reference_pool(W=60, H=55):
    from distributed_multi_agent_reinforcement_learning_b200.map_pool import MapPool
    grids, bnds, bxy, cnt, raser = [], [], [], [], []
    O = 176
    for name in env_fixture_names():
        fx = golden(name)
        if fx["grid"].shape != (W, H) or int(fx["param_O"]) != O:
            continue
        n = len(fx["boundary_xy"])
        pad = np.zeros((O, 2), np.int64)
        pad[:n] = fx["boundary_xy"]
        r = np.zeros((W, H, O), np.uint8)
        r[..., :n] = unpack_bits(fx["raser_packed"], n)
        grids.append(fx["grid"]); bnds.append(fx["boundary"]); bxy.append(pad); cnt.append(n); raser.append(r)
    assert len(grids) >= 3
    return MapPool.from_tables(np.stack(grids), np.stack(bnds), np.stack(bxy), np.array(cnt), np.stack(raser))


def test_pool_files_round_trip_and_driver_slicing(tmp_path):
    from distributed_multi_agent_reinforcement_learning_b200.map_pool import FILES, MapPool, MapPoolError
    pool = _reference_pool()
    d = str(tmp_path / "map_data")
    pool.save(d)
    assert sorted(os.listdir(d)) == sorted(f + ".npy" for f in FILES)
    back = MapPool.load(d)
    M, W, H, O = back.shape
    assert (M, W, H, O) == (len(pool), 60, 55, 176)
    # the driver's own slicing (MAPPO_parallel_main.py:109-121), restated on the raw files
    raw = {f: np.load(os.path.join(d, f + ".npy")) for f in FILES}
    for idx in range(M):
        lo, hi = raw["obstacle_num_list"][:idx].sum(), raw["obstacle_num_list"][:idx + 1].sum()
        blo, bhi = raw["boundary_obstacle_num_list"][:idx].sum(), raw["boundary_obstacle_num_list"][:idx + 1].sum()
        info = back.map_info(idx)
        assert np.array_equal(info[0], raw["obstacle_map_list"][idx]) and np.array_equal(info[1], raw["boundary_map_list"][idx])
        assert info[2] == raw["obstacle_list"][lo:hi].tolist() and info[3] == raw["boundary_obstacle_list"][blo:bhi].tolist()
        assert np.array_equal(info[4], raw["hash_map_list"][idx])
        # semantics: the occupied cells are the grid's ones, the boundary cells come in np.argwhere order and index the hash map
        assert np.array_equal(np.argwhere(info[0] != 0), np.asarray(info[2]))
        assert np.array_equal(np.argwhere(info[1] != 0), np.asarray(info[3]))
        assert not info[4][..., len(info[3]):].any()
        assert set(map(tuple, info[3])) <= set(map(tuple, info[2]))
    # a pool whose prefix sums do not describe the ragged list is refused
    np.save(os.path.join(d, "obstacle_num_list.npy"), raw["obstacle_num_list"] + 1)
    with pytest.raises(MapPoolError):
        MapPool.load(d)
    os.remove(os.path.join(d, "hash_map_list.npy"))
    with pytest.raises(MapPoolError):
        MapPool.load(d)


@pytest.mark.gpu
def test_pool_install_verifies_against_reference_tables_and_exports_them(tmp_path):
    torch = pytest.importorskip("torch")
    from distributed_multi_agent_reinforcement_learning_b200 import default_config
    from distributed_multi_agent_reinforcement_learning_b200.map_pool import MapPool, MapPoolError
    from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import BatchedPursuitEnv
    pool = _reference_pool()
    M = len(pool)
    cfg = default_config(env__num_defender=8)
    env = BatchedPursuitEnv(cfg, 2 * M, device="cuda:0", num_maps=M)
    pool.install(env, verify=True)                       # GPU boundary / raser kernels == the reference's tables, via the pool
    out = MapPool.from_engine(env)
    for f in ("obstacle_map_list", "boundary_map_list", "hash_map_list", "obstacle_list", "obstacle_num_list",
              "boundary_obstacle_list", "boundary_obstacle_num_list"):
        assert np.array_equal(getattr(out, f), getattr(pool, f)), f
    # a subset in another order, through the files
    d = str(tmp_path / "pool")
    out.save(d)
    env2 = BatchedPursuitEnv(cfg, 4, device="cuda:0", num_maps=2)
    MapPool.load(d, mmap=True).install(env2, indices=[M - 1, 0], verify=True)
    assert torch.equal(env2.raser_bits[0], env.raser_bits[M - 1]) and torch.equal(env2.raser_bits[1], env.raser_bits[0])
    # a tampered visibility table is caught
    bad = MapPool.load(d)
    hm = np.array(bad.hash_map_list)
    x, y, k = np.argwhere(hm[0])[0]
    hm[0, x, y, k] = 0
    bad.hash_map_list = hm
    with pytest.raises(MapPoolError):
        bad.install(env, verify=True)
