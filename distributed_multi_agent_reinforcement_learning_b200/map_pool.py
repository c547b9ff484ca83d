"""On-disk map pool of the reference's multi-node driver (MAPPO_parallel_main.py:65-75 reads it, :109-121 slices it).

Seven `.npy` files in one directory describe M pre-generated maps (the authors use M = 2000):

    obstacle_map_list.npy            [M, W, H]     occupancy grid of every map (OccupiedGridMap.grid_map, 0/1)
    boundary_map_list.npy            [M, W, H]     inner-boundary grid (get_boundary_map, pursuit_env.py:18-27)
    hash_map_list.npy                [M, W, H, O]  "raser" visibility table (get_raser_map, pursuit_env.py:29-53): entry
                                                   [x, y, k] = 1 iff boundary cell k is the first one a beam from (x, y) meets;
                                                   k indexes the map's boundary cells in np.argwhere order, zero-padded to O
    obstacle_list.npy                [sum n, 2]    occupied cells of all maps, concatenated (ragged)
    obstacle_num_list.npy            [M]           n per map  -> map i owns rows [num[:i].sum(), num[:i+1].sum())
    boundary_obstacle_list.npy       [sum nb, 2]   boundary cells of all maps, concatenated, np.argwhere (row-major) order
    boundary_obstacle_num_list.npy   [M]           nb per map

`MapPool.from_engine` writes a pool out of the device tables of a `BatchedPursuitEnv` (whose boundary / raser kernels are
bit-exact against the reference's functions, tests/test_gpu_env.py); `MapPool.load` + `install` put a pool into an engine: the
occupancy grids are uploaded, the sensor tables are rebuilt on the GPU and — `verify=True` — compared with the file's own
boundary grids, boundary lists and hash maps, so a pool produced by the reference's generator is checked entry by entry
before it is trusted.  `map_info(i)` returns what the driver hands to its workers (:109-121).
"""
import os

import numpy as np

from . import maps

FILES = ("obstacle_map_list", "boundary_map_list", "hash_map_list", "obstacle_list", "obstacle_num_list",
         "boundary_obstacle_list", "boundary_obstacle_num_list")


class MapPoolError(ValueError):
    pass


class MapPool:
    def __init__(self, obstacle_map_list, boundary_map_list, hash_map_list, obstacle_list, obstacle_num_list,
                 boundary_obstacle_list, boundary_obstacle_num_list):
        self.obstacle_map_list = np.asarray(obstacle_map_list)
        self.boundary_map_list = np.asarray(boundary_map_list)
        self.hash_map_list = np.asarray(hash_map_list)
        self.obstacle_list = np.asarray(obstacle_list).reshape(-1, 2)
        self.obstacle_num_list = np.asarray(obstacle_num_list).reshape(-1)
        self.boundary_obstacle_list = np.asarray(boundary_obstacle_list).reshape(-1, 2)
        self.boundary_obstacle_num_list = np.asarray(boundary_obstacle_num_list).reshape(-1)
        self.check()

    # ------------------------------------------------------------------------------------------------ structure
    def __len__(self):
        return int(self.obstacle_map_list.shape[0])

    @property
    def shape(self):
        """(M, W, H, O)."""
        return tuple(int(v) for v in self.hash_map_list.shape)

    def check(self):
        om, bm, hm = self.obstacle_map_list, self.boundary_map_list, self.hash_map_list
        if om.ndim != 3 or bm.shape != om.shape or hm.ndim != 4 or hm.shape[:3] != om.shape:
            raise MapPoolError(f"inconsistent shapes: obstacle_map {om.shape}, boundary_map {bm.shape}, hash_map {hm.shape}")
        M = om.shape[0]
        for name, num, lst in (("obstacle", self.obstacle_num_list, self.obstacle_list),
                               ("boundary_obstacle", self.boundary_obstacle_num_list, self.boundary_obstacle_list)):
            if num.shape != (M,) or (num < 0).any() or int(num.sum()) != lst.shape[0]:
                raise MapPoolError(f"{name}_num_list does not describe {name}_list ({num.shape}, sum {int(num.sum())}, rows {lst.shape[0]})")
        if int(self.boundary_obstacle_num_list.max(initial=0)) > hm.shape[3]:
            raise MapPoolError("a map has more boundary cells than the hash map's last dimension")

    def _rows(self, lst, num, i):
        lo = int(num[:i].sum())                                  # the driver's own prefix-sum slicing (:112-116)
        return lst[lo:lo + int(num[i])]

    def obstacles(self, i):
        return self._rows(self.obstacle_list, self.obstacle_num_list, i)

    def boundary_obstacles(self, i):
        return self._rows(self.boundary_obstacle_list, self.boundary_obstacle_num_list, i)

    def map_info(self, i):
        """[obstacle_map, boundary_map, obstacles, boundary_obstacles, hash_map] of map i (MAPPO_parallel_main.py:109-121)."""
        return [self.obstacle_map_list[i], self.boundary_map_list[i], self.obstacles(i).tolist(),
                self.boundary_obstacles(i).tolist(), self.hash_map_list[i]]

    # ------------------------------------------------------------------------------------------------ files
    def save(self, directory):
        os.makedirs(directory, exist_ok=True)
        for f in FILES:
            np.save(os.path.join(directory, f + ".npy"), getattr(self, f))

    @classmethod
    def load(cls, directory, mmap=False):
        missing = [f for f in FILES if not os.path.exists(os.path.join(directory, f + ".npy"))]
        if missing:
            raise MapPoolError(f"{directory}: missing {', '.join(m + '.npy' for m in missing)}")
        return cls(*(np.load(os.path.join(directory, f + ".npy"), mmap_mode="r" if mmap else None) for f in FILES))

    # ------------------------------------------------------------------------------------------------ reference semantics on the host
    @classmethod
    def from_tables(cls, grids, boundaries, boundary_xy, boundary_count, raser, dtype=np.uint8):
        """grids / boundaries u8 [M,W,H]; boundary_xy int [M,O,2] (first boundary_count[m] rows valid, argwhere order);
        raser u8 [M,W,H,O]."""
        grids = np.asarray(grids)
        M = grids.shape[0]
        counts = np.asarray(boundary_count).astype(np.int64).reshape(M)
        occ = [np.argwhere(grids[m] != 0) for m in range(M)]        # OccupiedGridMap.obstacles are (x, y) cells; row-major here
        return cls(grids.astype(dtype), np.asarray(boundaries).astype(dtype), np.asarray(raser).astype(dtype),
                   np.concatenate(occ).astype(np.int64) if M else np.zeros((0, 2), np.int64),
                   np.array([len(o) for o in occ], np.int64),
                   np.concatenate([np.asarray(boundary_xy[m][:counts[m]]) for m in range(M)]).astype(np.int64) if M else np.zeros((0, 2), np.int64),
                   counts)

    # ------------------------------------------------------------------------------------------------ device engine
    @classmethod
    def from_engine(cls, engine, dtype=np.uint8):
        """Pool of the M maps an engine currently holds (its tables were built by csrc/sensor_maps.cu)."""
        p = engine.params
        grids = maps.unpack_words(engine.grid_bits.cpu().numpy(), p.H)
        bnd = maps.unpack_words(engine.boundary_bits.cpu().numpy(), p.H)
        counts = engine.boundary_count.cpu().numpy()
        if int(counts.max(initial=0)) > p.O:
            raise MapPoolError(f"a map has {int(counts.max())} boundary cells > map.num_max_obstacle = {p.O}")
        raser = maps.unpack_words(engine.raser_bits.cpu().numpy(), p.O).reshape(engine.M, p.W, p.H, p.O)
        return cls.from_tables(grids, bnd, engine.boundary_xy.cpu().numpy(), counts, raser, dtype=dtype)

    def install(self, engine, indices=None, verify=True):
        """Uploads maps `indices` (default: the first engine.M) into the engine, rebuilds the sensor tables on the GPU and, with
        verify=True, requires them to equal the pool's boundary grids, boundary lists and hash maps entry by entry."""
        p = engine.params
        M, W, H, O = self.shape
        if (W, H) != (p.W, p.H):
            raise MapPoolError(f"pool maps are {W}x{H}, the engine's config says {p.W}x{p.H}")
        idx = np.arange(engine.M) if indices is None else np.asarray(indices, dtype=np.int64).reshape(-1)
        if idx.shape[0] != engine.M or (idx < 0).any() or (idx >= M).any():
            raise MapPoolError(f"need {engine.M} map indices in [0, {M})")
        grids = (np.asarray(self.obstacle_map_list[idx]) != 0).astype(np.uint8)
        engine.set_maps(grids)
        if not verify:
            return idx
        counts = engine.boundary_count.cpu().numpy()
        want_counts = self.boundary_obstacle_num_list[idx]
        if not np.array_equal(counts, want_counts):
            raise MapPoolError("boundary cell counts differ from boundary_obstacle_num_list")
        if int(counts.max(initial=0)) > p.O:
            raise MapPoolError(f"a map has {int(counts.max())} boundary cells > map.num_max_obstacle = {p.O}")
        bnd = maps.unpack_words(engine.boundary_bits.cpu().numpy(), p.H)
        if not np.array_equal(bnd, (np.asarray(self.boundary_map_list[idx]) != 0).astype(np.uint8)):
            raise MapPoolError("boundary grids differ from boundary_map_list")
        bxy = engine.boundary_xy.cpu().numpy()
        raser = maps.unpack_words(engine.raser_bits.cpu().numpy(), p.O).reshape(engine.M, p.W, p.H, p.O)
        for j, i in enumerate(idx):
            n = int(counts[j])
            if not np.array_equal(bxy[j, :n], self.boundary_obstacles(int(i))):
                raise MapPoolError(f"map {int(i)}: boundary cell order differs from boundary_obstacle_list")
            hm = np.asarray(self.hash_map_list[int(i)])
            if hm[..., n:].any() or not np.array_equal(raser[j, ..., :n], (hm[..., :n] != 0).astype(np.uint8)):
                raise MapPoolError(f"map {int(i)}: visibility table differs from hash_map_list")
        return idx
