"""Host-side map / initial-state generation and bit packing for the pursuit-evasion family.

Episode initialisation mirrors the reference's `Pursuit_Env.reset` call order
(environment/pursuit_evasion_game/pursuit_env.py:60-73 -> base_env.py:37-162, Occupied_Grid_Map.py:46-62) and, in
`RefRng` mode, consumes the *global* `random` / `np.random` streams in exactly the reference's order, so that
`random.seed(s); np.random.seed(s); Pursuit_Env(cfg).reset()` yields the reference's map, pursuers, evader and
target bit for bit (checked against tests/golden in tests/test_host_logic.py).  `GenRng` is the fast
generator-backed source used for large synthetic batches.

The per-map sensor tables (boundary list, "raser" visibility table) are built on the GPU
(csrc/sensor_maps.cu via marl_raser_map_build); nothing here walks beams.
"""
import random as _pyrandom

import numpy as np


# ------------------------------------------------------------------------------------------------ bit packing
def pack_grid(grid):
    """u8 [..., W, H] -> int32 bit words [..., W, HW], bit (y & 31) of word y >> 5 (include/marl_b200.h)."""
    g = np.asarray(grid, dtype=np.uint8)
    H = g.shape[-1]
    HW = (H + 31) // 32
    pad = HW * 32 - H
    if pad:
        g = np.concatenate([g, np.zeros(g.shape[:-1] + (pad,), np.uint8)], axis=-1)
    b = np.packbits(g.reshape(g.shape[:-1] + (HW, 32)), axis=-1, bitorder="little")
    return np.ascontiguousarray(b).view(np.uint32).reshape(g.shape[:-1] + (HW,)).view(np.int32)


def unpack_words(words, n):
    """int32/uint32 bit words [..., NW] -> u8 [..., n]."""
    w = np.ascontiguousarray(np.asarray(words)).view(np.uint32)
    bits = np.unpackbits(w.view(np.uint8).reshape(w.shape[:-1] + (w.shape[-1] * 4,)), axis=-1, bitorder="little")
    return bits[..., :n]


def dilate(grid, e):
    """Chebyshev inflation by e cells inside the map (Occupied_Grid_Map.py:157-166 as a set operation)."""
    g = np.asarray(grid, dtype=bool)
    out = g.copy()
    W, H = g.shape[-2:]
    for dx in range(-e, e + 1):
        for dy in range(-e, e + 1):
            xs0, xs1 = max(0, dx), min(W, W + dx)
            ys0, ys1 = max(0, dy), min(H, H + dy)
            out[..., xs0:xs1, ys0:ys1] |= g[..., xs0 - dx:xs1 - dx, ys0 - dy:ys1 - dy]
    return out.astype(np.uint8)


# ------------------------------------------------------------------------------------------------ random sources
class RefRng:
    """Draws from the interpreter-global `random` and `np.random` streams, like the reference does."""

    def block_kind(self):
        return _pyrandom.randrange(1)                    # Occupied_Grid_Map.py:60 `random.randrange(len(shape))`

    def normal2(self, center, variance):
        return np.random.normal(center, variance, 2)     # Occupied_Grid_Map.py:61

    def randint(self, lo, hi):
        return _pyrandom.randint(lo, hi)                 # base_env.py:65-66

    def rand2(self):
        return np.random.rand(2)                         # base_env.py:89,137


class GenRng:
    """numpy Generator-backed source (throughput runs; no reference counterpart for the stream itself)."""

    def __init__(self, seed):
        self.g = np.random.default_rng(seed)

    def block_kind(self):
        return 0

    def normal2(self, center, variance):
        return self.g.normal(center, variance, 2)

    def randint(self, lo, hi):
        return int(self.g.integers(lo, hi + 1))

    def rand2(self):
        return self.g.random(2)


# ------------------------------------------------------------------------------------------------ generation
def make_obstacle_grid(map_cfg, rng):
    """`num_obstacle_block` rectangular blocks around N(center, variance) (base_env.py:37-50).  The block
    footprint is x,y in [-3,3) — 6x6, not the nominal 6x7 (Occupied_Grid_Map.py:48-49 uses data[0] twice)."""
    W, H = int(map_cfg.map_size[0]), int(map_cfg.map_size[1])
    grid = np.zeros((W, H), np.uint8)
    shape = (6, 7)
    for _ in range(int(map_cfg.num_obstacle_block)):
        rng.block_kind()
        c = rng.normal2(map_cfg.center, map_cfg.variance)
        for ox in range(int(-shape[0] / 2), int(shape[0] / 2)):
            for oy in range(int(-shape[1] / 2), int(shape[0] / 2)):
                cx, cy = round(float(ox + c[0])), round(float(oy + c[1]))
                if 0 <= cx < W and 0 <= cy < H:
                    grid[cx, cy] = 1
    return grid


def draw_target(inflated, rng):
    """base_env.py:52-70: uniform integer cell, rejected while occupied in the inflated map."""
    W, H = inflated.shape
    while True:
        t = (rng.randint(0, W - 1), rng.randint(0, H - 1))
        if not inflated[t[0], t[1]]:
            return t


def _cell(v):
    return round(float(v))   # Python round(): half to even


def place_pursuers(inflated, n, comm_range, rng, min_dist=4, extend=2):
    """base_env.py:72-120.  `inflated` is updated in place with every accepted pursuer's inflated footprint.
    Returns (positions [n,2] f64, pursuer cells list)."""
    W, H = inflated.shape
    scale = np.array([W - 1, H - 1])
    pos, cells = [], []
    while len(pos) < n:
        p = tuple(rng.rand2() * scale)
        if inflated[_cell(p[0]), _cell(p[1])]:
            continue
        ok = not pos
        if pos:
            d = [float(np.linalg.norm((p[0] - q[0], p[1] - q[1]))) for q in pos]
            close = sum(1 for v in d if v < min_dist)
            linked = sum(1 for v in d if v < comm_range)
            ok = close == 0 and 0 < linked <= 2
        if ok:
            pos.append(p)
            c = (_cell(p[0]), _cell(p[1]))
            if c not in cells:
                cells.append(c)
            inflated[max(0, c[0] - extend):c[0] + extend + 1, max(0, c[1] - extend):c[1] + extend + 1] = 1
    return np.array(pos, dtype=np.float64), cells


def place_evader(inflated, cells, sen_range, rng):
    """base_env.py:122-152 with is_percepted=True: a free point within sen_range of some pursuer cell."""
    W, H = inflated.shape
    scale = np.array([W - 1, H - 1])
    while True:
        p = tuple(rng.rand2() * scale)
        if inflated[_cell(p[0]), _cell(p[1])]:
            continue
        for c in cells:
            if float(np.linalg.norm([c[0] - p[0], c[1] - p[1]])) < sen_range:
                return np.array(p, dtype=np.float64)


def reset_one(cfg, rng):
    """One full episode initialisation in the reference's draw order.  Returns a dict of numpy arrays."""
    grid = make_obstacle_grid(cfg.map, rng)
    inflated_static = dilate(grid, 2)
    target = draw_target(inflated_static, rng)
    work = inflated_static.copy()
    p_xy, cells = place_pursuers(work, int(cfg.env.num_defender), float(cfg.defender.comm_range), rng)
    e_xy = place_evader(work, cells, float(cfg.defender.sen_range), rng)
    N = int(cfg.env.num_defender)
    p_state = np.zeros((N, 4), np.float64)
    p_state[:, :2] = p_xy
    e_state = np.zeros(4, np.float64)
    e_state[:2] = e_xy
    return dict(grid=grid, inflated=inflated_static, target=np.array(target, np.int32), p_state=p_state,
                e_state=e_state)


def action_table(vmax):
    """Agent.actions_mat (agent.py:57-60) — computed with numpy on the host so cos/sin are the values the
    reference would see on this machine."""
    theta = [i * np.pi / 4 for i in range(0, 8)]
    tab = [[np.cos(t) * vmax, np.sin(t) * vmax] for t in theta]
    tab.append([0.0, 0.0])
    return np.array(tab, dtype=np.float64)


def beam_directions(num_beams):
    """(cos, sin)(beam * 2*pi / num_beams) with the reference's operation order (pursuit_env.py:37-39)."""
    return np.array([[np.cos(b * 2 * np.pi / num_beams), np.sin(b * 2 * np.pi / num_beams)] for b in range(num_beams)],
                    dtype=np.float64)
