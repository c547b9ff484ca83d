"""TEST INFRASTRUCTURE — imports the *unmodified* reference (read-only at /root/reference) in this
container so that golden vectors can be generated from it.  Never imported by the product package.

The reference is pure Python but needs packages that are absent here.  None of the stubs
below touch arithmetic except ``find_boundaries`` (scikit-image 0.19.3, pinned in the reference's
environment.yml, not vendored): that one is *restated* — parity for it is "unpinned" by the reference's
own tests (there are none) and is property-checked in tests/test_oracle_golden.py.

This module only works where /root/reference exists (the build container); it cannot travel to the GPU
box, which is why its outputs are committed under tests/golden/ (see oracle/gen_golden.py).
"""
import os
import sys
import types
import random
from types import SimpleNamespace as NS

import numpy as np

REFERENCE_ROOT = os.environ.get("MARL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "environment", "pursuit_evasion_game"))


def _stub(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


def find_boundaries(img, mode="inner"):
    """Restatement of skimage.segmentation.find_boundaries(mode='inner') for a 2-D label image with
    connectivity 1 and background 0 (reference call site: environment/pursuit_evasion_game/pursuit_env.py:20).
    A cell is an inner boundary iff it is foreground and its 4-neighbourhood (edge-replicated) is not constant."""
    assert mode == "inner"
    a = np.asarray(img)
    h, w = a.shape
    pad = np.pad(a, 1, mode="edge")
    mx = a.copy()
    mn = a.copy()
    for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1)):
        nb = pad[1 + dx:1 + dx + h, 1 + dy:1 + dy + w]
        mx = np.maximum(mx, nb)
        mn = np.minimum(mn, nb)
    return (mx != mn) & (a != 0)


_loaded = {}


def load_reference():
    """Returns a namespace with the reference modules: pe (pursuit_env), agent, astar, ogm, base_env,
    mappo (DHGN.mappo_parallel), replay_buffer, normalization."""
    if _loaded:
        return NS(**_loaded)
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "hydra" not in sys.modules:
        hy = _stub("hydra")
        hy.main = lambda **kw: (lambda f: f)
        hy.utils = _stub("hydra.utils", instantiate=None)
    for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "imageio", "imageio.v2", "sko"):
        if n not in sys.modules:
            _stub(n)
    if "matplotlib.animation" not in sys.modules:
        _stub("matplotlib.animation", FuncAnimation=None)
        _stub("matplotlib.path", Path=None)
        _stub("matplotlib.font_manager", FontProperties=None)
        sys.modules["matplotlib.patches"].PathPatch = None
        _stub("sko.PSO", PSO=None)
        _stub("sko.SA", SA=None)
    if "skimage" not in sys.modules:
        _stub("skimage")
        _stub("skimage.segmentation", find_boundaries=find_boundaries)
    import warnings
    warnings.filterwarnings("ignore")
    import environment.pursuit_evasion_game.pursuit_env as pe
    import environment.pursuit_evasion_game.agent as agent
    import environment.pursuit_evasion_game.astar as astar
    import environment.pursuit_evasion_game.Occupied_Grid_Map as ogm
    import environment.pursuit_evasion_game.base_env as base_env
    # numba >= 0.59 is nopython-only and cannot type the OccupiedGridMap argument; the reference relied on
    # numba 0.56's object-mode fallback, i.e. plain Python semantics.
    if hasattr(pe.get_raser_map, "py_func"):
        pe.get_raser_map = pe.get_raser_map.py_func
        pe.prange = range
    import DHGN.mappo_parallel as mappo
    import DHGN.replay_buffer as replay_buffer
    import DHGN.normalization as normalization
    _loaded.update(pe=pe, agent=agent, astar=astar, ogm=ogm, base_env=base_env, mappo=mappo,
                   replay_buffer=replay_buffer, normalization=normalization)
    return NS(**_loaded)


def make_cfg(num_defender=15, depth=1, max_steps=150, map_size=(60, 55), center=(30, 25),
             num_max_obstacle=176, extend_dis=1, use_reward_norm=True, sample_epi_num=1,
             embedding_dim=128, difficulty=10):
    """config.yaml:13-94 as nested namespaces (attribute access is all the reference uses)."""
    return NS(
        env=NS(name="Pursuit_Env", state_dim=4, action_dim=9, attacker_class="Evader", defender_class="Pursuer",
               max_steps=max_steps, num_attacker=1, num_defender=num_defender, num_target=1, step_size=0.1,
               difficulty=difficulty),
        sensor=NS(num_beams=36, radius=8),
        map=NS(center=list(center), map_size=list(map_size), num_obstacle_block=5, resolution=1, variance=10,
               num_max_obstacle=num_max_obstacle),
        attacker=NS(DOF=2, collision_radius=0.5, comm_range=16, sen_range=8, step_size=0.1, tau=0.2, vmax=4,
                    extend_dis=extend_dis),
        defender=NS(DOF=2, collision_radius=0.5, comm_range=16, sen_range=8, step_size=0.1, tau=0.2, vmax=2),
        algo=NS(learner_device="cpu", worker_device="cpu", evaluator_device="cpu", max_train_steps=20000000,
                lr=0.0005, gamma=0.99, lamda=0.95, epsilon=0.05, epochs=1, entropy_coef=0.05, save_cwd="./model",
                sample_epi_num=sample_epi_num, use_adv_norm=True, use_agent_specific=True, use_grad_clip=True,
                use_lr_decay=True, use_orthogonal_init=True, use_reward_norm=use_reward_norm,
                use_spectral_norm=True, use_value_clip=True, set_adam_eps=True, mlp_hidden_dim=128,
                rnn_hidden_dim=embedding_dim, embedding_dim=embedding_dim, num_layers=2,
                semantic_level_aggregator="mean", vertex_level_aggregator="mean", fcra_aggregator="mean",
                depth=depth, num_relation=3),
    )


def seed_all(seed: int):
    import torch
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
