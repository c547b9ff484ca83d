"""Checkpoint file set of the reference driver (main.py:146-172): eight pickled modules

    actor.pth critic.pth actor_gnn.pth critic_gnn.pth actor_gru.pth critic_gru.pth actor_mean.pth critic_mean.pth

(+ the same names with `_final` at the end of training) and `recorder.npy` (rows of
(total_step, avgR, stdR, expR, objC, objA), evaluator.py).  The modules here carry the reference's sub-module names and
state_dict keys, so `torch.load(f).state_dict()` of either side's file loads into the other side's class
(evaluator.py:329-330 relies on exactly that).  Additionally every file gets a `*.state_dict.pth` twin that loads with
`weights_only=True` (no pickled classes, no omegaconf)."""
import os

import numpy as np
import torch

FILES = ("actor", "critic", "actor_gnn", "critic_gnn", "actor_gru", "critic_gru", "actor_mean", "critic_mean")


def _parts(mappo):
    a, c = mappo.actor, mappo.critic
    return dict(actor=a, critic=c, actor_gnn=a.shared_net, critic_gnn=c.shared_net, actor_gru=a.GRU, critic_gru=c.GRU,
                actor_mean=a.Mean, critic_mean=c.Mean)


def _detached_copy(module):
    """A CPU deep copy whose parameters own their storage (the live ones are views into the flat optimizer arena)."""
    import copy
    m = copy.deepcopy(module).cpu()
    for p in m.parameters():
        p.data = p.data.clone()
        p.grad = None
    return m


def save_checkpoint(mappo, cwd, final=False, recorder=None):
    os.makedirs(cwd, exist_ok=True)
    suffix = "_final" if final else ""
    for name, mod in _parts(mappo).items():
        m = _detached_copy(mod)
        torch.save(m, os.path.join(cwd, f"{name}{suffix}.pth"))
        torch.save(m.state_dict(), os.path.join(cwd, f"{name}{suffix}.state_dict.pth"))
    if recorder is not None:
        np.save(os.path.join(cwd, "recorder.npy"), np.asarray(recorder))


def load_state(path):
    """state_dict from either a pickled module (reference or ours) or a state_dict file."""
    try:
        obj = torch.load(path, map_location="cpu", weights_only=True)
    except Exception:
        obj = torch.load(path, map_location="cpu", weights_only=False)
    return obj.state_dict() if hasattr(obj, "state_dict") else obj


def load_checkpoint(mappo, cwd, final=False, strict=True):
    """Loads whichever of the eight files exist (whole-network files first, then the per-part files override)."""
    suffix = "_final" if final else ""
    parts = _parts(mappo)
    loaded = []
    for name in FILES:
        for ext in (".state_dict.pth", ".pth"):
            path = os.path.join(cwd, f"{name}{suffix}{ext}")
            if os.path.exists(path):
                parts[name].load_state_dict(load_state(path), strict=strict)
                loaded.append(name)
                break
    return loaded
