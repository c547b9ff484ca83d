"""B200-native MAPPO rollout-and-update hot path (drop-in behind the plugin surface of
Desperodoo/distributed_multi_agent_reinforcement_learning).  See DESIGN.md / INTEGRATION.md.

Importing the package does not touch CUDA; every compute entry point goes through the C-ABI library
`libmarl_b200.so` (include/marl_b200.h) and raises if it is missing — there is no CPU fallback."""
from . import _lib  # noqa: F401
from .config import Cfg, default_config, env_params_dict, load_conf_dir, load_config  # noqa: F401

__all__ = ["Cfg", "default_config", "env_params_dict", "load_conf_dir", "load_config", "Pursuit_Env",
           "BatchedPursuitEnv", "RolloutArena"]


def __getattr__(name):   # torch is imported lazily so that config / ABI checks stay light
    if name in ("Pursuit_Env", "BatchedPursuitEnv", "RolloutArena"):
        from . import pursuit_env
        return getattr(pursuit_env, name)
    raise AttributeError(name)
