"""Config schema loader — accepts the reference's hydra files unchanged, without hydra/omegaconf.

* `load_config("config.yaml")` reads the monolithic file main.py uses (config.yaml:13-94); the hydra-only keys
  `defaults:` / `hydra:` / `main:` (config.yaml:1-11) are ignored.
* `load_conf_dir("conf/")` assembles the flat per-group dumps conf/{env,sensor,mao,attacker,defender,algo}.yaml
  (`mao.yaml` [sic] is the map group and uses `!!python/tuple`, which yaml.safe_load rejects).
* Every consumer only needs attribute access (`cfg.env.num_defender`), so any object tree works: OmegaConf,
  argparse.Namespace, or the `Cfg` namespaces built here.
"""
import os
from types import SimpleNamespace

import yaml


class Cfg(SimpleNamespace):
    """Attribute- and item-style access, like an OmegaConf node."""

    def __getitem__(self, k):
        return getattr(self, k)

    def __contains__(self, k):
        return hasattr(self, k)

    def get(self, k, default=None):
        return getattr(self, k, default)

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, Cfg) else v) for k, v in vars(self).items()}


class _Loader(yaml.SafeLoader):
    pass


_Loader.add_constructor("tag:yaml.org,2002:python/tuple", lambda loader, node: list(loader.construct_sequence(node)))


def _to_cfg(obj):
    if isinstance(obj, dict):
        return Cfg(**{str(k): _to_cfg(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return [_to_cfg(v) for v in obj]
    return obj


def _pair(v):
    """map.center / map.map_size appear as lists (config.yaml:31-32), tuples (conf/mao.yaml) or {x,y} dicts
    (outputs/*/.hydra/config.yaml)."""
    if isinstance(v, Cfg):
        return [v.x, v.y]
    if isinstance(v, dict):
        return [v["x"], v["y"]]
    return list(v)


_GROUPS = ("env", "sensor", "map", "attacker", "defender", "algo")


def _normalise(cfg):
    m = cfg.map
    if hasattr(m, "x_dim") and not hasattr(m, "map_size"):
        m.map_size = [m.x_dim, m.y_dim]
    m.map_size = _pair(m.map_size)
    m.center = _pair(m.center)
    if not hasattr(m, "num_max_obstacle") and hasattr(m, "max_num_obstacle"):   # conf/mao.yaml spelling
        m.num_max_obstacle = m.max_num_obstacle
    e = cfg.env
    # conf/env.yaml has no state_dim/action_dim/difficulty; conf/algo.yaml carries the first two
    for k, default in (("state_dim", 4), ("action_dim", 9)):
        if not hasattr(e, k):
            setattr(e, k, getattr(cfg.algo, k, default))
    if not hasattr(e, "difficulty"):
        e.difficulty = 10
    if not hasattr(cfg.attacker, "extend_dis"):
        cfg.attacker.extend_dis = 1
    a = cfg.algo
    if not hasattr(a, "epochs") and hasattr(a, "K_epochs"):
        a.epochs = a.K_epochs
    return cfg


def load_config(path):
    with open(path) as f:
        raw = yaml.load(f, Loader=_Loader)
    return _normalise(_to_cfg({g: raw[g] for g in _GROUPS if g in raw}))


def load_conf_dir(path):
    names = {"env": "env", "sensor": "sensor", "map": "mao", "attacker": "attacker", "defender": "defender",
             "algo": "algo"}
    raw = {}
    for group, fname in names.items():
        fp = os.path.join(path, fname + ".yaml")
        if group == "map" and not os.path.exists(fp):
            fp = os.path.join(path, "map.yaml")
        with open(fp) as f:
            raw[group] = yaml.load(f, Loader=_Loader)
    return _normalise(_to_cfg(raw))


def default_config(**over):
    """config.yaml:13-94 values (the configuration main.py actually runs), for code that has no YAML at hand."""
    cfg = _to_cfg(dict(
        env=dict(name="Pursuit_Env", state_dim=4, action_dim=9, attacker_class="Evader", defender_class="Pursuer",
                 max_steps=150, num_attacker=1, num_defender=15, num_target=1, step_size=0.1, difficulty=10),
        sensor=dict(num_beams=36, radius=8),
        map=dict(center=[30, 25], map_size=[60, 55], num_obstacle_block=5, resolution=1, variance=10,
                 num_max_obstacle=176),
        attacker=dict(DOF=2, collision_radius=0.5, comm_range=16, sen_range=8, step_size=0.1, tau=0.2, vmax=4,
                      extend_dis=1),
        defender=dict(DOF=2, collision_radius=0.5, comm_range=16, sen_range=8, step_size=0.1, tau=0.2, vmax=2),
        algo=dict(pretrain_model_cwd="./pretrain/experiment/pretrain_model_15", learner_device="cuda",
                  worker_device="cuda", evaluator_device="cuda", max_train_steps=20000000, lr=0.0005, gamma=0.99,
                  lamda=0.95, epsilon=0.05, epochs=1, entropy_coef=0.05, save_cwd="./model", sample_epi_num=1,
                  use_adv_norm=True, use_agent_specific=True, use_grad_clip=True, use_lr_decay=True,
                  use_orthogonal_init=True, use_reward_norm=True, use_spectral_norm=True, use_value_clip=True,
                  set_adam_eps=True, mlp_hidden_dim=128, rnn_hidden_dim=128, embedding_dim=128, num_layers=2,
                  semantic_level_aggregator="mean", vertex_level_aggregator="mean", fcra_aggregator="mean", depth=1,
                  num_relation=3)))
    for dotted, v in over.items():
        group, key = dotted.split("__")
        setattr(getattr(cfg, group), key, v)
    return cfg


def env_params_dict(cfg):
    """cfg -> the scalar fields of marl_env_params."""
    W, H = _pair(cfg.map.map_size)
    return dict(
        W=int(W), H=int(H), N=int(cfg.env.num_defender), O=int(cfg.map.num_max_obstacle),
        max_steps=int(cfg.env.max_steps), difficulty=int(getattr(cfg.env, "difficulty", 10)),
        sensor_beams=int(cfg.sensor.num_beams), sensor_radius=int(cfg.sensor.radius),
        e_extend_dis=int(getattr(cfg.attacker, "extend_dis", 1)), e_sen_range=int(cfg.attacker.sen_range),
        d_step=float(cfg.defender.step_size), d_tau=float(cfg.defender.tau), d_vmax=float(cfg.defender.vmax),
        d_collision_radius=float(cfg.defender.collision_radius), d_comm_range=float(cfg.defender.comm_range),
        d_sen_range=float(cfg.defender.sen_range), e_step=float(cfg.attacker.step_size), e_tau=float(cfg.attacker.tau),
        e_vmax=float(cfg.attacker.vmax), e_collision_radius=float(cfg.attacker.collision_radius),
        resolution=float(cfg.map.resolution))
