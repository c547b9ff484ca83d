"""A/B aid for the tools: MARL_AB_LIB=<path to an alternative libmarl_b200.so> makes the package load that build."""
import os

from distributed_multi_agent_reinforcement_learning_b200 import _lib

if os.environ.get("MARL_AB_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["MARL_AB_LIB"])
