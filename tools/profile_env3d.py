"""Profiling aid: a few fused launches of the 3-D particle env kernel at BASELINE config 4's per-GPU shape (for ncu)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_multi_agent_reinforcement_learning_b200.particle_env import BatchedParticleEnv, Env3dArena  # noqa: E402

B, N, T, K = 8192, 32, 200, 20
eng = BatchedParticleEnv(B, N)
eng.reset(seed=4)
arena = Env3dArena(N, B, T, eng.device)
for t0 in range(0, 3 * K, K):
    eng.rollout(arena, K, t0, None, None, seed=0xB200)
torch.cuda.synchronize()
print("done")
