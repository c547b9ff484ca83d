"""Timing aid: one nn.GRU layer over a sequence through _GRULayer (persistent sequence kernels vs the per-step path)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ab  # noqa: E402,F401
from distributed_multi_agent_reinforcement_learning_b200 import policy_ops as ops  # noqa: E402

E, T = 128, 150
for R in (3280, 32768):
    g = torch.Generator(device="cuda").manual_seed(0)
    mk = lambda *s: (torch.randn(*s, device="cuda", generator=g) * 0.2)   # noqa: E731
    w = [mk(3 * E, E).requires_grad_(True), mk(3 * E, E).requires_grad_(True), mk(3 * E).requires_grad_(True), mk(3 * E).requires_grad_(True)]
    x, h0 = mk(T, R, E).requires_grad_(True), mk(R, E)
    dout = mk(T, R, E)
    for seq in (True, False):
        if not seq and R > 4000:
            continue
        ops.USE_GRU_SEQ = seq
        for it in range(3):
            torch.cuda.synchronize()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            out = ops._GRULayer.apply(x, h0, *w)
            e[1].record()
            out.backward(dout)
            e[2].record()
            torch.cuda.synchronize()
        print(f"R={R} seq={seq}: fwd {e[0].elapsed_time(e[1]):.2f} ms  bwd {e[1].elapsed_time(e[2]):.2f} ms (incl. projections / weight-grad GEMMs)")

from torch.profiler import profile, ProfilerActivity  # noqa: E402
ops.USE_GRU_SEQ = True
R = 3280
x, h0 = (torch.randn(T, R, E, device="cuda") * 0.2).requires_grad_(True), torch.randn(R, E, device="cuda") * 0.2
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    out = ops._GRULayer.apply(x, h0, *w)
    out.backward(torch.randn_like(out))
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=60))
