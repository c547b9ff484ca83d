"""Micro-benchmark of the training-side row-tile GEMM (csrc/rowgemm_tf32x3.cu) at the bench's minibatch shape.
MARL_AB_LIB=<alternative libmarl_b200.so> times another build (tools/ab.py)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ab  # noqa: E402,F401
from distributed_multi_agent_reinforcement_learning_b200 import policy_ops  # noqa: E402

M = int(os.environ.get("M", 410 * 150 * 8))
dev = torch.device("cuda:0")
torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for N, K1, K2, relu in ((128, 128, 0, True), (128, 128, 128, True), (128, 384, 0, False), (384, 128, 0, False)):
    x = torch.randn(M, K1, device=dev)
    x2 = torch.randn(M, K2, device=dev) if K2 else None
    W = torch.randn(N, K1 + K2, device=dev) * 0.05
    b = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev)
    with policy_ops.pack_scope():
        for _ in range(3):
            policy_ops._gemm_tc(x, x2, W, b, None, relu, out=out)
        ts = []
        for _ in range(10):
            flush.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            policy_ops._gemm_tc(x, x2, W, b, None, relu, out=out)
            e.record()
            e.synchronize()
            ts.append(s.elapsed_time(e))
    ref = torch.nn.functional.linear(torch.cat([x, x2], 1) if K2 else x, W, b)
    ref = torch.relu(ref) if relu else ref
    us = 1e3 * sorted(ts)[len(ts) // 2]
    byts = 4 * M * (K1 + K2 + N)
    print(json.dumps({"M": M, "N": N, "K": K1 + K2, "us": us, "GBps": byts / us / 1e3, "TFLOPs_fp32_equiv": 2 * M * N * (K1 + K2) / us / 1e6,
                      "max_abs_err_vs_torch": float((out - ref).abs().max())}))
