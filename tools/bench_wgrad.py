import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_multi_agent_reinforcement_learning_b200 import policy_ops as ops
R = 492000
for N, K in ((128, 128), (384, 128)):
    dy, x = torch.randn(R, N, device="cuda"), torch.randn(R, K, device="cuda")
    db = torch.empty(N, device="cuda")
    for _ in range(3):
        ops.wgrad(dy, x, dbias=db)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ops.wgrad(dy, x, dbias=db)
    e.record(); e.synchronize()
    ms = s.elapsed_time(e) / 10
    print(f"wgrad R={R} N={N} K={K}: {ms:.3f} ms  {2 * R * N * K / ms / 1e9:.1f} TFLOP/s  {(R * (N + K) * 4) / ms / 1e6:.0f} GB/s")
