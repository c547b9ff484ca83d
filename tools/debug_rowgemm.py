import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_multi_agent_reinforcement_learning_b200 import policy_ops as ops
g = torch.Generator(device="cuda").manual_seed(1)
M, E = 2050, 128
mk = lambda *s: torch.randn(*s, device="cuda", generator=g)
x, x2, W, b, add = mk(M, E), mk(M, E), mk(E, 2 * E) * 0.1, mk(E) * 0.1, mk(M, E)
dout = mk(M, E)
for trial in range(3):
    a = [t.clone().requires_grad_(True) for t in (x, x2, W, b, add)]
    out = ops.linear(a[0], a[2], a[3], True, x2=a[1], add=a[4])
    r = [t.clone().requires_grad_(True) for t in (x, x2, W, b, add)]
    ref = torch.relu(torch.addmm(r[3], torch.cat([r[0], r[1]], 1), r[2].t()) + r[4])
    out.backward(dout)
    ref.backward(dout)
    for name, u, v in zip(("x", "x2", "W", "b", "add"), a, r):
        d = (u.grad - v.grad).abs()
        bad = (d > 1e-4 * max(1.0, float(v.grad.abs().max()))).nonzero()
        print(trial, name, "max diff", float(d.max()), "n bad", len(bad), "rows", sorted(set(bad[:, 0].tolist()))[:10] if len(bad) else [],
              "cols", (int(bad[:, 1].min()), int(bad[:, 1].max())) if len(bad) and bad.shape[1] > 1 else None)
# direct: dx via transposed rowgemm vs torch, many repeats
dy = mk(M, E)
for trial in range(5):
    got = ops._gemm_tc(dy, None, W[:, :E], None, None, False, transposed=True)
    ref = dy @ W[:, :E]
    d = (got - ref).abs()
    bad = (d > 1e-4).nonzero()
    print("direct", trial, float(d.max()), len(bad), sorted(set(bad[:, 0].tolist()))[:10])
