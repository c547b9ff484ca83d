import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributed_multi_agent_reinforcement_learning_b200 import policy_ops as ops
E, T, R = 128, 30, 3280
w = [(torch.randn(3 * E, E, device="cuda") * 0.2).requires_grad_(True), (torch.randn(3 * E, E, device="cuda") * 0.2).requires_grad_(True),
     (torch.randn(3 * E, device="cuda") * 0.2).requires_grad_(True), (torch.randn(3 * E, device="cuda") * 0.2).requires_grad_(True)]
x, h0 = (torch.randn(T, R, E, device="cuda") * 0.2).requires_grad_(True), torch.randn(R, E, device="cuda") * 0.2
for _ in range(2):
    out = ops._GRULayer.apply(x, h0, *w)
    out.backward(torch.randn_like(out))
torch.cuda.synchronize()
print("done")
