"""Stock-PyTorch-on-the-same-B200 baseline for the MAPPO update ("the real bar for kernel 5").

What the reference does when its learner runs with `learner_device='cuda'` (DHGN/mappo_parallel.py:676-715): eager
fp32 `nn.Linear` / `nn.GRU` (cuDNN) modules, autograd, `clip_grad_norm_`, Adam.  /root/reference does not exist on
the GPU box, so the architecture is restated here with stock torch.nn modules only (no kernel of this repo, no
oracle import): DHGN encoder (three relations, shared AGG_vertex_0, semantic layer), FCRA depth D, 2-layer GRU,
actor / critic heads, PPO clip loss with value clip, masked means.

Bounded sample: the reference's minibatch at the bench shape is 410 envs x 150 steps x 8 agents; its relation-2
activation [mb, T, N, O, 128] alone would be 44 GB there (and autograd keeps several of them), so the baseline runs
minibatches of --mb envs (default 32: 3.5 GB per such tensor) and reports samples / s, which does not depend on mb
once the GEMMs have millions of rows.  Prints one JSON line.

    python tools/bench_train_torch_cuda.py [--mb 32] [--iters 5] [--depth 1] [--agents 8] [--obstacles 176]
"""
import argparse
import json

import torch
import torch.nn as nn
import torch.nn.functional as F

E = 128


class Encoder(nn.Module):
    def __init__(self, depth):
        super().__init__()
        self.msg = nn.ModuleList([nn.Linear(8, E), nn.Linear(4, E), nn.Linear(4, E)])
        self.agg_vertex = nn.Linear(E, E)
        self.agg_fcra = nn.ModuleList([nn.Linear(E, E) for _ in range(depth)])
        self.fcra = nn.ModuleList([nn.Linear(2 * E, E) for _ in range(depth)])
        self.semantic = nn.Linear(4 + 3 * E, E)

    def mean_op(self, lin, adj, msg):
        return torch.relu(lin(torch.matmul(F.normalize(adj, p=1, dim=-1), msg)))

    def forward(self, p, e, o, p_adj, e_adj, o_adj, hist):
        rel_pp = p.unsqueeze(-2) - p.unsqueeze(-3)
        rel_pe = p.unsqueeze(-2) - e.unsqueeze(-3)
        rel_po = p.unsqueeze(-2) - o.unsqueeze(-3)
        a0 = torch.cat([rel_pp, rel_pe.expand(*rel_pp.shape[:-1], 4)], dim=-1)
        embs = []
        for r, (attr, adj) in enumerate(((a0, p_adj), (rel_pe, e_adj), (rel_po, o_adj))):
            embs.append(self.mean_op(self.agg_vertex, adj.unsqueeze(-2), torch.relu(self.msg[r](attr))))
        h = self.semantic(torch.cat([p.unsqueeze(-2)] + embs, dim=-1)).squeeze(-2)
        for k, hk in enumerate(hist):
            m = self.mean_op(self.agg_fcra[k], p_adj, hk)
            h = torch.relu(self.fcra[k](torch.cat([m, h], dim=-1)))
        return h


class Net(nn.Module):
    def __init__(self, encoder, out_dim, spectral):
        super().__init__()
        self.shared_net = encoder                      # one encoder instance shared by actor and critic (:582-616)
        self.GRU = nn.GRU(E, E, num_layers=2)
        head = nn.Linear(E, out_dim)
        self.Mean = nn.utils.spectral_norm(head) if spectral else head

    def forward(self, p, e, o, p_adj, e_adj, o_adj, hist, h=None):
        mb, T, N = p.shape[:3]
        x = self.shared_net(p, e, o, p_adj, e_adj, o_adj, hist)
        x = x.permute(1, 0, 2, 3).reshape(T, mb * N, E)
        y, h = self.GRU(x, h)
        out = self.Mean(y.reshape(T, mb, N, E).permute(1, 0, 2, 3))
        return out if h is None or not self.return_hidden else (out, h)

    return_hidden = False


def rollout_step_baseline(args, dev, actor, critic):
    """One network-in-the-loop policy step of the rollout (MAPPO.run_episode :760-790, batched over --envs envs the way
    this repo's fused kernel is): both encoders + GRUs + heads, Categorical sample + log-prob, no autograd."""
    B, N, O, D = args.envs, args.agents, args.obstacles, args.depth
    g = torch.Generator(device=dev).manual_seed(2)
    rnd = lambda *s: torch.rand(*s, device=dev, generator=g)
    p = torch.cat([rnd(B, 1, N, 2) * 55, rnd(B, 1, N, 2) * 4 - 2], -1)
    e = torch.cat([rnd(B, 1, 1, 2) * 55, rnd(B, 1, 1, 2) * 8 - 4], -1)
    o = torch.cat([torch.floor(rnd(B, 1, O, 2) * 55), torch.zeros(B, 1, O, 2, device=dev)], -1)
    o[:, :, 100:] = 0
    p_adj = ((rnd(B, 1, N, N) < 0.5) | torch.eye(N, device=dev, dtype=torch.bool)).float()
    e_adj = (rnd(B, 1, N, 1) < 0.3).float()
    o_adj = (rnd(B, 1, N, O) < 0.08).float()
    o_adj[..., 100:] = 0
    ones = [torch.ones_like(x) for x in (p_adj, e_adj, o_adj)]
    hist_a, hist_c = [rnd(B, 1, N, E) for _ in range(D)], [rnd(B, 1, N, E) for _ in range(D)]
    state = {"ha": torch.zeros(2, B * N, E, device=dev), "hc": torch.zeros(2, B * N, E, device=dev)}
    actor.return_hidden = critic.return_hidden = True

    @torch.no_grad()
    def step():
        logits, state["ha"] = actor(p, e, o, p_adj, e_adj, o_adj, hist_a, state["ha"])
        values, state["hc"] = critic(p, e, o, *ones, hist_c, state["hc"])
        dist = torch.distributions.Categorical(logits=logits)
        a = dist.sample()
        return a, dist.log_prob(a), values

    for _ in range(3):
        step()
    n = max(3, args.iters * 4)
    if dev.type == "cuda":
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(n):
            step()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / n
    else:
        import time

        t = time.perf_counter()
        for _ in range(n):
            step()
        ms = (time.perf_counter() - t) * 1e3 / n
    actor.return_hidden = critic.return_hidden = False
    print(json.dumps({"impl": f"stock torch.nn eager on {dev} (fp32, cuDNN GRU, no autograd)",
                      "metric": "policy_step_agent_env_steps_per_sec", "value": B * N / (ms * 1e-3),
                      "unit": "agent-env-steps/s (policy step only, env not included)", "ms_per_policy_step": ms,
                      "sample": f"{n} steps of {B} envs x {N} agents (O={O}, depth {D})"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=32)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--depth", type=int, default=1)
    ap.add_argument("--agents", type=int, default=8)
    ap.add_argument("--obstacles", type=int, default=176)
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--envs", type=int, default=4096, help="envs of the rollout-step baseline")
    ap.add_argument("--device", default="cuda:0", help="cpu only for a syntax check of this script")
    args = ap.parse_args()
    dev = torch.device(args.device)
    cuda = dev.type == "cuda"
    torch.manual_seed(0)
    mb, T, N, O, D = args.mb, args.steps, args.agents, args.obstacles, args.depth
    enc = Encoder(D)
    actor, critic = Net(enc, 9, False).to(dev), Net(enc, 1, True).to(dev)
    params = list({id(q): q for q in list(actor.parameters()) + list(critic.parameters())}.values())
    opt = torch.optim.Adam(params, lr=5e-4, eps=1e-5)

    g = torch.Generator(device=dev).manual_seed(1)
    rnd = lambda *s: torch.rand(*s, device=dev, generator=g)
    p = torch.cat([rnd(mb, T, N, 2) * 55, rnd(mb, T, N, 2) * 4 - 2], -1)
    e = torch.cat([rnd(mb, T, 1, 2) * 55, rnd(mb, T, 1, 2) * 8 - 4], -1)
    o = torch.cat([torch.floor(rnd(mb, T, O, 2) * 55), torch.zeros(mb, T, O, 2, device=dev)], -1)
    o[:, :, 100:] = 0                                    # 100 boundary cells, zero-padded to O (replay_buffer.py:33,52)
    p_adj = ((rnd(mb, T, N, N) < 0.5) | torch.eye(N, device=dev, dtype=torch.bool)).float()
    e_adj = (rnd(mb, T, N, 1) < 0.3).float()
    o_adj = (rnd(mb, T, N, O) < 0.08).float()
    o_adj[..., 100:] = 0
    hist_a, hist_c = rnd(mb, T + D, N, E), rnd(mb, T + D, N, E)
    a_n = torch.randint(0, 9, (mb, T, N), device=dev, generator=g)
    logp_old = -2.2 + 0.1 * rnd(mb, T, N)
    adv, v_target, v_old = rnd(mb, T, N) - 0.5, rnd(mb, T, N), rnd(mb, T, N)
    active = torch.ones(mb, T, N, device=dev)
    ones = [torch.ones_like(x) for x in (p_adj, e_adj, o_adj)]

    def minibatch():
        hist = lambda hb: [hb[:, D - 1 - k: D - 1 - k + T] for k in range(D)]
        logits = actor(p, e, o, p_adj, e_adj, o_adj, hist(hist_a))
        values = critic(p, e, o, *ones, hist(hist_c)).squeeze(-1)
        logp_all = torch.log_softmax(logits, -1)
        logp = logp_all.gather(-1, a_n.unsqueeze(-1)).squeeze(-1)
        entropy = -(logp_all.exp() * logp_all).sum(-1)
        ratios = torch.exp(logp - logp_old)
        s1, s2 = ratios * adv, torch.clamp(ratios, 0.95, 1.05) * adv
        la = ((-torch.min(s1, s2) - 0.05 * entropy) * active).sum() / active.sum()
        ec = torch.clamp(values - v_old, -0.05, 0.05) + v_old - v_target
        lc = (torch.max(ec ** 2, (values - v_target) ** 2) * active).sum() / active.sum()
        (la + lc).backward()
        torch.nn.utils.clip_grad_norm_(params, 5.0)

    def epoch(n):
        opt.zero_grad()
        for _ in range(n):
            minibatch()
        opt.step()

    epoch(1)                                             # warm-up (cuDNN plans, allocator)
    if cuda:
        torch.cuda.synchronize()
        peak_gb = torch.cuda.max_memory_allocated() / 2 ** 30
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        epoch(args.iters)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1)
    else:
        import time

        peak_gb, t = None, time.perf_counter()
        epoch(args.iters)
        ms = (time.perf_counter() - t) * 1e3
    rollout_step_baseline(args, dev, actor, critic)
    print(json.dumps({"impl": f"stock torch.nn eager on {dev} (fp32, cuDNN GRU, autograd, clip_grad_norm_, Adam)",
                      "metric": "mappo_train_samples_per_sec", "value": args.iters * mb * T * N / (ms * 1e-3),
                      "unit": "samples/s", "ms_per_minibatch": ms / args.iters,
                      "sample": f"{args.iters} minibatches of {mb} envs x {T} steps x {N} agents (O={O}, depth {D}) + 1 Adam step",
                      "peak_mem_GiB": peak_gb, "matmul_tf32": torch.backends.cuda.matmul.allow_tf32,
                      "cudnn_tf32": torch.backends.cudnn.allow_tf32}))


if __name__ == "__main__":
    main()
