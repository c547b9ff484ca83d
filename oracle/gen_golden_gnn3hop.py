"""TEST INFRASTRUCTURE — golden vectors of the second network family: `GnnExtractor`, `SharedActor`,
`SharedCritic` and `MAPPO.train` of obstacle_differ_3hop/mappo_parallel.py, produced by EXECUTING THE UNMODIFIED REFERENCE
classes on synthetic tensors of the reference's buffer layout (module/replay_buffer.py:21-30).  The env of that family is not
in the reference tree, and `MAPPO.__init__` loads ./model/*.pth pickles of an older layout, so the MAPPO object is created
without __init__ (`object.__new__`) and given the attributes `train` reads; `train` itself runs unmodified.
Re-run:  python -m oracle.gen_golden_gnn3hop
"""
import os
import sys
from types import SimpleNamespace as NS

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_bootstrap import load_reference, seed_all  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def make_args(n, emb):
    return NS(max_train_steps=int(2e8), lr=5e-4, gamma=0.99, lamda=0.95, epsilon=0.05, K_epochs=1, entropy_coef=0.05,
              use_grad_clip=True, use_lr_decay=True, use_adv_norm=True, use_value_clip=True, state_dim=8, num_layers=2,
              gnn_output_dim=emb, gnn_middle_dim=emb, rnn_hidden_dim=emb, n_hops=3, learner_device="cpu", worker_device="cpu",
              evaluator_device="cpu", use_reward_norm=False, use_spectral_norm=True, action_dim=9, pursuer_num=[n])


def synth_buffer(rng, B, T, N, O, emb):
    """Synthetic tensors with the structure run_episode builds (:472-500): agent rows [1, dp(4), de(4)*mask], obstacle rows
    [0, do(2), -phi, -v, 0,0,0,0], zero padding rows; adjacency = [p_p_adj | p_o_adj] 0/1."""
    J = N + O
    state = np.zeros((B, T, N, J, 9), np.float32)
    state[..., :N, 0] = 1.0
    state[..., :N, 1:5] = rng.normal(0, 3, (B, T, N, N, 4))
    mask = (rng.random((B, T, N, 1, 1)) < 0.4)
    state[..., :N, 5:9] = rng.normal(0, 3, (B, T, N, 1, 4)) * mask
    o_real = int(0.7 * O)
    state[..., N:N + o_real, 1:5] = rng.normal(0, 8, (B, T, N, o_real, 4))
    adj = np.zeros((B, T, N, J), np.float32)
    adj[..., :N] = rng.random((B, T, N, N)) < 0.6
    adj[..., np.arange(N), np.arange(N)] = 1.0
    adj[..., N:N + o_real] = rng.random((B, T, N, o_real)) < 0.15
    adj[:, :, 0, N:] = 0 if N > 2 else adj[:, :, 0, N:]          # an agent that sees no obstacle
    f = lambda *s: rng.normal(0, 0.5, s).astype(np.float32)     # noqa: E731
    active = np.ones((B, T, N), np.float32)
    active[0, T // 2:, 0] = 0.0
    return dict(state=state, adj=adj, actor_comm_embedding=f(B, T, N, 2 * emb), critic_comm_embedding=f(B, T, N, 2 * emb),
                v_n=f(B, T + 1, N), a_n=rng.integers(0, 9, (B, T, N)).astype(np.float32),
                a_logprob_n=(-2.2 + 0.1 * rng.normal(0, 1, (B, T, N))).astype(np.float32), r=f(B, T, N), active=active)


class _Big:
    def __init__(self, buf, T):
        self.buf, self.T = buf, T

    def get_training_data(self, num, device):
        return self.buf, self.T


def gen(m3, N, O, B, T, mb, emb, seed, compact=False):
    import torch
    seed_all(seed)
    args = make_args(N, emb)
    actor = m3.SharedActor(m3.GnnExtractor(9, emb, emb, 3, True), emb, 9, 2, emb, True)
    critic = m3.SharedCritic(m3.GnnExtractor(9, emb, emb, 3, True), emb, 1, 2, emb, True)
    rng = np.random.default_rng(seed)
    with torch.no_grad():                      # biases are zero-initialised: make them matter
        for p in list(actor.parameters()) + list(critic.parameters()):
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    buf_np = synth_buffer(rng, B, T, N, O, emb)
    buf = {k: torch.from_numpy(v) for k, v in buf_np.items()}
    w0 = {("actor." + k): v.clone().numpy() for k, v in actor.state_dict().items()}
    w0.update({("critic." + k): v.clone().numpy() for k, v in critic.state_dict().items()})
    fx = {"buf." + k: v for k, v in buf_np.items()}
    if compact:      # weights are regenerated from the seed by the test (same construction order); only checksums are stored
        fx.update({"wsum." + k: np.float64(np.abs(v.astype(np.float64)).sum()) for k, v in w0.items()})
    else:
        fx.update({"w." + k: v for k, v in w0.items()})
    # rollout-mode forward (mode 0) of sample (b=0, t=0): one launch of each network as run_episode does (:503-504)
    torch.set_grad_enabled(False)
    st, ad = buf["state"][0, 0], buf["adj"][0, 0]
    ha, hc = torch.zeros(2, N, emb), torch.zeros(2, N, emb)
    prob, ha2, comm_a = actor.forward(st, ad, ha, buf["actor_comm_embedding"][0, 0], mode=0)
    val, hc2, comm_c = critic.forward(st, ad, hc, buf["critic_comm_embedding"][0, 0], mode=0)
    fx.update({"roll.prob": prob.numpy(), "roll.ha": ha2.numpy(), "roll.comm_a": comm_a.numpy(), "roll.val": val.numpy(),
               "roll.hc": hc2.numpy(), "roll.comm_c": comm_c.numpy()})
    # the spectral-norm power iteration moved u/v during that forward: record the state train() starts from
    w1 = {("actor." + k): v.clone().numpy() for k, v in actor.state_dict().items()}
    w1.update({("critic." + k): v.clone().numpy() for k, v in critic.state_dict().items()})
    if not compact:
        fx.update({"w_train." + k: v for k, v in w1.items() if not np.array_equal(v, w0[k])})
    torch.set_grad_enabled(True)
    learner = object.__new__(m3.MAPPO)
    learner.__dict__.update(batch_size=B, mini_batch_size=mb, max_train_steps=args.max_train_steps, lr=args.lr, gamma=args.gamma,
                            lamda=args.lamda, epsilon=args.epsilon, K_epochs=1, entropy_coef=args.entropy_coef,
                            use_grad_clip=True, use_lr_decay=True, use_adv_norm=True, use_value_clip=True, num_layers=2,
                            gnn_output_dim=emb, rnn_input_dim=emb, rnn_hidden_dim=emb, n_hops=3, device=torch.device("cpu"),
                            actor=actor, critic=critic, args=args, minibuffer=None)
    learner.ac_parameters = (list(critic.shared_net.parameters()) + list(actor.shared_net.parameters()) + list(actor.GRU.parameters()) +
                             list(critic.GRU.parameters()) + list(critic.Mean.parameters()) + list(actor.Mean.parameters()))
    learner.ac_optimizer = torch.optim.Adam(learner.ac_parameters, lr=args.lr, eps=1e-5)
    objC, objA, ag, cg = learner.train(_Big(buf, T), total_steps=B * T)
    def put(key, g):
        g = np.asarray(g)
        if compact and g.size > 4096:       # large matrices: norm + a fixed strided sample (keeps the fixture small)
            fx["gnorm." + key] = np.float64(np.linalg.norm(g.astype(np.float64)))
            fx["gsample." + key] = g.reshape(-1)[::max(1, g.size // 2048)].copy()
        else:
            fx["grad." + key] = g

    for (name, _), g in zip(actor.named_parameters(), ag):
        if g is not None:
            put("actor." + name, g)
    for (name, _), g in zip(critic.named_parameters(), cg):
        if g is not None:
            put("critic." + name, g)
    fx["objC"], fx["objA"] = np.float64(objC), np.float64(objA)
    fx["meta"] = np.array([N, O, B, T, mb, emb, seed], np.int64)
    return fx


def main():
    load_reference()
    import obstacle_differ_3hop.mappo_parallel as m3
    for N, O, B, T, mb, emb, seed in ((3, 10, 4, 5, 2, 32, 41), (4, 12, 2, 4, 1, 128, 43)):
        fx = gen(m3, N, O, B, T, mb, emb, seed, compact=(emb == 128))
        path = os.path.join(GOLDEN_DIR, f"gnn3hop_n{N}_e{emb}.npz")
        np.savez_compressed(path, **fx)
        print(path, f"{os.path.getsize(path) / 1e6:.2f} MB objC={float(fx['objC']):.6f} objA={float(fx['objA']):.6f}")


if __name__ == "__main__":
    main()
