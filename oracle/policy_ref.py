"""TEST INFRASTRUCTURE — plain-PyTorch fp32 restatement of the reference's network math (DHGN encoder + FCRA + GRU +
heads, PPO loss), functional and loop-free, used ONLY as the checker for the CUDA policy kernels.

Pinned: tests/test_policy_ref_golden.py checks it against tests/golden/algo_*.npz, which were produced by executing
the unmodified reference `DHGN/mappo_parallel.py` (oracle/gen_golden_algo.py): per-minibatch log-probs, entropies,
values, both losses and every accumulated gradient.  Each function cites the reference lines it follows.
`w` is a dict of tensors keyed like the reference state_dicts with an "actor." / "critic." prefix.
"""
import torch
import torch.nn.functional as F


def _lin(w, name, x):
    return F.linear(x, w[name + ".weight"], w[name + ".bias"])


def _mean_op(w, name, adj, msg):
    """DHGN.mean_operator (DHGN/mappo_parallel.py:336-348): ReLU(Linear(L1norm(adj) @ msg)); adj [...,N,1,K]."""
    return torch.relu(_lin(w, name, torch.matmul(F.normalize(adj, p=1, dim=-1), msg)))


def dhgn_encoder(w, net, p, e, o, p_adj, e_adj, o_adj):
    """DHGN.encoder (:241-304).  p [...,N,4], e [...,1,4], o [...,O,4]; adjacency [...,N,N], [...,N,1], [...,N,O]."""
    pre = f"{net}.shared_net."
    rel_pp = p.unsqueeze(-2) - p.unsqueeze(-3)                       # coordinate(a1, a1): [.., N, N, 4]
    rel_pe = p.unsqueeze(-2) - e.unsqueeze(-3)                       # [.., N, 1, 4]
    rel_po = p.unsqueeze(-2) - o.unsqueeze(-3)                       # [.., N, O, 4]
    a0 = torch.cat([rel_pp, rel_pe.expand(*rel_pp.shape[:-1], 4)], dim=-1)
    embs = []
    for r, (attr, adj) in enumerate(((a0, p_adj), (rel_pe, e_adj), (rel_po, o_adj))):
        msg = torch.relu(_lin(w, pre + f"MSG_layers.{r}", attr))     # DHGN.message (:323-334)
        embs.append(_mean_op(w, pre + "AGG_layers.AGG_vertex_0", adj.unsqueeze(-2), msg))   # same AGG layer for all r (:278)
    v = torch.cat([p.unsqueeze(-2)] + embs, dim=-1)                  # [.., N, 1, 4+3E]
    return _lin(w, pre + "semantic_layer", v).squeeze(-2)            # no activation (:303)


def dhgn_fcra(w, net, h0, hist, adj):
    """DHGN.fcra (:204-233).  hist[k] [...,N,E] is the embedding k+1 steps back; adj [...,N,N]."""
    pre = f"{net}.shared_net."
    h = h0
    for k, hk in enumerate(hist):
        m = _mean_op(w, pre + f"AGG_layers.AGG_fcra_{k}", adj, hk)
        h = torch.relu(_lin(w, pre + f"FCRA_layers.{k}", torch.cat([m, h], dim=-1)))
    return h


def gru_layer(w, prefix, layer, x, h0):
    """nn.GRU one layer, gate order (r, z, n).  x [T,R,E], h0 [R,E] -> (out [T,R,E], hT)."""
    wi, wh = w[f"{prefix}.weight_ih_l{layer}"], w[f"{prefix}.weight_hh_l{layer}"]
    bi, bh = w[f"{prefix}.bias_ih_l{layer}"], w[f"{prefix}.bias_hh_l{layer}"]
    E = h0.shape[-1]
    gi_all = F.linear(x, wi, bi)
    h, outs = h0, []
    for t in range(x.shape[0]):
        gi, gh = gi_all[t], F.linear(h, wh, bh)
        r = torch.sigmoid(gi[:, :E] + gh[:, :E])
        z = torch.sigmoid(gi[:, E:2 * E] + gh[:, E:2 * E])
        n = torch.tanh(gi[:, 2 * E:] + r * gh[:, 2 * E:])
        h = (1 - z) * n + z * h
        outs.append(h)
    return torch.stack(outs), h


def gru(w, prefix, x, h0, num_layers=2):
    hs = []
    for layer in range(num_layers):
        x, hT = gru_layer(w, prefix, layer, x, h0[layer])
        hs.append(hT)
    return x, torch.stack(hs)


def critic_head_weight(w):
    """torch.nn.utils.spectral_norm on the [1,E] critic head (:485): one power iteration is exact for a one-row matrix,
    sigma = u^T W v with u, v constants of the differentiation."""
    W = w["critic.Mean.weight_orig"]
    with torch.no_grad():
        u = w["critic.Mean.weight_u"]
        v = F.normalize(torch.mv(W.t(), u), dim=0, eps=1e-12)
        u = F.normalize(torch.mv(W, v), dim=0, eps=1e-12)
    sigma = torch.dot(u, torch.mv(W, v))
    return W / sigma, u, v


def train_forward(w, batch, depth, num_layers=2):
    """Training-mode forward of both networks on a minibatch dict of [mb,T,...] tensors (reference buffer layout,
    :676-690).  Returns (logp [mb,T,N], entropy, values)."""
    p, e, o = batch["p_state"], batch["e_state"], batch["o_state"]
    mb, T, N = p.shape[:3]
    E = w["actor.shared_net.semantic_layer.weight"].shape[0]
    out = {}
    for net in ("actor", "critic"):
        if net == "actor":
            p_adj, e_adj, o_adj = batch["p_adj"], batch["e_adj"], batch["o_adj"]
        else:   # AttributeDataset(is_critic=True): all-ones adjacency over ALL padded obstacle slots (:65)
            p_adj, e_adj, o_adj = (torch.ones_like(batch[k]) for k in ("p_adj", "e_adj", "o_adj"))
        h0 = dhgn_encoder(w, net, p, e, o, p_adj, e_adj, o_adj)
        hb = batch[f"{net}_historical_embedding"]                                  # [mb, T+D, N, E]
        hist = [hb[:, depth - 1 - k: depth - 1 - k + T] for k in range(depth)]     # EmbeddingDataset2 (:106-113)
        emb = dhgn_fcra(w, net, h0, hist, p_adj)
        x = emb.permute(1, 0, 2, 3).reshape(T, mb * N, E)
        feat, _ = gru(w, f"{net}.GRU", x, torch.zeros(num_layers, mb * N, E, dtype=x.dtype, device=x.device), num_layers)
        feat = feat.reshape(T, mb, N, E).permute(1, 0, 2, 3)
        if net == "actor":
            logits = F.linear(feat, w["actor.Mean.weight"], w["actor.Mean.bias"])
            logp_all = torch.log_softmax(logits, dim=-1)
            a = batch["a_n"].long().unsqueeze(-1)
            out["logp"] = logp_all.gather(-1, a).squeeze(-1)
            out["entropy"] = -(logp_all.exp() * logp_all).sum(-1)
        else:
            Weff, _, _ = critic_head_weight(w)
            out["values"] = F.linear(feat, Weff, w["critic.Mean.bias"]).squeeze(-1)
    return out["logp"], out["entropy"], out["values"]


def ppo_losses(logp, entropy, values, batch, adv, v_target, epsilon=0.05, entropy_coef=0.05, use_value_clip=True):
    """:692-706."""
    active = batch["active"]
    ratios = torch.exp(logp - batch["a_logprob_n"])
    surr1, surr2 = ratios * adv, torch.clamp(ratios, 1 - epsilon, 1 + epsilon) * adv
    actor_loss = ((-torch.min(surr1, surr2) - entropy_coef * entropy) * active).sum() / active.sum()
    if use_value_clip:
        v_old = batch["v_n"][:, :-1]
        err_clip = torch.clamp(values - v_old, -epsilon, epsilon) + v_old - v_target
        critic_loss = torch.max(err_clip ** 2, (values - v_target) ** 2)
    else:
        critic_loss = (values - v_target) ** 2
    critic_loss = (critic_loss * active).sum() / active.sum()
    return actor_loss, critic_loss


def gae(batch, gamma=0.99, lamda=0.95, T=None):
    """:643-658."""
    v, r, active = batch["v_n"], batch["r"], batch["active"]
    deltas = (r + gamma * v[:, 1:] - v[:, :-1]) * active
    adv, g = [], 0
    for t in reversed(range(r.shape[1])):
        g = deltas[:, t] + gamma * lamda * g
        adv.insert(0, g)
    adv = torch.stack(adv, dim=1)
    v_target = adv + v[:, :-1]
    adv = (adv - adv.mean()) / (adv.std() + 1e-5) * active
    return adv, v_target


def rollout_step(w, obs, hist, ha, hc, depth, num_layers=2, generator=None):
    """One step of MAPPO.run_episode's network part (:773-784) for a batch of envs, dense tensors like the reference:
    obs = dict(p [B,N,4], e [B,1,4], o [B,O,4], p_adj [B,N,N], e_adj [B,N,1], o_adj [B,N,O], o_real [B,O] 0/1 mask of the
    O_b real boundary cells); hist = newest-first list of [B,N,E] (the aliased list: C(t-1), A(t-1), C(t-2), ...).
    Returns (action [B,N], logp, value, emb_a, emb_c, ha, hc)."""
    p, e, o = obs["p"], obs["e"], obs["o"]
    B, N = p.shape[:2]
    E = w["actor.shared_net.semantic_layer.weight"].shape[0]
    out = {}
    for net, h in (("actor", ha), ("critic", hc)):
        if net == "actor":
            p_adj, e_adj, o_adj = obs["p_adj"], obs["e_adj"], obs["o_adj"]
        else:   # all ones over the O_b REAL cells only: o_ten is unpadded during rollout (:756-757,774)
            p_adj, e_adj = torch.ones_like(obs["p_adj"]), torch.ones_like(obs["e_adj"])
            o_adj = obs["o_real"].unsqueeze(1).expand(B, N, -1).contiguous()
        h0 = dhgn_encoder(w, net, p, e, o, p_adj, e_adj, o_adj)
        emb = dhgn_fcra(w, net, h0, hist, p_adj)
        feat, hn = gru(w, f"{net}.GRU", emb.reshape(1, B * N, E), h, num_layers)
        out[net] = (emb, feat[0], hn)
    logits = F.linear(out["actor"][1], w["actor.Mean.weight"], w["actor.Mean.bias"])
    dist = torch.distributions.Categorical(probs=torch.softmax(logits, -1))
    a = dist.sample() if generator is None else torch.multinomial(dist.probs, 1, generator=generator)[:, 0]
    Weff, _, _ = critic_head_weight(w)
    val = F.linear(out["critic"][1], Weff, w["critic.Mean.bias"])[:, 0]
    return (a.view(B, N), dist.log_prob(a).view(B, N), val.view(B, N), out["actor"][0], out["critic"][0],
            out["actor"][2], out["critic"][2])
