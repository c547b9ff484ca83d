"""TEST INFRASTRUCTURE — golden vectors (embedding_dim 32 to keep the fixtures small) for the algorithm half of the path (DHGN forward, GAE, PPO loss/gradients,
Adam), produced by EXECUTING THE UNMODIFIED REFERENCE `DHGN/mappo_parallel.py` in the build container.

For depth 1 and 3: a Worker runs `explore_env` for a few short episodes on the reference env, the buffers are
concatenated with the reference `BigBuffer`, and a Learner with the same weights runs the reference `train`.
Recorded: the 13 buffer tensors, all weights/buffers before training, advantages / value targets / per-minibatch
log-probs, entropies and values (observed through a line tracer, no code is restated), losses, the accumulated
gradients, and the weights after one reference Adam step.
"""
import os
import sys
from copy import deepcopy

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_bootstrap import load_reference, make_cfg, seed_all  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _trace_train(learner, buffer, total_steps):
    """Runs learner.train under a line tracer and returns (result, captured locals)."""
    import torch
    cap = dict(mb=[])
    code = learner.train.__func__.__code__

    def local(frame, event, arg):
        if event == "line":
            ln = frame.f_lineno
            if ln == 663 and "adv" not in cap:      # self.ac_optimizer.zero_grad(): GAE block finished
                cap["adv"] = frame.f_locals["adv"].detach().clone()
                cap["v_target"] = frame.f_locals["v_target"].detach().clone()
            elif ln == 692:                          # ratios = ...: this minibatch's forward outputs exist
                f = frame.f_locals
                cap["mb"].append(dict(index=list(f["index"]), logp=f["a_logprob_n_now"].detach().clone(),
                                      ent=f["dist_entropy"].detach().clone(), val=f["values_now"].detach().clone()))
            elif ln == 713:                          # after backward + clip of this minibatch
                f = frame.f_locals
                cap["mb"][-1]["critic_loss"] = float(f["critic_loss"].item())
                cap["mb"][-1]["actor_loss"] = float(f["actor_loss"].item())
        return local

    def tracer(frame, event, arg):
        return local if frame.f_code is code else None

    sys.settrace(tracer)
    try:
        res = learner.train(buffer, total_steps)
    finally:
        sys.settrace(None)
    return res, cap


def gen_case(R, depth, n_def, T, episodes, mb, seed, emb=32):
    import torch
    cfg = make_cfg(num_defender=n_def, depth=depth, max_steps=T, embedding_dim=emb)
    seed_all(seed)
    torch.set_grad_enabled(False)
    worker = R.mappo.MAPPO(cfg, None, None, "Worker")
    env = R.pe.Pursuit_Env(cfg)
    big = R.replay_buffer.BigBuffer()
    ep_rewards, o_counts = [], []
    sn_uv = []   # spectral-norm power-iteration buffers evolve on every critic forward, rollout included
    w0 = {("actor." + k): v.clone() for k, v in worker.actor.state_dict().items()}
    w0.update({("critic." + k): v.clone() for k, v in worker.critic.state_dict().items()})
    for e in range(episodes):
        r, buf, steps = worker.explore_env(env, 1)
        assert steps == T
        ep_rewards.append(float(r))
        o_counts.append(len(env.boundary_map.obstacle_agent))
        big.concat_buffer(deepcopy(buf))
        sn_uv.append((worker.critic.Mean.weight_u.clone(), worker.critic.Mean.weight_v.clone()))
    buffer = {k: v.clone() for k, v in big.buffer.items()}
    learner = R.mappo.MAPPO(cfg, episodes, mb, "Learner")
    learner.actor.set_weights(worker.actor.get_weights())
    learner.critic.set_weights(worker.critic.get_weights())
    w1 = {("actor." + k): v.clone() for k, v in learner.actor.state_dict().items()}
    w1.update({("critic." + k): v.clone() for k, v in learner.critic.state_dict().items()})
    torch.set_grad_enabled(True)
    total_steps = episodes * T
    (objC, objA, ag, cg), cap = _trace_train(learner, big, total_steps)
    torch.set_grad_enabled(False)
    grads = {}
    for (name, p_), g in zip(learner.actor.named_parameters(), ag):
        grads["actor." + name] = None if g is None else np.asarray(g)
    for (name, p_), g in zip(learner.critic.named_parameters(), cg):
        grads["critic." + name] = None if g is None else np.asarray(g)
    # one reference optimizer step on those gradients (runner.py:72-78 without the cross-learner sum)
    learner.ac_optimizer.zero_grad()
    learner.actor.set_gradients(ag, torch.device("cpu"))
    learner.critic.set_gradients(cg, torch.device("cpu"))
    lr_used = learner.ac_optimizer.param_groups[0]["lr"]
    learner.ac_optimizer.step()
    w2 = {("actor." + k): v.clone() for k, v in learner.actor.state_dict().items()}
    w2.update({("critic." + k): v.clone() for k, v in learner.critic.state_dict().items()})

    fx = {}
    for k, v in buffer.items():
        fx["buf." + k] = v.numpy()
    for k, v in w0.items():                # weights used during the rollout: only the SN u/v buffers differ from w1
        if not np.array_equal(v.numpy(), w1[k].numpy()):
            fx["w_init." + k] = v.numpy()
    for k, v in w1.items():
        fx["w." + k] = v.numpy()           # weights at the start of train()
    for k, v in w2.items():
        fx["w_after." + k] = v.numpy()
    for k, v in grads.items():
        if v is not None:
            fx["grad." + k] = v
    fx["grad_none"] = np.array([k for k, v in grads.items() if v is None])
    fx["adv"] = cap["adv"].numpy()
    fx["v_target"] = cap["v_target"].numpy()
    for i, m in enumerate(cap["mb"]):
        fx[f"mb{i}.index"] = np.array(m["index"], np.int64)
        fx[f"mb{i}.logp"] = m["logp"].numpy()
        fx[f"mb{i}.ent"] = m["ent"].numpy()
        fx[f"mb{i}.val"] = m["val"].numpy()
        fx[f"mb{i}.losses"] = np.array([m["critic_loss"], m["actor_loss"]])
    fx["n_mb"] = np.int32(len(cap["mb"]))
    fx["objC"], fx["objA"] = np.float64(objC), np.float64(objA)
    fx["lr_used"] = np.float64(lr_used)
    fx["total_steps"] = np.int64(total_steps)
    fx["ep_rewards"] = np.array(ep_rewards)
    fx["o_counts"] = np.array(o_counts, np.int32)
    fx["meta"] = np.array([depth, n_def, T, episodes, mb, seed, emb], np.int64)
    for i, (u, v) in enumerate(sn_uv):
        fx[f"sn_u_after_ep{i}"] = u.numpy()
        fx[f"sn_v_after_ep{i}"] = v.numpy()
    return fx


# (depth, pursuers, T, episodes, minibatch episodes, seed, embedding_dim).  The E = 128 cases pin training at the PRODUCTION width,
# where every dense layer runs on the tcgen05 kernels (rowgemm / wgrad / gru_seq) instead of the library GEMMs the E = 32 cases
# fall back to; 5 episodes with minibatches of 2 give 3 minibatches (the last one ragged), so the running clip matters.
CASES = ((1, 4, 10, 3, 2, 21, 32), (3, 5, 8, 3, 2, 23, 32), (1, 8, 20, 5, 2, 25, 128), (3, 5, 16, 5, 2, 27, 128))


def gen_algo(only_emb=None):
    R = load_reference()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for depth, n_def, T, episodes, mb, seed, emb in CASES:
        if only_emb is not None and emb != only_emb:
            continue
        fx = gen_case(R, depth, n_def, T, episodes, mb, seed, emb)
        if emb != 32:
            # keep the wide fixtures small: the encoder is ONE module shared by both networks, so its critic.* copies (weights and
            # gradients) are dropped after checking they are identical, and the weights after the optimizer step are not stored
            # (the reference's optimizer is stock torch.optim.Adam(eps=1e-5): the test re-runs it on the stored gradients)
            for k in [k for k in fx if ".critic.shared_net." in k]:
                twin = k.replace(".critic.", ".actor.")
                assert np.array_equal(fx[k], fx[twin]), k
                del fx[k]
            for k in [k for k in fx if k.startswith("w_after.")]:
                del fx[k]
        path = os.path.join(GOLDEN_DIR, f"algo_d{depth}_n{n_def}" + ("" if emb == 32 else f"_e{emb}") + ".npz")
        np.savez_compressed(path, **fx)
        print(f"algo depth={depth} N={n_def} E={emb}: objC={float(fx['objC']):.6f} objA={float(fx['objA']):.6f} "
              f"minibatches={int(fx['n_mb'])} file={os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    gen_algo(int(sys.argv[1]) if len(sys.argv) > 1 else None)
