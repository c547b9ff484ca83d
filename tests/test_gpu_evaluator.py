"""GPU parity of the evaluator path (evaluator.py:20-201 of the reference) — SURVEY §8(f) rank 3.

tests/golden/eval_*.npz were produced by executing the unmodified reference `evaluate(env, actor, cfg)` (oracle/gen_golden_eval.py):
one arg-max episode, every action / reward vector, the returned [episode_reward, step] and the final fp64 states.  The mirror
must reproduce the whole closed loop: same seeds -> same initial weights and initial env state, then every arg-max action
(integers), every reward (integers), the collision flag and the final fp64 pursuer state bit for bit, the evader to 1e-12."""
import glob
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN_DIR

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

FIXTURES = sorted(glob.glob(os.path.join(GOLDEN_DIR, "eval_*.npz")))
IDS = [os.path.basename(p)[:-4] for p in FIXTURES]


def _cfg(n_def, depth, T, emb):
    from distributed_multi_agent_reinforcement_learning_b200 import default_config
    return default_config(env__num_defender=n_def, env__max_steps=T, algo__depth=depth, algo__embedding_dim=emb,
                          algo__rnn_hidden_dim=emb, algo__learner_device="cuda", algo__worker_device="cuda",
                          algo__evaluator_device="cuda")


@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_evaluate_reproduces_reference_episode(path):
    from distributed_multi_agent_reinforcement_learning_b200.evaluator import evaluate
    from distributed_multi_agent_reinforcement_learning_b200.mappo_parallel import MAPPO
    from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import Pursuit_Env
    fx = np.load(path)
    n_def, depth, T, seed, emb = (int(v) for v in fx["meta"])
    cfg = _cfg(n_def, depth, T, emb)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    agent = MAPPO(cfg, None, None, "Evaluator")          # same construction order as the generator: weights first, env second
    env = Pursuit_Env(cfg)
    actions, rewards = [], []
    orig_step = env.step

    def step(a):
        actions.append(np.asarray(a).astype(np.int32).copy())
        out = orig_step(a)
        rewards.append(np.asarray(out[0], dtype=np.int32).copy())
        return out
    env.step = step
    ret = evaluate(env, agent.actor, cfg)
    got_a, got_r = np.stack(actions), np.stack(rewards)
    assert got_a.shape == fx["actions"].shape
    same = (got_a == fx["actions"]).all(axis=1)
    assert same.all(), f"first differing step {int(np.argmin(same))}: {got_a[np.argmin(same)]} vs {fx['actions'][np.argmin(same)]}"
    assert np.array_equal(got_r, fx["rewards"])
    assert [int(ret[0]), int(ret[1])] == [int(v) for v in fx["ret"]]
    assert bool(env.collision) == bool(fx["collision"][0])
    p = np.asarray(env.get_state("defender"), np.float64)
    assert np.array_equal(p.view(np.int64), fx["p_final"].view(np.int64)), np.abs(p - fx["p_final"]).max()
    np.testing.assert_allclose(np.asarray(env.get_state("attacker"), np.float64), fx["e_final"], rtol=0, atol=1e-12)


def test_evaluator_bookkeeping_and_batched_episodes():
    """Evaluator (EvaluatorProc without Ray): recorder rows, the save rule, and the batched arg-max episodes it is fed with."""
    from distributed_multi_agent_reinforcement_learning_b200.evaluator import Evaluator
    from distributed_multi_agent_reinforcement_learning_b200.mappo_parallel import MAPPO
    cfg = _cfg(8, 1, 12, 128)
    torch.manual_seed(5)
    agent = MAPPO(cfg, None, None, "Evaluator")
    ev = Evaluator(cfg, 16, agent, seed=3)
    rs = ev.get_rewards_and_step()
    assert tuple(rs.shape) == (16, 2) and rs.dtype == torch.float32
    assert (rs[:, 1] == 11).all() and (rs[:, 0] == rs[:, 0].round()).all()
    aw, cw = agent.actor.get_weights(), agent.critic.get_weights()
    if_train, saved = ev.run(aw, cw, 1000, -3.5, (0.25, -0.11))
    assert if_train is True and len(saved) == 3 and saved[2] is ev.recorder            # first evaluation always "improves"
    row = ev.recorder[-1]
    assert row[0] == 1000 and row[3] == -3.5 and row[4:] == (0.25, -0.11) and len(row) == 6
    assert abs(row[1] - ev.last["avg_r"]) == 0 and ev.max_r == row[1]
    ev.max_r = row[1] + 1e9                                                              # a worse average is not saved
    _, saved2 = ev.run(aw, cw, 2000, 0.0, (0.0, 0.0))
    assert saved2 == [] and len(ev.recorder) == 2 and ev.total_step == 2000


@pytest.mark.parametrize("n_def,depth,T", [(8, 1, 20), (5, 3, 16)])
def test_batched_evaluator_episode_equals_evaluate(n_def, depth, T):
    """The Evaluator's batched arg-max episodes (`rollout_batched(deterministic=True, actor_only=True)`) are the reference's
    `evaluate()` episodes (evaluator.py:118-156): the actor alone, conditioned on its OWN previous embeddings A(t-1), A(t-2), ...
    (not on the training rollout's aliased actor / critic history).  One env, same initial state: same action at every step, same
    rewards, same return, bit-identical final pursuer state - for depth 1 and 3, where the two histories differ from t = 1 on."""
    from distributed_multi_agent_reinforcement_learning_b200.evaluator import evaluate
    from distributed_multi_agent_reinforcement_learning_b200.mappo_parallel import MAPPO
    from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import Pursuit_Env, RolloutArena
    cfg = _cfg(n_def, depth, T, 128)
    torch.manual_seed(11)
    agent = MAPPO(cfg, None, None, "Evaluator")
    with torch.no_grad():                                    # biases are zero-initialised: make them matter
        for p in agent.ac_parameters:
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    env = Pursuit_Env(cfg)

    def seed_all():
        random.seed(77)
        np.random.seed(77)

    seed_all()
    actions, rewards = [], []
    orig_step = env.step

    def step(a):
        actions.append(np.asarray(a).astype(np.int32).copy())
        out = orig_step(a)
        rewards.append(np.asarray(out[0], dtype=np.int32).copy())
        return out
    env.step = step
    ret = evaluate(env, agent.actor, cfg)
    p_final = np.asarray(env.get_state("defender"), np.float64)
    env.step = orig_step
    # the same episode as ONE batched rollout on the facade's engine
    seed_all()
    env.reset()
    env.begin_batched_episode()
    eng = env.engine
    arena = RolloutArena(eng.params, 1, T, eng.device)
    for fused in (True, False):
        snap = eng.snapshot()
        agent.rollout_batched(eng, arena, T, seed=0, deterministic=True, actor_only=True, use_fused=fused)
        torch.cuda.synchronize()
        got_a = arena.a_n[:T, 0].cpu().numpy().astype(np.int32)
        same = (got_a == np.stack(actions)).all(axis=1)
        assert same.all(), f"fused={fused}: first differing step {int(np.argmin(same))}"
        assert np.array_equal(arena.raw_reward[:T, 0].cpu().numpy(), np.stack(rewards))
        assert int(arena.raw_reward[:T].sum()) == int(ret[0]) and ret[1] == T - 1
        assert np.array_equal(eng.p_state[0].cpu().numpy().view(np.int64), p_final.view(np.int64))
        eng.restore(snap)
    # and the aliased training history gives a DIFFERENT actor input from t = 1 on (what the evaluator must not use)
    tb_eval = agent.rollout_batched(eng, arena, T, seed=0, deterministic=True, actor_only=True)
    emb_eval = tb_eval.hist_a.clone()
    eng.restore(snap)
    tb_train = agent.rollout_batched(eng, arena, T, seed=0, deterministic=True)
    assert torch.equal(emb_eval[depth], tb_train.hist_a[depth])                      # t = 0: no history yet
    assert not torch.allclose(emb_eval[depth + 1], tb_train.hist_a[depth + 1], rtol=1e-4, atol=1e-5)
