"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C-ABI, against
(1) the committed golden vectors from the reference and (2) the CPU oracle on seeded synthetic inputs.
Bit-exact for every integer/bit quantity and for the fp64 pursuer state; the evader's fp64 state (CUDA libm vs
numpy cos/sin/acos) is held to 1e-9 relative — north_star allows 1e-5 for float observations."""
import numpy as np
import pytest

from conftest import env_fixture_names, golden, unpack_bits

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

NAMES = env_fixture_names()


def _cfg_for(fx, **over):
    from distributed_multi_agent_reinforcement_learning_b200 import default_config
    g = lambda k: fx["param_" + k].item()
    W, H = g("W"), g("H")
    kw = dict(env__num_defender=g("N"), env__max_steps=g("max_steps"), map__map_size=[W, H],
              map__center=[30, 25] if H == 55 else [30, 30], map__num_max_obstacle=g("O"),
              attacker__extend_dis=g("e_extend_dis"))
    kw.update(over)
    return default_config(**kw)


def _engine(fx, B):
    from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import BatchedPursuitEnv
    env = BatchedPursuitEnv(_cfg_for(fx), B, num_maps=1)
    env.action_table.copy_(torch.from_numpy(fx["action_table"]))
    env.beam_dir.copy_(torch.from_numpy(fx["beam_dir"]))
    env.set_maps(fx["grid"][None], fx["inflated"][None])
    return env


def _bits(a):
    return np.ascontiguousarray(a).view(np.int64)


@pytest.mark.parametrize("name", NAMES)
def test_step_golden(name):
    """All T golden steps as T independent envs (teacher forced): rewards, can_apply, next fp64 state."""
    fx = golden(name)
    T = fx["action"].shape[0]
    env = _engine(fx, T)
    env.set_state(fx["p_state"][:T], fx["e_after"], map_id=np.zeros(T, np.int32))
    env.time_step.copy_(torch.arange(T, dtype=torch.int32))
    reward, done = env.step(torch.from_numpy(fx["action"]).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(reward.cpu().numpy(), fx["reward"])
    assert np.array_equal(env.can_apply.cpu().numpy(), fx["can_apply"])
    assert np.array_equal(_bits(env.p_state.cpu().numpy()), _bits(fx["p_state"][1:]))
    assert np.array_equal(env.collision.cpu().numpy(), (fx["can_apply"] == 0).any(axis=1).astype(np.uint8))
    assert np.array_equal(done.cpu().numpy(), fx["done"])
    assert np.array_equal(env.time_step.cpu().numpy(), np.arange(1, T + 1))


@pytest.mark.parametrize("name", NAMES)
def test_closed_loop_golden(name):
    """One env stepped T times from the initial state only, evader from the tape."""
    fx = golden(name)
    env = _engine(fx, 1)
    env.set_state(fx["p_state"][:1], fx["e_after"][:1], map_id=np.zeros(1, np.int32))
    act = torch.from_numpy(fx["action"]).cuda()
    for t in range(fx["action"].shape[0]):
        env.e_state.copy_(torch.from_numpy(fx["e_after"][t:t + 1]))
        env.step(act[t:t + 1])
    assert np.array_equal(_bits(env.p_state.cpu().numpy()[0]), _bits(fx["p_state"][-1]))
    assert int(env.collision.item()) == int(fx["collision"][-1])


@pytest.mark.parametrize("name", NAMES)
def test_sensor_tables_golden(name):
    fx = golden(name)
    env = _engine(fx, 1)
    from distributed_multi_agent_reinforcement_learning_b200 import maps
    p = env.params
    n_b = int(env.boundary_count.item())
    assert n_b == len(fx["boundary_xy"])
    assert np.array_equal(env.boundary_xy[0, :n_b].cpu().numpy(), fx["boundary_xy"])
    assert not env.boundary_xy[0, n_b:].any()
    assert np.array_equal(maps.unpack_words(env.boundary_bits[0].cpu().numpy(), p.H), fx["boundary"])
    raser = maps.unpack_words(env.raser_bits[0].cpu().numpy(), p.O).reshape(p.W, p.H, p.O)
    assert np.array_equal(raser[..., :n_b], unpack_bits(fx["raser_packed"], n_b))
    assert not raser[..., n_b:].any()


@pytest.mark.parametrize("name", NAMES)
def test_observe_golden(name):
    fx = golden(name)
    from distributed_multi_agent_reinforcement_learning_b200 import maps
    T1 = fx["p_state"].shape[0]
    env = _engine(fx, T1)
    e = np.concatenate([fx["e_before"], fx["e_after"][-1:]], axis=0)
    env.set_state(fx["p_state"], e, map_id=np.zeros(T1, np.int32))
    p_adj, e_adj, o_adj = env.observe(dense=True)
    torch.cuda.synchronize()
    n_b = int(fx["raser_ob"])
    o_gold = unpack_bits(fx["o_adj_packed"], n_b)
    assert np.array_equal(p_adj.cpu().numpy(), fx["p_adj"].astype(np.float32))
    assert np.array_equal(e_adj.cpu().numpy()[..., 0], fx["e_adj"].astype(np.float32))
    assert np.array_equal(o_adj.cpu().numpy()[..., :n_b], o_gold.astype(np.float32))
    assert not o_adj[..., n_b:].any()
    # packed canonical outputs carry the same bits
    N, O = env.N, env.O
    assert np.array_equal(maps.unpack_words(env.p_adj_bits.cpu().numpy(), N), fx["p_adj"])
    assert np.array_equal(env.e_adj.cpu().numpy(), fx["e_adj"])
    assert np.array_equal(maps.unpack_words(env.o_adj_bits.cpu().numpy(), O)[..., :n_b], o_gold)


@pytest.mark.parametrize("name", NAMES)
def test_evader_teacher_forced_golden(name):
    """Every golden attacker_step as its own env: A* paths exact, target exact, fp64 state to 1e-9."""
    fx = golden(name)
    T = fx["action"].shape[0]
    env = _engine(fx, T)
    env.set_state(fx["p_state"][:T], fx["e_before"], fx["e_target_attr"], np.zeros(T, np.int32))
    env.time_step.copy_(torch.arange(T, dtype=torch.int32))
    # previous path of env t = golden path after step t-1
    offs = np.concatenate([[0], np.cumsum(fx["path_len"])])
    path = np.zeros((T, env.PATH_CAP, 2), np.int16)
    plen = np.zeros(T, np.int32)
    for t in range(1, T):
        L = int(fx["path_len"][t - 1])
        path[t, :L] = fx["path_flat"][offs[t - 1]:offs[t]]
        plen[t] = L
    env.path.copy_(torch.from_numpy(path))
    env.path_len.copy_(torch.from_numpy(plen))
    env.set_target_tape(fx["target"][1:T + 1].reshape(T, 1, 2))   # the draw the reference made, if it made one
    env.evader_step()
    torch.cuda.synchronize()
    status = env.evader_status.cpu().numpy()
    assert not (status & 3).any(), "search heap / path overflow"
    got_len = env.path_len.cpu().numpy()
    got_path = env.path.cpu().numpy()
    assert np.array_equal(got_len, fx["path_len"])
    for t in range(T):
        assert np.array_equal(got_path[t, :got_len[t]].astype(np.int32), fx["path_flat"][offs[t]:offs[t + 1]]), (name, t)
    np.testing.assert_allclose(env.e_state.cpu().numpy(), fx["e_after"], rtol=1e-9, atol=1e-9)
    assert np.array_equal(env.target.cpu().numpy(), fx["target"][1:T + 1])


@pytest.mark.parametrize("name", NAMES)
def test_evader_closed_loop_golden(name):
    fx = golden(name)
    T = fx["action"].shape[0]
    env = _engine(fx, 1)
    env.set_state(fx["p_state"][:1], fx["e_before"][:1], fx["e_target_attr"][:1], np.zeros(1, np.int32))
    tape = fx["targets_drawn"].reshape(1, -1, 2)
    env.set_target_tape(tape)
    worst = 0.0
    for t in range(T):
        env.p_state.copy_(torch.from_numpy(fx["p_state"][t:t + 1]))
        env.time_step.fill_(t)
        env.evader_step()
        got = env.e_state.cpu().numpy()[0]
        worst = max(worst, float(np.abs(got - fx["e_after"][t]).max()))
    assert worst < 1e-7, worst
    assert np.array_equal(env.target.cpu().numpy()[0], fx["target"][-1])
    assert int(env.evader_status.item()) == 0


@pytest.mark.parametrize("name", NAMES)
def test_welford_golden(name):
    fx = golden(name)
    T, N = fx["reward"].shape
    env = _engine(fx, 1)
    for t in range(T):
        env.reward.copy_(torch.from_numpy(fx["reward"][t:t + 1]))
        out = env.normalize_reward().cpu().numpy()[0]
        assert np.array_equal(out, fx["r_norm"][t].astype(np.float32)), (name, t)


# ------------------------------------------------------------------------------------------------------------
# synthetic batches vs the CPU oracle


def _random_batch(cfg, B, M, seed, crowd=False):
    from distributed_multi_agent_reinforcement_learning_b200 import maps
    from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import BatchedPursuitEnv
    env = BatchedPursuitEnv(cfg, B, num_maps=M)
    if not crowd:
        env.reset(seed=seed)
        return env
    # many pursuers: the reference's chain-connectivity placement rule cannot seat them, so draw
    # positions directly, crowded into a band at the map edge so that collisions and edge clips actually happen
    rng = maps.GenRng(seed)
    grids = np.stack([maps.make_obstacle_grid(cfg.map, rng) for _ in range(M)])
    env.set_maps(grids)
    g = np.random.default_rng(seed)
    ps = np.zeros((B, env.N, 4))
    ps[:, :, 0] = g.uniform(0, 3.0, (B, env.N))
    ps[:, :, 1] = g.uniform(0, env.params.H - 1, (B, env.N))
    es = np.zeros((B, 4))
    es[:, 0] = g.uniform(0, 4.0, B)
    es[:, 1] = g.uniform(0, env.params.H - 1, B)
    env.set_state(ps, es, np.zeros((B, 2), np.int32), np.arange(B) % M, time_step=0)
    env.start_episode()
    return env


def _oracle_state(env, orc, K, e_tape, actions):
    from distributed_multi_agent_reinforcement_learning_b200 import maps
    p = orc.EnvParams.from_dict(env.params.as_dict())
    B, N, O, M = env.B, env.N, env.O, env.M
    grid = maps.unpack_words(env.grid_bits.cpu().numpy(), p.H)                                   # [M,W,H]
    raser = maps.unpack_words(env.raser_bits.cpu().numpy(), O).reshape(M, p.W * p.H, O)
    st = dict(p_state=env.p_state.cpu().numpy().copy(), grid=np.ascontiguousarray(grid),
              raser=np.ascontiguousarray(raser), ob_count=np.minimum(env.boundary_count.cpu().numpy(), O).astype(np.int32),
              map_id=env.map_id.cpu().numpy().copy(), action_table=env.action_table.cpu().numpy().copy(),
              p_adj=np.zeros((B, N, N), np.uint8), o_adj=np.zeros((B, N, O), np.uint8), e_adj=np.zeros((B, N), np.uint8),
              reward=np.zeros((B, N), np.int32), can_apply=np.zeros((B, N), np.uint8), collision=np.zeros(B, np.uint8),
              time_step=np.zeros(B, np.int32), done=np.zeros(B, np.uint8), wf_n=np.zeros(B, np.int64),
              wf_mean=np.zeros((B, N)), wf_S=np.zeros((B, N)), wf_std=np.zeros((B, N)), r_norm=np.zeros((B, N), np.float32))
    rec = dict(p_state=[], p_adj=[], o_adj=[], e_adj=[], reward=[], r_norm=[])
    for k in range(K):
        st["e_before"] = np.ascontiguousarray(e_tape[k])
        st["e_after"] = np.ascontiguousarray(e_tape[k + 1])
        st["action"] = np.ascontiguousarray(actions[k])
        rec["p_state"].append(st["p_state"].copy())
        orc.rollout_iteration(p, st)
        for key in ("p_adj", "o_adj", "e_adj", "reward", "r_norm"):
            rec[key].append(st[key].copy())
    return st, {k: np.stack(v) for k, v in rec.items()}


def _evader_tape(env, K, seed):
    """Synthetic but plausible evader motion: a bounded random walk (the evader kernel has its own tests)."""
    rng = np.random.default_rng(seed)
    e0 = env.e_state.cpu().numpy()
    tape = np.zeros((K + 1, env.B, 4))
    tape[0] = e0
    p = env.params
    for k in range(K):
        step = rng.normal(0, 0.25, (env.B, 2))
        tape[k + 1, :, :2] = np.clip(tape[k, :, :2] + step, 0, [p.W - 1, p.H - 1])
        tape[k + 1, :, 2:] = step / 0.1
    return tape


def _rand_actions_numpy(seed, B, N, t0, K):
    """numpy re-implementation of rand_action (csrc/common.cuh) — integer arithmetic, must match bit for bit."""
    def sm(z):
        z = (z + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))
    with np.errstate(over="ignore"):
        agent = np.arange(B * N, dtype=np.uint64).reshape(1, B, N)
        t = np.arange(t0, t0 + K, dtype=np.uint64).reshape(K, 1, 1)
        h = sm(np.uint64(seed) ^ sm(agent * np.uint64(0x100000001B3) + t))
        return (((h >> np.uint64(32)) * np.uint64(9)) >> np.uint64(32)).astype(np.int32)


@pytest.mark.parametrize("N,B,M", [(8, 64, 8), (4, 37, 5), (15, 33, 4), (32, 9, 3), (64, 5, 2), (100, 3, 1)])
def test_fused_rollout_vs_oracle(oracle, N, B, M):
    from distributed_multi_agent_reinforcement_learning_b200 import default_config, maps
    from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import RolloutArena
    K = 24
    cfg = default_config(env__num_defender=N, env__max_steps=K)
    env = _random_batch(cfg, B, M, seed=N, crowd=N >= 32)
    e_tape = _evader_tape(env, K, seed=1)
    actions = _rand_actions_numpy(1234, B, N, 0, K)
    st, rec = _oracle_state(env, oracle, K, e_tape, actions)
    arena = RolloutArena(env.params, B, K, env.device, dense=True)
    env.rollout(arena, K, 0, e_tape=torch.from_numpy(e_tape).cuda(), action_tape=None, seed=1234)
    torch.cuda.synchronize()
    assert np.array_equal(arena.a_n.cpu().numpy(), actions.astype(np.float32)), "device action generator != numpy"
    assert np.array_equal(arena.raw_reward.cpu().numpy(), rec["reward"])
    assert np.array_equal(arena.r.cpu().numpy(), rec["r_norm"])
    assert np.array_equal(arena.p_state_f32.cpu().numpy(), rec["p_state"].astype(np.float32))
    assert np.array_equal(maps.unpack_words(arena.p_adj_bits.cpu().numpy(), N), rec["p_adj"])
    assert np.array_equal(maps.unpack_words(arena.o_adj_bits.cpu().numpy(), env.O), rec["o_adj"])
    assert np.array_equal(arena.e_adj.cpu().numpy(), rec["e_adj"])
    assert np.array_equal(arena.p_adj_f32.cpu().numpy(), rec["p_adj"].astype(np.float32))
    assert np.array_equal(arena.o_adj_f32.cpu().numpy(), rec["o_adj"].astype(np.float32))
    assert np.array_equal(arena.e_adj_f32.cpu().numpy()[..., 0], rec["e_adj"].astype(np.float32))
    assert np.array_equal(_bits(env.p_state.cpu().numpy()), _bits(st["p_state"]))
    assert np.array_equal(env.collision.cpu().numpy(), st["collision"])
    assert np.array_equal(env.time_step.cpu().numpy(), st["time_step"])
    assert np.array_equal(_bits(env.wf_mean.cpu().numpy()), _bits(st["wf_mean"]))
    assert np.array_equal(_bits(env.wf_std.cpu().numpy()), _bits(st["wf_std"]))
    # reference-layout views are permutations of the arena, not copies
    assert arena.reference_view("p_state").shape == (B, K, N, 4)
    assert arena.reference_view("o_adj").data_ptr() == arena.o_adj_f32.data_ptr()
    if N >= 32:
        assert (rec["reward"] < 0).any(), "scenario did not exercise collisions"


def test_stepwise_kernels_equal_fused_rollout_at_full_size():
    """BASELINE config 2 size (4096 envs x 8 pursuers): the per-step kernels (observe/step/welford) and the fused
    rollout kernel must agree bit for bit; plus the domain invariants."""
    from distributed_multi_agent_reinforcement_learning_b200 import default_config, maps
    from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import RolloutArena
    B, N, M, K = 4096, 8, 64, 30
    cfg = default_config(env__num_defender=N, env__max_steps=K)
    env = _random_batch(cfg, B, M, seed=5)
    p0 = env.p_state.clone()
    e_tape = torch.from_numpy(_evader_tape(env, K, seed=2)).cuda()
    actions = torch.from_numpy(_rand_actions_numpy(77, B, N, 0, K)).cuda()
    arena = RolloutArena(env.params, B, K, env.device)
    env.rollout(arena, K, 0, e_tape=e_tape, action_tape=actions)
    fused_state = env.p_state.clone()
    fused_wf = env.wf_mean.clone()
    # replay with the per-step kernels
    env.p_state.copy_(p0)
    for t in (env.wf_n, env.wf_mean, env.wf_S, env.wf_std, env.time_step, env.collision):
        t.zero_()
    for k in range(K):
        env.e_state.copy_(e_tape[k])
        pa, ea, oa = env.observe()
        assert torch.equal(pa, arena.p_adj_bits[k]) and torch.equal(ea, arena.e_adj[k]) and torch.equal(oa, arena.o_adj_bits[k])
        before = env.p_state.clone()
        env.e_state.copy_(e_tape[k + 1])
        reward, _ = env.step(actions[k])
        assert torch.equal(reward, arena.raw_reward[k])
        assert torch.equal(env.normalize_reward(), arena.r[k])
        # invariant: a rejected move leaves the agent untouched; accepted moves stay inside the clip box
        rej = env.can_apply == 0
        assert torch.equal(env.p_state[rej], before[rej])
        assert (env.p_state[..., 0] >= 0).all() and (env.p_state[..., 0] <= env.params.W - 1).all()
        assert (env.p_state[..., 1] >= 0).all() and (env.p_state[..., 1] <= env.params.H - 1).all()
    assert torch.equal(env.p_state, fused_state)
    assert torch.equal(env.wf_mean, fused_wf)
    assert (env.time_step == K).all() and env.done.all()
    # adjacency quirk survives at scale: column 1 all ones, strictly-lower triangle zero elsewhere
    adj = maps.unpack_words(arena.p_adj_bits[0].cpu().numpy(), N)
    assert (adj[:, :, 1] == 1).all() and (adj[:, np.arange(N), np.arange(N)] == 1).all()
    low = np.tril(np.ones((N, N), bool), -1)
    low[:, 1] = False
    assert not adj[:, low].any()


def test_closed_loop_rollout_vs_oracle(oracle):
    """Whole closed loop on the GPU (A* replanning + fused evader move + env kernels) against the oracle's closed
    loop.  Discrete records exact; the evader's fp64 state differs only by libm ulps."""
    from distributed_multi_agent_reinforcement_learning_b200 import default_config, maps
    from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import RolloutArena
    import bench
    B, N, M, K = 96, 8, 12, 45
    cfg = default_config(env__num_defender=N, env__max_steps=K)
    wl = bench.host_workload(cfg, B, M, seed=21)
    from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import BatchedPursuitEnv
    env = BatchedPursuitEnv(cfg, B, num_maps=M)
    env.set_maps(wl["grids"], wl["inflated"])
    env.set_state(wl["p_state"], wl["e_state"], wl["target"], wl["map_id"], time_step=0)
    env.set_target_tape(wl["tape"])
    env.start_episode()
    actions = _rand_actions_numpy(99, B, N, 0, K)
    # oracle
    p = oracle.EnvParams.from_dict(env.params.as_dict())
    O = env.O
    st = dict(p_state=wl["p_state"].copy(), e_state=wl["e_state"].copy(), target=wl["target"].copy(),
              path=np.zeros((B, 512, 2), np.int16), path_len=np.zeros(B, np.int32),
              grid=np.ascontiguousarray(wl["grids"]), inflated=np.ascontiguousarray(wl["inflated"]),
              raser=np.ascontiguousarray(maps.unpack_words(env.raser_bits.cpu().numpy(), O).reshape(M, p.W * p.H, O)),
              ob_count=np.minimum(env.boundary_count.cpu().numpy(), O).astype(np.int32), map_id=wl["map_id"].copy(),
              action_table=env.action_table.cpu().numpy().copy(), tape=np.ascontiguousarray(wl["tape"]),
              tape_pos=np.zeros(B, np.int32), p_adj=np.zeros((B, N, N), np.uint8), o_adj=np.zeros((B, N, O), np.uint8),
              e_adj=np.zeros((B, N), np.uint8), reward=np.zeros((B, N), np.int32), can_apply=np.zeros((B, N), np.uint8),
              collision=np.zeros(B, np.uint8), time_step=np.zeros(B, np.int32), done=np.zeros(B, np.uint8),
              wf_n=np.zeros(B, np.int64), wf_mean=np.zeros((B, N)), wf_S=np.zeros((B, N)), wf_std=np.zeros((B, N)),
              r_norm=np.zeros((B, N), np.float32), status=np.zeros(B, np.int32))
    rec = dict(reward=[], e_adj=[], o_adj=[], p_adj=[], e_before=[])
    for k in range(K):
        st["action"] = np.ascontiguousarray(actions[k])
        rec["e_before"].append(st["e_state"].copy())
        oracle.rollout_iteration_closed(p, st)
        for key in ("reward", "e_adj", "o_adj", "p_adj"):
            rec[key].append(st[key].copy())
    assert not st["status"].any()
    arena = RolloutArena(env.params, B, K, env.device)
    env.rollout_closed(arena, K, 0, action_tape=torch.from_numpy(actions).cuda())
    torch.cuda.synchronize()
    assert int(env.evader_status.max().item()) == 0
    assert np.array_equal(env.path_len.cpu().numpy(), st["path_len"])
    assert np.array_equal(env.target.cpu().numpy(), st["target"])
    assert np.array_equal(env.tape_pos.cpu().numpy(), st["tape_pos"])
    np.testing.assert_allclose(env.e_state.cpu().numpy(), st["e_state"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(arena.e_state_f32.cpu().numpy()[:, :, 0], np.stack(rec["e_before"]).astype(np.float32), rtol=1e-6)
    assert np.array_equal(arena.raw_reward.cpu().numpy(), np.stack(rec["reward"]))
    assert np.array_equal(arena.e_adj.cpu().numpy(), np.stack(rec["e_adj"]))
    assert np.array_equal(maps.unpack_words(arena.o_adj_bits.cpu().numpy(), O), np.stack(rec["o_adj"]))
    assert np.array_equal(maps.unpack_words(arena.p_adj_bits.cpu().numpy(), N), np.stack(rec["p_adj"]))
    assert np.array_equal(_bits(env.p_state.cpu().numpy()), _bits(st["p_state"]))
    assert (env.time_step == K).all()


def test_evader_hard_searches_vs_oracle(oracle):
    """Searches that run past the reachability-check budget: (a) goal sealed inside a ring of obstacles — the reference's
    A* exhausts the component and returns [start]; (b) a serpentine maze with a long but existing path."""
    from distributed_multi_agent_reinforcement_learning_b200 import default_config, maps
    from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import BatchedPursuitEnv
    cfg = default_config(env__num_defender=4)
    W, H = cfg.map.map_size
    ring = np.zeros((W, H), np.uint8)
    ring[36:47, 26] = ring[36:47, 36] = 1
    ring[36, 26:37] = ring[46, 26:37] = 1                   # closed ring around (41, 31)
    maze = np.zeros((W, H), np.uint8)
    for i, x in enumerate(range(8, 52, 6)):                  # walls with alternating gaps
        if i % 2 == 0:
            maze[x, 0:H - 6] = 1
        else:
            maze[x, 6:H] = 1
    grids = np.stack([ring, maze])
    B = 2
    env = BatchedPursuitEnv(cfg, B, num_maps=2)
    env.set_maps(grids)
    p_state = np.zeros((B, 4, 4))
    p_state[:, :, 0] = [[2, 3, 4, 5]] * B                    # pursuers parked out of the way
    p_state[:, :, 1] = 50.0
    e_state = np.array([[5.2, 5.1, 0, 0], [2.3, 2.2, 0, 0]], np.float64)
    target = np.array([[41, 31], [57, 3]], np.int32)
    env.set_state(p_state, e_state, target, np.arange(B), time_step=0)
    env.start_episode()
    env.set_target_tape(np.zeros((B, 0, 2), np.int32))
    env.evader_step()
    torch.cuda.synchronize()
    assert not (env.evader_status.cpu().numpy() & 3).any()
    p = oracle.EnvParams.from_dict(env.params.as_dict())
    infl = maps.dilate(grids, 2)
    for b in range(B):
        ev = oracle.EvaderState(e_state[b], target[b])
        assert oracle.evader_step(p, ev, p_state[b], 0, grids[b], infl[b], np.zeros((0, 2), np.int32)) in (0, -2)
        L = ev.path_len.value
        assert int(env.path_len[b]) == L, (b, int(env.path_len[b]), L)
        assert np.array_equal(env.path[b, :L].cpu().numpy(), ev.path[:L])
        np.testing.assert_allclose(env.e_state[b].cpu().numpy(), ev.e_state, rtol=1e-9, atol=1e-9)
    assert int(env.path_len[0]) == 1                          # sealed goal -> [start]
    assert int(env.path_len[1]) > 100                         # the maze path is long


@pytest.mark.parametrize("size", [(60, 55), (120, 60), (200, 62)])
def test_evader_cup_trap_search_vs_oracle(oracle, size):
    """The evader starts inside a cup that opens away from its target: the weighted A* floods the cup before it finds the way
    round (1.5 K pops and ~160 OPEN entries on the default map, 4.5 K pops and ~420 entries on 200 x 62 - stripes of the OPEN
    list many entries long, dead-end pops, every column loop with more columns than lanes).  Path, length and the evader's move
    must equal the oracle's."""
    from distributed_multi_agent_reinforcement_learning_b200 import default_config, maps
    from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import BatchedPursuitEnv
    W, H = size
    cfg = default_config(env__num_defender=4, map__map_size=[W, H])
    grid = np.zeros((W, H), np.uint8)
    x0 = W // 2
    grid[x0, 4:H - 4] = 1
    grid[x0 - W // 4:x0 + 1, 4] = 1
    grid[x0 - W // 4:x0 + 1, H - 5] = 1                        # (< 176 boundary cells: the obstacle slots of the sensor tables)
    grids = grid[None]
    env = BatchedPursuitEnv(cfg, 1, num_maps=1)
    env.set_maps(grids)
    p_state = np.zeros((1, 4, 4))
    p_state[0, :, 0] = [1, 2, 3, 4]                            # pursuers parked in a corner
    p_state[0, :, 1] = 1.0
    e_state = np.array([[x0 - 5.2, H // 2 + 0.1, 0, 0]], np.float64)
    target = np.array([[W - 4, H // 2]], np.int32)
    env.set_state(p_state, e_state, target, np.arange(1), time_step=0)
    env.start_episode()
    env.set_target_tape(np.zeros((1, 0, 2), np.int32))
    env.evader_step()
    torch.cuda.synchronize()
    assert not (env.evader_status.cpu().numpy() & 3).any()
    p = oracle.EnvParams.from_dict(env.params.as_dict())
    infl = maps.dilate(grids, 2)
    ev = oracle.EvaderState(e_state[0], target[0])
    assert oracle.evader_step(p, ev, p_state[0], 0, grids[0], infl[0], np.zeros((0, 2), np.int32)) in (0, -2)
    L = ev.path_len.value
    assert L > W // 4                                          # out of the cup and round the wall
    assert int(env.path_len[0]) == L
    assert np.array_equal(env.path[0, :L].cpu().numpy(), ev.path[:L])
    np.testing.assert_allclose(env.e_state[0].cpu().numpy(), ev.e_state, rtol=1e-9, atol=1e-9)


def test_grouped_multistream_rollout_is_identical():
    """Sub-batches on separate streams (and inside a CUDA graph) give bit-identical arenas and final states."""
    from distributed_multi_agent_reinforcement_learning_b200 import default_config
    from distributed_multi_agent_reinforcement_learning_b200.pursuit_env import BatchedPursuitEnv, EpisodeGraph, RolloutArena
    import bench
    B, N, M, K = 203, 8, 16, 37
    cfg = default_config(env__num_defender=N, env__max_steps=K)
    wl = bench.host_workload(cfg, B, M, seed=5)
    env = BatchedPursuitEnv(cfg, B, num_maps=M)
    env.set_maps(wl["grids"], wl["inflated"])
    env.set_state(wl["p_state"], wl["e_state"], wl["target"], wl["map_id"], time_step=0)
    env.set_target_tape(wl["tape"])
    env.start_episode()
    snap = env.snapshot()
    fields = ("p_state_f32", "e_state_f32", "p_adj_bits", "e_adj", "o_adj_bits", "a_n", "r", "raw_reward")
    a1 = RolloutArena(env.params, B, K, env.device)
    env.rollout_closed(a1, K, 0, seed=7, groups=1)
    torch.cuda.synchronize()
    ref_state = {k: v.clone() for k, v in env.snapshot().items()}
    for groups in (5, 16):
        env.restore(snap)
        a2 = RolloutArena(env.params, B, K, env.device)
        env.rollout_closed(a2, K, 0, seed=7, groups=groups)
        torch.cuda.synchronize()
        for f in fields:
            assert torch.equal(getattr(a1, f), getattr(a2, f)), (groups, f)
        for k, v in env.snapshot().items():
            assert torch.equal(v, ref_state[k]), (groups, k)
    # with an action tape and through a captured graph
    env.restore(snap)
    tape = torch.from_numpy(_rand_actions_numpy(7, B, N, 0, K)).cuda()
    a3 = RolloutArena(env.params, B, K, env.device)
    env.rollout_closed(a3, K, 0, action_tape=tape, groups=4)
    torch.cuda.synchronize()
    for f in fields:
        assert torch.equal(getattr(a1, f), getattr(a3, f)), f
    env.restore(snap)
    a4 = RolloutArena(env.params, B, K, env.device)
    g = EpisodeGraph(env, a4, K, seed=7, groups=6)
    a4.raw_reward.zero_()
    g.replay()
    torch.cuda.synchronize()
    for f in fields:
        assert torch.equal(getattr(a1, f), getattr(a4, f)), f
