"""The data-parallel protocol on the GPU path, two ranks ("all-reduced grads == sum of single-GPU grads";
main.py:73-75,99-129 and runner.py:72-78 are the reference's protocol: every learner starts from learner 0's weights, accumulates
and clips its own gradients, the driver SUMS them, every learner applies the same Adam step).

Each rank runs `sync_weights -> train(its shard) -> update` through NCCL when the box has two GPUs (one rank per GPU), otherwise
through gloo with both ranks on the one GPU (same kernels, same protocol, host-staged collective).  Checked: the flat parameter
arena is bit-identical on both ranks afterwards and equals a single-process run that sums the two shards' gradient arenas; the
north-star mode `train(allreduce="minibatch")` (one all-reduce per PPO minibatch, the clip acting on the reduced running sum)
equals a single-process emulation built from the two shards' per-minibatch gradients."""
import os
import socket

import numpy as np
import pytest

from conftest import GOLDEN_DIR, AlgoFixture

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

FIXTURE = os.path.join(GOLDEN_DIR, "algo_d1_n8_e128.npz")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup(device):
    from distributed_multi_agent_reinforcement_learning_b200 import default_config
    from distributed_multi_agent_reinforcement_learning_b200.mappo_parallel import MAPPO
    fx = AlgoFixture(FIXTURE)
    depth, n_def, T, episodes, mb, seed, emb = (int(v) for v in fx["meta"])
    cfg = default_config(env__num_defender=n_def, env__max_steps=T, algo__depth=depth, algo__embedding_dim=emb, algo__rnn_hidden_dim=emb,
                         algo__learner_device=str(device), algo__worker_device=str(device))
    m = MAPPO(cfg, 2, 1, "Learner")                      # a shard = 2 episodes, minibatches of 1 episode
    buf = {k[4:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("buf.")}
    return fx, m, buf


def _load_weights(m, fx):
    for net, mod in (("actor", m.actor), ("critic", m.critic)):
        mod.load_state_dict({k[len("w." + net + "."):]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("w." + net + ".")})


class _Big:
    def __init__(self, d):
        self.buffer = d

    def get_training_data(self, device):
        return {k: v.to(device) for k, v in self.buffer.items()}


def _shard(buf, rank):
    return {k: v[2 * rank: 2 * rank + 2].contiguous() for k, v in buf.items()}


def _worker(rank, world_size, port, backend, mode, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world_size, **(dict(device_id=dev) if backend == "nccl" else {}))
    try:
        fx, m, buf = _setup(dev)
        if rank == 0:
            _load_weights(m, fx)                          # rank 1 keeps its own random initial weights until the broadcast
        else:
            torch.manual_seed(1234)
            with torch.no_grad():
                m.ac_optimizer.flat_param.add_(0.01 * torch.randn_like(m.ac_optimizer.flat_param))
        m.sync_weights(0)
        steps = int(fx["total_steps"])
        m.train(_Big(_shard(buf, rank)), steps, return_numpy=False, allreduce=mode)
        m.update(steps)
        torch.cuda.synchronize()
        out[rank] = (m.ac_optimizer.flat_param.cpu().numpy(), m.ac_optimizer.flat_grad.cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", [None, "minibatch"], ids=["allreduce_per_update", "allreduce_per_minibatch"])
def test_two_rank_update_equals_single_process_sum(mode):
    import torch.multiprocessing as mp
    from distributed_multi_agent_reinforcement_learning_b200 import policy_ops as ops
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), backend, mode, out), nprocs=2, join=True)
    (p0, g0), (p1, g1) = out[0], out[1]
    assert np.array_equal(p0, p1), "replicas diverged"
    assert np.array_equal(g0, g1), "replicas hold different reduced gradients"

    # single process: both shards' gradients, summed the way the protocol says
    fx, m, buf = _setup(torch.device("cuda", 0))
    _load_weights(m, fx)
    u0, v0 = m.critic.Mean.weight_u.clone(), m.critic.Mean.weight_v.clone()
    steps = int(fx["total_steps"])
    arenas, per_mb = [], []
    for r in range(2):
        m.critic.Mean.weight_u.copy_(u0)                  # each replica starts its epoch from the broadcast power-iteration state
        m.critic.Mean.weight_v.copy_(v0)
        trace = {}
        m.train(_Big(_shard(buf, r)), steps, return_numpy=False, trace=trace)
        arenas.append(m.ac_optimizer.flat_grad.clone())
        per_mb.append(trace["mb_grads"])
    flat = m.ac_optimizer.flat_grad
    if mode is None:                                      # main.py:121-126: sum of the per-learner accumulated-and-clipped gradients
        flat.copy_(arenas[0] + arenas[1])
    else:                                                 # per minibatch: reduce, accumulate, clip the running sum
        flat.zero_()
        for mb_i in range(per_mb[0].shape[0]):
            flat.add_(per_mb[0][mb_i] + per_mb[1][mb_i])
            ops.clip_grad_norm_(flat, 5.0)
    np.testing.assert_allclose(g0, flat.cpu().numpy(), rtol=1e-6, atol=1e-9)
    m.ac_optimizer.step()
    np.testing.assert_allclose(p0, m.ac_optimizer.flat_param.cpu().numpy(), rtol=1e-6, atol=1e-6)
    if mode == "minibatch" and float(torch.linalg.norm(per_mb[0][0] + per_mb[1][0])) > 5.0:
        # the clip acted on the GLOBAL gradient of the first minibatch: not the per-update result
        assert not np.allclose(g0, (arenas[0] + arenas[1]).cpu().numpy(), rtol=1e-4, atol=1e-8)
