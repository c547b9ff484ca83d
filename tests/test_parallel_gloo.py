"""world_size-2 `gloo` tests (CPU) of the data-parallel plumbing: env sharding without a data-path collective, rank-0
weight broadcast, and the SUM all-reduce of the flat gradient arena followed by identical optimizer steps."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from distributed_multi_agent_reinforcement_learning_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        # 1. sharding covers every env exactly once
        lo, hi = parallel.shard_range(4097)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world_size)]
        dist.all_gather(sizes, torch.tensor([hi - lo]))
        assert sum(int(s) for s in sizes) == 4097 and max(int(s) for s in sizes) - min(int(s) for s in sizes) <= 1
        # 2. replicas start from rank 0's weights
        torch.manual_seed(100 + rank)
        w = torch.randn(1000)
        parallel.broadcast_(w, src=0)
        torch.manual_seed(100)
        assert torch.equal(w, torch.randn(1000))
        # 3. gradient arena: SUM (reference semantics), then the same Adam step everywhere
        torch.manual_seed(7 + rank)
        g_local = torch.randn(1000)
        g = g_local.clone()
        parallel.allreduce_sum_(g)
        torch.manual_seed(7)
        expect = torch.randn(1000)
        torch.manual_seed(8)
        expect = expect + torch.randn(1000)
        assert torch.allclose(g, expect)
        p = torch.nn.Parameter(w.clone())
        p.grad = g
        torch.optim.Adam([p], lr=5e-4, eps=1e-5).step()
        gathered = [torch.zeros(1000) for _ in range(world_size)]
        dist.all_gather(gathered, p.data)
        assert torch.equal(gathered[0], gathered[1])
        gm = g_local.clone()
        parallel.allreduce_sum_(gm, mean=True)
        assert torch.allclose(gm, expect / world_size)
        assert parallel.max_over_ranks(rank + 1.5, "cpu") == world_size + 0.5
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_single_process_is_identity():
    assert parallel.shard_range(10) == (0, 10)
    assert parallel.shard_range(10, 1, 3) == (4, 7) and parallel.shard_range(10, 2, 3) == (7, 10)
    t = torch.ones(4)
    assert parallel.allreduce_sum_(t) is t and torch.equal(t, torch.ones(4))
