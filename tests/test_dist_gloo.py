"""world_size-2 gloo tests of the data-parallel host logic (CPU only)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from point2cyl_b200 import dist as pd
from point2cyl_b200 import synthetic


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_partitions_exactly():
    for total in (1, 7, 32, 33, 256):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = pd.shard_range(total, r, world)
                assert 0 <= lo <= hi <= total
                seen += list(range(lo, hi))
            assert seen == list(range(total))
            sizes = [pd.shard_range(total, r, world)[1] - pd.shard_range(total, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w, _ = pd.init("gloo")
    assert (r, w) == (rank, world)
    # 1) shards of one global batch tile it exactly
    batch = synthetic.s_cyl(5, 64, 4, seed=9)
    mine = pd.shard_batch(batch, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine["pcs"])
    assert torch.equal(torch.cat(gathered), batch["pcs"])
    # 2) timing reduction: max over ranks
    assert pd.reduce_max(10.0 + rank) == 10.0 + world - 1
    assert pd.reduce_sum(1.0) == float(world)
    # 3) flat gradient all-reduce == gradient of the mean loss over the global batch
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(3, 8), torch.nn.ReLU(), torch.nn.Linear(8, 2))
    flat = pd.FlatGradAllReduce(net.parameters())
    x = batch["pcs"].reshape(-1, 3)
    lo, hi = pd.shard_range(x.shape[0], rank, world)       # equal shard sizes: mean of means == global mean
    flat.zero()
    net(x[lo:hi]).pow(2).mean().backward()
    flat.sync()
    ref = torch.nn.Sequential(torch.nn.Linear(3, 8), torch.nn.ReLU(), torch.nn.Linear(8, 2))
    ref.load_state_dict(net.state_dict())
    ref(x).pow(2).mean().backward()
    for p, q in zip(net.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-6), (p.grad - q.grad).abs().max()
    # 4) buffers follow rank 0
    bn = torch.nn.BatchNorm1d(4)
    bn.running_mean.fill_(float(rank + 1))
    pd.broadcast_buffers(bn, 0)
    assert float(bn.running_mean[0]) == 1.0
    pd.barrier()
    dist.destroy_process_group()
    out.put(rank)


@pytest.mark.timeout(120)
def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    assert sorted(q.get() for _ in range(2)) == [0, 1]
