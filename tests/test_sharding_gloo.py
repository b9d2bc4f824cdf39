"""CPU, world_size 2 over gloo: the N>1 host logic (episode sharding, the single flat gradient
all-reduce, logits gathering)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnn_pathplanning_b200 import sharding


def test_shard_ranges_partition_the_batch():
    for B in (0, 1, 7, 64, 513):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(B, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                       # same replica on every rank
        model = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.ReLU(), torch.nn.Linear(8, 5))
        g = torch.Generator().manual_seed(1)
        X = torch.randn(B, 6, generator=g)
        Y = torch.randint(0, 5, (B,), generator=g)
        # full-batch gradient (what the single-process reference computes)
        model.zero_grad()
        torch.nn.functional.cross_entropy(model(X), Y).backward()
        want = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
        # sharded: local mean loss, one weighted flat all-reduce
        bucket = sharding.GradientBucket(model)
        bucket.zero()
        lo, hi = sharding.shard_range(B, rank, world)
        torch.nn.functional.cross_entropy(model(X[lo:hi]), Y[lo:hi]).backward()
        bucket.all_reduce(hi - lo, B)
        assert bucket.numel == want.numel()
        assert torch.allclose(bucket.flat, want, rtol=1e-5, atol=1e-6), (bucket.flat - want).abs().max()
        for p in model.parameters():               # p.grad are views of the flat buffer
            assert p.grad.data_ptr() >= bucket.flat.data_ptr()
        # logits gather: agent-major [N, b_r, 5] shards -> [N, B, 5]
        full = torch.arange(3 * B * 5, dtype=torch.float32).reshape(3, B, 5)
        got = sharding.gather_logits(full[:, lo:hi].contiguous(), B)
        assert torch.equal(got, full)
        xs, Ss = sharding.shard_batch(torch.zeros(B, 2), torch.ones(B, 3), rank, world)
        assert xs.shape[0] == hi - lo and Ss.shape[0] == hi - lo
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7])
def test_flat_gradient_allreduce_and_gather_world2(B):
    mp.spawn(_worker, args=(2, _free_port(), B), nprocs=2, join=True)
