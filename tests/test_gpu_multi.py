"""GPU, 2 ranks over NCCL (skipped on a 1-GPU box): episode-sharded inference equals the unsharded result;
one sharded training step equals the weighted average of per-shard oracle gradients."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import gnn_pathplanning_b200 as gp
        from gnn_pathplanning_b200 import sharding, synthetic
        from oracle import planner_oracle as po

        class Cfg:
            num_agents, nGraphFilterTaps, device = 5, 3, torch.device("cuda", rank)

        N, K, B = 5, 3, 13
        sd = po.init_state_dict(K, seed=9)
        po.randomize_bn_stats(sd, seed=1)
        x, S = synthetic.make_batch(B, N, 12, seed=3)
        tgt = synthetic.random_targets(B, N, seed=8)
        xt, St, tt = torch.from_numpy(x), torch.from_numpy(S), torch.from_numpy(tgt)
        m = gp.DecentralPlannerNet(Cfg())
        m.load_state_dict(sd)
        m = m.cuda()
        lo, hi = sharding.shard_range(B, rank, world)
        # --- inference: shard, run, gather; no collective on the data path
        m.eval()
        with torch.no_grad():
            m.addGSO(St[lo:hi].cuda())
            local = torch.stack(m(xt[lo:hi].cuda()))
            full = sharding.gather_logits(local, B)
            ref = torch.stack(po.planner_forward(sd, St, xt)).numpy()
        err_inf = float(np.abs(full.cpu().numpy() - ref).max() / np.abs(ref).max())
        # --- training: one flat weighted all-reduce; reference = weighted mean of per-shard oracle gradients
        m.train()
        bucket = sharding.GradientBucket(m)
        opt = torch.optim.SGD(m.parameters(), lr=0.0)
        sharding.train_step(m, opt, bucket, xt[lo:hi].cuda(), St[lo:hi].cuda(), tt[lo:hi].cuda(), B)
        want = None
        for r in range(world):
            l2, h2 = sharding.shard_range(B, r, world)
            leaf = {k: (v.clone().requires_grad_(True) if (v.is_floating_point() and "running" not in k) else v)
                    for k, v in sd.items()}
            bn = {k: v.clone() for k, v in sd.items() if "running" in k or "tracked" in k}
            loss = po.planner_loss(po.planner_forward(leaf, St[l2:h2], xt[l2:h2], True, bn), tt[l2:h2])
            loss.backward()
            g = torch.cat([leaf[n].grad.reshape(-1) for n, _ in m.named_parameters()]) * ((h2 - l2) / B)
            want = g if want is None else want + g
        got = bucket.flat.cpu()
        err_tr = float((got - want).abs().max() / want.abs().max())
        if rank == 0:
            q.put((err_inf, err_tr))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_inference_and_training_nccl():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    mp.spawn(_worker, args=(2, _free_port(), q), nprocs=2, join=True)
    err_inf, err_tr = q.get(timeout=10)
    assert err_inf <= 1e-5, err_inf
    assert err_tr <= 5e-5, err_tr
