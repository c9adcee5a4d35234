"""CPU, world_size 2, gloo: the data-parallel gradient exchange (rave_b200/ddp.py) -- host logic only."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, bucket_bytes, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rave_b200 import ddp
        torch.manual_seed(1234 + rank)                      # ranks start different
        model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3), torch.nn.Linear(3, 1))
        model.register_buffer("buf", torch.full((4,), float(rank)))
        ddp.broadcast_module(model)
        flat0 = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        gathered = [torch.zeros_like(flat0) for _ in range(world)]
        dist.all_gather(gathered, flat0)
        assert all(torch.equal(g, gathered[0]) for g in gathered), "broadcast_module did not sync params"
        assert torch.equal(model.buf, torch.zeros(4)), "buffers follow rank 0"
        # per-rank gradients g_r = (rank + 1) * ones ; expected average = (1 + 2) / 2
        for p in model.parameters():
            p.grad = torch.full_like(p, float(rank + 1))
        list(model.parameters())[2].grad = None             # unused parameter: skipped, stays None
        red = ddp.GradientAllReducer(bucket_bytes=bucket_bytes)
        red(list(model.parameters()))
        for i, p in enumerate(model.parameters()):
            if i == 2:
                assert p.grad is None
            else:
                assert torch.allclose(p.grad, torch.full_like(p, 1.5)), (i, p.grad)
        # buffers that moved on each rank (the RVQ's EMA state) follow rank 0
        model.buf.fill_(float(10 + rank))
        assert ddp.broadcast_buffers(model) == 4
        assert torch.equal(model.buf, torch.full((4,), 10.0))
        n_with_grad = sum(p.numel() for i, p in enumerate(model.parameters()) if i != 2)
        assert red.bytes_reduced == 4 * n_with_grad
        ret[rank] = red.n_collectives
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bucket_bytes", [64 << 20, 64])
def test_gradient_allreduce_world2(bucket_bytes):
    world = 2
    port = _free_port()
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, bucket_bytes, ret), nprocs=world, join=True)
    assert len(ret) == world
    if bucket_bytes == 64:
        assert ret[0] > 1          # tiny buckets -> several collectives
    else:
        assert ret[0] == 1


def test_single_process_is_noop():
    from rave_b200 import ddp
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    ddp.GradientAllReducer()([p])
    assert torch.equal(p.grad, torch.full((3,), 2.0))
