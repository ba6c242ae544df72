"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row sharding + gather, global argmin
over ranks, gradient all-reduce equivalence (N-rank mean == single-process mean)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from active_inference_diffusion_b200.distributed import (allreduce_grads, gather_rows, global_argmin, shard_bounds,
                                                             shard_rows)
    g = torch.Generator().manual_seed(0)
    total = 1001
    efe = torch.randn(total, generator=g)
    efe[[17, 900]] = -9.0                      # tie across ranks -> lowest global index wins
    rows = torch.randn(total, 5, generator=g)
    lo, hi = shard_bounds(total, rank, world)
    local = shard_rows(rows, rank, world)
    back = gather_rows(local * 2.0, total)
    idx, val = global_argmin(efe[lo:hi], lo)
    # data-parallel gradient: each rank differentiates its shard's mean loss
    w = torch.nn.Parameter(torch.ones(5))
    ((local @ w) ** 2).mean().backward()
    w.grad.mul_((hi - lo) * world / total)     # weight shards by size so the average equals the global mean
    allreduce_grads([w])
    q.put((rank, torch.equal(back, rows * 2.0), idx, val, w.grad.clone()))
    dist.destroy_process_group()


def test_two_rank_sharding_argmin_and_grad_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    efe = torch.randn(1001, generator=g)
    efe[[17, 900]] = -9.0
    rows = torch.randn(1001, 5, generator=g)
    w = torch.nn.Parameter(torch.ones(5))
    ((rows @ w) ** 2).mean().backward()
    for rank, gathered_ok, idx, val, grad in out:
        assert gathered_ok
        assert idx == int(torch.argmin(efe)) == 17 and val == -9.0
        assert torch.allclose(grad, w.grad, rtol=1e-5, atol=1e-6)
