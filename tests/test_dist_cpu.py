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


def _worker_plumbing(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from active_inference_diffusion_b200.distributed import FlatGrads, allreduce_mean, gather_time_loss
    from oracle import restatement as R
    g = torch.Generator().manual_seed(3)
    B = 64
    t_all, loss_all = torch.rand(world * B, generator=g), torch.rand(world * B, generator=g) * 3
    t, loss = t_all[rank * B:(rank + 1) * B], loss_all[rank * B:(rank + 1) * B]
    # (t, loss) all-gather in rank order -> the sequential time-importance EMA of the GLOBAL batch
    tg, lg = gather_time_loss(t, loss)
    w = R.update_time_importance(torch.ones(100), tg, lg)
    # scalar all-reduce: mean t of the global batch (KL weight) from the local means
    t_mean = allreduce_mean(t.mean())
    # flat gradient buffer: autograd accumulates into the views, one collective, no staging copies
    lin = torch.nn.Linear(5, 3)
    torch.manual_seed(0)
    with torch.no_grad():
        for p in lin.parameters():
            p.copy_(torch.randn(p.shape, generator=torch.Generator().manual_seed(1)))
    fg = FlatGrads(lin.parameters())
    fg.zero_and_attach()
    x = torch.randn(world * 16, 5, generator=torch.Generator().manual_seed(2))[rank * 16:(rank + 1) * 16]
    (lin(x) ** 2).mean().backward()
    assert lin.weight.grad.data_ptr() == fg.views[0].data_ptr()      # accumulated in place, still a view
    fg.allreduce()
    # sharded MINE statistic: one 3-float all-reduce = the statistic of the concatenated batch
    from active_inference_diffusion_b200.distributed import sharded_mine_statistic
    gj = torch.Generator().manual_seed(9)
    tj_all, tm_all = torch.randn(world * 40, 1, generator=gj), torch.randn(world * 40, 1, generator=gj)
    mi, joint, marg, t_exp = sharded_mine_statistic(tj_all[rank * 40:(rank + 1) * 40], tm_all[rank * 40:(rank + 1) * 40])
    # the same statistic from the per-rank partials the fused kernel writes (aid_epistemic_forward):
    # (sum T_joint, max T_marg, sum exp(T_marg - max), count) -> scalar MAX + 3-double SUM all-reduce
    from active_inference_diffusion_b200.distributed import merge_mine_partials
    tj, tm = tj_all[rank * 40:(rank + 1) * 40].double().reshape(-1), tm_all[rank * 40:(rank + 1) * 40].double().reshape(-1)
    part = torch.stack([tj.sum(), tm.max(), torch.exp(tm - tm.max()).sum(), torch.tensor(40.0, dtype=torch.float64)])
    mi2, joint2, marg2, _ = merge_mine_partials(part)
    assert abs(float(mi2) - float(mi)) < 1e-6 and abs(float(joint2) - float(joint)) < 1e-6 and abs(float(marg2) - float(marg)) < 1e-6
    q.put((rank, torch.equal(tg, t_all) and torch.equal(lg, loss_all), w, float(t_mean), fg.flat.clone(),
           (float(mi), float(joint), float(marg))))
    dist.destroy_process_group()


def test_two_rank_training_plumbing_time_importance_gather_scalar_mean_flat_grads():
    """the (t_b, loss_b) all-gather gives every rank the identical time-importance EMA,
    the mean-t scalar all-reduce equals the global mean, and the flat-buffer gradient exchange equals the
    single-process gradient of the concatenated batch."""
    from oracle import restatement as R
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_plumbing, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(3)
    t_all, loss_all = torch.rand(world * 64, generator=g), torch.rand(world * 64, generator=g) * 3
    w_want = R.update_time_importance(torch.ones(100), t_all, loss_all)
    lin = torch.nn.Linear(5, 3)
    with torch.no_grad():
        for p in lin.parameters():
            p.copy_(torch.randn(p.shape, generator=torch.Generator().manual_seed(1)))
    x = torch.randn(world * 16, 5, generator=torch.Generator().manual_seed(2))
    (lin(x) ** 2).mean().backward()
    want = torch.cat([lin.weight.grad.reshape(-1), lin.bias.grad.reshape(-1)])
    gj = torch.Generator().manual_seed(9)
    tj_all, tm_all = torch.randn(world * 40, 1, generator=gj), torch.randn(world * 40, 1, generator=gj)
    marg_want = float(tm_all.double().exp().mean().log())
    for rank, gathered_ok, w, t_mean, flat, mine in out:
        assert abs(mine[1] - float(tj_all.mean())) < 1e-6 and abs(mine[2] - marg_want) < 1e-6
        assert abs(mine[0] - (float(tj_all.mean()) - marg_want)) < 1e-6
        assert gathered_ok
        assert torch.equal(w, w_want)                      # bit-identical EMA on every rank
        assert abs(t_mean - float(t_all.mean())) < 1e-6
        assert torch.allclose(flat, want, rtol=1e-5, atol=1e-6)
