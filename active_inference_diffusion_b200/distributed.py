"""Multi-GPU plumbing: one process per GPU (torchrun), rows sharded across ranks.

Sampling, EFE scoring and the belief update are row-independent, so the data path has
NO collective: each rank scores its slice.  The only exchanges are
  * `global_argmin`   — all-gather of (min EFE, global row index) pairs, 8 bytes per rank;
  * `allreduce_grads` — sum/world of the score-net + diffusion parameter gradients for the
                        data-parallel training step (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced row range of `rank` (first `total % world` ranks get one extra row)."""
    base, extra = divmod(total_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def gather_rows(local: torch.Tensor, total_rows: int) -> torch.Tensor:
    """All-gather ragged row shards back into the global order (used by tests / final reporting)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(total_rows, r, world) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


def global_argmin(local_efe: torch.Tensor, row_offset: int) -> Tuple[int, float]:
    """Index (in the global row order) and value of the minimum EFE over all ranks.  Ties resolve to
    the lowest global index, as torch.argmin on the concatenated tensor does."""
    idx = torch.argmin(local_efe)
    pair = torch.stack([local_efe[idx].double(), (idx + row_offset).double()])
    if dist.is_initialized() and dist.get_world_size() > 1:
        pairs = [torch.empty_like(pair) for _ in range(dist.get_world_size())]
        dist.all_gather(pairs, pair)
        pairs = torch.stack(pairs)
    else:
        pairs = pair.unsqueeze(0)
    best = pairs[:, 0].min()
    cand = pairs[pairs[:, 0] == best]
    return int(cand[:, 1].min().item()), float(best.item())


class FlatGrads:
    """Persistent flat gradient buffer for a set of parameters: `.grad` of every parameter is a view
    into one contiguous fp32 tensor, so the data-parallel exchange is ONE all-reduce on that tensor with
    no staging copies (`torch.cat` / `copy_` back), and autograd accumulates straight into it."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def zero_and_attach(self) -> None:
        """Zero the buffer and make it the parameters' `.grad` (call before backward)."""
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v

    def allreduce(self, group=None) -> None:
        if self.flat.numel() == 0 or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return
        dist.all_reduce(self.flat, group=group)
        self.flat.mul_(1.0 / dist.get_world_size(group))


def allreduce_mean(t: torch.Tensor, group=None) -> torch.Tensor:
    """Average of a small device tensor over ranks (mean t of the KL weight, the loss metrics)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    out = t.detach().clone()
    dist.all_reduce(out, group=group)
    return out / dist.get_world_size(group)


def gather_time_loss(t: torch.Tensor, loss: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather of the per-sample (t_b, loss_b) pairs in rank order: every rank then applies the
    identical sequential time-importance EMA (core/active_inference.py:750-771) over the GLOBAL batch,
    exactly as one process would on the concatenated batch (equal shard sizes)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t, loss
    world = dist.get_world_size(group)
    pair = torch.stack([t.detach().float().reshape(-1), loss.detach().float().reshape(-1)], dim=1).contiguous()
    out = torch.empty(world * pair.shape[0], 2, dtype=pair.dtype, device=pair.device)
    dist.all_gather_into_tensor(out, pair, group=group)
    return out[:, 0].contiguous(), out[:, 1].contiguous()


def sharded_mine_statistic(t_joint: torch.Tensor, t_marg: torch.Tensor, group=None):
    """MINE statistic of a batch sharded over ranks (core/active_inference.py:1040-1053 is
    a batch-GLOBAL mean): every rank contributes (sum T_joint, sum exp(T_marg), count) and ONE 3-float
    all-reduce gives all ranks the statistic of the global batch:
        joint = sum T_joint / n,  t_exp = sum exp(T_marg) / n,  mi = joint - log(t_exp).
    Returns (mi, joint, log t_exp, t_exp) as 0-dim tensors.  The marginal permutation stays inside a
    rank's shard (documented deviation from the single-process randperm over the whole batch)."""
    stat = torch.stack([t_joint.detach().double().sum(), t_marg.detach().double().exp().sum(),
                        torch.tensor(float(t_joint.numel()), dtype=torch.float64, device=t_joint.device)])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stat, group=group)
    joint = stat[0] / stat[2]
    t_exp = stat[1] / stat[2]
    return (joint - t_exp.log()).float(), joint.float(), t_exp.log().float(), t_exp.float()


def merge_mine_partials(partial: torch.Tensor, group=None):
    """The same statistic from the per-rank partials `aid_epistemic_forward` writes (sum T_joint, max T_marg,
    sum exp(T_marg - max), N): the sums of exponentials are brought to the global maximum (one MAX
    all-reduce of a scalar) and then ONE 3-double SUM all-reduce follows.  Returns (mi, joint, log t_exp,
    t_exp) as 0-dim float tensors."""
    gmax = partial[1].clone()
    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    if multi:
        dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=group)
    stat = torch.stack([partial[0], partial[2] * torch.exp(partial[1] - gmax), partial[3]])
    if multi:
        dist.all_reduce(stat, group=group)
    joint = stat[0] / stat[2]
    log_t = gmax + torch.log(stat[1] / stat[2])
    return (joint - log_t).float(), joint.float(), log_t.float(), torch.exp(log_t).float()


def allreduce_grads(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, group=None) -> None:
    """Average the gradients of `params` over ranks.  Gradients that already live in one contiguous
    buffer (FlatGrads) go out as one collective; others are exchanged tensor by tensor, largest first
    (NVSwitch: the cost is latency-, not link-bound, and no staging copy is made)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    small = [g for g in grads if g.numel() * g.element_size() < (1 << 20)]
    large = [g for g in grads if g.numel() * g.element_size() >= (1 << 20)]
    for g in large:
        if g.is_contiguous():
            dist.all_reduce(g, group=group)
            g.mul_(1.0 / world)
        else:
            c = g.contiguous()
            dist.all_reduce(c, group=group)
            g.copy_(c.mul_(1.0 / world))
    if small:
        flat = torch.cat([g.reshape(-1) for g in small])          # < 1 MB each: one coalesced collective
        dist.all_reduce(flat, group=group)
        flat.mul_(1.0 / world)
        off = 0
        for g in small:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
