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


def allreduce_grads(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20) -> None:
    """Average gradients over ranks in flat buckets (one collective per ~64 MB: NVSwitch makes
    cost latency-, not link-bound, so few large buckets)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    world = dist.get_world_size()
    grads = [p.grad for p in params if p.grad is not None]
    bucket, size = [], 0
    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat)
        flat.div_(world)
        off = 0
        for g in bucket:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
        bucket, size = [], 0
    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
