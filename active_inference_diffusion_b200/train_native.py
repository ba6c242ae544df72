"""Native forward / backward / double backward of the score network's trunk for the training loss
(`compute_diffusion_elbo`, core/active_inference.py:584-606 and :709-729).

`trunk(net, z, cond, time_weight, folds)` returns `(s, g)`:
    s = s_theta(z | cond)                     models/score_networks.py:151-171
    g = d(sum s)/dz                           the gradient penalty's inner gradient (:717-723)
as ONE autograd node whose backward produces, for any loss L(s, g), the gradients of every trunk
parameter, of the conditioning embedding `cond` and of z (score-matching stream only: the reference
detaches the penalty's input, :711).  Forward, the inner VJP, the adjoint of that VJP and the
backward through the forward graph are four passes of `libaid_sm100*.so` over one workspace
(csrc/train.inc: aid_dsm_forward, aid_gp_forward_backward, aid_dsm_backward); the derivation is
written out and checked against autograd in oracle/manual_score_grad.py /
tests/test_manual_score_grad.py.

The caller keeps the conditioning path (time embeddings, observation encoder) and the fold
W_f = W_o W_v in torch autograd; their gradients arrive through `cond` / the folded tensors.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib

# TrainParam / TrainBlockParam order of csrc/train.inc
HEAD_PARAMS = 8
BLOCK_PARAMS = 6


def supported(net) -> bool:
    """Dims the native training kernels cover (everything the reference instantiates: H in
    {128, 256, 512}, L in {32, 64, 128})."""
    return bool(getattr(net, "use_attention", False)) and net.hidden_dim % 128 == 0 and net.latent_dim % 8 == 0


def _dims(net) -> _lib.AidScoreDims:
    return _lib.AidScoreDims(net.latent_dim, net.observation_dim, net.hidden_dim, net.time_embed_dim, net.num_blocks)


def trunk_parameters(net, folds) -> List[torch.Tensor]:
    """The trunk's weights in TrainParam order; the S = 2*blocks+1 adaLN modulation Linears are
    concatenated in forward order (differentiable views: autograd splits the gradients back)."""
    mods = [m for blk in net.transformer_blocks for m in (blk.norm1, blk.norm2)] + [net.norm_final]
    w_mod = torch.cat([m.adaLN_modulation[1].weight for m in mods], dim=0)
    b_mod = torch.cat([m.adaLN_modulation[1].bias for m in mods], dim=0)
    out = [net.latent_proj.weight, net.latent_proj.bias, net.output_proj[0].weight, net.output_proj[0].bias,
           net.output_proj[2].weight, net.output_multiplier, w_mod, b_mod]
    for blk, (w_f, b_f) in zip(net.transformer_blocks, folds):
        out += [w_f, b_f, blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight, blk.mlp[2].bias]
    return out


def _table(tensors: Sequence[torch.Tensor]):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


KEEP_LAST = False      # tests: keep a reference to the last (packed weights, workspace) pair


class _Trunk(torch.autograd.Function):
    last = None

    @staticmethod
    def forward(ctx, dims, operand, z, cond, tw, *params):
        dev = _lib.require_cuda(z, cond, tw, *params)
        l = _lib.lib(operand)
        B = z.shape[0]
        z, cond = _lib.f32c(z.detach()), _lib.f32c(cond.detach())
        tw = None if tw is None else _lib.f32c(tw.detach().reshape(-1))
        keep = [_lib.f32c(p.detach()) for p in params]
        n = l.aid_train_num_params(ctypes.byref(dims))
        if n != len(keep):
            raise RuntimeError(f"trunk: expected {n} parameter tensors, got {len(keep)}")
        pbytes = l.aid_train_packed_bytes(ctypes.byref(dims))
        wbytes = l.aid_train_workspace_bytes(ctypes.byref(dims), B)
        if pbytes == 0 or wbytes == 0:
            _lib.check(-1, "aid_train_*_bytes", l)
        packed = torch.empty(pbytes, dtype=torch.uint8, device=dev)
        ws = torch.empty(wbytes, dtype=torch.uint8, device=dev)
        s = torch.empty(B, dims.latent_dim, dtype=torch.float32, device=dev)
        g = torch.empty_like(s)
        st = _lib.stream_ptr(dev)
        with torch.cuda.device(dev):
            _lib.check(l.aid_train_pack(ctypes.byref(dims), _table(keep), len(keep), packed.data_ptr(), pbytes, st),
                       "aid_train_pack", l)
            _lib.check(l.aid_dsm_forward(ctypes.byref(dims), packed.data_ptr(), ws.data_ptr(), wbytes, B, z.data_ptr(),
                                         cond.data_ptr(), _lib.ptr(tw), s.data_ptr(), st), "aid_dsm_forward", l)
            _lib.check(l.aid_gp_forward_backward(ctypes.byref(dims), packed.data_ptr(), ws.data_ptr(), wbytes, B, 0,
                                                 _lib.ptr(tw), g.data_ptr(), None, None, None, st),
                       "aid_gp_forward_backward(phase 0)", l)
        ctx.dims, ctx.operand, ctx.lib = dims, operand, l
        ctx.packed, ctx.ws, ctx.shapes = packed, ws, [tuple(p.shape) for p in params]
        ctx.save_for_backward(cond, tw if tw is not None else torch.empty(0, device=dev))
        ctx.has_tw = tw is not None
        if KEEP_LAST:
            _Trunk.last = (packed, ws)
        return s, g

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, s_bar, g_bar):
        cond, tw = ctx.saved_tensors
        tw = tw if ctx.has_tw else None
        l, dims = ctx.lib, ctx.dims
        dev = cond.device
        B = cond.shape[0]
        s_bar = torch.zeros(B, dims.latent_dim, device=dev) if s_bar is None else _lib.f32c(s_bar)
        with_penalty = g_bar is not None
        nb = dims.num_blocks
        # One flat gradient buffer, laid out stage by stage of aid_dsm_backward (output head, blocks from
        # the last to the first, latent_proj + modulation Linear): a stage's gradients are contiguous,
        # so the data-parallel all-reduce of a finished stage is ONE collective on a slice, with no
        # staging copies; the tensors handed to autograd are views of it.
        order = [2, 3, 4, 5] + [HEAD_PARAMS + BLOCK_PARAMS * i + j for i in reversed(range(nb)) for j in range(BLOCK_PARAMS)] \
            + [0, 1, 6, 7]
        sizes = [int(torch.Size(ctx.shapes[i]).numel()) for i in order]
        padded = [(n + 3) // 4 * 4 for n in sizes]            # every tensor starts 16-byte aligned (128-bit stores)
        flat = torch.zeros(sum(padded), dtype=torch.float32, device=dev)
        grads, off, bounds = [None] * len(order), 0, {}
        for i, n, m in zip(order, sizes, padded):
            grads[i] = flat[off:off + n].view(ctx.shapes[i])
            bounds[i] = (off, off + m)
            off += m
        stage_slices = [(bounds[2][0], bounds[5][1])] + \
            [(bounds[HEAD_PARAMS + BLOCK_PARAMS * i][0], bounds[HEAD_PARAMS + BLOCK_PARAMS * i + BLOCK_PARAMS - 1][1])
             for i in reversed(range(nb))] + [(bounds[0][0], bounds[7][1])]
        table = _table(grads)
        need_dz = ctx.needs_input_grad[2]
        dz = torch.empty(B, dims.latent_dim, dtype=torch.float32, device=dev) if need_dz else None
        dcond = torch.empty_like(cond)
        st = _lib.stream_ptr(dev)
        wbytes = ctx.ws.numel()
        group = DATA_PARALLEL_GROUP if (DATA_PARALLEL_GROUP is not None and dist.is_initialized()
                                        and dist.get_world_size(DATA_PARALLEL_GROUP) > 1) else None

        def run(lo, hi):
            _lib.check(l.aid_dsm_backward(ctypes.byref(dims), ctx.packed.data_ptr(), ctx.ws.data_ptr(), wbytes, B,
                                          s_bar.data_ptr(), _lib.ptr(tw), cond.data_ptr(), int(with_penalty), table,
                                          _lib.ptr(dz), dcond.data_ptr(), lo, hi, st), "aid_dsm_backward", l)

        with torch.cuda.device(dev):
            if with_penalty:
                g_bar = _lib.f32c(g_bar)
                _lib.check(l.aid_gp_forward_backward(ctypes.byref(dims), ctx.packed.data_ptr(), ctx.ws.data_ptr(), wbytes,
                                                     B, 1, _lib.ptr(tw), None, g_bar.data_ptr(), s_bar.data_ptr(), table,
                                                     st), "aid_gp_forward_backward(phase 1)", l)
            if group is None:
                run(0, nb + 2)
            else:
                # gradient all-reduce overlapped with the remaining stages: NCCL on a side stream that
                # waits for the stage just enqueued; the main stream joins before the gradients are used
                main = torch.cuda.current_stream(dev)
                side = _side_stream(dev)
                world = dist.get_world_size(group)
                for stage, (a, b) in enumerate(stage_slices):
                    run(stage, stage + 1)
                    side.wait_stream(main)
                    with torch.cuda.stream(side):
                        dist.all_reduce(flat[a:b], group=group)
                main.wait_stream(side)
                flat.mul_(1.0 / world)
        ctx.ws = ctx.packed = None           # the saved activations are dead: release them to the allocator
        return (None, None, dz, dcond, None) + tuple(grads)


# Data-parallel training: when set (train_graph.GraphedElboStep does it under torchrun), the trunk's
# backward all-reduces (averages) its own gradients stage by stage, overlapped with the computation of
# the following stages.  The parameters covered are `reduced_parameters(net)`; the caller all-reduces
# the rest (conditioning path, diffusion parameters) after backward.
DATA_PARALLEL_GROUP = None
_SIDE = {}


def _side_stream(dev: torch.device) -> torch.cuda.Stream:
    s = _SIDE.get(dev)
    if s is None:
        s = _SIDE[dev] = torch.cuda.Stream(device=dev)
    return s


def reduced_parameters(net) -> List[torch.nn.Parameter]:
    """Parameters whose gradients come exclusively through the trunk node (and are therefore already
    averaged over ranks when DATA_PARALLEL_GROUP is set): latent_proj, the DiT blocks (incl. the
    attention tensors reached through the fold and the adaLN modulations), norm_final, output_proj,
    output_multiplier."""
    out = list(net.latent_proj.parameters()) + list(net.transformer_blocks.parameters()) + \
        list(net.norm_final.parameters()) + list(net.output_proj.parameters()) + [net.output_multiplier]
    return out


def trunk(net, z: torch.Tensor, cond: torch.Tensor, time_weight: Optional[torch.Tensor], folds,
          operand: Optional[str] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(s, g) of the score net's z-dependent part; `cond` = time embedding + observation embedding
    (pre-SiLU, [B,H]), `time_weight` [B,1] for the continuous-time branch or None, `folds` from
    `autograd_path.fold_attention`.  `operand`: "f16" (default: TF32-class, the rel-1e-3 contract) or
    "bf16"."""
    if not supported(net):
        raise ValueError("native training kernels need hidden_dim % 128 == 0 and latent_dim % 8 == 0")
    operand = operand or "f16"
    if operand not in _lib.LIB_PATHS:
        raise ValueError(f"unknown operand type {operand!r}")
    return _Trunk.apply(_dims(net), operand, z, cond, time_weight, *trunk_parameters(net, folds))


# ---------------------------------------------------------------------------------------------
# test / debug access to the saved tensors of a workspace (layout: csrc/train.inc, train_ws_layout)
def untile(ws: torch.Tensor, offset: int, rows: int, cols: int) -> torch.Tensor:
    """tiled fp32 [row tile][cols/4][128 rows] float4 at byte `offset` -> [rows, cols]."""
    rt, cp = (rows + 127) // 128, (cols + 127) // 128 * 128
    v = ws[offset:offset + rt * cp * 128 * 4].view(torch.float32).view(rt, cp // 4, 128, 4)
    return v.permute(0, 2, 1, 3).reshape(rt * 128, cp)[:rows, :cols]


def unpack(ws: torch.Tensor, offset: int, rows: int, cols: int, operand: str, kb: Optional[int] = None) -> torch.Tensor:
    """packed 16-bit operand [row tile][64-col block][8-col chunk][128 rows][8] -> fp32 [rows, cols]."""
    rt = (rows + 127) // 128
    kb = kb or (cols + 63) // 64
    dt = torch.float16 if operand == "f16" else torch.bfloat16
    v = ws[offset:offset + rt * kb * 16384].view(dt).view(rt, kb, 8, 128, 8)
    return v.permute(0, 3, 1, 2, 4).reshape(rt * 128, kb * 64)[:rows, :cols].float()


def debug_offset(net, batch: int, name: str, index: int = 0, operand: str = "f16") -> int:
    d = _dims(net)
    off = _lib.lib(operand).aid_train_debug_offset(ctypes.byref(d), batch, name.encode(), index)
    if off < 0:
        raise KeyError(name)
    return int(off)
