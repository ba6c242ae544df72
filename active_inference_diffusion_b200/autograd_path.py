"""Differentiable torch-op evaluation of the score network, used ONLY to build the autograd graph
of the training loss (`compute_diffusion_elbo`: backward and the gradient penalty's double
backward, core/active_inference.py:584-606,709-729).

Status ("training path"): sampling, score forward and EFE run on the hand-written
sm_100a kernels; the training loss still differentiates through these torch ops (cuBLAS on the
device).  Native fwd/bwd kernels are SURVEY §8 row a9/a10, next round.  Nothing on the sampling or
EFE path imports this module's `score_forward`.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F


def _sinusoid(time: torch.Tensor, dim: int, freq_scale: torch.Tensor) -> torch.Tensor:
    half = dim // 2
    k = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half, device=time.device) * -k) * freq_scale
    arg = time[:, None] * freqs[None, :]
    return torch.cat((arg.sin(), arg.cos()), dim=-1)


def _time_embed(net, time: torch.Tensor) -> torch.Tensor:
    e = _sinusoid(time, net.time_embed_dim, net.time_embed[0].freq_scale)
    return net.time_embed[3](F.silu(net.time_embed[1](e)))


def _ada_ln(mod, x: torch.Tensor, cond: torch.Tensor) -> torch.Tensor:
    scale, shift = mod.adaLN_modulation(cond).chunk(2, dim=-1)
    return F.layer_norm(x, (x.shape[-1],), None, None, 1e-5) * (1 + scale) + shift


def score_forward(net, z_t: torch.Tensor, time: torch.Tensor,
                  observation: Optional[torch.Tensor] = None) -> torch.Tensor:
    """models/score_networks.py:101-171 with torch ops (eval-mode obs_encoder: Dropout = identity,
    the same contract as the fused path)."""
    H = net.hidden_dim
    continuous = bool(time.max() <= 1.0 and time.min() >= 0.0)
    if continuous:
        t_emb = _time_embed(net, time * 999.0) + net.time_scale * net.continuous_time_embed(2.0 * time.view(-1, 1) - 1.0)
        time_weight = torch.sqrt(1.0 / (1e-5 + time.view(-1, 1)))
    else:
        t_emb, time_weight = _time_embed(net, time), None
    if observation is not None:
        enc = net.obs_encoder
        o = F.silu(enc[1](enc[0](observation)))
        o = F.silu(enc[5](enc[4](o)))
        o = enc[8](enc[7](o))
    else:
        o = torch.zeros(z_t.shape[0], H, device=z_t.device)
    cond = t_emb + o
    h = net.latent_proj(z_t)
    for blk in net.transformer_blocks:
        att = blk.attention
        x = _ada_ln(blk.norm1, h, cond)
        v = F.linear(x, att.in_proj_weight[2 * H:], att.in_proj_bias[2 * H:])   # seq-len-1 attention == out(V x)
        h = h + att.out_proj(v)
        h = h + blk.mlp(_ada_ln(blk.norm2, h, cond))
    s = net.output_proj(_ada_ln(net.norm_final, h, cond))
    s = torch.clamp(s, min=-10, max=10) * net.output_multiplier
    return s * time_weight if time_weight is not None else s


class EMALogMeanExp(torch.autograd.Function):
    """log(mean(exp(x))) with the MINE running-mean gradient (core/active_inference.py:815-826)."""

    @staticmethod
    def forward(ctx, x, running_mean):
        ctx.save_for_backward(x, running_mean)
        return x.exp().mean().log()

    @staticmethod
    def backward(ctx, grad_output):
        x, running_mean = ctx.saved_tensors
        return grad_output * x.exp().detach() / (running_mean + 1e-6) / x.shape[0], None
