"""Differentiable evaluation of the score network for the training loss
(`compute_diffusion_elbo`: backward and the gradient penalty's double backward,
core/active_inference.py:584-606,709-729).

Every dense contraction of the graph -- forward x W^T, input gradient dY W, weight gradient
dY^T X, and the same three again inside the double backward of the gradient penalty -- is
`MatmulNT`, one autograd Function over the library's `aid_gemm_nt` (tcgen05, bf16 operands,
fp32 accumulation, split-K for weight gradients).  Its backward is expressed with `MatmulNT`
itself, so `create_graph=True` differentiates through it to any order without a second set of
formulas.  The element-wise glue between the GEMMs (LayerNorm, SiLU/GELU, adaLN modulation,
clamp) is memory-bound and stays on torch device ops, whose double-backward formulas autograd
already has (DESIGN.md "training path").
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F


from . import _lib

# Precision of the training-graph GEMMs: "bf16x3" (default; hi/lo operand split, products exact to
# ~2^-16, meets the rel-1e-3 loss/gradient contract) or "bf16" (3x fewer tensor-core FLOPs;
# measured gradient bound stated in DESIGN.md).
PRECISION = "bf16x3"


class precision:
    """`with precision("bf16"):` — GEMM precision for a region (e.g. inference-only estimators)."""

    def __init__(self, name: str):
        if name not in _lib.PRECISIONS:
            raise ValueError(f"unknown precision {name!r}")
        self.name = name

    def __enter__(self):
        global PRECISION
        self.prev, PRECISION = PRECISION, self.name

    def __exit__(self, *exc):
        global PRECISION
        PRECISION = self.prev


def set_precision(name: str) -> None:
    global PRECISION
    if name not in _lib.PRECISIONS:
        raise ValueError(f"unknown precision {name!r}; expected one of {sorted(_lib.PRECISIONS)}")
    PRECISION = name


class MatmulNT(torch.autograd.Function):
    """out[M,N] = a[M,K] @ b[N,K]^T through `aid_gemm_nt`.  d/da = g @ b, d/db = g^T @ a, both again
    MatmulNT on transposed views, hence differentiable to any order.  The backward GEMMs run at the
    precision the forward ran at (a `with precision(...)` region may have ended by then)."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        ctx.prec = PRECISION
        return _lib.gemm_nt(a, b, precision=PRECISION)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        with precision(ctx.prec):
            ga = MatmulNT.apply(g, b.t()) if ctx.needs_input_grad[0] else None
            gb = MatmulNT.apply(g.t(), a.t()) if ctx.needs_input_grad[1] else None
        return ga, gb


class ColSum(torch.autograd.Function):
    """g.sum(0) (bias gradients) through `aid_colsum`; its backward is a broadcast view, so the
    gradient penalty's double backward differentiates through it."""

    @staticmethod
    def forward(ctx, g):
        ctx.rows = g.shape[0]
        return _lib.colsum(g)

    @staticmethod
    def backward(ctx, gg):
        return gg.unsqueeze(0).expand(ctx.rows, gg.shape[0])


class LinearNT(torch.autograd.Function):
    """x W^T + b with the bias added in the GEMM epilogue; gradients again through MatmulNT."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.prec = PRECISION
        return _lib.gemm_nt(x, weight, bias, precision=PRECISION)

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        with precision(ctx.prec):
            gx = MatmulNT.apply(g, weight.t()) if ctx.needs_input_grad[0] else None
            gw = MatmulNT.apply(g.t(), x.t()) if ctx.needs_input_grad[1] else None
        gb = ColSum.apply(g) if ctx.needs_input_grad[2] else None
        return gx, gw, gb


class GeluBackward(torch.autograd.Function):
    """g * gelu'(x) (torch's fused kernel) whose own backward -- needed only by the gradient penalty's
    double backward -- is two kernels: gelu_backward(gg, x) and aid_gelu_double_backward.  Autograd's
    composite formula for the same derivative is ~10 element-wise kernels on [B, 4H] tensors."""

    @staticmethod
    def forward(ctx, g, x):
        ctx.save_for_backward(g, x)
        return torch.ops.aten.gelu_backward(g, x)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gg):
        g, x = ctx.saved_tensors
        return torch.ops.aten.gelu_backward(gg, x), _lib.gelu_double_backward(gg, g, x)


class Gelu(torch.autograd.Function):
    """Exact (erf) GELU, models/score_networks.py:199, differentiable twice through GeluBackward."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return F.gelu(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return GeluBackward.apply(g, x)


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """F.linear with the contraction (and the bias add) on the tcgen05 path."""
    if bias is None:
        return MatmulNT.apply(x, weight)
    return LinearNT.apply(x, weight, bias)


def _seq(mods, x: torch.Tensor) -> torch.Tensor:
    """nn.Sequential forward with every nn.Linear routed through `linear` (nested Sequentials too)."""
    for m in mods:
        if isinstance(m, torch.nn.Linear):
            x = linear(x, m.weight, m.bias)
        elif isinstance(m, torch.nn.Sequential):
            x = _seq(m, x)
        elif isinstance(m, torch.nn.GELU) and m.approximate == "none" and x.is_cuda:
            x = Gelu.apply(x)
        else:
            x = m(x)
    return x


seq = _seq


def _sinusoid(time: torch.Tensor, dim: int, freq_scale: torch.Tensor) -> torch.Tensor:
    half = dim // 2
    k = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half, device=time.device) * -k) * freq_scale
    arg = time[:, None] * freqs[None, :]
    return torch.cat((arg.sin(), arg.cos()), dim=-1)


def _time_embed(net, time: torch.Tensor) -> torch.Tensor:
    e = _sinusoid(time, net.time_embed_dim, net.time_embed[0].freq_scale)
    t1, t3 = net.time_embed[1], net.time_embed[3]
    return linear(F.silu(linear(e, t1.weight, t1.bias)), t3.weight, t3.bias)


def _ada_ln(x: torch.Tensor, scale1_shift: torch.Tensor) -> torch.Tensor:
    """LN(x) * (1 + scale) + shift as one fused multiply-add on top of the LayerNorm; the `1 +` is
    already inside `scale1_shift` (see _all_modulations)."""
    scale1, shift = scale1_shift.chunk(2, dim=-1)
    return torch.addcmul(shift, F.layer_norm(x, (x.shape[-1],), None, None, 1e-5), scale1)


def _all_modulations(net, cond: torch.Tensor):
    """Every adaLN modulation of the forward (2 per block + the final one) reads the SAME
    conditioning, so they are ONE GEMM: SiLU(cond) x [W_1; W_2; ...]^T.  One operand pack and one
    launch instead of 13 (and one input-gradient GEMM instead of 13 in the backward).  The `1 +` of
    `(1 + scale)` (models/score_networks.py:269) is added to the scale half of the concatenated bias
    (a 13,312-element op) instead of to 13 [B, H] tensors; gradients are unchanged (a constant)."""
    mods = [m for blk in net.transformer_blocks for m in (blk.norm1, blk.norm2)] + [net.norm_final]
    H = net.hidden_dim
    w = torch.cat([m.adaLN_modulation[1].weight for m in mods], dim=0)
    b = torch.cat([m.adaLN_modulation[1].bias for m in mods], dim=0)
    one = torch.zeros(2 * H, device=b.device, dtype=b.dtype)
    one[:H] = 1.0
    b = b + one.repeat(len(mods))
    return linear(F.silu(cond), w, b).split(2 * H, dim=-1)


def score_cond_embedding(net, time: torch.Tensor, observation: Optional[torch.Tensor], batch: int,
                         continuous: Optional[bool] = None):
    """Time embeddings (models/score_networks.py:117-141) + observation encoder (:143-149): the
    conditioning vector `cond` [B,H] every adaLN modulation reads, and the time weight of the
    continuous branch (or None).  `continuous` states which time branch (:121) applies when the caller
    already knows it (the ELBO draws t in [0,1) itself); None reads it from `time` as the reference
    does, which costs a host sync and cannot be captured in a CUDA graph.  The observation encoder's
    Dropout(0.1) follows `net.training` as in the reference."""
    H = net.hidden_dim
    if continuous is None:
        continuous = bool(time.max() <= 1.0 and time.min() >= 0.0)
    if continuous:
        # continuous_time_embed.0 is Linear(1 -> E): an outer product, kept element-wise
        c0 = net.continuous_time_embed[0]
        t_norm = 2.0 * time.view(-1, 1) - 1.0
        t_cont = _seq(list(net.continuous_time_embed)[1:], t_norm * c0.weight.view(1, -1) + c0.bias)
        t_emb = _time_embed(net, time * 999.0) + net.time_scale * t_cont
        time_weight = torch.sqrt(1.0 / (1e-5 + time.view(-1, 1)))
    else:
        t_emb, time_weight = _time_embed(net, time), None
    if observation is not None:
        enc = net.obs_encoder
        o = enc[3](F.silu(enc[1](linear(observation, enc[0].weight, enc[0].bias))))     # enc[3]: Dropout(0.1)
        o = F.silu(enc[5](linear(o, enc[4].weight, enc[4].bias)))
        o = enc[8](linear(o, enc[7].weight, enc[7].bias))
    else:
        o = torch.zeros(batch, H, device=time.device)
    return t_emb + o, time_weight


def score_conditioning(net, time: torch.Tensor, observation: Optional[torch.Tensor], batch: int,
                       continuous: Optional[bool] = None):
    """Everything of models/score_networks.py:101-171 that does not depend on z_t: the conditioning
    embedding and the 13 adaLN modulations of it.  Returns (modulations, time_weight or None).  The
    ELBO evaluates the score net twice on the same (t, observation) -- the score-matching term and the
    gradient penalty (core/active_inference.py:584,717) -- so it computes this once and shares it: one
    [B,512]x[512,13312] modulation GEMM (16 % of a forward's FLOPs) and one backward of it instead of
    two; autograd sums both branches' gradients into the shared tensor, so the parameter gradients are
    the same sums."""
    cond, time_weight = score_cond_embedding(net, time, observation, batch, continuous)
    return _all_modulations(net, cond), time_weight


def fold_attention(net):
    """nn.MultiheadAttention over ONE token has softmax == 1, so attn(x) = W_o (W_v x + b_v) + b_o =
    (W_o W_v) x + (W_o b_v + b_o) (models/score_networks.py:214-224; the inference kernels use the
    same fold).  The fold is a differentiable H x H x H product of the parameters, so the training
    graph runs ONE batch-sized GEMM per block instead of two (and one input-/weight-gradient pair
    instead of two); the chain rule through the fold gives in_proj_weight[2H:], in_proj_bias[2H:],
    out_proj.weight and out_proj.bias exactly the gradients of the two-GEMM form.
    Returns [(W_f [H,H], b_f [H])] per block."""
    H = net.hidden_dim
    folds = []
    for blk in net.transformer_blocks:
        att = blk.attention
        w_v, b_v = att.in_proj_weight[2 * H:], att.in_proj_bias[2 * H:]
        w_o, b_o = att.out_proj.weight, att.out_proj.bias
        folds.append((linear(w_o, w_v.t()), torch.mv(w_o, b_v) + b_o))
    return folds


def score_from_conditioning(net, z_t: torch.Tensor, mod, time_weight: Optional[torch.Tensor],
                            folds=None) -> torch.Tensor:
    """The z_t-dependent part of the forward: latent_proj, the DiT blocks, output head (:151-171)."""
    if folds is None:
        folds = fold_attention(net)
    h = linear(z_t, net.latent_proj.weight, net.latent_proj.bias)
    for i, blk in enumerate(net.transformer_blocks):
        w_f, b_f = folds[i]
        h = h + linear(_ada_ln(h, mod[2 * i]), w_f, b_f)          # seq-len-1 attention, folded
        h = h + _seq(blk.mlp, _ada_ln(h, mod[2 * i + 1]))
    s = _seq(net.output_proj, _ada_ln(h, mod[-1]))
    s = torch.clamp(s, min=-10, max=10) * net.output_multiplier
    return s * time_weight if time_weight is not None else s


def score_forward(net, z_t: torch.Tensor, time: torch.Tensor,
                  observation: Optional[torch.Tensor] = None, continuous: Optional[bool] = None) -> torch.Tensor:
    """models/score_networks.py:101-171 with torch ops (eval-mode obs_encoder: Dropout = identity,
    the same contract as the fused path)."""
    mod, time_weight = score_conditioning(net, time, observation, z_t.shape[0], continuous)
    return score_from_conditioning(net, z_t, mod, time_weight)


# ---------------------------------------------------------------------------------------------
# Differentiable EFE heads / rollout: used when a caller records a graph through the heads
# (policy / value / dynamics training, agents/state_agent.py:162-238).  Same math as the fused
# rollout kernel (aid_efe_rollout); every Linear on aid_gemm_nt.

def needs_graph(module, *tensors) -> bool:
    """True when autograd is recording and an input or a parameter of `module` requires grad."""
    if not torch.is_grad_enabled():
        return False
    if any(t is not None and torch.is_tensor(t) and t.requires_grad for t in tensors):
        return True
    return any(p.requires_grad for p in module.parameters())


def policy_forward(pol, z: torch.Tensor, eps: Optional[torch.Tensor]):
    """models/policy_networks.py:95-146 -> (action, mean, log_std clamped, std)."""
    h = _seq(pol.latent_encoder, z)
    h = h + _seq(pol.trunk, h)
    mean = _seq(pol.mean_head, h)
    log_std = torch.clamp(_seq(pol.log_std_head, h), pol.log_std_min, pol.log_std_max)
    std = torch.exp(log_std)
    action = mean if eps is None else mean + std * eps
    return action, mean, log_std, std


def dynamics_forward(dyn, state: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
    """models/dynamics_models.py:47-67 (residual)."""
    return state + _seq(dyn.network, torch.cat([state, action], dim=-1))


def value_forward(val, state: torch.Tensor, time: torch.Tensor) -> torch.Tensor:
    """models/value_networks.py:47-60 -> [B, 1]."""
    emb = val.time_embed[0]
    e = _sinusoid(time, val.time_embed_dim, emb.freq_scale)
    t_emb = F.relu(linear(e, val.time_embed[1].weight, val.time_embed[1].bias))
    return _seq(val.network, torch.cat([state, t_emb], dim=-1))


def efe_rollout(ai, latent: torch.Tensor, horizon: int, num_trajectories: int, policy_noise: torch.Tensor,
                reparam_noise: torch.Tensor, epistemic: Optional[torch.Tensor]):
    """compute_expected_free_energy_diffusion (core/active_inference.py:314-396) as a differentiable
    graph.  Returns (efe [B], first_action [B,A], pragmatic_last [K,B], consistency_last [K,B])."""
    cfg = ai.config
    ew, pw, cw, gamma = (float(cfg.epistemic_weight), float(cfg.pragmatic_weight),
                         float(cfg.consistency_weight), float(cfg.discount_factor))
    B = latent.shape[0]
    std_next = math.exp(0.5 * math.log(0.1))
    total = torch.zeros(B, device=latent.device)
    first_action, prag_l, cons_l = None, [], []
    for k in range(num_trajectories):
        cur, traj = latent, torch.zeros(B, device=latent.device)
        for t in range(horizon):
            d = k * horizon + t
            action, _, _, std = policy_forward(ai.policy_network, cur, policy_noise[d])
            if d == 0:
                first_action = action
            mean = cur + dynamics_forward(ai.latent_dynamics, cur, action)       # 2z + f(z,a), SURVEY fact 10
            nxt = mean + reparam_noise[d] * std_next
            r = _seq(ai.reward_predictor, nxt)[:, 0]
            prag = pw * (r / ai.preference_temperature)
            prag = prag + value_forward(ai.value_network, nxt, torch.full((B,), float(t), device=latent.device)).squeeze(-1)
            cons = -(0.5 + 0.5 * math.log(2 * math.pi) + torch.log(std)).sum(-1)
            step = pw * prag + cw * cons
            if epistemic is not None:
                step = step + ew * epistemic[d]
            traj = traj + (gamma ** t) * step
            cur = nxt
        total = total + traj / num_trajectories
        prag_l.append(prag)
        cons_l.append(cons)
    return total, first_action, torch.stack(prag_l), torch.stack(cons_l)


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
