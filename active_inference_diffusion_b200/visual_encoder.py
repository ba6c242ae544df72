"""DrQ-v2 visual encoder on the sm_100a path.

Mirror of the reference's `DrQV2Encoder` / `SpatialAttention`
(`active_inference_diffusion/encoder/visual_encoders.py:13-224`): same constructor signature, same
parameter/buffer registration order and initialisation draws (so the same seed gives the same
`state_dict`, spectral-norm `weight_orig / weight_u / weight_v` entries included), same input
conventions.  The forward pass is one `aid_encoder_forward` call: the four 3x3 convolutions and the
`Linear(conv_out_dim -> 2*feature_dim)` run as tcgen05 GEMMs, GroupNorm + Mish, the spatial
attention, the LayerNorms and the small projection tail are fused element-wise kernels around them
(layout in `csrc/encoder.inc`).

The fused kernels are the inference path (eval mode under no recorded graph: what
`DiffusionPixelAgent.act`, the collector and evaluation use).  In training mode -- or when the
input requires grad -- the module evaluates the reference's forward as a differentiable graph
(`_forward_autograd`): Dropout2d / Dropout masks and the spectral-norm power iteration behave as in
the reference, the 117 M-parameter `Linear(conv_out_dim -> 2F)` and `Linear(2F -> F)` run forward,
input-gradient and weight-gradient on the tcgen05 `aid_gemm_nt` path; the 3x3 convolutions and
their gradients use torch's conv ops in this round (hand-written conv backward kernels are the
remaining part of SURVEY §8 f-1).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib


class SpatialAttention(nn.Module):
    """Parameters of `visual_encoders.py:192-208`; evaluated inside `aid_encoder_forward`.

    `channel_reduce` is registered (it is part of the reference's `state_dict`) but the reference's
    forward never uses it (`:210-224`), so neither does the kernel."""

    def __init__(self, channels: int):
        super().__init__()
        self.channel_reduce = nn.Conv2d(channels, max(channels // 8, 16), 1)
        self.spatial_conv = nn.Conv2d(2, 1, kernel_size=7, padding=3)
        self.temperature = nn.Parameter(torch.ones(1))


class DrQV2Encoder(nn.Module):
    def __init__(self, obs_shape: Tuple[int, int, int], feature_dim: int = 50, frame_stack: int = 1,
                 num_layers: int = 4, num_filters: int = 32, use_spectral_norm: bool = True,
                 use_attention: bool = True):
        super().__init__()
        c, h, w = obs_shape
        self.base_channels = c
        self.frame_stack = frame_stack
        self.input_channels = c * frame_stack
        self.use_attention = use_attention
        self.use_spectral_norm = use_spectral_norm
        self.obs_shape = (self.input_channels, h, w)
        self.feature_dim = feature_dim
        self.num_filters = num_filters
        self.num_layers = num_layers

        widths = [self.input_channels] + [num_filters * 2 ** min(i, 3) for i in range(num_layers)]
        self.convs = nn.ModuleList()
        self.norms = nn.ModuleList()
        self.dropouts = nn.ModuleList()
        for i in range(num_layers):
            conv = nn.Conv2d(widths[i], widths[i + 1], kernel_size=3, stride=2 if i == 0 else 1, padding=1,
                             bias=False)
            self.convs.append(nn.utils.spectral_norm(conv) if use_spectral_norm else conv)
            self.norms.append(nn.GroupNorm(min(32, widths[i + 1] // 4), widths[i + 1]))
            self.dropouts.append(nn.Dropout2d(0.1 * (i / num_layers)))
        if use_attention:
            self.attention = SpatialAttention(widths[-1])

        # The reference sizes the projection with a training-mode dry run on a zero image (:92-102).
        # That run is part of the initial state: it advances every spectral-norm u/v by one power
        # iteration and draws the Dropout2d masks from the global generator.  Reproduced here with
        # torch ops (construction time only) so that a seed gives the reference's state_dict.
        probe = torch.zeros(1, *self.obs_shape)
        for i in range(num_layers):
            probe = torch.nn.functional.mish(self.norms[i](self.convs[i](probe)))
            if i + 1 < num_layers:
                probe = self.dropouts[i](probe)
        self.conv_out_dim = probe.reshape(1, -1).shape[1]
        self._out_hw = (probe.shape[2], probe.shape[3])

        self.ln = nn.LayerNorm(self.conv_out_dim)
        self.output_layers = nn.Sequential(
            nn.Linear(self.conv_out_dim, feature_dim * 2), nn.LayerNorm(feature_dim * 2), nn.Mish(),
            nn.Dropout(0.1), nn.Linear(feature_dim * 2, feature_dim), nn.LayerNorm(feature_dim), nn.Tanh())
        self._initialize_weights()
        object.__setattr__(self, "_cache", _lib.PackedCache())
        self.precision = "bf16"

    def _initialize_weights(self) -> None:
        # same walk and the same draws as visual_encoders.py:121-134 (for a spectral-normed conv
        # `weight` is the derived tensor, so the draw is spent without touching weight_orig)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, (nn.LayerNorm, nn.GroupNorm)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    # ------------------------------------------------------------------------------------------
    def dims(self) -> _lib.AidEncoderDims:
        if self.precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        return _lib.AidEncoderDims(self.input_channels, self.obs_shape[1], self.obs_shape[2], self.num_filters,
                                   self.num_layers, self.feature_dim, int(self.use_attention),
                                   _lib.PRECISIONS[self.precision][1])

    def _lib_handle(self):
        """The encoder's own `precision` names the library (operand type) and the hi/lo split."""
        return _lib.lib(_lib.PRECISIONS[self.precision][0])

    def invalidate_packed(self) -> None:
        """Required after in-place `.data` edits of encoder parameters (see _lib.PackedCache)."""
        self._cache.invalidate()

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._cache.invalidate()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._cache.invalidate()
        return out

    def _param_table(self):
        sd = dict(self.named_parameters())
        sd.update(dict(self.named_buffers()))
        table = []
        for i in range(self.num_layers):
            if self.use_spectral_norm:
                table += [sd[f"convs.{i}.weight_orig"], sd[f"convs.{i}.weight_u"], sd[f"convs.{i}.weight_v"]]
            else:
                table += [sd[f"convs.{i}.weight"], None, None]
            table += [sd[f"norms.{i}.weight"], sd[f"norms.{i}.bias"]]
        if self.use_attention:
            table += [sd["attention.spatial_conv.weight"], sd["attention.spatial_conv.bias"],
                      sd["attention.temperature"]]
        else:
            table += [None, None, None]
        table += [sd["ln.weight"], sd["ln.bias"]]
        for j in (0, 1, 4, 5):
            table += [sd[f"output_layers.{j}.weight"], sd[f"output_layers.{j}.bias"]]
        return table

    def packed_weights(self, verify: bool = False) -> torch.Tensor:
        params = self._param_table()
        dev = _lib.require_cuda(*params)

        def build():
            l = self._lib_handle()
            d = self.dims()
            nbytes = l.aid_encoder_packed_bytes(ctypes.byref(d))
            if nbytes == 0:
                _lib.check(-1, "aid_encoder_packed_bytes", l)
            keep = [None if p is None else _lib.f32c(p.detach()) for p in params]
            table = (ctypes.c_void_p * len(keep))(*[None if t is None else t.data_ptr() for t in keep])
            packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                _lib.check(l.aid_encoder_pack(ctypes.byref(d), table, len(keep), packed.data_ptr(), nbytes,
                                              _lib.stream_ptr(dev)), "aid_encoder_pack", l)
            return packed

        return self._cache.get(("encoder", self.precision), params, build, verify)

    def _canonical_input(self, x: torch.Tensor) -> torch.Tensor:
        # input conventions of visual_encoders.py:149-162
        if x.dim() == 5:
            b, t, c, h, w = x.shape
            if t != self.frame_stack:
                raise AssertionError(f"Expected {self.frame_stack} frames, got {t}")
            x = x.reshape(b, t * c, h, w)
        elif x.dim() == 4:
            c = x.shape[1]
            if c == self.base_channels and self.frame_stack > 1:
                x = x.repeat(1, self.frame_stack, 1, 1)
            elif c != self.input_channels:
                raise ValueError(f"Expected {self.input_channels} channels, got {c}")
        elif x.dim() == 3:
            x = x.unsqueeze(0)
        else:
            raise ValueError(f"Unexpected observation shape: {x.shape}")
        if tuple(x.shape[1:]) != tuple(self.obs_shape):
            raise ValueError(f"Expected observations of shape {self.obs_shape}, got {tuple(x.shape[1:])}")
        return x

    # Convolutions of the differentiable graph: "native" = conv_ops.conv3x3 (forward / input gradient /
    # weight gradient on the library's tcgen05 GEMMs, csrc/conv_train.inc); "torch" = nn.Conv2d (cuDNN),
    # kept as the cross-check.
    conv_backend = "native"

    def _conv_weight(self, i: int, x: torch.Tensor) -> torch.Tensor:
        """The weight nn.Conv2d would use: for a spectral-normed layer the forward pre-hook derives
        W / sigma from weight_orig (and advances the power iteration in training mode, as the module's
        own forward does); gradients reach weight_orig through that expression."""
        conv = self.convs[i]
        for hook in conv._forward_pre_hooks.values():
            hook(conv, (x,))
        return conv.weight

    def _forward_autograd(self, x: torch.Tensor) -> torch.Tensor:
        """visual_encoders.py:166-189 as a differentiable graph (training mode / input gradients)."""
        from . import autograd_path, conv_ops
        F = torch.nn.functional
        _lib.require_cuda(x)
        x = x.float() / 255.0 if x.dtype == torch.uint8 else x.float()
        native = self.conv_backend == "native"
        for i in range(self.num_layers):
            if native:
                x = conv_ops.conv3x3(x, self._conv_weight(i, x), 2 if i == 0 else 1, self.precision)
            else:
                x = self.convs[i](x)
            x = F.mish(self.norms[i](x))
            if i + 1 < self.num_layers:
                x = self.dropouts[i](x)
        if self.use_attention:           # SpatialAttention.forward, :210-224
            att = self.attention
            pooled = torch.cat([x.mean(dim=1, keepdim=True), x.max(dim=1, keepdim=True)[0]], dim=1)
            if native:
                # the 7x7, 2 -> 1 channel convolution (98 MACs per pixel) as a weighted sum over sliding-window
                # views of the padded pools (no copy, no library convolution)
                win = F.pad(pooled, (3, 3, 3, 3)).unfold(2, 7, 1).unfold(3, 7, 1)       # [n, 2, hh, ww, 7, 7]
                logits = (win * att.spatial_conv.weight.view(1, 2, 1, 1, 7, 7)).sum(dim=(1, 4, 5)).unsqueeze(1) \
                    + att.spatial_conv.bias.view(1, 1, 1, 1)
            else:
                logits = att.spatial_conv(pooled)
            x = x + x * torch.sigmoid(logits / att.temperature)
        x = self.ln(x.reshape(x.shape[0], -1))
        with autograd_path.precision(self.precision):
            return autograd_path.seq(self.output_layers, x)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self._canonical_input(x)
        if self.training or (torch.is_grad_enabled() and x.requires_grad):
            return self._forward_autograd(x)
        dev = _lib.require_cuda(x)
        is_u8 = x.dtype == torch.uint8
        x = x.contiguous() if is_u8 else _lib.f32c(x)
        packed = self.packed_weights()
        l = self._lib_handle()
        d = self.dims()
        batch = x.shape[0]
        need = l.aid_encoder_workspace_bytes(ctypes.byref(d), batch)
        if need == 0:
            _lib.check(-1, "aid_encoder_workspace_bytes", l)
        ws = self._cache.workspace(need, dev)
        out = torch.empty(batch, self.feature_dim, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(l.aid_encoder_forward(ctypes.byref(d), packed.data_ptr(), ws.data_ptr(), ws.numel(),
                                             batch, x.data_ptr(), int(is_u8), out.data_ptr(), _lib.stream_ptr(dev)),
                       "aid_encoder_forward", l)
        return out
