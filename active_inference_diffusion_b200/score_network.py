"""Drop-in `LatentScoreNetwork` whose forward runs on the sm_100a library.

Mirrors the reference's constructor signature, parameter registration order, initialisation
calls and `state_dict` keys (models/score_networks.py:12-99, 174-212, 238-255, 273-281) so that
(a) the same `torch.manual_seed` produces identical weights and (b) reference checkpoints load
unchanged.  The submodules are parameter containers only: `forward` packs the parameters into
tcgen05 operand tiles (a derived cache, rebuilt when any parameter changes) and calls
`aid_score_forward`.
"""
from __future__ import annotations

import ctypes
import warnings
from typing import Optional

import torch
import torch.nn as nn

from . import _lib


class SinusoidalPositionEmbeddings(nn.Module):
    """Holds the learnable `freq_scale` (models/score_networks.py:273-281); the embedding itself
    is computed by `k_sincos_pack` on the device."""

    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim
        self.freq_scale = nn.Parameter(torch.ones(1))


class AdaptiveLayerNorm(nn.Module):
    """adaLN parameters (models/score_networks.py:238-255): zero-initialised modulation."""

    def __init__(self, hidden_dim: int):
        super().__init__()
        self.norm = nn.LayerNorm(hidden_dim, elementwise_affine=False)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_dim, 2 * hidden_dim))
        nn.init.zeros_(self.adaLN_modulation[1].weight)
        nn.init.zeros_(self.adaLN_modulation[1].bias)


class DiTBlock(nn.Module):
    """Parameters of one DiT block (models/score_networks.py:174-212)."""

    def __init__(self, hidden_dim: int, num_heads: int, mlp_ratio: float = 4.0):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.num_heads = num_heads
        self.norm1 = AdaptiveLayerNorm(hidden_dim)
        self.norm2 = AdaptiveLayerNorm(hidden_dim)
        self.attention = nn.MultiheadAttention(embed_dim=hidden_dim, num_heads=num_heads,
                                               batch_first=True, dropout=0.0)
        inner = int(hidden_dim * mlp_ratio)
        self.mlp = nn.Sequential(nn.Linear(hidden_dim, inner), nn.GELU(), nn.Linear(inner, hidden_dim))
        nn.init.xavier_uniform_(self.mlp[0].weight)
        nn.init.xavier_uniform_(self.mlp[2].weight)
        nn.init.zeros_(self.mlp[0].bias)
        nn.init.zeros_(self.mlp[2].bias)


class LatentScoreNetwork(nn.Module):
    """s_theta(z_t, t, o) — same call surface as the reference (models/score_networks.py:12-171)."""

    _warned_dropout = False

    def __init__(self, latent_dim: int, observation_dim: int, hidden_dim: int = 256,
                 time_embed_dim: int = 128, num_layers: int = 6, use_attention: bool = True,
                 output_scale: float = 1e-3):
        super().__init__()
        self.latent_dim = latent_dim
        self.observation_dim = observation_dim
        self.hidden_dim = hidden_dim
        self.time_embed_dim = time_embed_dim
        self.num_heads = 8
        self.mlp_ratio = 4.0
        self.use_attention = use_attention
        self.output_scale = output_scale
        self.time_embed = nn.Sequential(
            SinusoidalPositionEmbeddings(time_embed_dim),
            nn.Linear(time_embed_dim, hidden_dim * 2), nn.SiLU(), nn.Linear(hidden_dim * 2, hidden_dim))
        self.obs_encoder = nn.Sequential(
            nn.Linear(observation_dim, hidden_dim), nn.LayerNorm(hidden_dim), nn.SiLU(), nn.Dropout(0.1),
            nn.Linear(hidden_dim, hidden_dim), nn.LayerNorm(hidden_dim), nn.SiLU(),
            nn.Linear(hidden_dim, hidden_dim), nn.LayerNorm(hidden_dim))
        self.continuous_time_embed = nn.Sequential(
            nn.Linear(1, time_embed_dim), nn.SiLU(), nn.Linear(time_embed_dim, time_embed_dim), nn.SiLU(),
            nn.Linear(time_embed_dim, hidden_dim))
        self.time_scale = nn.Parameter(torch.tensor(1.0))
        self.register_buffer("grad_norm_ema", torch.tensor(1.0))
        self.grad_norm_decay = 0.999
        self.latent_proj = nn.Linear(latent_dim, hidden_dim)
        if use_attention:
            self.transformer_blocks = nn.ModuleList(
                [DiTBlock(hidden_dim, self.num_heads, self.mlp_ratio) for _ in range(num_layers)])
        self.norm_final = AdaptiveLayerNorm(hidden_dim)
        self.output_proj = nn.Sequential(
            nn.Linear(hidden_dim, hidden_dim // 2), nn.SiLU(), nn.Linear(hidden_dim // 2, latent_dim, bias=False))
        self.output_multiplier = nn.Parameter(torch.ones(1) * output_scale)
        nn.init.zeros_(self.output_proj[-1].weight)
        # derived caches (not part of the state_dict)
        object.__setattr__(self, "_cache", _lib.PackedCache())

    @torch.no_grad()
    def randomize_zero_init(self, seed: int = 123, output_multiplier: Optional[float] = None) -> None:
        """Give the tensors the reference zero-initialises (adaLN modulations :254-255, the last output
        weight :99, the DiT MLP biases :211-212) seeded random values.  At construction the score is
        identically zero, so synthetic-weight throughput and parity runs call this to
        exercise real operand values (tensor-pipe power is data dependent).  Every tensor gets its own
        generator stream keyed by (seed, position in the sorted parameter names)."""
        named = dict(self.named_parameters())
        for i, k in enumerate(sorted(named)):
            p = named[k]
            g = torch.Generator().manual_seed(seed * 100003 + i)
            if k.endswith("adaLN_modulation.1.weight"):
                v = torch.randn(p.shape, generator=g) * (0.5 / p.shape[1] ** 0.5)
            elif k.endswith("adaLN_modulation.1.bias"):
                v = torch.randn(p.shape, generator=g) * 0.1
            elif k.endswith("output_proj.2.weight"):
                v = torch.randn(p.shape, generator=g) * (1.0 / p.shape[1] ** 0.5)
            elif k.endswith("mlp.0.bias") or k.endswith("mlp.2.bias"):
                v = torch.randn(p.shape, generator=g) * 0.02
            elif k.endswith("output_multiplier") and output_multiplier is not None:
                v = torch.full(p.shape, float(output_multiplier))
            else:
                continue
            p.copy_(v.to(p.device, p.dtype))
        self._cache.invalidate()

    # ---- derived cache -------------------------------------------------------------------
    @property
    def num_blocks(self) -> int:
        return len(self.transformer_blocks) if self.use_attention else 0

    def dims(self) -> _lib.AidScoreDims:
        return _lib.AidScoreDims(self.latent_dim, self.observation_dim, self.hidden_dim,
                                 self.time_embed_dim, self.num_blocks)

    def _param_table(self):
        named = dict(self.named_parameters())
        keys = list(_lib.SCORE_PARAM_KEYS)
        for i in range(self.num_blocks):
            keys += [f"transformer_blocks.{i}.{k}" for k in _lib.SCORE_BLOCK_KEYS]
        return [named[k] for k in keys]

    def invalidate_packed(self) -> None:
        """Drop the packed operand tiles: required after in-place edits of parameters through `.data`
        (they do not bump the version counters the cache is keyed on; see _lib.PackedCache)."""
        self._cache.invalidate()

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._cache.invalidate()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._cache.invalidate()
        return out

    def packed_weights(self, verify: bool = False) -> torch.Tensor:
        """tcgen05 operand tiles of the current parameters in the current operand type
        (`_lib.operand_type()`), rebuilt when the parameters change."""
        params = self._param_table()
        dev = _lib.require_cuda(*params)

        def build():
            l = _lib.lib()
            d = self.dims()
            nbytes = l.aid_score_packed_bytes(ctypes.byref(d))
            if nbytes == 0:
                _lib.check(-1, "aid_score_packed_bytes")
            keep = [_lib.f32c(p.detach()) for p in params]
            table = (ctypes.c_void_p * len(keep))(*[t.data_ptr() for t in keep])
            packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                _lib.check(l.aid_score_pack(ctypes.byref(d), table, len(keep), packed.data_ptr(), nbytes,
                                            _lib.stream_ptr(dev)), "aid_score_pack")
            return packed

        return self._cache.get(("score", _lib.operand_type()), params, build, verify)

    def workspace(self, batch: int, table_rows: int, device: torch.device) -> torch.Tensor:
        l = _lib.lib()
        d = self.dims()
        need = l.aid_score_workspace_bytes(ctypes.byref(d), batch, table_rows)
        if need == 0:
            _lib.check(-1, "aid_score_workspace_bytes")
        return self._cache.workspace(need, device)

    # ---- forward -------------------------------------------------------------------------
    def forward(self, z_t: torch.Tensor, time: torch.Tensor,
                observation: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Score [B, L].  The continuous/discrete branch is decided per BATCH from `time`
        exactly as the reference does (models/score_networks.py:121) — one host sync, as there."""
        if self.training and observation is not None and not LatentScoreNetwork._warned_dropout:
            # Parity contract: the fused path evaluates obs_encoder's Dropout(0.1) as the
            # identity; the reference's train()-mode masks come from a fused native op and cannot be
            # injected.  Warn once instead of failing so agents that call .train() keep working.
            warnings.warn("LatentScoreNetwork (b200): obs_encoder Dropout(0.1) is evaluated as identity "
                          "in train() mode")
            LatentScoreNetwork._warned_dropout = True
        if not self.use_attention:
            raise NotImplementedError("use_attention=False is never instantiated by the reference")
        dev = _lib.require_cuda(z_t, time, observation)
        if torch.is_grad_enabled() and (z_t.requires_grad or time.requires_grad
                                        or (observation is not None and observation.requires_grad)
                                        or any(p.requires_grad for p in self.parameters())):
            # A caller that records a graph (the reference's own compute_diffusion_elbo :584,717, or
            # FreeEnergyComputation :77) gets the differentiable evaluation: same math, every GEMM on
            # aid_gemm_nt, differentiable to any order.  Under torch.no_grad() the fused inference
            # kernels below run.
            from . import autograd_path
            return autograd_path.score_forward(self, _lib.f32c(z_t), _lib.f32c(time), _lib.f32c(observation))
        z_t, time, observation = _lib.f32c(z_t), _lib.f32c(time), _lib.f32c(observation)
        batch = z_t.shape[0]
        if batch == 0:                  # empty batch: empty score, as the reference module gives
            return torch.empty(0, self.latent_dim, dtype=torch.float32, device=dev)
        continuous = bool(time.max() <= 1.0 and time.min() >= 0.0)
        packed = self.packed_weights()
        ws = self.workspace(batch, batch, dev)
        out = torch.empty(batch, self.latent_dim, dtype=torch.float32, device=dev)
        d = self.dims()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().aid_score_forward(
                ctypes.byref(d), packed.data_ptr(), ws.data_ptr(), ws.numel(), z_t.data_ptr(), time.data_ptr(),
                _lib.ptr(observation), batch, int(continuous), out.data_ptr(), _lib.stream_ptr(dev)),
                "aid_score_forward")
        return out
