"""Drop-in `DiffusionActiveInference` (core/active_inference.py:19-771) for state observations.

Constructor signature, sub-module names/registration order/initialisation and `state_dict` keys
follow the reference so seeded weights and checkpoints are interchangeable.  The three hot
methods run on the sm_100a library:

  update_belief_via_diffusion              -> aid_sample           (reverse diffusion)
  compute_expected_free_energy_diffusion   -> aid_efe_rollout      (policy/dynamics/reward/value rollout)
  latent_score_network(...)                -> aid_score_forward

`compute_diffusion_elbo` (training) evaluates the same loss as the reference; in this round its
autograd graph is built from torch ops on the device (see DESIGN.md "training path") — the native
fwd/bwd kernels are the next row of SURVEY §8.

Pixel observations (ConvDecoder, SpatialAttentionAggregator) are SURVEY §8(f) "next" and raise.
"""
from __future__ import annotations

import contextlib
import ctypes
import math
from typing import Dict, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .diffusion import LatentDiffusionProcess
from .heads import (DiffusionConditionedPolicy, HeadsBundle, LatentDynamicsModel, ValueNetwork,
                    make_reward_predictor)
from .score_network import LatentScoreNetwork
from . import _lib, autograd_path, distributed, train_native


class FunctionSpaceEpistemicEstimator(nn.Module):
    """State-mode MINE epistemic estimator (core/active_inference.py:839-1063), parameters and
    forward.  The reference forward calls `self.decoder(z)` on an nn.ModuleList, which raises
    (:953); here the decoder is applied with `decode_observation`'s skip forward (:237-242), the
    documented shim of SURVEY §8c(1).  All its dense layers (decoder x5 passes, feature extractor,
    projector, latent processor, MINE network) run on `aid_gemm_nt` in the bf16x3 mode: the value is
    a difference of two batch means, so it is kept at fp32-class precision.  It yields ONE scalar
    per call for the whole batch (:1050-1053)."""

    def __init__(self, decoder: nn.Module, latent_dim: int, observation_shape, hidden_dim: int = 256,
                 spatial_aggregator_output_dim: int = 256, is_pixel_observation: bool = True,
                 device: Union[str, torch.device] = "cuda"):
        super().__init__()
        if is_pixel_observation:
            raise NotImplementedError("pixel-mode epistemic estimator: SURVEY §8(f) 'next'")
        self.decoder = decoder
        self.latent_dim = latent_dim
        self.is_pixel = False
        self.device = torch.device(device) if isinstance(device, str) else device
        self.ntk_samples = 4
        self.perturbation_scale = nn.Parameter(torch.tensor(0.1)).to(self.device)
        self.state_dim = observation_shape
        self.feature_extractor = nn.Sequential(nn.Linear(self.state_dim, 128), nn.ReLU(), nn.Linear(128, 256),
                                               nn.ReLU(), nn.Linear(256, 128))
        jacobian_dim = 128 * self.ntk_samples
        self.jacobian_projector = nn.Sequential(nn.Linear(jacobian_dim, 512), nn.LayerNorm(512), nn.ReLU(),
                                                nn.Dropout(0.1), nn.Linear(512, 256))
        self.latent_processor = nn.Sequential(nn.Linear(latent_dim, 128), nn.ReLU(), nn.Linear(128, 128))
        self.mine_network = nn.Sequential(nn.Linear(spatial_aggregator_output_dim + 128, 512), nn.ReLU(),
                                          nn.Dropout(0.1), nn.Linear(512, 512), nn.ReLU(), nn.Dropout(0.1),
                                          nn.Linear(512, 1))
        self.register_buffer("running_mean", torch.tensor(0.0))
        self.alpha = 0.01
        self.to(self.device)

    def to(self, device):
        super().to(device)
        self.device = device if isinstance(device, torch.device) else torch.device(device)
        return self

    # Every nn.Linear below runs through autograd_path.seq -> aid_gemm_nt (tcgen05); LayerNorm /
    # ReLU / SiLU / Dropout(eval) stay element-wise device ops.
    def _decode(self, z: torch.Tensor) -> torch.Tensor:
        m = self.decoder
        h1 = autograd_path.seq(m[0], z)
        h2 = autograd_path.seq(m[1], h1) + h1
        return autograd_path.linear(autograd_path.seq(m[2], h2), m[3].weight, m[3].bias)

    def compute_jacobian_features(self, z: torch.Tensor, dir_noise=None) -> torch.Tensor:
        was_training = self.decoder.training
        self.decoder.eval()
        with torch.no_grad():
            f_z = self._decode(z)
        eps = self.perturbation_scale
        feats = []
        for i in range(self.ntk_samples):
            d = torch.randn_like(z) if dir_noise is None else dir_noise[i]
            delta = F.normalize(d, dim=-1) * eps
            with torch.no_grad():
                f_p = self._decode(z + delta)
            feats.append(autograd_path.seq(self.feature_extractor, (f_p - f_z) / eps))
        if was_training:
            self.decoder.train()
        return autograd_path.seq(self.jacobian_projector, torch.cat(feats, dim=1))

    def forward(self, next_latent_mean: torch.Tensor, next_latent_logvar: torch.Tensor, num_samples: int = 5,
                *, z_noise=None, dir_noise=None, perms=None):
        """Keyword-only arguments inject the reference's draws in its order: S x
        randn_like [B,L], 4 x randn_like [S*B,L], S x randperm(B).  Without a graph being recorded the
        batched evaluation below runs (no host read until the metrics are converted, once)."""
        if not torch.is_grad_enabled():
            value, stats = self.forward_device(next_latent_mean, next_latent_logvar, num_samples,
                                               z_noise=z_noise, dir_noise=dir_noise, perms=perms)
            return value, self.metrics_from(stats)
        B = next_latent_mean.shape[0]
        std = torch.exp(0.5 * next_latent_logvar)
        zs = [next_latent_mean + (torch.randn_like(next_latent_mean) if z_noise is None else z_noise[i]) * std
              for i in range(num_samples)]
        z_all = torch.cat(zs, dim=0)
        jac = self.compute_jacobian_features(z_all, dir_noise)
        lat = autograd_path.seq(self.latent_processor, z_all)
        t_joint = autograd_path.seq(self.mine_network, torch.cat([jac, lat], dim=1))
        marg = torch.cat([jac[i * B:(i + 1) * B][torch.randperm(B, device=jac.device) if perms is None else perms[i]]
                          for i in range(num_samples)], dim=0)
        t_marg = autograd_path.seq(self.mine_network, torch.cat([marg, lat], dim=1))
        # ema_loss (:828-836): forward value log(mean(exp(T))), running mean updated on the side
        t_exp = torch.exp(torch.logsumexp(t_marg, 0) - math.log(t_marg.shape[0])).detach().reshape(())
        self.running_mean = torch.where(self.running_mean == 0, t_exp,
                                        self.alpha * t_exp + (1.0 - self.alpha) * self.running_mean).reshape(())
        marginal_term = autograd_path.EMALogMeanExp.apply(t_marg, self.running_mean)
        mi = t_joint.mean() - marginal_term
        stats = torch.stack([mi.detach(), t_joint.mean().detach(), marginal_term.detach(), self.running_mean])
        return torch.clamp(mi.expand(B), min=0.0), self.metrics_from(stats)

    # The fused library path (csrc/epistemic.inc) serves eval-mode calls (its Dropout layers are the
    # identity); set False to evaluate through aid_gemm_nt in the bf16x3 mode instead.
    fused = True
    FUSED_OPERAND = "f16"
    _FUSED_KEYS = ("feature_extractor.0", "feature_extractor.2", "feature_extractor.4", "jacobian_projector.0",
                   "jacobian_projector.1", "jacobian_projector.4", "latent_processor.0", "latent_processor.2",
                   "mine_network.0", "mine_network.3", "mine_network.6")

    def _fused_params(self):
        dec = self.decoder
        out = []
        for i in range(3):
            out += [dec[i][0].weight, dec[i][0].bias, dec[i][1].weight, dec[i][1].bias]
        out += [dec[3].weight, dec[3].bias]
        for key in self._FUSED_KEYS:
            m = self.get_submodule(key)
            out += [m.weight, m.bias]
        return out

    def invalidate_packed(self) -> None:
        self._cache.invalidate()

    def _fused_dims(self):
        L = self.latent_processor[0].in_features
        H = self.decoder[2][0].out_features
        return _lib.AidEpistemicDims(L, H, self.decoder[3].out_features, self.jacobian_projector[4].out_features)

    def fused_packed(self, dev: torch.device) -> torch.Tensor:
        """Packed operand image of the estimator's weights (cached per weight version)."""
        if not hasattr(self, "_cache"):
            object.__setattr__(self, "_cache", _lib.PackedCache())
        d = self._fused_dims()
        l = _lib.lib(self.FUSED_OPERAND)
        params = self._fused_params()

        def build():
            nbytes = l.aid_epistemic_packed_bytes(ctypes.byref(d))
            if nbytes == 0:
                _lib.check(-1, "aid_epistemic_packed_bytes", l)
            keep = [_lib.f32c(p.detach()) for p in params]
            table = (ctypes.c_void_p * len(keep))(*[t.data_ptr() for t in keep])
            packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                _lib.check(l.aid_epistemic_pack(ctypes.byref(d), table, len(keep), packed.data_ptr(), nbytes,
                                                _lib.stream_ptr(dev)), "aid_epistemic_pack", l)
            return packed

        return self._cache.get(("epistemic", self.FUSED_OPERAND), params, build)

    def _forward_fused(self, mean, logvar, S, z_noise, dir_noise, perms):
        """`aid_epistemic_forward`: the whole estimator as one launch sequence on IEEE fp16 tensor-core
        operands (11-bit significand; fp32 accumulation, LayerNorm, activations and statistic).  The
        finite differences (f(z + delta) - f(z)) / 0.1 amplify operand rounding ~10x, which is why bf16
        operands are not offered here: with fp16 the MINE statistic stays within 3e-4 of the fp32
        oracle at the reference dims (tests/test_gpu_round2.py)."""
        dev = mean.device
        B, L = mean.shape
        N = S * B
        d = self._fused_dims()
        l = _lib.lib(self.FUSED_OPERAND)
        packed = self.fused_packed(dev)
        need = l.aid_epistemic_workspace_bytes(ctypes.byref(d), B, S)
        if need == 0:
            _lib.check(-1, "aid_epistemic_workspace_bytes", l)
        ws = self._cache.workspace(need, dev)
        eps_z = torch.randn(S, B, L, device=dev) if z_noise is None else torch.stack([_lib.f32c(e) for e in z_noise])
        dirs = (torch.randn(self.ntk_samples, N, L, device=dev) if dir_noise is None
                else torch.stack([_lib.f32c(e) for e in dir_noise]))
        if perms is None:
            # S uniform random permutations of range(B) in one launch sequence (argsort of iid uniforms)
            # instead of S `randperm` calls (each a multi-kernel sort: 5 of 33 ms at B = 8,192, S = 10)
            perm = torch.rand(S, B, device=dev).argsort(dim=1)
        else:
            perm = torch.stack([p.to(dev) for p in perms])
        idx = (perm + torch.arange(S, device=dev).unsqueeze(1) * B).reshape(-1).contiguous()
        group = getattr(self, "data_parallel_group", None)
        stats = torch.empty(4, dtype=torch.float32, device=dev)
        partial = torch.empty(4, dtype=torch.float64, device=dev) if group is not None else None
        rm = self.running_mean.detach().reshape(1).float().clone() if group is not None else self.running_mean.data.view(1)
        mean, logvar = _lib.f32c(mean.detach()), _lib.f32c(logvar.detach())
        eps = self.perturbation_scale.detach().reshape(1).float()
        with torch.cuda.device(dev):
            _lib.check(l.aid_epistemic_forward(
                ctypes.byref(d), packed.data_ptr(), ws.data_ptr(), ws.numel(), B, S, mean.data_ptr(), logvar.data_ptr(),
                eps_z.data_ptr(), dirs.data_ptr(), idx.data_ptr(), eps.data_ptr(), float(self.alpha), rm.data_ptr(),
                stats.data_ptr(), None, _lib.ptr(partial), _lib.stream_ptr(dev)), "aid_epistemic_forward", l)
        if group is not None:
            mi, joint, marginal_term, t_exp = distributed.merge_mine_partials(partial, group)
            self.running_mean = torch.where(self.running_mean == 0, t_exp,
                                            self.alpha * t_exp + (1.0 - self.alpha) * self.running_mean).reshape(())
            stats = torch.stack([mi.reshape(()), joint.reshape(()), marginal_term.reshape(()), self.running_mean])
        return torch.clamp(stats[0].expand(B), min=0.0), stats

    def forward_grouped(self, mean: torch.Tensor, logvar: torch.Tensor, num_samples: int, groups: int,
                        *, z_noise=None, dir_noise=None, perms=None) -> torch.Tensor:
        """`groups` independent estimator evaluations in one launch sequence (`aid_epistemic_forward_grouped`):
        rows g*Bg .. g*Bg+Bg-1 of `mean` / `logvar` [groups*Bg, L] are the batch of evaluation g.  Returns
        group_stats [groups, 4] = (mi, joint, marginal term, exp(marginal term)) on the device; the running
        mean is left to `apply_running_mean` (the reference updates it once per evaluation, in order).
        Injected draws (tests): z_noise [S, groups*Bg, L], dir_noise [4, S*groups*Bg, L], perms [S, groups, Bg]."""
        dev = mean.device
        Bt, L = mean.shape
        S, Bg = num_samples, Bt // groups
        N = S * Bt
        d = self._fused_dims()
        l = _lib.lib(self.FUSED_OPERAND)
        packed = self.fused_packed(dev)
        need = l.aid_epistemic_workspace_bytes(ctypes.byref(d), Bt, S)
        if need == 0:
            _lib.check(-1, "aid_epistemic_workspace_bytes", l)
        ws = self._cache.workspace(need, dev)
        eps_z = torch.randn(S, Bt, L, device=dev) if z_noise is None else _lib.f32c(z_noise)
        dirs = torch.randn(self.ntk_samples, N, L, device=dev) if dir_noise is None else _lib.f32c(dir_noise)
        perm = torch.rand(S, groups, Bg, device=dev).argsort(dim=2) if perms is None else perms.to(dev)
        # row s*Bt + g*Bg + b of the marginal pass reads row s*Bt + g*Bg + perm[s, g, b]
        idx = (perm + (torch.arange(S, device=dev).view(S, 1, 1) * Bt
                       + torch.arange(groups, device=dev).view(1, groups, 1) * Bg)).reshape(-1).contiguous()
        stats = torch.empty(groups, 4, dtype=torch.float32, device=dev)
        mean, logvar = _lib.f32c(mean.detach()), _lib.f32c(logvar.detach())
        eps = self.perturbation_scale.detach().reshape(1).float()
        with torch.cuda.device(dev):
            _lib.check(l.aid_epistemic_forward_grouped(
                ctypes.byref(d), packed.data_ptr(), ws.data_ptr(), ws.numel(), Bt, S, groups, mean.data_ptr(),
                logvar.data_ptr(), eps_z.data_ptr(), dirs.data_ptr(), idx.data_ptr(), eps.data_ptr(),
                stats.data_ptr(), _lib.stream_ptr(dev)), "aid_epistemic_forward_grouped", l)
        return stats

    def apply_running_mean(self, t_exp: torch.Tensor) -> None:
        """The running-mean updates (:828-836) of the evaluations whose exp(marginal term) are `t_exp`, in order."""
        t_exp = _lib.f32c(t_exp.reshape(-1))
        dev = t_exp.device
        with torch.cuda.device(dev):
            _lib.check(_lib.lib(self.FUSED_OPERAND).aid_ema_sequence(
                t_exp.data_ptr(), t_exp.numel(), float(self.alpha), self.running_mean.data.view(1).data_ptr(),
                _lib.stream_ptr(dev)), "aid_ema_sequence")

    METRIC_KEYS = ("epistemic/mi_estimate", "epistemic/joint_term", "epistemic/marginal_term",
                   "epistemic/running_mean")

    @classmethod
    def metrics_from(cls, stats: torch.Tensor) -> Dict[str, float]:
        """The reference's four `.item()` reads (:1055-1060) as ONE device->host transfer."""
        return dict(zip(cls.METRIC_KEYS, stats.tolist()))

    @torch.no_grad()
    def forward_device(self, next_latent_mean: torch.Tensor, next_latent_logvar: torch.Tensor, num_samples: int = 5,
                       *, z_noise=None, dir_noise=None, perms=None):
        """The estimator's forward (:996-1063) without a host synchronisation and with its repeated
        passes batched: the base and the 4 perturbed decoder evaluations are ONE pass over 5*S*B rows,
        the 4 feature-extractor calls one pass over 4*S*B rows, the joint and the marginal MINE
        evaluations one pass over 2*S*B rows (15 dense layers are launched once instead of 42 times).
        Row r of every pass sees exactly the inputs the reference's call sequence gives it.  The
        running mean (:828-836, first-call rule included) is updated by a device-side select.  With
        `data_parallel_group` set the batch is a shard and the statistic of the GLOBAL batch comes from
        one 3-float all-reduce.  Returns (epistemic[B], stats[4] on the device in
        METRIC_KEYS order)."""
        mean, logvar = next_latent_mean.to(self.device), next_latent_logvar.to(self.device)
        B, L = mean.shape
        S, N = num_samples, num_samples * mean.shape[0]
        if self.fused and mean.is_cuda and not self.training and B > 0:
            return self._forward_fused(mean, logvar, S, z_noise, dir_noise, perms)
        eps_z = torch.randn(S, B, L, device=mean.device) if z_noise is None else torch.stack(list(z_noise))
        z_all = (mean.unsqueeze(0) + eps_z * torch.exp(0.5 * logvar).unsqueeze(0)).reshape(N, L)
        d = torch.randn(self.ntk_samples, N, L, device=mean.device) if dir_noise is None else torch.stack(list(dir_noise))
        eps = self.perturbation_scale
        delta = F.normalize(d, dim=-1) * eps
        was_training = self.decoder.training
        self.decoder.eval()
        f = self._decode(torch.cat([z_all.unsqueeze(0), z_all.unsqueeze(0) + delta], dim=0).reshape(-1, L))
        if was_training:
            self.decoder.train()
        f = f.view(self.ntk_samples + 1, N, -1)
        diff = ((f[1:] - f[:1]) / eps).reshape(self.ntk_samples * N, -1)
        feats = autograd_path.seq(self.feature_extractor, diff).view(self.ntk_samples, N, -1)
        jac = autograd_path.seq(self.jacobian_projector, feats.permute(1, 0, 2).reshape(N, -1))
        lat = autograd_path.seq(self.latent_processor, z_all)
        if perms is None:
            perms = [torch.randperm(B, device=mean.device) for _ in range(S)]
        idx = torch.cat([p.to(mean.device) + i * B for i, p in enumerate(perms)])
        t = autograd_path.seq(self.mine_network, torch.cat([torch.cat([jac, lat], dim=1),
                                                            torch.cat([jac[idx], lat], dim=1)], dim=0))
        t_joint, t_marg = t[:N], t[N:]
        group = getattr(self, "data_parallel_group", None)
        if group is not None:
            mi, joint, marginal_term, t_exp = distributed.sharded_mine_statistic(t_joint, t_marg, group)
        else:
            joint = t_joint.mean()
            marginal_term = torch.logsumexp(t_marg.reshape(-1), 0) - math.log(N)
            t_exp = torch.exp(marginal_term)
            mi = joint - marginal_term
        self.running_mean = torch.where(self.running_mean == 0, t_exp,
                                        self.alpha * t_exp + (1.0 - self.alpha) * self.running_mean).reshape(())
        stats = torch.stack([mi.reshape(()), joint.reshape(()), marginal_term.reshape(()), self.running_mean])
        return torch.clamp(mi.reshape(()).expand(B), min=0.0), stats


class DiffusionActiveInference(nn.Module):
    def __init__(self, observation_dim: int, action_dim: int, latent_dim: int, config,
                 pixel_shape: Optional[Tuple[int, int, int]] = None):
        super().__init__()
        self.observation_dim, self.action_dim, self.latent_dim = observation_dim, action_dim, latent_dim
        self.config = config
        self.pixel_shape = pixel_shape
        self.is_pixel_observation = config.pixel_observation
        if self.is_pixel_observation:
            raise NotImplementedError("pixel observations (DrQ-v2 encoder / ConvDecoder) are SURVEY §8(f) 'next'")
        self.epistemic_dropout_rate = 0.2
        self.device = torch.device(config.device)
        self.raw_observation_shape = None
        self.use_epistemic = True     # b200 switch: False -> epistemic term = 0
        self._build_models()
        self.to(self.device)
        self.current_latent = None
        self.latent_trajectory = []

    # ---- construction: same order as core/active_inference.py:59-171 -------------------------
    def _build_models(self) -> None:
        cfg, H, L = self.config, self.config.hidden_dim, self.latent_dim
        self.latent_diffusion = LatentDiffusionProcess(cfg.diffusion, latent_dim=L)
        self.register_buffer("reward_mean", torch.tensor(0.0))
        self.register_buffer("reward_var", torch.tensor(1.0))
        self.register_buffer("preference_temperature", torch.tensor(cfg.preference_temperature))
        # NB the reference hard-codes observation_dim = latent_dim for the score net (:75-80)
        self.latent_score_network = LatentScoreNetwork(latent_dim=L, observation_dim=L, hidden_dim=H,
                                                       use_attention=True)
        self.policy_network = DiffusionConditionedPolicy(latent_dim=L, action_dim=self.action_dim, hidden_dim=H,
                                                         use_state_dependent_std=True)
        self.value_network = ValueNetwork(state_dim=L, hidden_dim=H, time_embed_dim=128, num_layers=3)
        self.latent_dynamics = LatentDynamicsModel(state_dim=L, action_dim=self.action_dim, hidden_dim=H, num_layers=3)
        p = self.epistemic_dropout_rate
        self.observation_decoder = nn.ModuleList([
            nn.Sequential(nn.Linear(L, H * 2), nn.LayerNorm(H * 2), nn.SiLU(), nn.Dropout(p)),
            nn.Sequential(nn.Linear(H * 2, H * 2), nn.LayerNorm(H * 2), nn.SiLU(), nn.Dropout(p)),
            nn.Sequential(nn.Linear(H * 2, H), nn.LayerNorm(H), nn.SiLU(), nn.Dropout(p)),
            nn.Linear(H, self.observation_dim)])
        self.epistemic_estimator = FunctionSpaceEpistemicEstimator(
            decoder=self.observation_decoder, latent_dim=L, observation_shape=self.observation_dim, hidden_dim=H,
            spatial_aggregator_output_dim=cfg.spatial_aggregator_output_dim, is_pixel_observation=False,
            device=self.device)
        self.reward_predictor = make_reward_predictor(L, H)
        self._heads = HeadsBundle(self.policy_network, self.latent_dynamics, self.value_network,
                                  self.reward_predictor)

    def to(self, device):
        if isinstance(device, str):
            device = torch.device(device)
        super().to(device)
        self.device = device
        self.epistemic_estimator.device = device
        return self

    def _apply(self, fn, *args, **kwargs):
        # .cuda() / .cpu() / .to() all funnel through _apply: keep `self.device` (used like the
        # reference's `config.device`, core/active_inference.py:44) in step with the parameters
        out = super()._apply(fn, *args, **kwargs)
        p = next(self.parameters(), None)
        if p is not None:
            self.device = p.device
            if hasattr(self, "epistemic_estimator"):
                self.epistemic_estimator.device = p.device
        if hasattr(self, "_heads"):
            self._heads.invalidate_packed()
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_packed()
        return out

    def invalidate_packed(self) -> None:
        """Drop every derived packed-weight cache (score net, EFE heads).  Needed after in-place
        parameter edits through `.data` (EMA / target sync idiom), which no version counter sees."""
        self.latent_score_network.invalidate_packed()
        self._heads.invalidate_packed()

    # ---- small heads ---------------------------------------------------------------------
    def decode_observation(self, latent: torch.Tensor, decode_to_pixels: bool = True) -> torch.Tensor:
        latent = latent.to(self.device)
        m = self.observation_decoder
        h1 = autograd_path.seq(m[0], latent)
        h2 = autograd_path.seq(m[1], h1) + h1
        return autograd_path.linear(autograd_path.seq(m[2], h2), m[3].weight, m[3].bias)

    def predict_reward_from_latent(self, latent: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        latent = latent.to(self.device)
        if autograd_path.needs_graph(self.reward_predictor, latent):
            out = autograd_path.seq(self.reward_predictor, latent)
        else:
            out = self._heads.head_forward(3, latent)
        return out[:, 0], torch.exp(torch.clamp(out[:, 1], min=-5, max=2))

    def predict_next_latent(self, latent: torch.Tensor, action: torch.Tensor):
        latent, action = latent.to(self.device), action.to(self.device)
        next_mean = latent + self.latent_dynamics(latent, action)      # 2z + f(z,a), SURVEY fact 10 (differentiable when recording)
        return next_mean, torch.full_like(next_mean, float(np.log(0.1)))

    def reparameterize(self, mean: torch.Tensor, logvar: torch.Tensor) -> torch.Tensor:
        std = torch.exp(0.5 * logvar)
        return mean + torch.randn_like(std) * std

    # ---- belief update (core/active_inference.py:256-312) ---------------------------------
    def update_belief_via_diffusion(self, observation: torch.Tensor,
                                    raw_observation: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        observation = observation.to(self.device)
        if observation.dim() == 1:
            observation = observation.unsqueeze(0)
        batch_size = observation.shape[0]
        keep = getattr(self, "keep_trajectory", False)
        trajectory = self.latent_diffusion.generate_latent_trajectory(
            score_network=self.latent_score_network, batch_size=batch_size, observation=observation,
            deterministic=False, return_trajectory=keep)
        self.current_latent = trajectory[-1]
        self.latent_trajectory = trajectory
        if batch_size == 1:
            latent_mean = self.current_latent.squeeze(0)
            latent_std = torch.zeros_like(latent_mean)
        else:
            latent_mean = self.current_latent.mean(dim=0)
            latent_std = self.current_latent.std(dim=0)
        with torch.no_grad():
            reconstruction_error = F.mse_loss(self.decode_observation(self.current_latent), observation)
        return {"latent": self.current_latent, "latent_mean": latent_mean, "latent_std": latent_std,
                "trajectory_length": int(self.config.diffusion.num_diffusion_steps) + 1,
                "reconstruction_error": reconstruction_error, "observation": observation,
                "raw_observation": raw_observation}

    # ---- expected free energy (core/active_inference.py:314-396) ----------------------------
    def compute_expected_free_energy_diffusion(self, latent: torch.Tensor, horizon: int = 5,
                                               num_trajectories: int = 10, num_ambiguity_samples: int = 10,
                                               *, policy_noise: Optional[torch.Tensor] = None,
                                               reparam_noise: Optional[torch.Tensor] = None,
                                               epistemic: Optional[torch.Tensor] = None):
        """Keyword-only extensions inject the reference's random draws (parity tests): policy_noise
        [K*h,B,A] (Normal.rsample), reparam_noise [K*h,B,L]; `epistemic` [K*h] overrides the MINE
        values.  With `self.use_epistemic` the estimator is evaluated per (k,t) on the rollout's
        next-latent means exactly as the reference orders it, which needs one extra dynamics pass."""
        latent = latent.to(self.device)
        B, K, h = latent.shape[0], num_trajectories, horizon
        dev = latent.device
        drawn_here = policy_noise is None and reparam_noise is None and epistemic is None
        if policy_noise is None:
            policy_noise = torch.randn(K * h, B, self.action_dim, device=dev)
        if reparam_noise is None:
            reparam_noise = torch.randn(K * h, B, self.latent_dim, device=dev)
        cfg = {"epistemic_weight": float(self.config.epistemic_weight),
               "pragmatic_weight": float(self.config.pragmatic_weight),
               "consistency_weight": float(self.config.consistency_weight),
               "discount_factor": float(self.config.discount_factor)}
        metrics: Dict[str, float] = {}
        if drawn_here and self._efe_graph_ok(latent, B):
            try:
                return self._efe_graphed(latent, h, K, num_ambiguity_samples, cfg)
            except RuntimeError as e:          # an operation that cannot be captured on this torch build
                import warnings
                warnings.warn(f"EFE CUDA-graph capture failed ({e}); launching directly from now on")
                self.efe_graph = False
        if epistemic is None and self.use_epistemic:
            epistemic, stats = self._epistemic_sequence(latent, h, K, policy_noise, reparam_noise,
                                                        num_ambiguity_samples)
            metrics = self.epistemic_estimator.metrics_from(stats) if stats is not None else {}
        if any(autograd_path.needs_graph(m, latent) for m in (self.policy_network, self.latent_dynamics,
                                                               self.value_network, self.reward_predictor)):
            # a graph is being recorded (policy training, agents/state_agent.py:165-177): differentiable
            # evaluation of the same rollout, every Linear on aid_gemm_nt
            efe, first_action, prag, cons = autograd_path.efe_rollout(
                self, latent, h, K, policy_noise, reparam_noise, epistemic)
        else:
            efe, first_action, prag, cons = self._heads.efe_rollout(
                latent, h, K, cfg, self.preference_temperature, policy_noise, reparam_noise, epistemic)
        self.last_first_action = first_action
        if epistemic is not None:
            last = epistemic.view(K, h)[:, -1]
            epi_mean = last.clamp(min=0.0).mean() if epistemic.numel() else torch.zeros((), device=dev)
        else:
            epi_mean = torch.zeros((), device=dev)
        info = {"epistemic_mean": epi_mean, "pragmatic_mean": prag.mean(), "consistency_mean": cons.mean(),
                "num_trajectories": num_trajectories, "horizon": horizon, **metrics}
        return efe, info

    def _epistemic_sequence(self, latent, h, K, policy_noise, reparam_noise, num_samples):
        """MINE value for every (k,t) in rollout order.  The estimator needs the next-latent mean of
        each step, so the rollout is replayed step by step through the stand-alone head forwards.  No
        host read inside the loop: the metrics of the last (k,t) -- what the reference's loop leaves in
        its dict -- are converted once at the end."""
        vals, stats = [], None
        std = math.exp(0.5 * math.log(0.1))
        A = self.action_dim
        est = self.epistemic_estimator
        if (self.epistemic_grouped and K > 1 and latent.is_cuda and est.fused and not est.training
                and getattr(est, "data_parallel_group", None) is None and latent.shape[0] > 0):
            return self._epistemic_sequence_grouped(latent, h, K, policy_noise, reparam_noise, num_samples)
        with torch.no_grad():
            for k in range(K):
                cur = latent
                for t in range(h):
                    d = k * h + t
                    out = self._heads.head_forward(0, cur)
                    action = out[:, :A] + torch.exp(torch.clamp(out[:, A:], -20, 2)) * policy_noise[d]
                    mean, logvar = self.predict_next_latent(cur, action)
                    e, stats = self.epistemic_estimator.forward_device(mean, logvar, num_samples)
                    vals.append(e[0])
                    cur = mean + reparam_noise[d] * std
        return torch.stack(vals), stats

    # The K rollouts are independent given the noise: per step t the K estimator evaluations of the reference's
    # (k, t) loop are ONE grouped launch sequence over K*B rows (h sequences instead of K*h; every evaluation
    # still sees exactly its own B rows, its own draws and its own permutations), and the running mean is
    # advanced afterwards through the K*h values in the reference's k-major order.
    epistemic_grouped = True

    def _epistemic_sequence_grouped(self, latent, h, K, policy_noise, reparam_noise, num_samples):
        est = self.epistemic_estimator
        B, A, L = latent.shape[0], self.action_dim, self.latent_dim
        std = math.exp(0.5 * math.log(0.1))
        pn = policy_noise.view(K, h, B, A)
        rn = reparam_noise.view(K, h, B, L)
        per_step = []
        with torch.no_grad():
            cur = latent.repeat(K, 1)                                            # row k*B + b
            for t in range(h):
                out = self._heads.head_forward(0, cur)
                action = out[:, :A] + torch.exp(torch.clamp(out[:, A:], -20, 2)) * pn[:, t].reshape(K * B, A)
                mean, logvar = self.predict_next_latent(cur, action)
                per_step.append(est.forward_grouped(mean, logvar, num_samples, K))      # [K, 4]
                cur = mean + rn[:, t].reshape(K * B, L) * std
            gs = torch.stack(per_step, dim=1)                                    # [K, h, 4]: (k, t) in k-major order
            est.apply_running_mean(gs[:, :, 3])
            vals = torch.clamp(gs[:, :, 0], min=0.0).reshape(K * h)
            last = gs[K - 1, h - 1]
            stats = torch.stack([last[0], last[1], last[2], est.running_mean.reshape(())])
        return vals, stats

    # Small batches -- act() scores ONE observation with K = 10 rollouts x horizon 5 = 50 estimator
    # evaluations of ~40 launches each, plus the rollout: ~2,300 launches whose issue time (Python + ctypes +
    # driver, ~14 us each), not their execution, set the latency (32 ms).  Without injected draws and without
    # autograd recording the whole evaluation is captured once per (batch, K, h, S) and replayed.
    efe_graph = "auto"
    efe_graph_max_batch = 256

    def _efe_graph_ok(self, latent: torch.Tensor, B: int) -> bool:
        if not (self.efe_graph is True or (self.efe_graph == "auto" and 0 < B <= self.efe_graph_max_batch)):
            return False
        if not latent.is_cuda or B == 0 or torch.cuda.is_current_stream_capturing():
            return False
        if any(autograd_path.needs_graph(m, latent) for m in (self.policy_network, self.latent_dynamics,
                                                               self.value_network, self.reward_predictor)):
            return False
        est = self.epistemic_estimator
        if self.use_epistemic and not (est.fused and not est.training
                                       and getattr(est, "data_parallel_group", None) is None):
            return False          # the unfused / training-mode / sharded estimator rebinds its running mean
        return True

    def _efe_graphed(self, latent, h, K, S, cfg):
        dev = latent.device
        est = self.epistemic_estimator
        B = latent.shape[0]
        graphs = self.__dict__.setdefault("_efe_graphs", {})
        key = (dev, _lib.operand_type(), tuple(latent.shape), h, K, S, bool(self.use_epistemic),
               tuple(sorted(cfg.items())))

        def guard():
            # device buffers the captured launches point at: re-packed weights / moved buffers -> capture again
            return (self._heads.packed_weights().data_ptr(), self.preference_temperature.data_ptr(),
                    est.running_mean.data_ptr(), est.fused_packed(dev).data_ptr() if self.use_epistemic else 0)

        gams = self.__dict__.setdefault("_efe_discounts", {})

        def body(lat):
            pn = torch.randn(K * h, B, self.action_dim, device=dev)
            rn = torch.randn(K * h, B, self.latent_dim, device=dev)
            if not self.use_epistemic:
                efe, first, prag, cons = self._heads.efe_rollout(lat, h, K, cfg, self.preference_temperature, pn, rn, None)
                return efe, first, prag.mean(), cons.mean(), torch.zeros((), device=dev), None
            # The estimator sequences and the rollout only meet in the final sum (the epistemic term is one
            # batch-constant scalar per (k, t): G_k += gamma^t * epistemic_weight * e[k, t], efe = mean_k G_k), so
            # they run side by side: the estimator on a forked stream (its own workspaces: the caches are keyed by
            # stream), the rollout here, joined before the scalar is added.
            gkey = (dev, float(cfg["discount_factor"]), h)
            if gkey not in gams:                 # first (uncaptured) pass: a host->device copy cannot be captured
                gams[gkey] = torch.tensor([cfg["discount_factor"] ** t for t in range(h)], dtype=torch.float32, device=dev)
            main = torch.cuda.current_stream(dev)
            fork = torch.cuda.Stream(device=dev)
            fork.wait_stream(main)
            with torch.cuda.stream(fork):
                epi, stats = self._epistemic_sequence(lat, h, K, pn, rn, S)
                term = float(cfg["epistemic_weight"]) * (epi.view(K, h) * gams[gkey]).sum(dim=1).mean()
                last = epi.view(K, h)[:, -1].clamp(min=0.0).mean()
            efe, first, prag, cons = self._heads.efe_rollout(lat, h, K, cfg, self.preference_temperature, pn, rn, None)
            main.wait_stream(fork)
            for tns in (pn, rn, lat):                 # allocated on `main`, read on `fork`
                tns.record_stream(fork)
            for tns in (epi, stats, term, last):      # allocated on `fork`, read on `main`
                tns.record_stream(main)
            return efe + term, first, prag.mean(), cons.mean(), last, stats

        g = graphs.get(key)
        if g is not None and g["guard"] != guard():
            g = None
        if g is None:
            g = {"latent": latent.detach().clone()}
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            # the warm-up pass must leave no trace: generator state and the estimator's running mean are put back
            rng = torch.cuda.get_rng_state(dev)
            rm = est.running_mean.detach().clone()
            with torch.cuda.stream(side):
                body(g["latent"])
                est.running_mean.data.copy_(rm)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.set_rng_state(rng, dev)
            g["guard"] = guard()
            g["graph"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g["graph"]):
                g["out"] = body(g["latent"])
            if len(graphs) >= 8:
                graphs.pop(next(iter(graphs)))
            graphs[key] = g
        g["latent"].copy_(latent, non_blocking=True)
        g["graph"].replay()
        efe, first, prag, cons, last, stats = g["out"]
        self.last_first_action = first.clone()
        metrics = est.metrics_from(stats) if stats is not None else {}
        info = {"epistemic_mean": last.clone(), "pragmatic_mean": prag.clone(), "consistency_mean": cons.clone(),
                "num_trajectories": K, "horizon": h, **metrics}
        return efe.clone(), info

    # set by the agent: `agent.active_inference.epistemic_optimizer = Adam(...)` (agents/base_agent.py:134-139)
    epistemic_optimizer = None

    def train_epistemic_estimator(self, latents: torch.Tensor, actions: torch.Tensor, next_latents: torch.Tensor,
                                  **draws):
        """One MINE training step (core/active_inference.py:420-445, called by the agents every 5th
        `train_step`, agents/state_agent.py:217-220): loss = -mean(MI estimate) on the predicted
        next-latent distribution, gradient clipping at `config.gradient_clip`, one optimizer step.
        `next_latents` is accepted and unused, as in the reference.  Returns (mi, metrics).  Keyword
        arguments are forwarded to the estimator (parity tests inject its draws)."""
        if self.epistemic_optimizer is None:
            raise RuntimeError("train_epistemic_estimator: set `epistemic_optimizer` first (the reference's agents "
                               "assign it, agents/base_agent.py:134-139)")
        latents, actions = latents.to(self.device), actions.to(self.device)
        next_mean, next_logvar = self.predict_next_latent(latents, actions)
        mi_estimate, metrics = self.epistemic_estimator(next_mean, next_logvar, **draws)
        loss = -mi_estimate.mean()
        self.epistemic_optimizer.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.epistemic_estimator.parameters(), self.config.gradient_clip)
        self.epistemic_optimizer.step()
        self._heads.invalidate_packed()         # the dynamics head may be among the optimised parameters
        return mi_estimate.mean().item(), metrics

    def compute_epistemic_value(self, next_latent_mean, next_latent_logvar, num_samples: int = 5):
        with torch.no_grad():
            return self.epistemic_estimator(next_latent_mean.to(self.device), next_latent_logvar.to(self.device),
                                            num_samples)

    # ---- act (core/active_inference.py:478-531) -------------------------------------------
    @torch.no_grad()
    def act(self, observation: torch.Tensor, deterministic: bool = False,
            raw_observation: Optional[torch.Tensor] = None):
        """Runs without recording a graph, as the agents call it (agents/state_agent.py:94-95: under
        `torch.no_grad()`): belief update, EFE and policy head on the fused kernels; one device->host
        transfer for every scalar of `info`."""
        observation = observation.to(self.device)
        if observation.dim() == 1:
            observation = observation.unsqueeze(0)
        belief = self.update_belief_via_diffusion(observation, raw_observation)
        latent = belief["latent"]
        efe, efe_info = self.compute_expected_free_energy_diffusion(latent, horizon=self.config.efe_horizon)
        action, log_prob, policy_dist = self.policy_network(latent, deterministic=deterministic)
        # one device->host transfer for all the scalars the reference reads with .item() (:511-528)
        scalars = torch.stack([efe.mean(), log_prob.mean(), policy_dist.entropy().sum(dim=-1).mean()] +
                              [v.float().reshape(()) for v in efe_info.values() if torch.is_tensor(v)]).cpu()
        action = action.cpu()
        if action.dim() == 2 and action.shape[0] == 1:
            action = action.squeeze(0)
        names = [k for k, v in efe_info.items() if torch.is_tensor(v)]
        info = {**belief, "expected_free_energy": float(scalars[0]), "action_log_prob": float(scalars[1]),
                "policy_entropy": float(scalars[2]),
                **{k: v for k, v in efe_info.items() if not torch.is_tensor(v)},
                **{k: float(scalars[3 + i]) for i, k in enumerate(names)}}
        return action, info

    # ---- training loss (core/active_inference.py:533-636, 709-771) --------------------------
    def _compute_latent_kl(self, latent: torch.Tensor, prior_latent: torch.Tensor) -> torch.Tensor:
        return 0.5 * torch.sum((latent - prior_latent) ** 2, dim=-1)

    # Training path of compute_diffusion_elbo: "native" = hand-written forward / backward / double
    # backward kernels for the trunk (csrc/train.inc) with `training_operand` tensor-core operands
    # ("f16": TF32-class, the rel-1e-3 contract; "bf16": stated bf16 bound); "autograd" = the torch
    # graph over aid_gemm_nt at autograd_path.PRECISION (any dims, any order of differentiation).
    training_path = "native"
    training_operand = "f16"

    def _native_training(self) -> bool:
        return self.training_path == "native" and train_native.supported(self.latent_score_network)

    ELBO_KEYS = ["reconstruction_loss", "kl_loss", "score_matching_loss", "elbo", "reward_loss", "grad_penalty",
                 "mean_time", "loss_weight_mean"]

    def compute_diffusion_elbo(self, observations: torch.Tensor, rewards: torch.Tensor,
                               latents: Optional[torch.Tensor] = None,
                               raw_observations: Optional[torch.Tensor] = None, *,
                               t: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
                               prior_eps: Optional[torch.Tensor] = None):
        """-ELBO as written in the reference (signs included).  Keyword-only `t`,
        `noise`, `prior_eps` inject the reference's draws (rand / randn_like / randn_like)."""
        loss, vals = self.elbo_device(observations, rewards, latents, raw_observations, t=t, noise=noise,
                                      prior_eps=prior_eps)
        self._join_time_importance()
        vals = vals.cpu()                                                                   # one D2H
        return loss, {k: float(v) for k, v in zip(self.ELBO_KEYS, vals)}

    def elbo_device(self, observations: torch.Tensor, rewards: torch.Tensor,
                    latents: Optional[torch.Tensor] = None, raw_observations: Optional[torch.Tensor] = None, *,
                    t: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
                    prior_eps: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """`compute_diffusion_elbo` without its device->host read: returns (loss, metrics[8] in
        ELBO_KEYS order, both on the device).  With t drawn here there is no host synchronisation
        anywhere in the loss, so forward + backward can be captured in a CUDA graph
        (train_graph.GraphedElboStep)."""
        observations, rewards = observations.to(self.device), rewards.to(self.device)
        B, dev = observations.shape[0], self.device
        if latents is None:
            latents = self.update_belief_via_diffusion(observations, raw_observations)["latent"]
        # decoder / reward-head terms: their gradients are dropped by the reference's training step
        # (agents/state_agent.py:225); `elbo_score_only` evaluates them without recording a graph
        side = torch.no_grad() if getattr(self, "elbo_score_only", False) else contextlib.nullcontext()
        with side:
            recon = F.mse_loss(self.decode_observation(latents), observations)
        # t drawn here lies in [0,1): the score net's continuous-time branch (score_networks.py:121)
        # is known without reading the tensor back; an injected t is inspected as the reference does
        continuous = True if t is None else None
        if t is None:
            t = self._importance_sample_time(B, dev) if hasattr(self, "time_importance_weights") \
                else torch.rand(B, device=dev)
        if noise is None:
            noise = torch.randn_like(latents)
        diff = self.latent_diffusion
        noisy, _, info_q = diff.continuous_q_sample(latents, t, noise)
        if continuous is None:
            continuous = bool(t.max() <= 1.0 and t.min() >= 0.0)
        net = self.latent_score_network
        if self._native_training():
            # Native path (csrc/train.inc through train_native.trunk): ONE forward serves the
            # score-matching term and the gradient penalty (both evaluate s_theta on the same numbers,
            # :584 and :717); s and g = d(sum s)/dz come out of one autograd node whose backward runs
            # the hand-written double backward.  The conditioning embedding and the attention fold stay
            # in torch autograd (small GEMMs, fp32-class operands: their gradients carry no stream scale).
            with autograd_path.precision("bf16" if self.training_operand == "bf16" else "bf16x3"):
                cond, time_weight = autograd_path.score_cond_embedding(net, t, observations, B, continuous)
                folds = autograd_path.fold_attention(net)
            pred, g = train_native.trunk(net, noisy, cond, time_weight, folds, operand=self.training_operand)
            gp = torch.mean((g.norm(2, dim=1) - 1.0) ** 2)
        else:
            # the conditioning path (time embeddings, observation encoder, adaLN modulations) is the
            # same for the score-matching forward and the gradient penalty's forward: evaluated once
            mod, time_weight = autograd_path.score_conditioning(net, t, observations, B, continuous)
            folds = autograd_path.fold_attention(net)       # W_o W_v per block, shared too
            pred = autograd_path.score_from_conditioning(net, noisy, mod, time_weight, folds)
            gp = self._compute_gradient_penalty(noisy, t, observations, continuous, conditioning=(mod, time_weight, folds))
        sigma = info_q["sigma"]
        true_score = -noise / (sigma + 1e-8)
        w = diff.compute_loss_weight(t)
        per_sample = w.view(-1) * torch.sum((pred - true_score) ** 2, dim=1)
        sm = per_sample.mean()
        if prior_eps is None:
            prior = diff.sample_latent_prior(B, dev)
        else:
            prior = diff.latent_prior_mean.unsqueeze(0) + torch.exp(diff.latent_prior_log_std).unsqueeze(0) * prior_eps
        kl = self._compute_latent_kl(latents, prior).mean()
        # data parallel: the KL weight reads the mean t of the GLOBAL batch (one scalar all-reduce), so
        # the sharded step is the single-process step on the concatenated batch
        group = getattr(self, "data_parallel_group", None)
        t_mean = distributed.allreduce_mean(t.mean(), group) if group is not None else t.mean()
        klw = torch.exp(-5.0 * t_mean)
        with side:
            pr = autograd_path.seq(self.reward_predictor, latents)
        r_std = torch.exp(torch.clamp(pr[:, 1], min=-5, max=2))
        # -log N(r; m, s) written out (torch.distributions validates its arguments with a host sync)
        # (same expression as Normal.log_prob)
        log_prob = -((rewards - pr[:, 0]) ** 2) / (2 * r_std ** 2) - r_std.log() - math.log(math.sqrt(2 * math.pi))
        rl = -log_prob.mean()
        c = self.config
        elbo = -recon + c.kl_weight * kl * klw + c.diffusion_weight * sm + 0.1 * gp - c.reward_weight * rl
        self._update_time_importance(t, per_sample.detach())
        vals = torch.stack([recon, kl, sm, elbo, rl, gp, t.mean(), w.mean()]).detach()
        if group is not None:
            vals = distributed.allreduce_mean(vals, group)       # the metrics of the global batch
        return -elbo, vals

    def compute_lambda_returns(self, rewards: torch.Tensor, values: torch.Tensor, next_values: torch.Tensor,
                               dones: torch.Tensor, lambda_: float = 0.95, n_steps: int = 5,
                               exclude_immediate_rewards: bool = False) -> torch.Tensor:
        """Reference :638-707 (value targets of `train_step`, agents/state_agent.py:198-205): an
        O(B n^2) Python loop over 0-dim tensors there, one kernel here with the same fp32 operation
        order.  `values` is accepted and ignored, as in the reference."""
        return _lib.lambda_returns(rewards, next_values, dones, self.config.discount_factor, lambda_, n_steps,
                                   exclude_immediate_rewards)

    def _compute_gradient_penalty(self, noisy_latents, t, observations, continuous: Optional[bool] = None,
                                  conditioning=None) -> torch.Tensor:
        x = noisy_latents.detach().requires_grad_(True)
        if conditioning is None:
            conditioning = autograd_path.score_conditioning(self.latent_score_network, t, observations,
                                                            x.shape[0], continuous)
        s = autograd_path.score_from_conditioning(self.latent_score_network, x, *conditioning)
        g = torch.autograd.grad(outputs=s.sum(), inputs=x, create_graph=True, retain_graph=True)[0]
        return torch.mean((g.norm(2, dim=1) - 1.0) ** 2)

    def _importance_sample_time(self, batch_size: int, device: torch.device) -> torch.Tensor:
        if not hasattr(self, "time_importance_weights"):
            self.time_importance_weights = torch.ones(100, device=device)
        self._join_time_importance()
        probs = F.softmax(self.time_importance_weights, dim=0)
        if getattr(self, "graph_safe_time_sampling", False) or torch.cuda.is_current_stream_capturing():
            # torch.multinomial validates `probs` with a host read; inside a CUDA-graph capture (or
            # when `graph_safe_time_sampling` is set) the same categorical draw is taken by inverse
            # CDF: same distribution, a different use of the generator than torch.multinomial
            cdf = torch.cumsum(probs, dim=0)
            idx = torch.searchsorted(cdf, torch.rand(batch_size, device=device) * cdf[-1]).clamp_(max=probs.numel() - 1)
        else:
            idx = torch.multinomial(probs, batch_size, replacement=True)
        return (idx.float() + torch.rand(batch_size, device=device)) / 100.0

    def _update_time_importance(self, t: torch.Tensor, loss: torch.Tensor) -> None:
        """Sequential per-sample EMA over 100 bins (:750-771).  The reference does B x 3 host syncs in
        a Python loop; here one kernel (one thread per bin) replays the batch in order with the same
        double arithmetic and fp32 rounding -- bit-identical, no host round trip."""
        if not hasattr(self, "time_importance_weights"):
            self.time_importance_weights = torch.ones(100, device=t.device)
        if loss.dim() > 1:
            loss = loss.view(loss.shape[0], -1).sum(dim=1)
        w = self.time_importance_weights
        if w.device != t.device or w.dtype != torch.float32 or not w.is_contiguous():
            w = w.to(t.device, torch.float32).contiguous()
            self.time_importance_weights = w
        # The EMA only feeds the NEXT call's time sampling, and its per-bin chains are sequential
        # (1.8 ms at 32,768 samples once importance sampling has concentrated the batch in a few
        # bins), so it runs on a side stream beside the backward pass; `_join_time_importance` orders
        # it before anything that reads the weights.
        _lib.require_cuda(t, loss)
        main = torch.cuda.current_stream(t.device)
        side = getattr(self, "_ti_stream", None)
        if side is None or side.device != t.device:
            side = self._ti_stream = torch.cuda.Stream(device=t.device)
        t, loss = t.detach(), loss.detach()
        group = getattr(self, "data_parallel_group", None)
        if group is not None:
            t, loss = distributed.gather_time_loss(t, loss, group)   # identical EMA on every rank
        side.wait_stream(main)
        with torch.cuda.stream(side):
            _lib.time_importance_update(t, loss, w)
        if not torch.cuda.is_current_stream_capturing():
            t.record_stream(side)
            loss.record_stream(side)
        self._ti_pending = True

    def _join_time_importance(self) -> None:
        """Make the current stream wait for a pending time-importance update."""
        if getattr(self, "_ti_pending", False):
            torch.cuda.current_stream(self._ti_stream.device).wait_stream(self._ti_stream)
            self._ti_pending = False
