"""Drop-in EFE heads: `DiffusionConditionedPolicy`, `ValueNetwork`, `LatentDynamicsModel`
(models/policy_networks.py:12-151, value_networks.py:9-60, dynamics_models.py:9-67) plus the
packed-weight cache shared by the expected-free-energy rollout.

The modules reproduce the reference's parameter registration order, initialisation calls and
`state_dict` keys; their forwards run through `aid_head_forward` (tcgen05 GEMMs + fused
LayerNorm kernels).  Only the configuration `DiffusionActiveInference` instantiates is
supported (3 trunk layers, state-dependent std, no squashing, residual dynamics).
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributions as dist
import torch.nn as nn

from . import _lib
from .score_network import SinusoidalPositionEmbeddings

POLICY_KEYS = [
    "latent_encoder.0.weight", "latent_encoder.0.bias", "latent_encoder.1.weight", "latent_encoder.1.bias",
    "latent_encoder.3.weight", "latent_encoder.3.bias",
    "trunk.0.weight", "trunk.0.bias", "trunk.1.weight", "trunk.1.bias",
    "trunk.3.weight", "trunk.3.bias", "trunk.4.weight", "trunk.4.bias",
    "trunk.6.weight", "trunk.6.bias", "trunk.7.weight", "trunk.7.bias",
    "mean_head.0.weight", "mean_head.0.bias", "mean_head.2.weight", "mean_head.2.bias",
    "log_std_head.0.weight", "log_std_head.0.bias", "log_std_head.2.weight", "log_std_head.2.bias",
]
DYNAMICS_KEYS = [f"network.{i}.{p}" for i in (0, 1, 3, 4, 6, 7, 9) for p in ("weight", "bias")]
VALUE_KEYS = ["time_embed.0.freq_scale", "time_embed.1.weight", "time_embed.1.bias"] + DYNAMICS_KEYS
REWARD_KEYS = [f"{i}.{p}" for i in (0, 1, 3, 5) for p in ("weight", "bias")]


class DiffusionConditionedPolicy(nn.Module):
    """p_phi(pi|z) = N(mu_phi(z), Sigma_phi(z)) — models/policy_networks.py:12-151."""

    def __init__(self, latent_dim: int, action_dim: int, hidden_dim: int = 256, num_layers: int = 3,
                 log_std_min: float = -20, log_std_max: float = 2, use_state_dependent_std: bool = True,
                 squash_output: bool = False):
        super().__init__()
        if num_layers != 3 or not use_state_dependent_std or squash_output:
            raise NotImplementedError("b200 policy supports the configuration the reference instantiates: "
                                      "num_layers=3, state-dependent std, no squashing")
        self.latent_dim, self.action_dim, self.hidden_dim = latent_dim, action_dim, hidden_dim
        self.log_std_min, self.log_std_max = log_std_min, log_std_max
        self.use_state_dependent_std, self.squash_output = use_state_dependent_std, squash_output
        self.latent_encoder = nn.Sequential(nn.Linear(latent_dim, hidden_dim), nn.LayerNorm(hidden_dim), nn.ReLU(),
                                            nn.Linear(hidden_dim, hidden_dim))
        layers: List[nn.Module] = []
        for _ in range(num_layers):
            layers += [nn.Linear(hidden_dim, hidden_dim), nn.LayerNorm(hidden_dim), nn.ReLU()]
        self.trunk = nn.Sequential(*layers)
        self.mean_head = nn.Sequential(nn.Linear(hidden_dim, hidden_dim // 2), nn.ReLU(),
                                       nn.Linear(hidden_dim // 2, action_dim))
        self.log_std_head = nn.Sequential(nn.Linear(hidden_dim, hidden_dim // 2), nn.ReLU(),
                                          nn.Linear(hidden_dim // 2, action_dim))
        self._initialize_weights()
        self._owner = None  # set by HeadsBundle

    def _initialize_weights(self) -> None:
        # same calls, same order as models/policy_networks.py:76-92 (RNG stream parity)
        torch.nn.init.orthogonal_(self.mean_head[-1].weight, gain=torch.tensor(1.0))
        nn.init.zeros_(self.mean_head[-1].bias)
        torch.nn.init.orthogonal_(self.log_std_head[-1].weight, gain=torch.tensor(1.0))
        nn.init.zeros_(self.log_std_head[-1].bias)
        for m in [self.latent_encoder, self.trunk, self.mean_head[:-1]]:
            for layer in m.modules():
                if isinstance(layer, nn.Linear):
                    nn.init.xavier_uniform_(layer.weight)
                    if layer.bias is not None:
                        nn.init.zeros_(layer.bias)

    def forward(self, z: torch.Tensor, deterministic: bool = False
                ) -> Tuple[torch.Tensor, torch.Tensor, dist.Distribution]:
        from . import autograd_path
        A = self.action_dim
        if autograd_path.needs_graph(self, z):      # training: differentiable evaluation, GEMMs on aid_gemm_nt
            _, mean, log_std, _ = autograd_path.policy_forward(self, z, None)
        else:
            out = _bundle_of(self).head_forward(0, z)
            mean, log_std = out[:, :A], out[:, A:]
        log_std = torch.clamp(log_std, self.log_std_min, self.log_std_max)
        distribution = dist.Normal(mean, torch.exp(log_std))
        action = mean if deterministic else distribution.rsample()
        return action, distribution.log_prob(action).sum(dim=-1), distribution

    def get_policy_entropy(self, z: torch.Tensor) -> torch.Tensor:
        _, _, d = self.forward(z, deterministic=True)
        return d.entropy().sum(dim=-1)


def _ln_relu_stack(input_dim: int, hidden_dim: int, num_layers: int, out_dim: int) -> nn.Sequential:
    layers: List[nn.Module] = []
    for i in range(num_layers):
        layers += [nn.Linear(input_dim if i == 0 else hidden_dim, hidden_dim), nn.LayerNorm(hidden_dim), nn.ReLU()]
    layers.append(nn.Linear(hidden_dim, out_dim))
    return nn.Sequential(*layers)


class ValueNetwork(nn.Module):
    """V(s, t) — models/value_networks.py:9-60."""

    def __init__(self, state_dim: int, hidden_dim: int = 256, time_embed_dim: int = 128, num_layers: int = 3):
        super().__init__()
        if num_layers != 3:
            raise NotImplementedError("b200 ValueNetwork supports num_layers=3 (the reference's instantiation)")
        self.state_dim, self.hidden_dim, self.time_embed_dim = state_dim, hidden_dim, time_embed_dim
        self.time_embed = nn.Sequential(SinusoidalPositionEmbeddings(time_embed_dim),
                                        nn.Linear(time_embed_dim, time_embed_dim), nn.ReLU())
        self.network = _ln_relu_stack(state_dim + time_embed_dim, hidden_dim, num_layers, 1)
        self._owner = None

    def forward(self, state: torch.Tensor, time: torch.Tensor) -> torch.Tensor:
        from . import autograd_path
        if autograd_path.needs_graph(self, state):
            return autograd_path.value_forward(self, state, time)
        return _bundle_of(self).head_forward(2, state, time)


class LatentDynamicsModel(nn.Module):
    """f(s, a) -> s' — models/dynamics_models.py:9-67 (residual)."""

    def __init__(self, state_dim: int, action_dim: int, hidden_dim: int = 256, num_layers: int = 3,
                 residual: bool = True):
        super().__init__()
        if num_layers != 3 or not residual:
            raise NotImplementedError("b200 LatentDynamicsModel supports num_layers=3, residual=True")
        self.state_dim, self.action_dim, self.hidden_dim, self.residual = state_dim, action_dim, hidden_dim, residual
        self.network = _ln_relu_stack(state_dim + action_dim, hidden_dim, num_layers, state_dim)
        nn.init.uniform_(self.network[-1].weight, -1e-3, 1e-3)
        nn.init.zeros_(self.network[-1].bias)
        self._owner = None

    def forward(self, state: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
        from . import autograd_path
        if autograd_path.needs_graph(self, state, action):
            return autograd_path.dynamics_forward(self, state, action)
        return _bundle_of(self).head_forward(1, state, action)


def make_reward_predictor(latent_dim: int, hidden_dim: int) -> nn.Sequential:
    """core/active_inference.py:160-167."""
    return nn.Sequential(nn.Linear(latent_dim, hidden_dim), nn.LayerNorm(hidden_dim), nn.ReLU(),
                         nn.Linear(hidden_dim, hidden_dim // 2), nn.ReLU(), nn.Linear(hidden_dim // 2, 2))


F16_MAX_HORIZON = 10


def _bundle_of(module: nn.Module) -> "HeadsBundle":
    owner = getattr(module, "_owner", None)
    if owner is None:
        raise RuntimeError(f"{type(module).__name__} (b200) runs as part of a HeadsBundle / "
                           "DiffusionActiveInference: the four heads share one packed-weight cache")
    return owner


class HeadsBundle:
    """Packed-weight cache + workspace for (policy, dynamics, value, reward).  Not an nn.Module:
    it owns no parameters, only the derived tcgen05 operand tiles."""

    def __init__(self, policy: DiffusionConditionedPolicy, dynamics: LatentDynamicsModel,
                 value: ValueNetwork, reward: nn.Sequential):
        self.policy, self.dynamics, self.value, self.reward = policy, dynamics, value, reward
        for m in (policy, dynamics, value):
            object.__setattr__(m, "_owner", self)
        self._cache = _lib.PackedCache()

    def dims(self) -> _lib.AidHeadsDims:
        return _lib.AidHeadsDims(self.policy.latent_dim, self.policy.action_dim, self.policy.hidden_dim,
                                 self.value.time_embed_dim)

    def _params(self) -> List[torch.Tensor]:
        out: List[torch.Tensor] = []
        for module, keys in ((self.policy, POLICY_KEYS), (self.dynamics, DYNAMICS_KEYS),
                             (self.value, VALUE_KEYS), (self.reward, REWARD_KEYS)):
            named = dict(module.named_parameters())
            out += [named[k] for k in keys]
        return out

    def invalidate_packed(self) -> None:
        """Required after in-place `.data` edits of head parameters (see _lib.PackedCache)."""
        self._cache.invalidate()

    def packed_weights(self, verify: bool = False) -> torch.Tensor:
        params = self._params()
        dev = _lib.require_cuda(*params)

        def build():
            l, d = _lib.lib(), self.dims()
            nbytes = l.aid_heads_packed_bytes(ctypes.byref(d))
            if nbytes == 0:
                _lib.check(-1, "aid_heads_packed_bytes")
            keep = [_lib.f32c(p.detach()) for p in params]
            table = (ctypes.c_void_p * len(keep))(*[t.data_ptr() for t in keep])
            packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                _lib.check(l.aid_heads_pack(ctypes.byref(d), table, len(keep), packed.data_ptr(), nbytes,
                                            _lib.stream_ptr(dev)), "aid_heads_pack")
            return packed

        return self._cache.get(("heads", _lib.operand_type()), params, build, verify)

    def workspace(self, batch: int, device: torch.device) -> torch.Tensor:
        l, d = _lib.lib(), self.dims()
        need = l.aid_heads_workspace_bytes(ctypes.byref(d), batch)
        if need == 0:
            _lib.check(-1, "aid_heads_workspace_bytes")
        return self._cache.workspace(need, device)

    def head_forward(self, which: int, z: torch.Tensor, aux: Optional[torch.Tensor] = None) -> torch.Tensor:
        dev = _lib.require_cuda(z, aux)
        z, aux = _lib.f32c(z.detach()), _lib.f32c(None if aux is None else aux.detach())
        B = z.shape[0]
        d = self.dims()
        width = {0: 2 * d.action_dim, 1: d.latent_dim, 2: 1, 3: 2}[which]
        out = torch.empty(B, width, dtype=torch.float32, device=dev)
        if B == 0:                      # empty batch: empty result, as the reference's modules give
            return out
        packed, ws = self.packed_weights(), self.workspace(B, dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().aid_head_forward(ctypes.byref(d), packed.data_ptr(), ws.data_ptr(), ws.numel(),
                                                   which, B, z.data_ptr(), _lib.ptr(aux), out.data_ptr(),
                                                   _lib.stream_ptr(dev)), "aid_head_forward")
        return out

    # Small batches: the K rollouts of a candidate are independent rows, so K*B rows go through ONE
    # rollout of h steps (18 GEMM launches per step) instead of K sequential ones -- the regime is
    # launch-latency bound (act(): B = 1, K = 10 in the reference's defaults).  Above this many rows the
    # kernels are already full and the sequential form keeps the workspace at B rows.
    TRAJECTORY_ROWS_MAX = 32768

    def _efe_rollout_trajectories_as_rows(self, latent, h, K, cfg, tau, policy_noise, reparam_noise, epistemic):
        """Same numbers as the K-sequential rollout: a row's arithmetic does not depend on the other rows,
        and the mean over trajectories is accumulated in the kernel's order (total += G_k / K)."""
        B, A, L = latent.shape[0], policy_noise.shape[-1], latent.shape[1]
        rows = latent.repeat(K, 1)                                                      # row k*B + b
        pn = policy_noise.view(K, h, B, A).permute(1, 0, 2, 3).reshape(h, K * B, A).contiguous()
        rn = reparam_noise.view(K, h, B, L).permute(1, 0, 2, 3).reshape(h, K * B, L).contiguous()
        g, first, prag, cons = self.efe_rollout(rows, h, 1, cfg, tau, pn, rn, None)
        g = g.view(K, B)
        if epistemic is not None:
            # batch-constant scalar per (k, t): sum_t gamma^t * epistemic_weight * e[k, t] on top of G_k
            # (cached per device / discount / horizon: a host->device copy cannot be captured in a CUDA graph)
            gams = self.__dict__.setdefault("_discount_tables", {})
            gkey = (g.device, float(cfg["discount_factor"]), h)
            gam = gams.get(gkey)
            if gam is None:
                gam = gams[gkey] = torch.tensor([cfg["discount_factor"] ** t for t in range(h)], dtype=torch.float32,
                                                device=g.device)
            g = g + (cfg["epistemic_weight"] * (epistemic.view(K, h) * gam).sum(dim=1)).unsqueeze(1)
        efe = torch.zeros(B, dtype=torch.float32, device=g.device)
        for k in range(K):
            efe = efe + g[k] / float(K)
        return efe, first[:B].contiguous(), prag.view(K, B), cons.view(K, B)

    def efe_rollout(self, latent: torch.Tensor, horizon: int, num_trajectories: int, cfg: Dict[str, float],
                    preference_temperature: torch.Tensor, policy_noise: torch.Tensor, reparam_noise: torch.Tensor,
                    epistemic: Optional[torch.Tensor] = None):
        """compute_expected_free_energy_diffusion's rollout (core/active_inference.py:337-378).
        Returns (efe[B], first_action[B,A], pragmatic_last[K,B], consistency_last[K,B])."""
        dev = _lib.require_cuda(latent, policy_noise, reparam_noise, epistemic, preference_temperature)
        latent = _lib.f32c(latent.detach())
        policy_noise, reparam_noise = _lib.f32c(policy_noise), _lib.f32c(reparam_noise)
        epistemic = _lib.f32c(epistemic)
        tau = _lib.f32c(preference_temperature.detach().reshape(1))
        B, d = latent.shape[0], self.dims()
        K, h = num_trajectories, horizon
        if _lib.operand_type() == "f16" and h > F16_MAX_HORIZON:
            # predict_next_latent doubles the latent every step (mean = 2 z + f(z, a)):
            # after ~14 steps |z| passes the largest fp16 number (65504) and the operand conversion would
            # saturate.  bf16 operands (8-bit exponent) have no such limit.
            raise RuntimeError(f"fp16 tensor-core operands cover rollouts of at most {F16_MAX_HORIZON} steps "
                               f"(got horizon {h}): use the bf16 operand type for longer horizons")
        assert tuple(policy_noise.shape) == (K * h, B, d.action_dim), policy_noise.shape
        assert tuple(reparam_noise.shape) == (K * h, B, d.latent_dim), reparam_noise.shape
        if K > 1 and 0 < B * K <= self.TRAJECTORY_ROWS_MAX:
            return self._efe_rollout_trajectories_as_rows(latent, h, K, cfg, tau, policy_noise, reparam_noise, epistemic)
        efe = torch.empty(B, dtype=torch.float32, device=dev)
        first = torch.empty(B, d.action_dim, dtype=torch.float32, device=dev)
        prag = torch.empty(K, B, dtype=torch.float32, device=dev)
        cons = torch.empty(K, B, dtype=torch.float32, device=dev)
        if B == 0:
            return efe, first, prag, cons
        c = _lib.AidEfeConfig(cfg["epistemic_weight"], cfg["pragmatic_weight"], cfg["consistency_weight"],
                              cfg["discount_factor"])
        packed, ws = self.packed_weights(), self.workspace(B, dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().aid_efe_rollout(
                ctypes.byref(d), packed.data_ptr(), ws.data_ptr(), ws.numel(), B, h, K, ctypes.byref(c),
                tau.data_ptr(), latent.data_ptr(), policy_noise.data_ptr(), reparam_noise.data_ptr(),
                _lib.ptr(epistemic), efe.data_ptr(), first.data_ptr(), prag.data_ptr(), cons.data_ptr(),
                _lib.stream_ptr(dev)), "aid_efe_rollout")
        return efe, first, prag, cons
