"""`CandidateScorer`: the fused public call of the hot path — observation -> reverse diffusion
(belief latent) -> expected-free-energy rollout -> (efe, first action) per candidate row.

It composes the two library calls the reference's `act()` makes in sequence
(core/active_inference.py:492,501) for arbitrary score-net observation width (the reference's
state agent hard-codes observation_dim = latent_dim) and is the unit that shards
across ranks: rows are independent, so a rank scores its slice with no collective.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from .configs import ActiveInferenceConfig, DiffusionConfig
from .diffusion import LatentDiffusionProcess
from .heads import (DiffusionConditionedPolicy, HeadsBundle, LatentDynamicsModel, ValueNetwork,
                    make_reward_predictor)
from .score_network import LatentScoreNetwork
from . import _lib


class CandidateScorer(nn.Module):
    def __init__(self, observation_dim: int, action_dim: int, config: Optional[ActiveInferenceConfig] = None):
        super().__init__()
        self.config = config or ActiveInferenceConfig()
        c = self.config
        L, H = c.latent_dim, c.hidden_dim
        self.observation_dim, self.action_dim, self.latent_dim = observation_dim, action_dim, L
        self.latent_diffusion = LatentDiffusionProcess(c.diffusion, latent_dim=L)
        self.register_buffer("preference_temperature", torch.tensor(c.preference_temperature))
        self.latent_score_network = LatentScoreNetwork(L, observation_dim, H, use_attention=True)
        self.policy_network = DiffusionConditionedPolicy(L, action_dim, H, use_state_dependent_std=True)
        self.value_network = ValueNetwork(L, H, time_embed_dim=128, num_layers=3)
        self.latent_dynamics = LatentDynamicsModel(L, action_dim, H, num_layers=3)
        self.reward_predictor = make_reward_predictor(L, H)
        self.heads = HeadsBundle(self.policy_network, self.latent_dynamics, self.value_network, self.reward_predictor)
        self.latent_diffusion.noise_source = "philox"     # in-kernel noise: no [T-1,B,L] tensor in HBM

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if hasattr(self, "heads"):
            self.heads.invalidate_packed()
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_packed()
        return out

    def invalidate_packed(self) -> None:
        """Drop the derived packed-weight caches (needed after in-place `.data` edits)."""
        self.latent_score_network.invalidate_packed()
        self.heads.invalidate_packed()

    def efe_config(self) -> Dict[str, float]:
        c = self.config
        return {"epistemic_weight": float(c.epistemic_weight), "pragmatic_weight": float(c.pragmatic_weight),
                "consistency_weight": float(c.consistency_weight), "discount_factor": float(c.discount_factor)}

    # Whole-call CUDA graph (sampler + noise draws + EFE rollout, ~2,900 launches at T=50, K=1, h=5) for
    # batches in the launch-latency-bound regime: "auto" = at most `graph_max_batch` rows.
    use_graph = "auto"
    graph_max_batch = 16384

    def _score(self, observation: torch.Tensor, h: int, K: int, epistemic: Optional[torch.Tensor]):
        B, dev = observation.shape[0], observation.device
        traj = self.latent_diffusion.generate_latent_trajectory(
            self.latent_score_network, B, observation, deterministic=False, return_trajectory=False)
        latent = traj[-1]
        policy_noise = torch.randn(K * h, B, self.action_dim, device=dev)
        reparam_noise = torch.randn(K * h, B, self.latent_dim, device=dev)
        efe, first_action, _, _ = self.heads.efe_rollout(latent, h, K, self.efe_config(), self.preference_temperature,
                                                         policy_noise, reparam_noise, epistemic)
        return efe, first_action, latent

    @torch.no_grad()
    def forward(self, observation: torch.Tensor, horizon: Optional[int] = None, num_trajectories: int = 1,
                epistemic: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """observation [B,O] (device) -> (efe [B], first_action [B,A], latent [B,L]).
        All noise is drawn on the device, one stream of standard normals per candidate row: z_T and
        the T-1 step noises inside the sampler kernels (`latent_diffusion.noise_source == "philox"`,
        the default here) and the policy / reparameterisation draws by torch."""
        h = int(horizon or self.config.efe_horizon)
        K = int(num_trajectories)
        B, dev = observation.shape[0], observation.device
        graphed = self.use_graph is True or (self.use_graph == "auto" and 0 < B <= self.graph_max_batch)
        if not graphed or torch.cuda.is_current_stream_capturing():
            return self._score(observation, h, K, epistemic)
        graphs = self.__dict__.setdefault("_graphs", {})
        packed = (self.latent_score_network.packed_weights().data_ptr(), self.heads.packed_weights().data_ptr())
        key = (dev, _lib.operand_type(), tuple(observation.shape), h, K, epistemic is None,
               self.latent_diffusion.noise_source, int(self.latent_diffusion.row_offset))
        g = graphs.get(key)
        if g is not None and g["packed"] != packed:
            g = None                    # weights were re-packed into new buffers: capture again
        if g is None:
            g = {"packed": packed, "obs": observation.clone(),
                 "epi": None if epistemic is None else epistemic.clone()}
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            prev = self.latent_diffusion.use_graph
            self.latent_diffusion.use_graph = False          # the sampler is captured as part of THIS graph
            # the warm-up pass must not advance the noise streams: a captured call then draws what the
            # directly launched call would have drawn from the same generator state
            rng = torch.cuda.get_rng_state(dev)
            philox = self.latent_diffusion.philox_state(dev).clone() \
                if self.latent_diffusion.noise_source == "philox" else None
            try:
                with torch.cuda.stream(side):
                    self._score(g["obs"], h, K, g["epi"])    # warm-up outside the capture
                    if philox is not None:
                        self.latent_diffusion.philox_state(dev).copy_(philox)
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.set_rng_state(rng, dev)
                g["graph"] = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g["graph"]):
                    g["out"] = self._score(g["obs"], h, K, g["epi"])
            finally:
                self.latent_diffusion.use_graph = prev
            if len(graphs) >= 8:
                graphs.pop(next(iter(graphs)))
            graphs[key] = g
        g["obs"].copy_(observation, non_blocking=True)
        if epistemic is not None:
            g["epi"].copy_(epistemic, non_blocking=True)
        g["graph"].replay()
        efe, first_action, latent = g["out"]
        return efe.clone(), first_action.clone(), latent.clone()

    @torch.no_grad()
    def forward_pixels(self, encoder: nn.Module, pixels: torch.Tensor, **kw
                       ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Pixel observations (BASELINE cfg#5): `DrQV2Encoder` features are the score net's
        observation, as in `DiffusionPixelAgent.act` (agents/pixel_agent.py:193-205: encode ->
        belief update -> EFE); requires observation_dim == encoder.feature_dim."""
        features = encoder(pixels)
        if features.shape[1] != self.observation_dim:
            raise ValueError(f"encoder feature_dim {features.shape[1]} != observation_dim {self.observation_dim}")
        return self.forward(features, **kw)

    @torch.no_grad()
    def collect_actions(self, observation_host: torch.Tensor, max_diffusion_steps: int = 20, *,
                        deterministic: bool = False, z_init: Optional[torch.Tensor] = None,
                        noise: Optional[torch.Tensor] = None, policy_noise: Optional[torch.Tensor] = None
                        ) -> Tuple[torch.Tensor, torch.Tensor]:
        """The collector's batched inference (utils/async_collector.py:486-595, `_inference_impl`):
        observations of all environments (host memory, pinned if possible) -> one H2D copy ->
        truncated reverse diffusion (t = step/(T-1), the continuous branch) -> policy head ->
        rsample -> actions back on the host in one D2H copy.  Unlike the reference loop there is no
        per-step isnan/isinf host sync (:591-593): non-finite latents are detected once at the end
        and re-initialised as the reference does (0.1 * randn).  Returns (actions [n,A] on the host,
        latents [n,L] on the device).  Keyword-only tensors inject the reference's draws."""
        dev = next(self.parameters()).device
        obs = observation_host.to(dev, non_blocking=True)
        latents = self.latent_diffusion.collector_sample(self.latent_score_network, obs, max_diffusion_steps,
                                                         z_init=z_init, noise=noise)
        bad = ~torch.isfinite(latents).all(dim=1, keepdim=True)
        latents = torch.where(bad, 0.1 * torch.randn_like(latents), latents)
        out = self.heads.head_forward(0, latents)
        A = self.action_dim
        mean, log_std = out[:, :A], torch.clamp(out[:, A:], self.policy_network.log_std_min,
                                                self.policy_network.log_std_max)
        if deterministic:
            action = mean
        else:
            eps = torch.randn_like(mean) if policy_noise is None else policy_noise.to(dev)
            action = mean + torch.exp(log_std) * eps
        host = torch.empty(action.shape, dtype=action.dtype, pin_memory=True)
        host.copy_(action, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return host, latents

    @torch.no_grad()
    def select(self, observation: torch.Tensor, **kw) -> Tuple[torch.Tensor, torch.Tensor]:
        """argmin-EFE candidate (index, action) — new capability (the reference only logs EFE,
        SURVEY fact 5); index parity vs torch.argmin is exact by construction."""
        efe, first_action, _ = self.forward(observation, **kw)
        idx = torch.argmin(efe)
        return idx, first_action[idx]
