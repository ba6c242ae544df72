"""`train_step` helpers without host round trips (callers: agents/state_agent.py:110-243).

* `RunningMeanStd`: the reference's reward normaliser (agents/base_agent.py:24-52) with its float64
  statistics kept on the device: `update` / `normalize` take device tensors and enqueue a handful of
  element-wise kernels -- the reference moves the reward batch to the host and back twice per step
  (agents/state_agent.py:126-133).
* `update_belief_batched`: the three `update_belief_via_diffusion` calls of a training step (observations,
  next observations, next observations again: :136,139,195) as ONE reverse-diffusion run over the stacked
  rows -- rows are independent, so each row's latent is what its own call would have produced from the
  same draws; one library call instead of three fills the SMs three times better at training batch sizes.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch


class RunningMeanStd:
    def __init__(self, epsilon: float = 1e-4, shape=(), device="cuda"):
        self.mean = torch.zeros(shape, dtype=torch.float64, device=device)
        self.var = torch.ones(shape, dtype=torch.float64, device=device)
        self.count = torch.full((), float(epsilon), dtype=torch.float64, device=device)

    def update(self, x: torch.Tensor) -> None:
        x = x.detach().to(self.mean.device, torch.float64)
        self.update_from_moments(x.mean(dim=0), x.var(dim=0, unbiased=False), x.shape[0])

    def update_from_moments(self, batch_mean: torch.Tensor, batch_var: torch.Tensor, batch_count: int) -> None:
        delta = batch_mean - self.mean
        tot = self.count + batch_count
        m2 = self.var * self.count + batch_var * batch_count + delta * delta * self.count * batch_count / tot
        self.mean = self.mean + delta * batch_count / tot
        self.var = m2 / tot
        self.count = tot

    def normalize(self, x: torch.Tensor) -> torch.Tensor:
        """float32 result like the reference's `torch.tensor(..., dtype=torch.float32)`."""
        return ((x.to(self.mean.device, torch.float64) - self.mean) / torch.sqrt(self.var + 1e-8)).float()


@torch.no_grad()
def update_belief_batched(ai, observation_sets: Sequence[torch.Tensor]) -> List[Dict[str, torch.Tensor]]:
    """`[ai.update_belief_via_diffusion(o) for o in observation_sets]` as one reverse-diffusion run.
    Returns one info dict per set (keys of core/active_inference.py:304-312); `ai.current_latent` /
    `ai.latent_trajectory` are left as the LAST call of the reference sequence would leave them."""
    sizes = [int(o.shape[0]) for o in observation_sets]
    stacked = torch.cat([o.to(ai.device) for o in observation_sets], dim=0)
    info = ai.update_belief_via_diffusion(stacked)
    out, lo = [], 0
    for o, n in zip(observation_sets, sizes):
        lat = info["latent"][lo:lo + n]
        lo += n
        rec = torch.nn.functional.mse_loss(ai.decode_observation(lat), o.to(ai.device))
        out.append({"latent": lat, "latent_mean": lat.mean(dim=0) if n > 1 else lat.squeeze(0),
                    "latent_std": lat.std(dim=0) if n > 1 else torch.zeros_like(lat.squeeze(0)),
                    "trajectory_length": info["trajectory_length"], "reconstruction_error": rec,
                    "observation": o.to(ai.device), "raw_observation": None})
    ai.current_latent = out[-1]["latent"]
    return out


class EMAModel:
    """`EMAModel` of core/active_inference.py:779-813 (the agents keep one over the score network,
    agents/base_agent.py:73, and call `update()` every training step, agents/state_agent.py:158) with the
    same interface and the same arithmetic -- shadow = fl(fl(decay * shadow) + fl((1 - decay) * param)) --
    evaluated with multi-tensor kernels: three launches per update instead of three per parameter tensor
    (~150 tensors for the score network).  `apply_shadow` / `restore` swap `param.data` like the reference
    and drop the module's derived packed-weight caches (`invalidate_packed`), so the fused kernels never
    see stale operands after a swap."""

    def __init__(self, model, decay: float = 0.9999, device=None):
        self.model, self.decay, self.device = model, decay, device
        self.shadow, self.backup = {}, {}
        for name, param in model.named_parameters():
            if param.requires_grad:
                self.shadow[name] = param.data.clone().to(device)

    def _live(self):
        return [(n, p) for n, p in self.model.named_parameters() if p.requires_grad]

    @torch.no_grad()
    def update(self) -> None:
        live = self._live()
        shadows = [self.shadow[n] for n, _ in live]
        scaled = torch._foreach_mul([p.data.to(s.device) for (_, p), s in zip(live, shadows)], 1 - self.decay)
        torch._foreach_mul_(shadows, self.decay)
        torch._foreach_add_(shadows, scaled)

    def _invalidate(self) -> None:
        if hasattr(self.model, "invalidate_packed"):
            self.model.invalidate_packed()

    def apply_shadow(self) -> None:
        for name, param in self._live():
            self.backup[name] = param.data
            param.data = self.shadow[name]
        self._invalidate()

    def restore(self) -> None:
        for name, param in self._live():
            param.data = self.backup[name]
        self._invalidate()
