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
