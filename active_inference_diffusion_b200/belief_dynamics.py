"""Drop-in `BeliefDynamics` (core/belief_dynamics.py:12-408) and `FreeEnergyComputation`
(core/free_energy.py:11-103).

`BeliefDynamics.update` raises in the reference as shipped (undefined `_record_state_enhanced`
:170; Hessian of a detached gradient :210,:234,:261 — SURVEY §8c(2)).  The update here follows
the method's own arithmetic (:110-167) for the default Gaussian observation model, where the
free-energy gradient and Hessian are closed-form.  The diagonal branch runs in the fp64 CUDA kernel
`aid_fp_belief_update`; the full-covariance branch (matrix exponential, eigen-clamp, inverse) uses
torch.linalg on the device.  `update_batch` is the batched entry point used for candidate rows.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib


class BeliefDynamics(nn.Module):
    def __init__(self, latent_dim: int, config):
        super().__init__()
        self.latent_dim, self.config = latent_dim, config
        self.register_buffer("mean", torch.zeros(latent_dim, dtype=torch.float64))
        if config.use_full_covariance:
            self.register_buffer("covariance", torch.eye(latent_dim, dtype=torch.float64))
            self.register_buffer("precision", torch.eye(latent_dim, dtype=torch.float64))
        else:
            self.register_buffer("variance", torch.ones(latent_dim, dtype=torch.float64))
            self.register_buffer("precision", torch.ones(latent_dim, dtype=torch.float64))
        self.history = {k: [] for k in ("means", "covariances", "entropies", "free_energies", "condition_numbers",
                                        "numerical_warnings")}
        self.min_eigenvalue = max(config.min_variance, 1e-8)
        self.max_condition_number = 1e6

    def reset(self, initial_mean: Optional[torch.Tensor] = None, initial_cov: Optional[torch.Tensor] = None) -> None:
        dev = self.mean.device
        self.mean = initial_mean.to(dev).to(torch.float64) if initial_mean is not None else torch.zeros_like(self.mean)
        eye = torch.eye(self.latent_dim, device=dev, dtype=torch.float64)
        if self.config.use_full_covariance:
            if initial_cov is not None:
                self.covariance = self._stabilise(initial_cov.to(dev).to(torch.float64))
                self.precision = torch.linalg.inv(self.covariance + self.min_eigenvalue * eye)
            else:
                self.covariance, self.precision = eye.clone(), eye.clone()
        else:
            if initial_cov is not None:
                self.variance = torch.clamp(torch.diag(initial_cov).to(dev).to(torch.float64), min=self.min_eigenvalue)
                self.precision = 1.0 / self.variance
            else:
                self.variance, self.precision = torch.ones_like(self.mean), torch.ones_like(self.mean)
        self.history = {k: [] for k in self.history}

    # ---- batched kernel entry point ----------------------------------------------------------
    def update_batch(self, mean: torch.Tensor, variance: torch.Tensor, observation: torch.Tensor,
                     score: torch.Tensor, noise: Optional[torch.Tensor] = None
                     ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """[N, L] float64 beliefs -> (mean', variance', precision'), one warp per row on the device."""
        dev = _lib.require_cuda(mean, variance, observation, score, noise)
        f64 = lambda t: None if t is None else t.to(torch.float64).contiguous()
        mean, variance, observation, score, noise = map(f64, (mean, variance, observation, score, noise))
        N, L = mean.shape
        out = [torch.empty_like(mean) for _ in range(3)]
        c = self.config
        _lib.check(_lib.lib().aid_fp_belief_update(
            mean.data_ptr(), variance.data_ptr(), observation.data_ptr(), score.data_ptr(), _lib.ptr(noise), N, L,
            float(c.dt), float(c.diffusion_coefficient), float(c.learning_rate), float(c.noise_scale),
            float(c.min_variance), float(c.max_variance), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
            _lib.stream_ptr(dev)), "aid_fp_belief_update")
        return out[0], out[1], out[2]

    # ---- reference call surface ---------------------------------------------------------------
    def update(self, observation: torch.Tensor, score_function: torch.Tensor, action: torch.Tensor = None,
               observation_model: Optional[nn.Module] = None, noise: Optional[torch.Tensor] = None):
        if observation_model is not None:
            raise NotImplementedError("custom observation_model: only the default Gaussian model is closed-form")
        o = observation.to(torch.float64).reshape(-1)
        s = score_function.to(torch.float64).reshape(-1)
        if noise is None:
            noise = torch.randn_like(self.mean)
        c = self.config
        if not c.use_full_covariance:
            m, v, p = self.update_batch(self.mean[None], self.variance[None], o[None], s[None], noise[None])
            self.mean, self.variance, self.precision = m[0], v[0], p[0]
        else:
            g = -(self.mean - o) / c.noise_scale ** 2 - self.mean + s
            adaptive_dt = c.dt / (1 + 0.1 * g.norm())
            self.mean = self.mean + (-c.learning_rate * g) * adaptive_dt + math.sqrt(2 * c.diffusion_coefficient * c.dt) * noise * c.noise_scale
            eye = torch.eye(self.latent_dim, device=self.mean.device, dtype=torch.float64)
            H = -(1.0 / c.noise_scale ** 2 + 1.0) * eye
            E = torch.matrix_exp((-H - H.T + 2 * c.diffusion_coefficient * eye) * c.dt)
            self.covariance = self._stabilise(E @ self.covariance @ E.T)
            self.precision = torch.linalg.inv(self.covariance + self.min_eigenvalue * eye)
        return self.get_parameters()

    def _stabilise(self, m: torch.Tensor) -> torch.Tensor:
        w, V = torch.linalg.eigh(m)
        w = torch.clamp(w, min=self.min_eigenvalue)
        if (w.max() / w.min()) > self.max_condition_number:
            w = w + w.mean() * 1e-6
        return V @ torch.diag(w) @ V.T

    def get_parameters(self) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.config.use_full_covariance:
            return self.mean.to(torch.float32), self.covariance.to(torch.float32)
        return self.mean.to(torch.float32), torch.diag(self.variance).to(torch.float32)

    def entropy(self) -> torch.Tensor:
        k = self.latent_dim
        if self.config.use_full_covariance:
            return 0.5 * (k * math.log(2 * math.pi * math.e) + torch.logdet(self.covariance))
        return 0.5 * torch.sum(math.log(2 * math.pi * math.e) + torch.log(torch.clamp(self.variance, min=self.min_eigenvalue)))


    def get_diagnostics(self):
        """core/belief_dynamics.py:391-410, all values from ONE device->host transfer."""
        if self.config.use_full_covariance:
            w = torch.linalg.eigvals(self.covariance).real
            vals = torch.stack([w.min(), w.max(), w.max() / w.min(), torch.det(self.covariance),
                                self.mean.norm(), self.entropy()]).tolist()
            keys = ("min_eigenvalue", "max_eigenvalue", "condition_number", "determinant", "mean_norm", "entropy")
        else:
            vals = torch.stack([self.variance.min(), self.variance.max(), self.variance.mean(), self.mean.norm(),
                                self.entropy()]).tolist()
            keys = ("min_variance", "max_variance", "mean_variance", "mean_norm", "entropy")
        return dict(zip(keys, vals))


class FreeEnergyComputation(nn.Module):
    """F = complexity - accuracy + 0.01 |s_theta|^2 (core/free_energy.py:30-91); the score term runs
    on the sm_100a score forward (no gradient flows through it here)."""

    def __init__(self, precision_init: float = 1.0):
        super().__init__()
        self.log_precision = nn.Parameter(torch.log(torch.tensor(precision_init)))

    @property
    def precision(self) -> torch.Tensor:
        return torch.exp(self.log_precision)

    def compute_loss(self, states, observations, actions, score_network, current_time: float = 0.0,
                     prior_mean: Optional[torch.Tensor] = None, prior_std: float = 1.0):
        B, dev = states.shape[0], states.device
        if prior_mean is None:
            prior_mean = torch.zeros_like(states)
        complexity = 0.5 * torch.sum((states - prior_mean) ** 2 / (prior_std ** 2), dim=-1).mean()
        obs_err = torch.sum((observations - states) ** 2, dim=-1)
        accuracy = -0.5 * self.precision * obs_err.mean()
        t = torch.full((B,), current_time, device=dev)
        score = score_network(states, t, observations)
        score_reg = 0.01 * torch.sum(score ** 2, dim=-1).mean()
        fe = complexity - accuracy + score_reg
        return fe, {"complexity": complexity, "accuracy": -accuracy, "observation_error": obs_err.mean(),
                    "score_regularization": score_reg, "precision": self.precision}

    def update_precision(self, complexity: torch.Tensor, accuracy: torch.Tensor) -> None:
        with torch.no_grad():
            self.log_precision.data += 0.01 * (complexity - accuracy).clamp(-1, 1)
            self.log_precision.data = self.log_precision.data.clamp(-3, 3)
