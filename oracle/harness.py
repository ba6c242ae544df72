"""TEST INFRASTRUCTURE — helpers shared by the oracle, the golden generator and the tests.

* `perturb_state_dict`  — re-randomise the tensors the reference zero-initialises
  (adaLN modulations models/score_networks.py:254-255, last output weight :99,
  `output_multiplier` :97) so that parity tests exercise real math (SURVEY fact 7:
  at construction the score is identically zero).
* `RecordingRNG`        — records every random tensor a block of reference code draws
  (torch.randn / randn_like / rand / randperm / multinomial / Tensor.normal_), in order,
  so the same noise can be injected into the oracle restatement and the CUDA path.
* shims for the two reference pieces that crash as shipped.
"""
from __future__ import annotations

import contextlib
from typing import Dict, List, Tuple

import torch


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def perturb_state_dict(sd: Dict[str, torch.Tensor], seed: int = 123,
                       output_multiplier: float = 0.1) -> Dict[str, torch.Tensor]:
    """Return a copy of a LatentScoreNetwork state_dict with the zero-init tensors replaced.

    Keys are visited in sorted order and each gets its own generator stream so the result
    depends only on (key, shape, seed)."""
    out = {k: v.clone() for k, v in sd.items()}
    for i, k in enumerate(sorted(out)):
        g = _gen(seed * 100003 + i)
        v = out[k]
        if "adaLN_modulation.1.weight" in k:
            out[k] = torch.randn(v.shape, generator=g) * (0.5 / v.shape[1] ** 0.5)
        elif "adaLN_modulation.1.bias" in k:
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        elif k.endswith("output_proj.2.weight"):
            out[k] = torch.randn(v.shape, generator=g) * (1.0 / v.shape[1] ** 0.5)
        elif k.endswith("output_multiplier"):
            out[k] = torch.full_like(v, output_multiplier)
        elif k.endswith("mlp.0.bias") or k.endswith("mlp.2.bias"):
            # DiT MLP biases are zero-init (:211-212); give them values so bias paths are tested
            out[k] = torch.randn(v.shape, generator=g) * 0.02
        elif k.endswith("time_scale"):
            out[k] = torch.tensor(0.7)
        elif k.endswith("freq_scale"):
            out[k] = torch.full_like(v, 1.1)
    return out


def perturb_generic(sd: Dict[str, torch.Tensor], seed: int, scale: float = 0.05) -> Dict[str, torch.Tensor]:
    """Add small seeded noise to every floating tensor (heads: zero biases, tiny last layers)."""
    out = {}
    for i, k in enumerate(sorted(sd)):
        v = sd[k]
        if v.is_floating_point() and v.numel() > 0:
            g = _gen(seed * 7919 + i)
            out[k] = v + torch.randn(v.shape, generator=g).to(v.dtype) * scale
        else:
            out[k] = v.clone()
    return out


class RecordingRNG(contextlib.AbstractContextManager):
    """Record the random tensors drawn inside the `with` block, in draw order.

    `self.draws` is a list of (kind, tensor) with kind in
    {'randn','randn_like','rand','randperm','multinomial','normal_'}.
    `normal_` records the STANDARD normal underlying `Normal.rsample` only when called
    with the default mean=0,std=1 (how torch.distributions uses `_standard_normal`).
    """

    def __init__(self):
        self.draws: List[Tuple[str, torch.Tensor]] = []
        self._saved = {}

    def _wrap(self, owner, name, kind):
        orig = getattr(owner, name)
        self._saved[(owner, name)] = orig
        rec = self.draws

        def wrapped(*a, **k):
            out = orig(*a, **k)
            rec.append((kind, out.detach().clone()))
            return out

        setattr(owner, name, wrapped)

    def __enter__(self):
        self._wrap(torch, "randn", "randn")
        self._wrap(torch, "randn_like", "randn_like")
        self._wrap(torch, "rand", "rand")
        self._wrap(torch, "randperm", "randperm")
        self._wrap(torch, "multinomial", "multinomial")
        self._wrap(torch.Tensor, "normal_", "normal_")
        return self

    def __exit__(self, *exc):
        for (owner, name), orig in self._saved.items():
            setattr(owner, name, orig)
        self._saved.clear()
        return False

    def of(self, kind: str) -> List[torch.Tensor]:
        return [t for k, t in self.draws if k == kind]


class StateDecoderShim(torch.nn.Module):
    """the state-mode epistemic estimator calls `self.decoder(z)` on an
    nn.ModuleList (core/active_inference.py:953,966), which raises.  The oracle substitutes a
    module that reproduces `decode_observation`'s skip forward (:237-242) over the SAME
    ModuleList parameters."""

    def __init__(self, module_list):
        super().__init__()
        self.m = module_list

    def forward(self, z):
        h1 = self.m[0](z)
        h2 = self.m[1](h1) + h1
        h3 = self.m[2](h2)
        return self.m[3](h3)

    def __len__(self):  # FunctionSpaceEpistemicEstimator.to() iterates ModuleLists only
        return 4
