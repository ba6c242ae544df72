"""Shared helpers for the parity tests."""
import os

import torch

from oracle.harness import perturb_state_dict


def _log(kind: str, value: float) -> None:
    """AID_LOG_ERRORS=<file>: append (test id, measured error) -- how the stated tolerances were set
    (<= 3x the largest value measured on a B200, profiles/r2_measured_errors.txt)."""
    path = os.environ.get("AID_LOG_ERRORS")
    if path:
        with open(path, "a") as f:
            f.write(f"{os.environ.get('PYTEST_CURRENT_TEST', '?')}\t{kind}\t{value:.4e}\n")


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    v = float((a - b).norm() / (b.norm() + 1e-30))
    _log("rel_l2", v)
    return v


def max_rel(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    v = float((a - b).abs().max() / (b.abs().max() + 1e-30))
    _log("max_rel", v)
    return v


def make_score_net(L, O, H, NB, seed=0, perturb_seed=123, output_multiplier=0.1, device="cpu"):
    """Mirror module with reference-identical seeded init + the documented perturbation of the
    zero-initialised tensors.  Returns (module, cpu state_dict for the oracle)."""
    from active_inference_diffusion_b200 import LatentScoreNetwork
    torch.manual_seed(seed)
    net = LatentScoreNetwork(L, O, H, num_layers=NB).eval()
    sd = perturb_state_dict(net.state_dict(), perturb_seed, output_multiplier)
    net.load_state_dict(sd)
    params = {k: v.clone() for k, v in net.state_dict().items()}
    return net.to(device), params


def gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g
