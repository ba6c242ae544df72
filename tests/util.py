"""Shared helpers for the parity tests."""
import torch

from oracle.harness import perturb_state_dict


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_rel(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


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
