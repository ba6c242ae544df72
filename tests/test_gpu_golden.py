"""CUDA path vs the golden vectors recorded from the unmodified reference (tests/golden/)."""
import os

import pytest
import torch

from tests.util import make_score_net, rel_l2

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["score_small_cosine", "score_small_linear", "score_default_dims"])
def test_score_and_sampler_vs_reference_golden(name):
    from active_inference_diffusion_b200 import DiffusionConfig, LatentDiffusionProcess, LatentScoreNetwork
    fx = torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)
    d = fx["dims"]
    if "weights" in fx:
        net = LatentScoreNetwork(d["L"], d["O"], d["H"], num_layers=d["NB"]).eval()
        net.load_state_dict(fx["weights"])
        net = net.cuda()
    else:
        net, _ = make_score_net(d["L"], d["O"], d["H"], d["NB"], seed=fx["seed"], perturb_seed=fx["perturb_seed"], device="cuda")
    for k, t in fx["times"].items():
        got = net(fx["z"].cuda(), t.cuda(), fx["obs"].cuda())
        assert rel_l2(got, fx["scores"][k]) < 1e-2, (k, rel_l2(got, fx["scores"][k]))
    diff = LatentDiffusionProcess(DiffusionConfig(num_diffusion_steps=d["T"], beta_schedule=d["sched"]), d["L"]).cuda()
    s = fx["sampler"]
    traj = diff.generate_latent_trajectory(net, d["B"], fx["obs"].cuda(), z_init=s["zT"].cuda(), noise=s["noise"].cuda())
    assert len(traj) == s["n_traj"]
    assert rel_l2(traj[-1], s["z_final"]) < 2e-2, rel_l2(traj[-1], s["z_final"])
    assert rel_l2(traj[d["T"] // 2], s["z_mid"]) < 2e-2


def test_efe_vs_reference_golden():
    from active_inference_diffusion_b200 import ActiveInferenceConfig, DiffusionActiveInference, DiffusionConfig
    fx = torch.load(os.path.join(GOLD, "active_inference_small.pt"), weights_only=False)
    d = fx["dims"]
    cfg = ActiveInferenceConfig(hidden_dim=d["H"], latent_dim=d["L"], device="cpu",
                                diffusion=DiffusionConfig(num_diffusion_steps=d["T"]))
    ai = DiffusionActiveInference(observation_dim=d["L"], action_dim=d["A"], latent_dim=d["L"], config=cfg).eval()
    ai.load_state_dict(fx["weights"], strict=False)
    ai = ai.to("cuda")
    ai.use_epistemic = False
    e = fx["efe_zero"]
    with torch.no_grad():        # the fused rollout kernel (a recorded graph takes the differentiable path)
        got, info = ai.compute_expected_free_energy_diffusion(fx["z"].cuda(), horizon=e["h"], num_trajectories=e["K"],
                                                              policy_noise=e["policy_noise"].cuda(),
                                                              reparam_noise=e["reparam_noise"].cuda())
    assert rel_l2(got, e["efe"]) < 2e-2, rel_l2(got, e["efe"])
    assert int(torch.argmin(got)) == int(torch.argmin(e["efe"]))
