"""The oracle restatement against golden vectors recorded from the UNMODIFIED reference
(oracle/gen_golden.py, run in the build container where /root/reference exists).  CPU only.

The restatement is plain torch fp32 with the same op order, so the bar is bit-exact for the
forward quantities and 1e-6-relative for gradients (autograd accumulation order may differ)."""
import os

import numpy as np
import pytest
import torch

from oracle import restatement as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)


def weights_for(fx):
    if "weights" in fx:
        return fx["weights"]
    from tests.util import make_score_net
    d = fx["dims"]
    _, params = make_score_net(d["L"], d["O"], d["H"], d["NB"], seed=fx["seed"], perturb_seed=fx["perturb_seed"])
    return params


@pytest.mark.parametrize("name", ["score_small_cosine", "score_small_linear", "score_default_dims"])
def test_score_forward_bit_exact(name):
    fx = load(name)
    p = weights_for(fx)
    with torch.no_grad():
        for k, t in fx["times"].items():
            got = R.score_forward(p, fx["z"], t, fx["obs"])
            assert torch.equal(got, fx["scores"][k]), (name, k, float((got - fx["scores"][k]).abs().max()))


def test_score_batch_global_branch_quirk():
    """the continuous/discrete branch is decided per BATCH.  The same t=0 row scores
    ~316x smaller when batched with a t=2 row (golden 'mixed_batch' vs 't0')."""
    fx = load("score_small_cosine")
    r_mixed = fx["scores"]["mixed_batch"][0].abs().max()
    r_t0 = fx["scores"]["t0"][0].abs().max()
    assert r_t0 > 50 * r_mixed


@pytest.mark.parametrize("name", ["score_small_cosine", "score_small_linear", "score_default_dims"])
def test_schedule_and_sampler_bit_exact(name):
    fx = load(name)
    d = fx["dims"]
    sched = R.make_schedule(d["T"], d["sched"])
    for k, v in fx["schedule"].items():
        assert torch.equal(sched[k], v), k
    p = weights_for(fx)
    s = fx["sampler"]
    with torch.no_grad():
        traj = R.generate_latent_trajectory(p, sched, s["zT"], fx["obs"], list(s["noise"]))
    assert len(traj) == s["n_traj"] == d["T"] + 1
    assert torch.equal(traj[-1], s["z_final"])
    assert torch.equal(traj[d["T"] // 2], s["z_mid"])


def test_default_dims_seeded_init_matches_reference_checksums():
    """The mirror LatentScoreNetwork must consume the RNG exactly like the reference constructor:
    per-tensor (sum, abs-sum) of the seed-0 init recorded from the reference."""
    from active_inference_diffusion_b200 import LatentScoreNetwork
    fx = load("score_default_dims")
    d = fx["dims"]
    torch.manual_seed(fx["seed"])
    net = LatentScoreNetwork(d["L"], d["O"], d["H"], num_layers=d["NB"])
    sd = net.state_dict()
    assert set(k for k in sd if sd[k].is_floating_point()) == set(fx["init_checksums"])
    for k, (s, a) in fx["init_checksums"].items():
        v = sd[k].double()
        assert float(v.sum()) == s and float(v.abs().sum()) == a, k


def _nets(fx):
    w = fx["weights"]
    return {n: R.sub(w, pre) for n, pre in [("policy", "policy_network"), ("dynamics", "latent_dynamics"),
                                           ("value", "value_network"), ("reward", "reward_predictor"),
                                           ("decoder", "observation_decoder"), ("epistemic", "epistemic_estimator"),
                                           ("score", "latent_score_network"), ("diffusion", "latent_diffusion")]}


EFE_CFG = dict(epistemic_weight=0.1, pragmatic_weight=1.0, consistency_weight=0.1, discount_factor=0.99,
               preference_temperature=1.0)


def test_efe_epistemic_zero_bit_exact():
    fx = load("active_inference_small")
    nets, e = _nets(fx), fx["efe_zero"]
    noise = [dict(policy=e["policy_noise"][i], reparam=e["reparam_noise"][i]) for i in range(e["K"] * e["h"])]
    with torch.no_grad():
        efe, info, _ = R.expected_free_energy(nets, EFE_CFG, fx["z"], e["h"], e["K"], noise)
    assert torch.equal(efe, e["efe"])
    assert torch.equal(info["pragmatic_mean"], e["pragmatic_mean"])
    assert torch.equal(info["consistency_mean"], e["consistency_mean"])


def test_efe_with_mine_estimator_and_decoder_shim():
    fx = load("active_inference_small")
    nets, e = _nets(fx), fx["efe_mine"]
    with torch.no_grad():
        efe, info, _ = R.expected_free_energy(nets, EFE_CFG, fx["z"], e["h"], e["K"], e["noise"], epistemic="mine")
    assert torch.equal(efe, e["efe"])


def test_elbo_loss_and_gradients():
    fx = load("active_inference_small")
    nets = _nets(fx)
    cfg = dict(kl_weight=0.1, diffusion_weight=1.0, reward_weight=0.5)
    tiw = torch.ones(100)
    for call, rec in enumerate(fx["elbo"]):
        kinds, draws = rec["kinds"], rec["draws"]
        if call == 0:
            assert kinds == ["rand", "randn_like", "randn_like"]
            t, n1, n2 = draws
        else:
            assert kinds == ["multinomial", "rand", "randn_like", "randn_like"]
            t = R.importance_sample_time(None, draws[0], draws[1])
            n1, n2 = draws[2], draws[3]
        sp = {k: v.clone().requires_grad_(True) for k, v in nets["score"].items() if v.is_floating_point()}
        dp = {k: nets["diffusion"][k].clone().requires_grad_(True)
              for k in ("latent_prior_mean", "latent_prior_log_std", "log_snr_min", "log_snr_max")}
        loss, info, per = R.diffusion_elbo(sp, dp, nets["decoder"], nets["reward"], cfg, fx["obs"], fx["rew"], fx["z"],
                                           t, n1, n2)
        loss.backward()
        assert torch.equal(loss.detach(), rec["loss"])
        for k in ("score_matching_loss", "grad_penalty", "kl_loss", "reward_loss", "reconstruction_loss"):
            assert float(info[k]) == pytest.approx(rec["info"][k], rel=1e-6)
        assert torch.allclose(sp["output_proj.2.weight"].grad, rec["grad_out2"], rtol=1e-5, atol=1e-7)
        assert torch.allclose(sp["latent_proj.weight"].grad, rec["grad_latent_proj"], rtol=1e-5, atol=1e-7)
        assert torch.allclose(dp["log_snr_min"].grad, rec["grad_log_snr_min"], rtol=1e-5)
        assert torch.allclose(dp["latent_prior_mean"].grad, rec["grad_prior_mean"], rtol=1e-5, atol=1e-8)
        tiw = R.update_time_importance(tiw, t, per.detach())
        assert torch.equal(tiw, rec["tiw"])      # integer bins + sequential EMA: exact


def test_free_energy_and_belief_gradient():
    fx = load("free_energy_belief")
    with torch.no_grad():
        f, info = R.free_energy_loss(fx["weights"], fx["log_precision"], fx["states"], fx["obs"], 0.3)
    assert torch.equal(f, fx["free_energy"])
    b = fx["belief"]
    c = b["cfg"]
    g = -(b["mean"] - b["obs"]) / c["noise_scale"] ** 2 - b["mean"] + b["score"]
    assert torch.allclose(g, b["grad"], rtol=1e-12, atol=1e-9)    # closed form == reference autodiff (fp64)


def test_belief_update_restatement_properties():
    """BeliefDynamics.update cannot run in the reference: the restatement is checked
    through its invariants — variance clamp, precision = 1/variance, noise-free fixed point."""
    fx = load("free_energy_belief")
    b, c = fx["belief"], fx["belief"]["cfg"]
    L = b["mean"].shape[0]
    mean, var, prec = R.belief_update_diag(b["mean"].numpy(), np.ones(L), b["obs"].numpy(), b["score"].numpy(),
                                           np.zeros(L), dt=c["dt"], D=c["D"], lr=c["lr"], noise_scale=c["noise_scale"],
                                           min_variance=c["min_variance"], max_variance=c["max_variance"])
    assert var.max() <= c["max_variance"] and var.min() >= max(c["min_variance"], 1e-8)
    assert np.allclose(prec, 1.0 / var)
    # H_ii = -(1/ns^2+1) < 0 -> variance grows by exp((2/ns^2 + 2 + 2D) dt) before the clamp
    assert np.allclose(var, min(c["max_variance"], float(np.exp((2 * (1 / c["noise_scale"] ** 2 + 1) + 2 * c["D"]) * c["dt"]))))
    g = -(b["mean"].numpy() - b["obs"].numpy()) / c["noise_scale"] ** 2 - b["mean"].numpy() + b["score"].numpy()
    assert np.allclose(mean, b["mean"].numpy() - c["lr"] * g * c["dt"] / (1 + 0.1 * np.linalg.norm(g)))
    m2, S, P = R.belief_update_full(b["mean"].numpy(), np.eye(L), b["obs"].numpy(), b["score"].numpy(), np.zeros(L),
                                    dt=c["dt"], D=c["D"], lr=c["lr"], noise_scale=c["noise_scale"],
                                    min_variance=c["min_variance"])
    assert np.allclose(m2, mean) and np.allclose(S, S.T)
    assert np.allclose(P @ (S + max(c["min_variance"], 1e-8) * np.eye(L)), np.eye(L), atol=1e-8)


@pytest.mark.parametrize("name", ["encoder_small", "encoder_small_odd"])
def test_encoder_forward_golden(name):
    """DrQ-v2 encoder restatement vs the unmodified reference (uint8, 5-D float and single-frame
    inputs; the second fixture has odd sizes and no attention)."""
    fx = load(name)
    d = fx["dims"]
    for kind, x in fx["inputs"].items():
        xin = R.encoder_canonical_input(x, d["obs_shape"][0], d["frame_stack"])
        with torch.no_grad():
            got = R.encoder_forward(fx["weights"], xin)
        assert torch.allclose(got, fx["outputs"][kind], rtol=0, atol=2e-6), (kind, float((got - fx["outputs"][kind]).abs().max()))


@pytest.mark.parametrize("name", ["encoder_small", "encoder_small_odd"])
def test_encoder_mirror_seeded_init_matches_reference(name):
    """The mirror's constructor spends the same draws in the same order as the reference's, so the
    same seed gives the same state_dict (spectral-norm u/v after the dry-run power iteration included)."""
    from active_inference_diffusion_b200 import DrQV2Encoder
    fx = load(name)
    d = fx["dims"]
    torch.manual_seed(fx["seed"])
    enc = DrQV2Encoder(tuple(d["obs_shape"]), feature_dim=d["feature_dim"], frame_stack=d["frame_stack"],
                       num_filters=d["num_filters"], use_attention=d["use_attention"])
    sd = enc.state_dict()
    assert list(sd.keys()) == list(fx["init_state"].keys())
    for k, v in fx["init_state"].items():
        assert torch.equal(sd[k], v), k


def test_lambda_returns_golden_bit_exact():
    """compute_lambda_returns restatement vs the unmodified reference (all batch sizes incl. 1 and 2,
    n_steps 0/1/5/8, both exclude_immediate_rewards settings, random terminal flags)."""
    fx = load("lambda_returns")
    assert len(fx["cases"]) == 40
    for c in fx["cases"]:
        got = R.lambda_returns(c["rewards"], c["next_values"], c["dones"], c["gamma"], c["lam"], c["n_steps"], c["exclude"])
        assert torch.equal(got, c["out"]), (c["n_steps"], c["exclude"], got, c["out"])
