"""TEST INFRASTRUCTURE — generates tests/golden/*.pt by executing the UNMODIFIED reference
(/root/reference, this container only) on seeded inputs with recorded noise.

    python -m oracle.gen_golden

Each fixture stores inputs, the recorded random draws and the reference outputs.  Weights are
stored only for the small configuration; for the default dims (H=512, L=128: 27 M parameters) the
fixture stores the seed recipe and per-tensor checksums, and the tests rebuild the weights with
the mirror modules (whose seeded init is checked tensor-by-tensor against the reference here).
"""
from __future__ import annotations

import os
import types

import torch

from oracle.harness import RecordingRNG, StateDecoderShim, perturb_generic, perturb_state_dict
from oracle.ref_import import import_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def checksums(sd):
    return {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in sd.items() if v.is_floating_point()}


def gen_score_and_sampler(name, L, O, H, NB, T, B, sched, store_weights):
    from active_inference_diffusion.configs.config import DiffusionConfig
    from active_inference_diffusion.core.diffusion import LatentDiffusionProcess
    from active_inference_diffusion.models.score_networks import LatentScoreNetwork
    torch.manual_seed(0)
    net = LatentScoreNetwork(L, O, H, num_layers=NB).eval()
    raw_sd = {k: v.clone() for k, v in net.state_dict().items()}
    net.load_state_dict(perturb_state_dict(net.state_dict(), 123, 0.1))
    g = torch.Generator().manual_seed(7)
    z = torch.randn(B, L, generator=g)
    obs = torch.randn(B, O, generator=g)
    times = {"discrete": torch.full((B,), 7.0), "t1": torch.full((B,), 1.0), "t0": torch.zeros(B),
             "uniform": torch.rand(B, generator=g), "mixed_batch": torch.tensor([0.0, 2.0] + [3.0] * (B - 2))}
    fx = {"dims": dict(L=L, O=O, H=H, NB=NB, T=T, B=B, sched=sched), "seed": 0, "perturb_seed": 123,
          "init_checksums": checksums(raw_sd), "z": z, "obs": obs, "times": times, "scores": {}}
    with torch.no_grad():
        for k, t in times.items():
            fx["scores"][k] = net(z, t, obs)
        # observation=None crashes in the reference (models/score_networks.py:147: LayerNorm has no
        # .out_features); the intended zero embedding is restated in oracle/restatement.py and not pinned.
        dp = LatentDiffusionProcess(DiffusionConfig(num_diffusion_steps=T, beta_schedule=sched), latent_dim=L)
        torch.manual_seed(11)
        with RecordingRNG() as rec:
            traj = dp.generate_latent_trajectory(net, B, obs)
        fx["sampler"] = {"zT": rec.of("randn")[0], "noise": torch.stack(rec.of("randn_like")),
                         "z_final": traj[-1], "z_mid": traj[T // 2], "n_traj": len(traj)}
        fx["schedule"] = {k: getattr(dp, k).clone() for k in
                          ("betas", "alphas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
                           "sqrt_one_minus_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped")}
    if store_weights:
        fx["weights"] = {k: v.clone() for k, v in net.state_dict().items()}
    torch.save(fx, os.path.join(OUT, f"{name}.pt"))
    print(name, "score max", float(fx["scores"]["discrete"].abs().max()), "z_final max", float(traj[-1].abs().max()))


def gen_active_inference(name, L, A, H, T, B):
    from active_inference_diffusion.configs.config import ActiveInferenceConfig, DiffusionConfig
    from active_inference_diffusion.core.active_inference import DiffusionActiveInference
    torch.manual_seed(0)
    cfg = ActiveInferenceConfig(hidden_dim=H, latent_dim=L, device="cpu", diffusion=DiffusionConfig(num_diffusion_steps=T))
    ai = DiffusionActiveInference(observation_dim=L, action_dim=A, latent_dim=L, config=cfg).eval()
    init_ck = checksums(ai.state_dict())
    ai.latent_score_network.load_state_dict(perturb_state_dict(ai.latent_score_network.state_dict()))
    for n in ["policy_network", "latent_dynamics", "value_network", "reward_predictor", "observation_decoder"]:
        m = getattr(ai, n)
        m.load_state_dict(perturb_generic(m.state_dict(), 7, 0.05))
    learn = {k: v for k, v in ai.latent_diffusion.state_dict().items()
             if k in ("latent_prior_mean", "latent_prior_log_std", "log_snr_min", "log_snr_max")}
    ai.latent_diffusion.load_state_dict(perturb_generic(learn, 7, 0.05), strict=False)
    g = torch.Generator().manual_seed(3)
    z = torch.randn(B, L, generator=g)
    obs = torch.randn(B, L, generator=g)
    rew = torch.randn(B, generator=g)
    fx = {"dims": dict(L=L, A=A, H=H, T=T, B=B), "init_checksums": init_ck,
          "weights": {k: v.clone() for k, v in ai.state_dict().items()
                      if not k.startswith("epistemic_estimator.decoder.")},   # aliases of observation_decoder.*
          "z": z, "obs": obs, "rew": rew}
    # EFE, epistemic shimmed to zero (documented shim: the state-mode estimator crashes)
    orig = ai.compute_epistemic_value
    ai.compute_epistemic_value = types.MethodType(lambda self, m, lv, num_samples=5: (torch.zeros(m.shape[0]), {}), ai)
    K, h = 3, 4
    with RecordingRNG() as rec, torch.no_grad():
        efe, info = ai.compute_expected_free_energy_diffusion(z, horizon=h, num_trajectories=K)
    draws = rec.draws
    fx["efe_zero"] = {"K": K, "h": h, "policy_noise": torch.stack([t for k, t in draws if k == "normal_"]),
                      "reparam_noise": torch.stack([t for k, t in draws if k == "randn_like"]), "efe": efe,
                      "pragmatic_mean": info["pragmatic_mean"], "consistency_mean": info["consistency_mean"]}
    # EFE with the MINE estimator and the decoder shim
    ai.compute_epistemic_value = orig
    ai.epistemic_estimator.decoder = StateDecoderShim(ai.observation_decoder).eval()
    K2, h2, S = 2, 2, 3
    with RecordingRNG() as rec, torch.no_grad():
        efe2, info2 = ai.compute_expected_free_energy_diffusion(z, horizon=h2, num_trajectories=K2, num_ambiguity_samples=S)
    it = iter(rec.draws)
    steps = []
    for _ in range(K2 * h2):
        p = next(it)[1]; r = next(it)[1]
        zs = [next(it)[1] for _ in range(S)]; ds = [next(it)[1] for _ in range(4)]; pm = [next(it)[1] for _ in range(S)]
        steps.append(dict(policy=p, reparam=r, z=zs, dir=ds, perm=pm))
    fx["efe_mine"] = {"K": K2, "h": h2, "S": S, "noise": steps, "efe": efe2, "mi": info2["epistemic/mi_estimate"],
                      "running_mean": float(ai.epistemic_estimator.running_mean)}
    # ELBO first call (uniform t) and second call (importance-sampled t)
    fx["elbo"] = []
    for call in range(2):
        ai.zero_grad()
        with RecordingRNG() as rec:
            loss, info = ai.compute_diffusion_elbo(obs, rew, z)
        loss.backward()
        rec_d = {"kinds": [k for k, _ in rec.draws], "draws": [t for _, t in rec.draws], "loss": loss.detach(),
                 "info": info, "tiw": ai.time_importance_weights.clone(),
                 "grad_out2": ai.latent_score_network.output_proj[2].weight.grad.clone(),
                 "grad_latent_proj": ai.latent_score_network.latent_proj.weight.grad.clone(),
                 "grad_log_snr_min": ai.latent_diffusion.log_snr_min.grad.clone(),
                 "grad_prior_mean": ai.latent_diffusion.latent_prior_mean.grad.clone()}
        fx["elbo"].append(rec_d)
    torch.save(fx, os.path.join(OUT, f"{name}.pt"))
    print(name, "efe", float(efe.abs().max()), "loss", float(fx["elbo"][0]["loss"]))


def gen_misc(name):
    from active_inference_diffusion.configs.config import BeliefDynamicsConfig
    from active_inference_diffusion.core.belief_dynamics import BeliefDynamics
    from active_inference_diffusion.core.free_energy import FreeEnergyComputation
    from active_inference_diffusion.models.score_networks import LatentScoreNetwork
    torch.manual_seed(0)
    L = 32
    net = LatentScoreNetwork(L, L, 64, num_layers=2).eval()
    net.load_state_dict(perturb_state_dict(net.state_dict()))
    g = torch.Generator().manual_seed(5)
    states, obs = torch.randn(6, L, generator=g), torch.randn(6, L, generator=g)
    fe = FreeEnergyComputation(1.3)
    with torch.no_grad():
        f, info = fe.compute_loss(states, obs, None, net, current_time=0.3)
    bc = BeliefDynamicsConfig()
    bd = BeliefDynamics(L, bc)
    bd.reset(torch.randn(L, generator=g))
    o = torch.randn(L, generator=g, dtype=torch.float64)
    s = torch.randn(L, generator=g, dtype=torch.float64)
    grad = bd._compute_free_energy_gradient_autodiff(bd.mean, o, s)
    torch.save({"weights": {k: v.clone() for k, v in net.state_dict().items()}, "states": states, "obs": obs,
                "free_energy": f, "fe_info": {k: v.detach() for k, v in info.items()}, "log_precision": fe.log_precision.detach(),
                "belief": {"mean": bd.mean.clone(), "obs": o, "score": s, "grad": grad,
                           "cfg": dict(dt=bc.dt, D=bc.diffusion_coefficient, lr=bc.learning_rate, noise_scale=bc.noise_scale,
                                       min_variance=bc.min_variance, max_variance=bc.max_variance)}},
               os.path.join(OUT, f"{name}.pt"))
    print(name, float(f))


def gen_encoder(name, obs_shape, frame_stack, feature_dim, num_filters, B, use_attention=True):
    """DrQV2Encoder (encoder/visual_encoders.py) at a size whose weights fit a fixture: the
    architecture is parametric in num_filters / feature_dim / image size."""
    from active_inference_diffusion.encoder.visual_encoders import DrQV2Encoder
    torch.manual_seed(5)
    enc = DrQV2Encoder(obs_shape, feature_dim=feature_dim, frame_stack=frame_stack, num_filters=num_filters,
                       use_attention=use_attention)
    raw_sd = {k: v.clone() for k, v in enc.state_dict().items()}
    enc.load_state_dict(perturb_generic(enc.state_dict(), 31, 0.05))
    enc.eval()
    g = torch.Generator().manual_seed(17)
    c, h, w = obs_shape
    u8 = torch.randint(0, 256, (B, frame_stack * c, h, w), generator=g, dtype=torch.uint8)
    f32 = torch.rand(B, frame_stack, c, h, w, generator=g)          # 5-D, separate frames
    single = torch.rand(B, c, h, w, generator=g)                      # one frame, repeated by the encoder
    with torch.no_grad():
        outs = {"u8": enc(u8), "f32_5d": enc(f32), "single_frame": enc(single)}
    torch.save({"dims": dict(obs_shape=obs_shape, frame_stack=frame_stack, feature_dim=feature_dim,
                             num_filters=num_filters, B=B, use_attention=use_attention),
                "seed": 5, "init_state": raw_sd, "weights": {k: v.clone() for k, v in enc.state_dict().items()},
                "inputs": {"u8": u8, "f32_5d": f32, "single_frame": single}, "outputs": outs},
               os.path.join(OUT, f"{name}.pt"))
    print(name, "features max", float(outs["u8"].abs().max()), "conv_out_dim", enc.conv_out_dim)


def gen_lambda_returns(name):
    """compute_lambda_returns (core/active_inference.py:638-707) called unbound on a stub carrying
    config.discount_factor: the method touches nothing else of the object."""
    from active_inference_diffusion.core.active_inference import DiffusionActiveInference
    stub = types.SimpleNamespace(config=types.SimpleNamespace(discount_factor=0.99))
    g = torch.Generator().manual_seed(5)
    cases = []
    for B in (1, 2, 3, 7, 40):
        for n_steps in (0, 1, 5, 8):
            for excl in (False, True):
                r, nv = torch.randn(B, generator=g), torch.randn(B, generator=g)
                d = torch.rand(B, generator=g) < 0.2
                out = DiffusionActiveInference.compute_lambda_returns(stub, r, None, nv, d, lambda_=0.95, n_steps=n_steps,
                                                                      exclude_immediate_rewards=excl)
                cases.append(dict(rewards=r, next_values=nv, dones=d, gamma=0.99, lam=0.95, n_steps=n_steps,
                                  exclude=excl, out=out))
    torch.save({"cases": cases}, os.path.join(OUT, f"{name}.pt"))
    print(name, len(cases), "cases")


def main():
    import_reference()
    os.makedirs(OUT, exist_ok=True)
    gen_score_and_sampler("score_small_cosine", 64, 17, 128, 2, 10, 7, "cosine", True)
    gen_score_and_sampler("score_small_linear", 32, 17, 128, 2, 8, 5, "linear", True)
    gen_score_and_sampler("score_default_dims", 128, 17, 512, 6, 50, 8, "cosine", False)
    gen_active_inference("active_inference_small", 32, 6, 64, 6, 9)
    gen_misc("free_energy_belief")
    gen_lambda_returns("lambda_returns")
    gen_encoder("encoder_small", (3, 12, 12), 3, 16, 8, 5)
    gen_encoder("encoder_small_odd", (1, 11, 14), 2, 8, 8, 3, use_attention=False)


if __name__ == "__main__":
    main()
