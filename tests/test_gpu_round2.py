"""Round-2 parity tests through the C ABI on a B200.

* fp16 tensor-core operands (libaid_sm100_f16.so): the TF32-class mode of the rel-1e-3 contract for
  the score, the 50-step sampled latent and the EFE (north_star: "rel 1e-3 for fp32/TF32").
* In-kernel Philox noise of the sampler: distribution, determinism, shard invariance, fresh draws per
  call / per graph replay.
* CUDA-graph replay of the sampler and of the whole scorer call == the directly launched sequence.
* a8/a8b element-wise diffusion functions (q_sample, continuous_q_sample, log-SNR, loss weight,
  prior sample, stand-alone p_sample) vs the oracle, incl. the t[0]==0 rule of p_sample.
* BASELINE cfg#4 dims together (O=376, A=17, h=15, L=128, H=512) and the cfg#5 chain
  (`CandidateScorer.forward_pixels`: DrQ-v2 encoder -> sampler -> EFE).
* packed-weight cache invalidation (`.data` edits do not bump version counters).
"""
import math

import pytest
import torch

from oracle import restatement as R
from oracle.harness import perturb_generic
from tests.util import gen, make_score_net, rel_l2
from tests.test_gpu_efe import make_ai

pytestmark = pytest.mark.gpu

F16_TOL = 1e-3          # north_star fp32/TF32 bound; fp16 operands carry the TF32 significand
BF16_SCORE_TOL = 1e-2   # measured 2-4e-3 (precision)


def _lib():
    from active_inference_diffusion_b200 import _lib
    return _lib


# ---------------------------------------------------------------------------------------------
# fp16-operand ("TF32-class") inference mode
@pytest.mark.parametrize("L,O,H,NB", [(128, 17, 512, 6), (64, 17, 128, 2), (64, 40, 192, 3)])
@pytest.mark.parametrize("B", [7, 256])
def test_f16_score_forward_meets_rel_1e3(L, O, H, NB, B):
    net, params = make_score_net(L, O, H, NB, device="cuda")
    g = gen(B + L)
    z = torch.randn(B, L, generator=g)
    obs = torch.randn(B, O, generator=g)
    worst = 0.0
    for name, t in {"discrete": torch.full((B,), 7.0), "t0_x316": torch.zeros(B),
                    "uniform": torch.rand(B, generator=g)}.items():
        with torch.no_grad():
            want = R.score_forward(params, z, t, obs)
            with _lib().operand("f16"):
                got = net(z.cuda(), t.cuda(), obs.cuda()).cpu()
            coarse = net(z.cuda(), t.cuda(), obs.cuda()).cpu()
        e = rel_l2(got, want)
        worst = max(worst, e)
        assert e < F16_TOL, (name, e)
        assert rel_l2(coarse, want) < BF16_SCORE_TOL       # the bf16 library is untouched by the switch
    print(f"f16 score forward L{L} H{H} B{B}: worst rel-L2 = {worst:.2e}")


@pytest.mark.parametrize("L,O,H,NB,T,B", [(128, 17, 512, 6, 50, 64), (64, 17, 128, 2, 10, 256),
                                          (128, 376, 512, 6, 8, 32)])
def test_f16_reverse_diffusion_meets_rel_1e3(L, O, H, NB, T, B):
    from active_inference_diffusion_b200 import DiffusionConfig, LatentDiffusionProcess
    net, params = make_score_net(L, O, H, NB, device="cuda")
    diff = LatentDiffusionProcess(DiffusionConfig(num_diffusion_steps=T), L).cuda()
    g = gen(T + B)
    obs = torch.randn(B, O, generator=g)
    zT = torch.randn(B, L, generator=g)
    noise = torch.randn(T - 1, B, L, generator=g)
    with torch.no_grad():
        want = R.generate_latent_trajectory(params, R.make_schedule(T), zT, obs, list(noise))
        with _lib().operand("f16"):
            got = diff.generate_latent_trajectory(net, B, obs.cuda(), z_init=zT.cuda(), noise=noise.cuda())
    errs = [rel_l2(got[i], want[i]) for i in (1, T // 2, T)]
    print(f"f16 reverse diffusion L{L} O{O} H{H} T{T}: rel-L2 at steps 1, T/2, T = {errs}")
    assert max(errs) < F16_TOL, errs


@pytest.mark.parametrize("L,A,H,B,K,h", [(128, 6, 512, 256, 2, 5), (32, 6, 128, 50, 3, 4)])
def test_f16_efe_rollout_meets_rel_1e3(L, A, H, B, K, h):
    ai, nets, cfg = make_ai(L, A, H)
    ai.use_epistemic = False
    g = gen(B + K)
    z = torch.randn(B, L, generator=g)
    pn = torch.randn(K * h, B, A, generator=g)
    rn = torch.randn(K * h, B, L, generator=g)
    noise = [dict(policy=pn[i], reparam=rn[i]) for i in range(K * h)]
    ecfg = dict(epistemic_weight=cfg.epistemic_weight, pragmatic_weight=cfg.pragmatic_weight,
                consistency_weight=cfg.consistency_weight, discount_factor=cfg.discount_factor,
                preference_temperature=float(cfg.preference_temperature))
    with torch.no_grad():
        want, _, wfirst = R.expected_free_energy(nets, ecfg, z, h, K, noise)
        with _lib().operand("f16"):
            got, _ = ai.compute_expected_free_energy_diffusion(z.cuda(), horizon=h, num_trajectories=K,
                                                               policy_noise=pn.cuda(), reparam_noise=rn.cuda())
            first = ai.last_first_action
    e, ef = rel_l2(got, want), rel_l2(first, wfirst)
    print(f"f16 EFE L{L} H{H} K{K} h{h}: efe rel-L2 {e:.2e}, first action {ef:.2e}")
    assert e < F16_TOL and ef < F16_TOL
    assert int(torch.argmin(got.cpu())) == int(torch.argmin(want))


# ---------------------------------------------------------------------------------------------
# BASELINE cfg#4 dims together: Humanoid-v4 state shape (obs 376, act 17), horizon 15
@pytest.mark.parametrize("operand,tol,h", [("bf16", 2e-2, 15), ("f16", 1e-3, 8)])
def test_cfg4_dims_sampler_and_efe(operand, tol, h):
    """obs 376 / act 17 / L128 / H512 together.  Horizon 15 runs on bf16 operands: the rollout's latents
    double every step (2z + f) and leave the fp16 range after ~14 steps, which the fp16
    library refuses loudly (checked below); fp16 operands are checked at horizon 8."""
    from active_inference_diffusion_b200 import ActiveInferenceConfig, CandidateScorer, DiffusionConfig
    from oracle.harness import perturb_generic, perturb_state_dict
    L, O, A, H, T, B = 128, 376, 17, 512, 6, 48
    torch.manual_seed(0)
    cfg = ActiveInferenceConfig(latent_dim=L, hidden_dim=H, efe_horizon=h, device="cpu",
                                diffusion=DiffusionConfig(num_diffusion_steps=T))
    m = CandidateScorer(O, A, cfg).eval()
    m.latent_score_network.load_state_dict(perturb_state_dict(m.latent_score_network.state_dict()))
    for name in ("policy_network", "latent_dynamics", "value_network", "reward_predictor"):
        mod = getattr(m, name)
        mod.load_state_dict(perturb_generic(mod.state_dict(), 7, 0.05))
    sub = lambda mod: {k: v.detach().cpu().clone() for k, v in mod.state_dict().items()}
    nets = dict(policy=sub(m.policy_network), dynamics=sub(m.latent_dynamics), value=sub(m.value_network),
                reward=sub(m.reward_predictor))
    sp = sub(m.latent_score_network)
    m = m.cuda()
    g = gen(4)
    obs = torch.randn(B, O, generator=g).clamp_(-1, 1)
    zT = torch.randn(B, L, generator=g)
    noise = torch.randn(T - 1, B, L, generator=g)
    pn = torch.randn(h, B, A, generator=g)
    rn = torch.randn(h, B, L, generator=g)
    ecfg = dict(epistemic_weight=cfg.epistemic_weight, pragmatic_weight=cfg.pragmatic_weight,
                consistency_weight=cfg.consistency_weight, discount_factor=cfg.discount_factor,
                preference_temperature=float(cfg.preference_temperature))
    with torch.no_grad():
        wlat = R.generate_latent_trajectory(sp, R.make_schedule(T), zT, obs, list(noise))[-1]
        want, _, wfirst = R.expected_free_energy(nets, ecfg, wlat, h, 1, [dict(policy=pn[i], reparam=rn[i]) for i in range(h)])
        with _lib().operand(operand):
            lat = m.latent_diffusion.generate_latent_trajectory(m.latent_score_network, B, obs.cuda(), z_init=zT.cuda(),
                                                                noise=noise.cuda(), return_trajectory=False)[-1]
            efe, first, _, _ = m.heads.efe_rollout(lat, h, 1, m.efe_config(), m.preference_temperature, pn.cuda(), rn.cuda())
    errs = (rel_l2(lat, wlat), rel_l2(efe, want), rel_l2(first, wfirst))
    print(f"cfg#4 dims [{operand}, h={h}]: latent {errs[0]:.2e}, efe {errs[1]:.2e}, first action {errs[2]:.2e}")
    assert max(errs) < tol, errs
    if operand == "f16":
        with _lib().operand("f16"), pytest.raises(RuntimeError, match="fp16 tensor-core operands"):
            m.heads.efe_rollout(lat, 15, 1, m.efe_config(), m.preference_temperature,
                                torch.zeros(15, B, A, device="cuda"), torch.zeros(15, B, L, device="cuda"))


# ---------------------------------------------------------------------------------------------
# cfg#5 chain: DrQ-v2 encoder -> sampler -> EFE through CandidateScorer.forward_pixels
def test_cfg5_forward_pixels_chain_matches_oracle_pieces():
    from active_inference_diffusion_b200 import (ActiveInferenceConfig, CandidateScorer, DiffusionConfig, DrQV2Encoder)
    from oracle.harness import perturb_generic, perturb_state_dict
    L, A, H, T, h, B, F = 64, 6, 128, 5, 3, 6, 64
    torch.manual_seed(0)
    enc = DrQV2Encoder((3, 36, 36), feature_dim=F, frame_stack=3, num_filters=8).eval()
    cfg = ActiveInferenceConfig(latent_dim=L, hidden_dim=H, efe_horizon=h, device="cpu",
                                diffusion=DiffusionConfig(num_diffusion_steps=T))
    m = CandidateScorer(F, A, cfg).eval()
    m.latent_score_network.load_state_dict(perturb_state_dict(m.latent_score_network.state_dict()))
    for name in ("policy_network", "latent_dynamics", "value_network", "reward_predictor"):
        mod = getattr(m, name)
        mod.load_state_dict(perturb_generic(mod.state_dict(), 7, 0.05))
    sub = lambda mod: {k: v.detach().cpu().clone() for k, v in mod.state_dict().items()}
    nets = dict(policy=sub(m.policy_network), dynamics=sub(m.latent_dynamics), value=sub(m.value_network),
                reward=sub(m.reward_predictor))
    sp, ep = sub(m.latent_score_network), sub(enc)
    m, enc = m.cuda(), enc.cuda()
    g = gen(55)
    pixels = torch.randint(0, 256, (B, 9, 36, 36), generator=g, dtype=torch.uint8)
    with torch.no_grad():
        wfeat = R.encoder_forward(ep, pixels.float() / 255.0)
        feat = enc(pixels.cuda())
        assert rel_l2(feat, wfeat) < 3e-2                       # bf16 encoder bound
        # the chain call: same features, library-drawn noise -> compare against the oracle fed with the
        # chain's own latent (noise is drawn in-kernel; the sampler itself is tested with injected noise)
        m.latent_diffusion.seed_philox(5, torch.device("cuda", 0))
        efe, first, lat = m.forward_pixels(enc, pixels.cuda(), horizon=h, num_trajectories=1)
        assert lat.shape == (B, L) and torch.isfinite(lat).all() and torch.isfinite(efe).all()
        # replay the same Philox stream through the plain sampler call on the oracle features: the chain
        # must have consumed exactly `feat` as its observation
        m.latent_diffusion.seed_philox(5, torch.device("cuda", 0))
        lat2 = m.latent_diffusion.generate_latent_trajectory(m.latent_score_network, B, feat, return_trajectory=False)[-1]
        assert torch.equal(lat, lat2)
        # and that latent agrees with the oracle sampler run on the oracle's features with the same draws
        st = m.latent_diffusion.philox_state(torch.device("cuda", 0)).clone()
        st[1] = 1
        zT = _philox(st, 0, B, L)
        noise = [_philox(st, 1 + i, B, L) for i in range(T - 1)]
        wlat = R.generate_latent_trajectory(sp, R.make_schedule(T), zT.cpu(), wfeat, [n.cpu() for n in noise])[-1]
        assert rel_l2(lat, wlat) < 5e-2, rel_l2(lat, wlat)
    with pytest.raises(ValueError):
        CandidateScorer(F + 1, A, cfg).cuda().forward_pixels(enc, pixels.cuda())


def _philox(state: torch.Tensor, draw: int, rows: int, cols: int, row_offset: int = 0) -> torch.Tensor:
    lib = _lib()
    out = torch.empty(rows, cols, dtype=torch.float32, device=state.device)
    lib.check(lib.lib().aid_philox_normal(state.data_ptr(), draw, row_offset, out.data_ptr(), rows, cols,
                                          lib.stream_ptr(state.device)), "aid_philox_normal")
    return out


# ---------------------------------------------------------------------------------------------
# Philox noise stream
def test_philox_normals_distribution_and_row_addressing():
    dev = torch.device("cuda", 0)
    st = torch.tensor([1234567, 3], dtype=torch.int64, device=dev)
    x = _philox(st, 0, 65536, 128)
    n = x.numel()
    assert abs(float(x.mean())) < 4.0 / math.sqrt(n) * 1.0 + 1e-4
    assert abs(float(x.var()) - 1.0) < 5e-3
    assert abs(float((x ** 4).mean()) - 3.0) < 5e-2                    # kurtosis of a normal
    assert float(x.abs().max()) < 6.5
    # rows and columns are uncorrelated
    assert abs(float((x[:-1] * x[1:]).mean())) < 1e-3
    assert abs(float((x[:, :-1] * x[:, 1:]).mean())) < 1e-3
    assert abs(float((x[:, 0::4] * x[:, 1::4]).mean())) < 2e-3         # the two Box-Muller outputs of one pair
    # a row's values depend on its GLOBAL index only: a shard starting at row 1000 reproduces rows 1000..
    y = _philox(st, 0, 512, 128, row_offset=1000)
    assert torch.equal(y, x[1000:1512])
    # other draws / call offsets / seeds are different streams
    assert not torch.equal(_philox(st, 1, 64, 128), x[:64])
    st2 = st.clone(); st2[1] += 1
    assert not torch.equal(_philox(st2, 0, 64, 128), x[:64])
    st3 = st.clone(); st3[0] += 1
    assert abs(float((_philox(st3, 0, 4096, 128) * x[:4096]).mean())) < 5e-3


@pytest.mark.parametrize("graph", [False, True])
def test_philox_sampler_is_deterministic_fresh_and_shard_invariant(graph):
    from active_inference_diffusion_b200 import DiffusionConfig, LatentDiffusionProcess
    # 700 rows in shards of 300 / 400: every call takes the tcgen05 launch chain, whose per-row arithmetic does
    # not depend on the batch (calls of <= 256 rows take the persistent kernel of csrc/small.inc, which draws the
    # SAME numbers but sums in another order: tests/test_gpu_score.py::test_small_batch_kernel_matches_launch_chain)
    L, O, H, NB, T, B = 64, 17, 128, 2, 6, 700
    dev = torch.device("cuda", 0)
    net, params = make_score_net(L, O, H, NB, device="cuda")
    diff = LatentDiffusionProcess(DiffusionConfig(num_diffusion_steps=T), L).cuda()
    diff.noise_source, diff.use_graph = "philox", graph
    obs = torch.randn(B, O, generator=gen(1)).cuda()

    def run(o, row_offset=0):
        diff.row_offset = row_offset
        return diff.generate_latent_trajectory(net, o.shape[0], o, return_trajectory=False)[-1]

    with torch.no_grad():
        diff.seed_philox(42, dev)
        a1, a2 = run(obs), run(obs)                    # two calls of one stream: fresh noise
        diff.seed_philox(42, dev)
        b1 = run(obs)
        assert torch.equal(a1, b1) and not torch.equal(a1, a2)
        assert torch.isfinite(a1).all() and float(a1.std()) > 0.05
        # the same rows scored as two shards (global row offsets 0 / 300) see the same draws
        diff.seed_philox(42, dev)
        lo = run(obs[:300])
        diff.seed_philox(42, dev)
        hi = run(obs[300:], row_offset=300)
        assert torch.equal(torch.cat([lo, hi]), a1)
        # the draws are the documented ones: feeding them as injected noise reproduces the latent
        st = diff.philox_state(dev).clone()
        st[1] = 1
        zT = _philox(st, 0, B, L)
        noise = torch.stack([_philox(st, 1 + i, B, L) for i in range(T - 1)])
        diff.noise_source = "torch"
        inj = diff.generate_latent_trajectory(net, B, obs, z_init=zT, noise=noise, return_trajectory=False)[-1]
        assert torch.equal(inj, a1)
        want = R.generate_latent_trajectory(params, R.make_schedule(T), zT.cpu(), obs.cpu(), list(noise.cpu()))[-1]
        assert rel_l2(a1, want) < 2e-2


# ---------------------------------------------------------------------------------------------
# CUDA-graph replay == direct launches
def test_graphed_sampler_and_scorer_match_direct_launches():
    from active_inference_diffusion_b200 import ActiveInferenceConfig, CandidateScorer, DiffusionConfig
    from oracle.harness import perturb_generic, perturb_state_dict
    L, O, A, H, T, h, B = 64, 17, 6, 128, 7, 3, 130
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    cfg = ActiveInferenceConfig(latent_dim=L, hidden_dim=H, efe_horizon=h, device="cpu",
                                diffusion=DiffusionConfig(num_diffusion_steps=T))
    m = CandidateScorer(O, A, cfg).eval()
    m.latent_score_network.load_state_dict(perturb_state_dict(m.latent_score_network.state_dict()))
    for name in ("policy_network", "latent_dynamics", "value_network", "reward_predictor"):
        mod = getattr(m, name)
        mod.load_state_dict(perturb_generic(mod.state_dict(), 7, 0.05))
    m = m.cuda()
    g = gen(9)
    obs = torch.randn(B, O, generator=g).cuda()
    zT = torch.randn(B, L, generator=g).cuda()
    noise = torch.randn(T - 1, B, L, generator=g).cuda()
    d = m.latent_diffusion
    with torch.no_grad():
        outs = []
        for use_graph in (False, True, True):               # capture, then a pure replay
            d.use_graph = use_graph
            outs.append(d.generate_latent_trajectory(m.latent_score_network, B, obs, z_init=zT, noise=noise,
                                                     return_trajectory=False)[-1])
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
        # new inputs through the captured graph
        obs2 = obs.flip(0).contiguous()
        d.use_graph = False
        want = d.generate_latent_trajectory(m.latent_score_network, B, obs2, z_init=zT, noise=noise, return_trajectory=False)[-1]
        d.use_graph = True
        got = d.generate_latent_trajectory(m.latent_score_network, B, obs2, z_init=zT, noise=noise, return_trajectory=False)[-1]
        assert torch.equal(want, got)
        # a weight update re-packs and re-captures
        m.latent_score_network.latent_proj.weight.mul_(1.01)
        got2 = d.generate_latent_trajectory(m.latent_score_network, B, obs2, z_init=zT, noise=noise, return_trajectory=False)[-1]
        d.use_graph = False
        want2 = d.generate_latent_trajectory(m.latent_score_network, B, obs2, z_init=zT, noise=noise, return_trajectory=False)[-1]
        assert torch.equal(want2, got2) and not torch.equal(got2, got)

        # whole scorer call: same torch + Philox generator state -> same (efe, action, latent)
        res = []
        for use_graph in (False, True, True):
            m.use_graph = use_graph
            d.use_graph = False
            d.seed_philox(77, dev)
            torch.cuda.manual_seed(5)
            res.append(m(obs))
        for a, b in zip(res[0], res[1]):
            assert torch.equal(a, b)
        for a, b in zip(res[0], res[2]):
            assert torch.equal(a, b)
        nl = _lib().launch_count()
        m(obs)
        assert _lib().launch_count() == nl                   # a replay launches nothing from the host side


# ---------------------------------------------------------------------------------------------
# a8 / a8b: element-wise diffusion functions
def test_forward_process_and_loss_weight_functions_match_oracle():
    from active_inference_diffusion_b200 import DiffusionConfig, LatentDiffusionProcess
    L, T, B = 32, 25, 77
    for sched in ("cosine", "linear"):
        diff = LatentDiffusionProcess(DiffusionConfig(num_diffusion_steps=T, beta_schedule=sched), L).cuda()
        with torch.no_grad():
            diff.log_snr_min.fill_(-9.3); diff.log_snr_max.fill_(8.7)
            diff.latent_prior_mean.copy_(torch.randn(L, generator=gen(1)) * 0.1)
            diff.latent_prior_log_std.copy_(torch.randn(L, generator=gen(2)) * 0.1)
        dp = {k: v.detach().cpu() for k, v in diff.state_dict().items()}
        s = R.make_schedule(T, sched)
        g = gen(3)
        z0 = torch.randn(B, L, generator=g)
        eps = torch.randn(B, L, generator=g)
        ti = torch.randint(0, T, (B,), generator=g)
        tc = torch.rand(B, generator=g)
        with torch.no_grad():
            got, n = diff.q_sample(z0.cuda(), ti.cuda(), eps.cuda())                 # a8b, core/diffusion.py:154-174
            want = R.q_sample(s, z0, ti, eps)
            assert torch.equal(n.cpu(), eps) and torch.allclose(got.cpu(), want, rtol=1e-6, atol=1e-7)
            got, n, info = diff.continuous_q_sample(z0.cuda(), tc.cuda(), eps.cuda())    # a8, :56-91
            want, wlam, walpha, wsigma = R.continuous_q_sample(dp, z0, tc, eps)
            assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-6)
            assert torch.allclose(info["sigma"].cpu(), wsigma, rtol=1e-5, atol=1e-8)
            assert torch.allclose(info["alpha"].cpu(), walpha, rtol=1e-5, atol=1e-8)
            assert torch.allclose(info["log_snr"].cpu(), wlam, rtol=1e-6, atol=1e-6)
            assert torch.allclose(diff.compute_log_snr(tc.cuda()).cpu(), R.log_snr(dp, tc), rtol=1e-6, atol=1e-6)
            assert torch.allclose(diff.compute_loss_weight(tc.cuda()).cpu(), R.loss_weight(dp, tc), rtol=1e-5, atol=1e-9)
            # sample_latent_prior draws randn_like(mean) on the device: replay the generator
            torch.cuda.manual_seed(9)
            prior = diff.sample_latent_prior(B, torch.device("cuda", 0))
            torch.cuda.manual_seed(9)
            e2 = torch.randn(B, L, device="cuda")
            assert torch.allclose(prior.cpu(), R.sample_latent_prior(dp, e2.cpu()), rtol=1e-6, atol=1e-6)


def test_standalone_p_sample_matches_oracle_including_t0_rule():
    from active_inference_diffusion_b200 import DiffusionConfig, LatentDiffusionProcess
    L, T, B = 32, 25, 40
    diff = LatentDiffusionProcess(DiffusionConfig(num_diffusion_steps=T), L).cuda()
    s = R.make_schedule(T)
    g = gen(6)
    z = torch.randn(B, L, generator=g)
    score = torch.randn(B, L, generator=g)
    for t in (T - 1, 7, 1, 0):
        tt = torch.full((B,), t, dtype=torch.long)
        with torch.no_grad():
            det = diff.p_sample(z.cuda(), tt.cuda(), score.cuda(), deterministic=True)
            assert torch.allclose(det.cpu(), R.p_sample(s, z, t, score, None, deterministic=True), rtol=1e-6, atol=1e-6)
            torch.cuda.manual_seed(3)
            sto = diff.p_sample(z.cuda(), tt.cuda(), score.cuda())
            torch.cuda.manual_seed(3)
            eps = torch.randn(B, L, device="cuda").cpu()
            want = R.p_sample(s, z, t, score, None if t == 0 else eps)     # no noise at t == 0 (core/diffusion.py:233)
            assert torch.allclose(sto.cpu(), want, rtol=1e-6, atol=1e-6), t
    # the rule reads t[0] only (batch-global, like the reference): a mixed batch led by 0 gets no noise anywhere
    mixed = torch.tensor([0] + [5] * (B - 1))
    with torch.no_grad():
        out = diff.p_sample(z.cuda(), mixed.cuda(), score.cuda())
        det = diff.p_sample(z.cuda(), mixed.cuda(), score.cuda(), deterministic=True)
    assert torch.equal(out, det)


# ---------------------------------------------------------------------------------------------
# packed-weight cache
def test_packed_cache_invalidation_after_data_edits():
    net, _ = make_score_net(64, 17, 128, 2, device="cuda")
    g = gen(2)
    z, t, obs = torch.randn(9, 64, generator=g).cuda(), torch.full((9,), 5.0).cuda(), torch.randn(9, 17, generator=g).cuda()
    with torch.no_grad():
        base = net(z, t, obs)
        net.output_proj[2].weight.data.mul_(2.0)          # .data edit: no version bump, cache is stale
        stale = net(z, t, obs)
        assert torch.equal(stale, base)
        net.invalidate_packed()
        fresh = net(z, t, obs)
        assert rel_l2(fresh, 2.0 * base) < 2e-2
        # verify mode detects the edit by content, without an explicit invalidate
        net.output_proj[2].weight.data.mul_(0.5)
        net.packed_weights(verify=True)
        assert torch.equal(net(z, t, obs), base)
        # ordinary in-place ops and load_state_dict are seen by the version counters / hooks
        net.output_proj[2].weight.mul_(2.0)
        assert rel_l2(net(z, t, obs), 2.0 * base) < 2e-2
        sd = {k: v.clone() for k, v in net.state_dict().items()}
        sd["output_proj.2.weight"] *= 0.5
        net.load_state_dict(sd)
        assert torch.equal(net(z, t, obs), base)


# ---------------------------------------------------------------------------------------------
# train_epistemic_estimator (the reference's agents call it every 5th train_step)
def test_train_epistemic_estimator_step_matches_oracle_gradient():
    """core/active_inference.py:420-445: loss = -mean(MI), clip at config.gradient_clip, optimizer step.
    With plain SGD (lr 1) the parameter change IS the clipped gradient; compared with autograd on the
    oracle restatement fed the same draws (first call: running mean 0, where the MINE EMA gradient equals
    the plain gradient of log-mean-exp up to its 1e-6 stabiliser)."""
    L, A, H, B, S = 32, 6, 128, 48, 3
    ai, nets, cfg = make_ai(L, A, H)
    est = ai.epistemic_estimator
    own = {k: v.detach().cpu() for k, v in est.state_dict().items()
           if not k.startswith("decoder.") and k not in ("perturbation_scale", "running_mean")}
    est.load_state_dict(perturb_generic(own, 9, 0.05), strict=False)
    with pytest.raises(RuntimeError, match="epistemic_optimizer"):
        ai.train_epistemic_estimator(torch.zeros(B, L).cuda(), torch.zeros(B, A).cuda(), torch.zeros(B, L).cuda())
    names = [k for k, _ in est.named_parameters() if not k.startswith("decoder.")]
    params = [p for k, p in est.named_parameters() if not k.startswith("decoder.")]
    ai.epistemic_optimizer = torch.optim.SGD(params, lr=1.0)
    before = [p.detach().clone() for p in params]
    g = gen(78)
    lat, act = torch.randn(B, L, generator=g), torch.randn(B, A, generator=g)
    z_eps = [torch.randn(B, L, generator=g) for _ in range(S)]
    dir_eps = [torch.randn(S * B, L, generator=g) for _ in range(4)]
    perms = [torch.randperm(B, generator=g) for _ in range(S)]
    mi, metrics = ai.train_epistemic_estimator(lat.cuda(), act.cuda(), lat.cuda(), num_samples=S,
                                               z_noise=[e.cuda() for e in z_eps], dir_noise=[e.cuda() for e in dir_eps],
                                               perms=[p.cuda() for p in perms])
    step = {k: (b - p.detach()).cpu() for k, b, p in zip(names, before, params)}
    # oracle: same draws, autograd
    ep = {k: v.detach().cpu().clone() for k, v in zip(names, before)}
    ep = {k: v.requires_grad_(True) for k, v in ep.items()}
    ep["perturbation_scale"] = est.perturbation_scale.detach().cpu()
    mean, logvar = R.predict_next_latent(nets["dynamics"], lat, act)
    want, wmi, _, _, _ = R.epistemic_value(ep, nets["decoder"], mean, logvar, z_eps, dir_eps, perms, 0.0)
    (-want.mean()).backward()
    assert abs(mi - float(want.mean())) < 2e-3
    grads = {k: ep[k].grad for k in names if ep[k].grad is not None}
    total = torch.sqrt(sum((v.double() ** 2).sum() for v in grads.values()))
    clip = min(1.0, float(cfg.gradient_clip) / (float(total) + 1e-6))
    checked = 0
    for k, gr in grads.items():
        if float(gr.abs().max()) == 0.0:
            assert float(step[k].abs().max()) < 1e-7
            continue
        assert rel_l2(step[k], gr * clip) < 5e-3, (k, rel_l2(step[k], gr * clip))
        checked += 1
    assert checked >= 8 or float(wmi) <= 0.0


def test_batched_belief_update_equals_separate_calls():
    """train_utils.update_belief_batched: the three diffusion runs of a training step as one call; with
    the in-kernel Philox stream addressed by global row, each set's latents equal what a separate call
    would draw for those rows."""
    from active_inference_diffusion_b200 import update_belief_batched
    L, A, H, T, B = 32, 6, 128, 6, 70
    ai, nets, _ = make_ai(L, A, H, T)
    dev = torch.device("cuda", 0)
    d = ai.latent_diffusion
    d.noise_source = "philox"
    g = gen(3)
    obs, nxt = torch.randn(B, L, generator=g).cuda(), torch.randn(B, L, generator=g).cuda()
    d.seed_philox(9, dev)
    infos = update_belief_batched(ai, [obs, nxt, nxt])
    assert [tuple(i["latent"].shape) for i in infos] == [(B, L)] * 3
    # separate calls over the same global rows of the same noise stream call
    outs = []
    for k, o in enumerate([obs, nxt, nxt]):
        d.seed_philox(9, dev)
        d.row_offset = k * B
        outs.append(ai.update_belief_via_diffusion(o)["latent"])
    d.row_offset = 0
    for a, b in zip(infos, outs):
        assert torch.equal(a["latent"], b)
    assert not torch.equal(infos[1]["latent"], infos[2]["latent"])      # same observations, independent noise rows
    rec = torch.nn.functional.mse_loss(ai.decode_observation(outs[0]), obs)
    assert abs(float(infos[0]["reconstruction_error"]) - float(rec)) < 1e-6 * (1 + float(rec))


def test_grouped_estimator_equals_sequential_evaluations_and_efe_graph_replays():
    """The EFE loop calls the MINE estimator once per (k, t) (core/active_inference.py:337-378); per step the K
    calls are independent, so they run as ONE grouped launch sequence (aid_epistemic_forward_grouped) and the
    running mean is advanced afterwards in the reference's order (aid_ema_sequence).  With the same draws every
    group's statistic must equal the stand-alone evaluation of its rows, and the running mean the sequential
    one.  Then: act()-style EFE (no injected draws) replays as a CUDA graph and keeps advancing the running mean."""
    L, A, H, K, Bg, S = 32, 6, 128, 3, 5, 4
    ai, _, cfg = make_ai(L, A, H)
    est = ai.epistemic_estimator
    assert est.fused and not est.training
    g = gen(11)
    Bt, N = K * Bg, S * K * Bg
    mean = torch.randn(Bt, L, generator=g).cuda()
    logvar = torch.full((Bt, L), math.log(0.1)).cuda()
    zn = torch.randn(S, Bt, L, generator=g).cuda()
    dn = torch.randn(4, N, L, generator=g).cuda()
    perms = torch.stack([torch.stack([torch.randperm(Bg, generator=g) for _ in range(K)]) for _ in range(S)]).cuda()
    with torch.no_grad():
        est.running_mean.zero_()
        grouped = est.forward_grouped(mean, logvar, S, K, z_noise=zn, dir_noise=dn, perms=perms)
        assert float(est.running_mean) == 0.0                       # untouched by the grouped call
        seq = []
        for k in range(K):
            rows = slice(k * Bg, (k + 1) * Bg)
            e, st = est.forward_device(mean[rows], logvar[rows], S, z_noise=[zn[s, rows] for s in range(S)],
                                       dir_noise=[dn[j].view(S, Bt, L)[:, rows].reshape(S * Bg, L) for j in range(4)],
                                       perms=[perms[s, k] for s in range(S)])
            seq.append(st.clone())
        rm_seq = float(est.running_mean)
        est.running_mean.zero_()
        est.apply_running_mean(grouped[:, 3])
        rm_grouped = float(est.running_mean)
    for k in range(K):
        for j in range(3):
            a, b = float(grouped[k, j]), float(seq[k][j])
            assert abs(a - b) <= 1e-5 * (1 + abs(b)), (k, j, a, b)
    assert abs(rm_grouped - rm_seq) <= 1e-6 * (1 + abs(rm_seq)), (rm_grouped, rm_seq)
    # the whole evaluation as a CUDA graph: two replays draw different noise and advance the running mean twice
    lat = torch.randn(2, L, generator=g).cuda()
    with torch.no_grad():
        est.running_mean.zero_()
        e1, i1 = ai.compute_expected_free_energy_diffusion(lat, horizon=3, num_trajectories=K, num_ambiguity_samples=S)
        rm1 = float(est.running_mean)
        e2, i2 = ai.compute_expected_free_energy_diffusion(lat, horizon=3, num_trajectories=K, num_ambiguity_samples=S)
        rm2 = float(est.running_mean)
    assert "_efe_graphs" in ai.__dict__ and len(ai._efe_graphs) == 1 and ai.efe_graph == "auto"
    assert torch.isfinite(e1).all() and torch.isfinite(e2).all() and not torch.equal(e1, e2)
    assert rm1 != 0.0 and rm2 != rm1
    assert abs(i1["epistemic/running_mean"] - rm1) < 1e-6 * (1 + abs(rm1))


# ---------------------------------------------------------------------------------------------
# fused epistemic estimator (csrc/epistemic.inc, aid_epistemic_forward)
@pytest.mark.parametrize("L,A,H,B,S,fused", [(32, 6, 128, 48, 3, True), (32, 6, 128, 48, 3, False),
                                             (128, 6, 512, 200, 10, True), (64, 17, 256, 130, 5, True)])
def test_epistemic_estimator_fused_and_batched_paths_vs_oracle(L, A, H, B, S, fused):
    """core/active_inference.py:940-1063 (+ decoder shim) with the reference's draws injected: the fused
    launch sequence on fp16 operands and the batched aid_gemm_nt (bf16x3) evaluation both reproduce the
    oracle's statistic; two consecutive calls exercise both branches of the running-mean rule (:828-836)."""
    ai, nets, cfg = make_ai(L, A, H)
    est = ai.epistemic_estimator
    est.fused = fused
    own = {k: v.detach().cpu() for k, v in est.state_dict().items()
           if not k.startswith("decoder.") and k not in ("perturbation_scale", "running_mean")}
    est.load_state_dict(perturb_generic(own, 9, 0.05), strict=False)
    ep = {k: v.detach().cpu().clone() for k, v in est.state_dict().items() if not k.startswith("decoder.")}
    g = gen(77)
    rm = 0.0
    for call in range(2):
        mean = torch.randn(B, L, generator=g)
        logvar = torch.full((B, L), float(torch.log(torch.tensor(0.1)))) + 0.1 * torch.randn(B, L, generator=g)
        z_eps = [torch.randn(B, L, generator=g) for _ in range(S)]
        dir_eps = [torch.randn(S * B, L, generator=g) for _ in range(4)]
        perms = [torch.randperm(B, generator=g) for _ in range(S)]
        with torch.no_grad():
            want, mi, joint, marg, rm = R.epistemic_value(ep, nets["decoder"], mean, logvar, z_eps, dir_eps, perms, rm)
            got, stats = est.forward_device(mean.cuda(), logvar.cuda(), S, z_noise=[e.cuda() for e in z_eps],
                                            dir_noise=[e.cuda() for e in dir_eps], perms=[p.cuda() for p in perms])
        m = est.metrics_from(stats)
        assert abs(m["epistemic/joint_term"] - float(joint)) < 1e-3 * (1 + abs(float(joint))), (call, m, float(joint))
        assert abs(m["epistemic/marginal_term"] - float(marg)) < 1e-3 * (1 + abs(float(marg))), (call, m, float(marg))
        assert abs(m["epistemic/mi_estimate"] - float(mi)) < 2e-3, (call, m, float(mi))
        assert torch.allclose(got.cpu(), want, atol=2e-3)
        assert abs(m["epistemic/running_mean"] - rm) < 1e-3 * (1 + abs(rm)), (call, m, rm)
        assert abs(float(est.running_mean) - rm) < 1e-3 * (1 + abs(rm))


def test_epistemic_fused_sharded_partials_merge_to_global_statistic():
    """SURVEY 8e row 2 on one process: the per-shard partials of aid_epistemic_forward, merged the way
    distributed.merge_mine_partials all-reduces them, give the statistic of the unsharded T values."""
    from active_inference_diffusion_b200 import _lib
    from active_inference_diffusion_b200.distributed import merge_mine_partials
    L, A, H, B, S = 32, 6, 128, 64, 3
    ai, nets, _ = make_ai(L, A, H)
    est = ai.epistemic_estimator
    g = gen(5)
    mean, logvar = torch.randn(B, L, generator=g).cuda(), torch.full((B, L), -2.3).cuda()
    est.data_parallel_group = None
    with torch.no_grad():
        _, stats = est.forward_device(mean, logvar, S)
    # merging a single shard's partials reproduces that shard's own statistic
    class _G: pass
    est.data_parallel_group = torch.distributed.group.WORLD if torch.distributed.is_initialized() else object()
    est.running_mean.zero_()
    torch.manual_seed(3)
    with torch.no_grad():
        _, s1 = est.forward_device(mean, logvar, S)
    est.data_parallel_group = None
    est.running_mean.zero_()
    torch.manual_seed(3)
    with torch.no_grad():
        _, s0 = est.forward_device(mean, logvar, S)
    assert torch.allclose(s0, s1, atol=1e-6), (s0, s1)
    # two synthetic shards: partials of halves -> global
    t = torch.randn(2, 500, dtype=torch.float64, generator=gen(1)).cuda()
    def part(tj, tm):
        mx = tm.max()
        return torch.stack([tj.sum(), mx, torch.exp(tm - mx).sum(), torch.tensor(float(tj.numel()), dtype=torch.float64, device=tj.device)])
    pa, pb = part(t[0, :200], t[1, :200]), part(t[0, 200:], t[1, 200:])
    gmax = torch.maximum(pa[1], pb[1])
    merged = torch.stack([pa[0] + pb[0], gmax, pa[2] * torch.exp(pa[1] - gmax) + pb[2] * torch.exp(pb[1] - gmax), pa[3] + pb[3]])
    mi, joint, log_t, t_exp = merge_mine_partials(merged)
    assert abs(float(joint) - float(t[0].mean())) < 1e-6
    assert abs(float(log_t) - float(t[1].exp().mean().log())) < 1e-6


def test_round2_edge_cases_tiny_and_ragged_batches():
    """Single-row and ragged batches through the round-2 entry points: the fused estimator at B = 1
    (N = S rows: less than one row tile), the batched belief update over sets of unequal sizes, an empty
    image batch through the training convolution."""
    from active_inference_diffusion_b200 import conv_ops, update_belief_batched
    L, A, H = 32, 6, 128
    ai, nets, _ = make_ai(L, A, H)
    est = ai.epistemic_estimator
    ep = {k: v.detach().cpu().clone() for k, v in est.state_dict().items() if not k.startswith("decoder.")}
    g = gen(31)
    for B, S in ((1, 10), (3, 1), (129, 2)):
        est.running_mean.zero_()
        mean, logvar = torch.randn(B, L, generator=g), torch.full((B, L), -2.0)
        z_eps = [torch.randn(B, L, generator=g) for _ in range(S)]
        dir_eps = [torch.randn(S * B, L, generator=g) for _ in range(4)]
        perms = [torch.randperm(B, generator=g) for _ in range(S)]
        with torch.no_grad():
            want, mi, joint, marg, rm = R.epistemic_value(ep, nets["decoder"], mean, logvar, z_eps, dir_eps, perms, 0.0)
            got, stats = est.forward_device(mean.cuda(), logvar.cuda(), S, z_noise=[e.cuda() for e in z_eps],
                                            dir_noise=[e.cuda() for e in dir_eps], perms=[p.cuda() for p in perms])
        m = est.metrics_from(stats)
        assert got.shape == (B,)
        assert abs(m["epistemic/mi_estimate"] - float(mi)) < 2e-3, (B, S, m, float(mi))
        assert abs(m["epistemic/joint_term"] - float(joint)) < 1e-3 * (1 + abs(float(joint)))
    d = ai.latent_diffusion
    d.noise_source = "philox"
    d.seed_philox(4, torch.device("cuda", 0))
    sets = [torch.randn(n, L, generator=g).cuda() for n in (1, 130, 7)]
    infos = update_belief_batched(ai, sets)
    assert [tuple(i["latent"].shape) for i in infos] == [(1, L), (130, L), (7, L)]
    assert all(torch.isfinite(i["latent"]).all() for i in infos)
    assert infos[0]["latent_std"].abs().max() == 0            # single-row set: the reference's batch_size == 1 rule
    y = conv_ops.conv3x3(torch.zeros(0, 8, 10, 10).cuda(), torch.zeros(16, 8, 3, 3).cuda(), 1, "f16")
    assert tuple(y.shape) == (0, 16, 10, 10)


def test_efe_trajectories_as_rows_equals_sequential_rollouts():
    """HeadsBundle: K rollouts per candidate evaluated as K*B rows of one rollout (small batches) give
    bit-identical efe / first action / last-step terms to the K sequential rollouts; with supplied
    epistemic scalars the two agree to rounding."""
    L, A, H, B, K, h = 32, 6, 128, 37, 4, 3
    ai, nets, cfg = make_ai(L, A, H)
    ai.use_epistemic = False
    g = gen(12)
    z = torch.randn(B, L, generator=g).cuda()
    pn, rn = torch.randn(K * h, B, A, generator=g).cuda(), torch.randn(K * h, B, L, generator=g).cuda()
    epi = torch.rand(K * h, generator=g).cuda()
    hb = ai._heads
    ecfg = dict(epistemic_weight=cfg.epistemic_weight, pragmatic_weight=cfg.pragmatic_weight,
                consistency_weight=cfg.consistency_weight, discount_factor=cfg.discount_factor)
    out = {}
    for mode, cap in (("rows", 32768), ("sequential", 0)):
        hb.TRAJECTORY_ROWS_MAX = cap
        with torch.no_grad():
            out[mode] = hb.efe_rollout(z, h, K, ecfg, ai.preference_temperature, pn, rn, None)
            out[mode + "_epi"] = hb.efe_rollout(z, h, K, ecfg, ai.preference_temperature, pn, rn, epi)
    del hb.TRAJECTORY_ROWS_MAX
    for a, b in zip(out["rows"], out["sequential"]):
        assert torch.equal(a, b)
    assert torch.allclose(out["rows_epi"][0], out["sequential_epi"][0], rtol=1e-6, atol=1e-6)
    assert not torch.equal(out["rows_epi"][0], out["rows"][0])


@pytest.mark.parametrize("use_epistemic", [False, True])
def test_act_composes_belief_update_efe_and_policy_like_the_reference(use_epistemic):
    """core/active_inference.py:478-531: act() = update_belief_via_diffusion -> compute_expected_free_energy_
    diffusion(horizon = config.efe_horizon, K = 10 default) -> policy(latent, deterministic); action on the
    host with the batch axis squeezed for a single observation; info = belief keys + expected_free_energy,
    action_log_prob, policy_entropy + the EFE info with every tensor read out as a float.  Replaying the same
    calls from the same generator state reproduces act()'s numbers (each piece is parity-tested against the
    oracle on its own)."""
    L, A, H, T = 32, 6, 128, 6
    ai, nets, cfg = make_ai(L, A, H, T)
    ai.use_epistemic = use_epistemic
    obs = torch.randn(L, generator=gen(2))
    torch.manual_seed(11)
    action, info = ai.act(obs, deterministic=True)
    assert not action.is_cuda and tuple(action.shape) == (A,)
    for k in ("latent", "latent_mean", "latent_std", "trajectory_length", "reconstruction_error", "observation",
              "raw_observation", "expected_free_energy", "action_log_prob", "policy_entropy", "epistemic_mean",
              "pragmatic_mean", "consistency_mean", "num_trajectories", "horizon"):
        assert k in info, k
    assert info["horizon"] == cfg.efe_horizon and info["num_trajectories"] == 10
    assert all(isinstance(info[k], float) for k in ("expected_free_energy", "action_log_prob", "policy_entropy",
                                                     "epistemic_mean", "pragmatic_mean", "consistency_mean"))
    if use_epistemic:
        assert "epistemic/mi_estimate" in info and isinstance(info["epistemic/mi_estimate"], float)
    # replay
    ai.epistemic_estimator.running_mean.zero_()
    torch.manual_seed(11)
    with torch.no_grad():
        belief = ai.update_belief_via_diffusion(obs.unsqueeze(0))
        efe, efe_info = ai.compute_expected_free_energy_diffusion(belief["latent"], horizon=cfg.efe_horizon)
        act2, logp, dist = ai.policy_network(belief["latent"], deterministic=True)
    assert torch.equal(info["latent"], belief["latent"])
    assert abs(info["expected_free_energy"] - float(efe.mean())) <= 1e-6 * (1 + abs(float(efe.mean())))
    assert torch.equal(action, act2.squeeze(0).cpu())
    assert abs(info["action_log_prob"] - float(logp.mean())) < 1e-6 * (1 + abs(float(logp.mean())))
    # batch of observations: action keeps its batch axis
    a3, _ = ai.act(torch.randn(5, L, generator=gen(3)))
    assert tuple(a3.shape) == (5, A)
