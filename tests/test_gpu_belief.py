"""Fokker-Planck belief update (fp64 CUDA kernel) and FreeEnergyComputation vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import restatement as R
from tests.util import gen, make_score_net, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,L", [(1, 128), (1000, 128), (7, 50), (300, 33)])
def test_belief_update_diag_fp64(N, L):
    from active_inference_diffusion_b200 import BeliefDynamics, BeliefDynamicsConfig
    cfg = BeliefDynamicsConfig()
    bd = BeliefDynamics(L, cfg)
    g = gen(N + L)
    mean = torch.randn(N, L, generator=g, dtype=torch.float64)
    var = torch.rand(N, L, generator=g, dtype=torch.float64) + 0.5
    var[0, 0] = 9.99        # hits the max-variance clamp
    obs = torch.randn(N, L, generator=g, dtype=torch.float64)
    score = torch.randn(N, L, generator=g, dtype=torch.float64)
    eps = torch.randn(N, L, generator=g, dtype=torch.float64)
    m, v, p = bd.update_batch(mean.cuda(), var.cuda(), obs.cuda(), score.cuda(), eps.cuda())
    for i in range(min(N, 16)):
        wm, wv, wp = R.belief_update_diag(mean[i].numpy(), var[i].numpy(), obs[i].numpy(), score[i].numpy(), eps[i].numpy(),
                                          dt=cfg.dt, D=cfg.diffusion_coefficient, lr=cfg.learning_rate,
                                          noise_scale=cfg.noise_scale, min_variance=cfg.min_variance,
                                          max_variance=cfg.max_variance)
        # fp64 state: agreement to ~1e-12 relative (exp/sqrt libm vs CUDA differ in the last ulps)
        assert np.allclose(m[i].cpu().numpy(), wm, rtol=1e-11, atol=1e-12)
        assert np.allclose(v[i].cpu().numpy(), wv, rtol=1e-12)
        assert np.allclose(p[i].cpu().numpy(), wp, rtol=1e-12)
    assert float(v.max()) <= cfg.max_variance


def test_belief_dynamics_update_surface():
    from active_inference_diffusion_b200 import BeliefDynamics, BeliefDynamicsConfig
    for full in (False, True):
        cfg = BeliefDynamicsConfig(use_full_covariance=full)
        bd = BeliefDynamics(32, cfg).cuda()
        g = gen(3)
        bd.reset(torch.randn(32, generator=g))
        o, s = torch.randn(32, generator=g), torch.randn(32, generator=g)
        eps = torch.randn(32, generator=g, dtype=torch.float64)
        m0 = bd.mean.cpu().numpy().copy()
        mean, cov = bd.update(o.cuda(), s.cuda(), None, noise=eps.cuda())
        assert mean.dtype == torch.float32 and cov.shape == (32, 32)
        if full:
            wm, wS, wP = R.belief_update_full(m0, np.eye(32), o.double().numpy(), s.double().numpy(), eps.numpy(),
                                              dt=cfg.dt, D=cfg.diffusion_coefficient, lr=cfg.learning_rate,
                                              noise_scale=cfg.noise_scale, min_variance=cfg.min_variance)
            assert np.allclose(bd.covariance.cpu().numpy(), wS, rtol=1e-9)
        else:
            wm, wv, _ = R.belief_update_diag(m0, np.ones(32), o.double().numpy(), s.double().numpy(), eps.numpy(),
                                             dt=cfg.dt, D=cfg.diffusion_coefficient, lr=cfg.learning_rate,
                                             noise_scale=cfg.noise_scale, min_variance=cfg.min_variance,
                                             max_variance=cfg.max_variance)
            assert np.allclose(bd.variance.cpu().numpy(), wv, rtol=1e-12)
        assert np.allclose(bd.mean.cpu().numpy(), wm, rtol=1e-11, atol=1e-12)


def test_free_energy_loss():
    from active_inference_diffusion_b200 import FreeEnergyComputation
    net, params = make_score_net(32, 32, 128, 2, device="cuda")
    g = gen(8)
    states, obs = torch.randn(20, 32, generator=g), torch.randn(20, 32, generator=g)
    fe = FreeEnergyComputation(1.3).cuda()
    got, info = fe.compute_loss(states.cuda(), obs.cuda(), None, net, current_time=0.3)
    want, winfo = R.free_energy_loss(params, fe.log_precision.detach().cpu(), states, obs, 0.3)
    assert abs(float(got) - float(want)) < 1e-3 * abs(float(want))
    assert rel_l2(info["score_regularization"].reshape(1), winfo["score_regularization"].reshape(1)) < 2e-2
