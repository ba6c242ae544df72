"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, the mirror modules keep the reference's state_dict surface, and the product path fails
loudly without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "aid_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aid_[a-z0-9_]+)\s*\(", src)))


@pytest.mark.parametrize("operand,code", [("bf16", 0), ("f16", 1)])
def test_library_exports_every_declared_symbol(operand, code):
    """Both builds of the library (bf16 and fp16 tensor-core operands) load and export the header."""
    from active_inference_diffusion_b200 import _lib
    lib = _lib.lib(operand)
    names = header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{_lib.LIB_PATHS[operand]} does not export {n}"
    assert lib.aid_abi_version() == _lib.ABI_VERSION == 4
    assert lib.aid_operand_type() == code


def test_operand_type_switch_is_scoped():
    from active_inference_diffusion_b200 import _lib
    assert _lib.operand_type() == "bf16"
    with _lib.operand("f16"):
        assert _lib.operand_type() == "f16" and _lib.lib().aid_operand_type() == 1
    assert _lib.operand_type() == "bf16" and _lib.lib().aid_operand_type() == 0
    with pytest.raises(ValueError):
        _lib.set_operand_type("fp8")


def test_size_queries_and_errors_without_gpu():
    from active_inference_diffusion_b200 import _lib
    lib = _lib.lib()
    d = _lib.AidScoreDims(128, 17, 512, 128, 6)
    assert lib.aid_score_num_params(ctypes.byref(d)) == len(_lib.SCORE_PARAM_KEYS) + 6 * len(_lib.SCORE_BLOCK_KEYS)
    packed = lib.aid_score_packed_bytes(ctypes.byref(d))
    # 21.2 M live bf16 parameters (attention folded) ~ 42 MB
    assert 40e6 < packed < 50e6
    assert lib.aid_score_workspace_bytes(ctypes.byref(d), 65536, 50) > 5e8
    bad = _lib.AidScoreDims(128, 17, 500, 128, 6)
    assert lib.aid_score_packed_bytes(ctypes.byref(bad)) == 0
    assert b"hidden_dim" in lib.aid_last_error()
    h = _lib.AidHeadsDims(128, 6, 512, 128)
    assert lib.aid_heads_packed_bytes(ctypes.byref(h)) > 0


def test_no_cpu_fallback():
    from active_inference_diffusion_b200 import LatentScoreNetwork
    net = LatentScoreNetwork(32, 17, 64, num_layers=1).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.randn(2, 32), torch.rand(2), torch.randn(2, 17))


def test_score_param_table_covers_all_parameters():
    from active_inference_diffusion_b200 import LatentScoreNetwork, _lib
    net = LatentScoreNetwork(32, 17, 64, num_layers=2)
    names = set(dict(net.named_parameters()))
    table = set(_lib.SCORE_PARAM_KEYS) | {f"transformer_blocks.{i}.{k}" for i in range(2) for k in _lib.SCORE_BLOCK_KEYS}
    assert table == names


def test_state_dict_surface_matches_reference_keys():
    """Key list recorded from the reference — checkpoints must stay loadable."""
    from active_inference_diffusion_b200 import DiffusionConfig, LatentDiffusionProcess, LatentScoreNetwork
    sd = LatentScoreNetwork(128, 17, 512).state_dict()
    assert list(sd)[:8] == ["time_scale", "output_multiplier", "grad_norm_ema", "time_embed.0.freq_scale",
                            "time_embed.1.weight", "time_embed.1.bias", "time_embed.3.weight", "time_embed.3.bias"]
    assert sd["transformer_blocks.5.attention.in_proj_weight"].shape == (1536, 512)
    assert sd["output_proj.2.weight"].shape == (128, 256) and "output_proj.2.bias" not in sd
    assert sum(v.numel() for k, v in sd.items() if k != "grad_norm_ema") == 27238531
    assert float(sd["output_proj.2.weight"].abs().max()) == 0.0
    dd = LatentDiffusionProcess(DiffusionConfig(num_diffusion_steps=50), 128).state_dict()
    assert list(dd) == ["latent_prior_mean", "latent_prior_log_std", "log_snr_min", "log_snr_max", "betas", "alphas",
                        "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
                        "posterior_variance", "posterior_log_variance_clipped", "loss_weight_cache"]


def test_unknown_schedule_raises_like_reference():
    from active_inference_diffusion_b200 import DiffusionConfig, LatentDiffusionProcess
    with pytest.raises(ValueError, match="Unknown schedule"):
        LatentDiffusionProcess(DiffusionConfig(num_diffusion_steps=10, beta_schedule="sigmoid"), 32)


def test_reverse_coefficients_match_oracle_tables():
    from active_inference_diffusion_b200 import DiffusionConfig, LatentDiffusionProcess
    from oracle import restatement as R
    for sched in ("cosine", "linear"):
        d = LatentDiffusionProcess(DiffusionConfig(num_diffusion_steps=50, beta_schedule=sched), 128)
        c = R.reverse_step_coefficients(R.make_schedule(50, sched))
        got = torch.from_numpy(d.reverse_coefficients())
        for i, k in enumerate(["sqrt_one_minus_ac", "sqrt_recip_alpha", "coef1", "coef2", "sigma"]):
            assert torch.equal(got[i], c[k]), (sched, k)


def test_active_inference_mirror_seeded_init_checksums():
    from active_inference_diffusion_b200 import ActiveInferenceConfig, DiffusionActiveInference, DiffusionConfig
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "active_inference_small.pt"), weights_only=False)
    d = fx["dims"]
    torch.manual_seed(0)
    cfg = ActiveInferenceConfig(hidden_dim=d["H"], latent_dim=d["L"], device="cpu",
                                diffusion=DiffusionConfig(num_diffusion_steps=d["T"]))
    ai = DiffusionActiveInference(observation_dim=d["L"], action_dim=d["A"], latent_dim=d["L"], config=cfg)
    sd = ai.state_dict()
    for k, (s, a) in fx["init_checksums"].items():
        v = sd[k].double()
        assert float(v.sum()) == s and float(v.abs().sum()) == a, k


def test_shard_bounds_partition():
    from active_inference_diffusion_b200.distributed import shard_bounds
    for total in (1, 7, 65536, 262144, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_training_graph_algebra_matches_oracle_on_cpu(monkeypatch):
    """The graph-level rewrites of the training forward (conditioning path shared between the two
    score evaluations, `1 + scale` folded into the modulation bias, single-token attention folded to
    W_o W_v) are algebra, independent of the GEMM kernel: with `autograd_path.linear` replaced by
    F.linear the CPU graph must reproduce the oracle's score and every parameter gradient, first
    order and through the gradient penalty's double backward (core/active_inference.py:709-729)."""
    import torch.nn.functional as F
    from active_inference_diffusion_b200 import autograd_path as AP
    from oracle import restatement as R
    from tests.util import gen, make_score_net, rel_l2
    monkeypatch.setattr(AP, "linear", lambda x, w, b=None: F.linear(x, w, b))
    L, O, H, NB, B = 16, 5, 32, 2, 9
    net, params = make_score_net(L, O, H, NB)
    net = net.double()
    g = gen(4)
    z = torch.randn(B, L, generator=g, dtype=torch.float64)
    obs = torch.randn(B, O, generator=g, dtype=torch.float64)
    for t in (torch.rand(B, generator=g, dtype=torch.float64), torch.full((B,), 7.0, dtype=torch.float64)):
        p = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in params.items()}

        def penalised(score_fn):
            x = z.clone().requires_grad_(True)
            s = score_fn(x)
            (gx,) = torch.autograd.grad(s.sum(), x, create_graph=True)
            return s, (s ** 2).mean() + ((gx.norm(2, dim=1) - 1.0) ** 2).mean()

        s_want, loss_want = penalised(lambda x: R.score_forward(p, x, t, obs))
        loss_want.backward()
        # one conditioning + one fold shared by the evaluation (as compute_diffusion_elbo does)
        mod, tw = AP.score_conditioning(net, t, obs, B)
        folds = AP.fold_attention(net)
        s_got, loss_got = penalised(lambda x: AP.score_from_conditioning(net, x, mod, tw, folds))
        for q in net.parameters():
            q.grad = None
        loss_got.backward()
        assert rel_l2(s_got, s_want) < 1e-10
        assert abs(float(loss_got) - float(loss_want)) < 1e-10 * abs(float(loss_want))
        checked = 0
        for k, q in net.named_parameters():
            if q.grad is None or p[k].grad is None:
                continue
            if float(p[k].grad.abs().max()) == 0.0:
                assert float(q.grad.abs().max()) < 1e-12, k
                continue
            assert rel_l2(q.grad, p[k].grad) < 1e-8, (k, rel_l2(q.grad, p[k].grad))
            checked += 1
        assert checked > 20


def test_graphed_step_and_colsum_fail_loudly_without_cuda():
    """No CPU fallback for the training step either."""
    from active_inference_diffusion_b200 import _lib
    from active_inference_diffusion_b200.train_graph import GraphedElboStep
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError):
        GraphedElboStep(object(), 8)
    with pytest.raises(RuntimeError):
        _lib.colsum(torch.zeros(4, 4))
    with pytest.raises(RuntimeError):
        _lib.gelu_double_backward(torch.zeros(4), torch.zeros(4), torch.zeros(4))
    assert _lib.lib().aid_colsum_workspace_bytes(1000, 512) >= 512 * 4


def test_device_running_mean_std_matches_reference_normaliser():
    """train_utils.RunningMeanStd (device tensors, float64) == the reference's numpy RunningMeanStd
    (agents/base_agent.py:24-52) update by update; runs on CPU tensors here, on the GPU in training."""
    import numpy as np
    from active_inference_diffusion_b200.train_utils import RunningMeanStd

    class Ref:                                   # restated from agents/base_agent.py:24-52
        def __init__(self, epsilon=1e-4):
            self.mean, self.var, self.count = np.zeros(()), np.ones(()), epsilon

        def update(self, x):
            bm, bv, bc = np.mean(x, axis=0), np.var(x, axis=0), x.shape[0]
            delta, tot = bm - self.mean, self.count + bc
            m2 = self.var * self.count + bv * bc + np.square(delta) * self.count * bc / tot
            self.mean, self.var, self.count = self.mean + delta * bc / tot, m2 / tot, tot

        def normalize(self, x):
            return (x - self.mean) / np.sqrt(self.var + 1e-8)

    ours, ref = RunningMeanStd(device="cpu"), Ref()
    g = torch.Generator().manual_seed(0)
    for i in range(5):
        r = torch.randn(256, generator=g) * (1 + i) + 0.3 * i
        ours.update(r)
        ref.update(r.double().numpy())
        want = torch.tensor(ref.normalize(r.double().numpy()), dtype=torch.float32)
        assert torch.allclose(ours.normalize(r), want, rtol=1e-6, atol=1e-6)
    assert abs(float(ours.mean) - float(ref.mean)) < 1e-12 and abs(float(ours.var) - float(ref.var)) < 1e-12


def test_ema_model_matches_reference_arithmetic_and_swaps():
    """train_utils.EMAModel == core/active_inference.py:779-813 bit for bit over several updates (the
    reference class itself when the reference tree is importable, its restated formula otherwise);
    apply_shadow / restore swap param.data and call the module's invalidate_packed."""
    from active_inference_diffusion_b200.train_utils import EMAModel
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.LayerNorm(5), torch.nn.Linear(5, 3))
    net[1].bias.requires_grad_(False)                       # frozen parameters are skipped, as in the reference
    calls = []
    net.invalidate_packed = lambda: calls.append(1)
    ours = EMAModel(net, decay=0.99)
    ref_cls = None
    try:
        from oracle import ref_import
        if ref_import.reference_available():
            ref_import.import_reference()
            from active_inference_diffusion.core.active_inference import EMAModel as ref_cls
    except Exception:
        ref_cls = None
    ref = ref_cls(net, decay=0.99) if ref_cls is not None else None
    want = {n: p.data.clone() for n, p in net.named_parameters() if p.requires_grad}
    g = torch.Generator().manual_seed(1)
    for _ in range(4):
        with torch.no_grad():
            for p in net.parameters():
                p.add_(torch.randn(p.shape, generator=g) * 0.1)
        ours.update()
        if ref is not None:
            ref.update()
        for n, p in net.named_parameters():
            if p.requires_grad:
                want[n] = 0.99 * want[n] + (1 - 0.99) * p.data
    assert set(ours.shadow) == set(want) and "1.bias" not in ours.shadow
    for n in want:
        assert torch.equal(ours.shadow[n], want[n]), n
        if ref is not None:
            assert torch.equal(ours.shadow[n], ref.shadow[n]), n
    before = {n: p.data.clone() for n, p in net.named_parameters()}
    ours.apply_shadow()
    assert torch.equal(net[0].weight.data, ours.shadow["0.weight"]) and len(calls) == 1
    ours.restore()
    assert all(torch.equal(p.data, before[n]) for n, p in net.named_parameters()) and len(calls) == 2


def test_belief_dynamics_diagnostics_keys_and_values():
    """core/belief_dynamics.py:391-410: get_diagnostics of the diagonal and the full-covariance belief."""
    import math
    from active_inference_diffusion_b200 import BeliefDynamics, BeliefDynamicsConfig
    for full, keys in ((False, {"min_variance", "max_variance", "mean_variance", "mean_norm", "entropy"}),
                       (True, {"min_eigenvalue", "max_eigenvalue", "condition_number", "determinant", "mean_norm", "entropy"})):
        c = BeliefDynamicsConfig()
        c.use_full_covariance = full
        b = BeliefDynamics(6, c)
        b.reset(torch.arange(6.0), torch.diag(torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0, 6.0])))
        d = b.get_diagnostics()
        assert set(d) == keys and all(isinstance(v, float) for v in d.values())
        assert abs(d["mean_norm"] - math.sqrt(55.0)) < 1e-12
        want_entropy = 0.5 * (6 * math.log(2 * math.pi * math.e) + math.log(720.0))
        assert abs(d["entropy"] - want_entropy) < 1e-9
        assert abs(d["min_eigenvalue" if full else "min_variance"] - 1.0) < 1e-9
