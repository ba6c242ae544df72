"""Training-graph GEMM (`aid_gemm_nt`) and the MatmulNT autograd Function: strided operands
(forward / input-gradient / weight-gradient forms), split-K, both precision modes, first and
second derivatives against torch fp64/fp32 autograd on the same inputs."""
import pytest
import torch

from tests.util import gen, max_rel, rel_l2

pytestmark = pytest.mark.gpu


def _ref(a, b, bias=None):
    y = a.double() @ b.double().T
    return y if bias is None else y + bias.double()


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (7, 5, 3), (200, 96, 100), (1000, 512, 512), (300, 512, 2048),
                                   (2048, 512, 8192),     # weight-gradient shape: split-K over the batch
                                   (130, 384, 4096)])     # odd n-tile count + split-K
@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
def test_gemm_nt_row_major(M, N, K, precision):
    from active_inference_diffusion_b200 import _lib
    g = gen(M + 3 * N + 7 * K)
    a = torch.randn(M, K, generator=g)
    b = torch.randn(N, K, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g)
    y = _lib.gemm_nt(a.cuda(), b.cuda(), bias.cuda(), precision=precision).cpu()
    if precision == "bf16":      # products of bf16-rounded operands are exact in fp32
        assert max_rel(y, _ref(a.bfloat16(), b.bfloat16(), bias)) < 3e-5
    else:                        # hi/lo split: fp32-class result from full-precision operands
        assert max_rel(y, _ref(a, b, bias)) < 3e-5, max_rel(y, _ref(a, b, bias))


@pytest.mark.parametrize("ta,tb", [(False, True), (True, False), (True, True)])
def test_gemm_nt_transposed_views(ta, tb):
    """dY W (b transposed view) and dY^T X (both transposed views) without materialising copies."""
    from active_inference_diffusion_b200 import _lib
    M, N, K = 333, 200, 150
    g = gen(11)
    a_store = torch.randn(K, M, generator=g) if ta else torch.randn(M, K, generator=g)
    b_store = torch.randn(K, N, generator=g) if tb else torch.randn(N, K, generator=g)
    a = a_store.cuda().t() if ta else a_store.cuda()
    b = b_store.cuda().t() if tb else b_store.cuda()
    y = _lib.gemm_nt(a, b, precision="bf16x3").cpu()
    assert max_rel(y, _ref(a.cpu(), b.cpu())) < 3e-5


def test_matmul_nt_first_and_second_derivatives():
    """MatmulNT's backward is built from MatmulNT, so create_graph=True must give the same second
    derivatives as torch's own matmul (the gradient-penalty path, core/active_inference.py:709-729)."""
    from active_inference_diffusion_b200 import autograd_path as AP
    g = gen(3)
    x0 = torch.randn(48, 40, generator=g)
    w0 = torch.randn(24, 40, generator=g) / 6
    v0 = torch.randn(16, 24, generator=g) / 5

    def run(linear, dev, dt):
        x = x0.to(dev, dt).requires_grad_(True)
        w = w0.to(dev, dt).requires_grad_(True)
        v = v0.to(dev, dt).requires_grad_(True)
        y = linear(torch.tanh(linear(x, w)), v)
        gx = torch.autograd.grad(y.sum(), x, create_graph=True)[0]
        pen = ((gx.norm(2, dim=1) - 1.0) ** 2).mean() + (y ** 2).mean()
        pen.backward()
        return [t.detach().double().cpu() for t in (y, gx, x.grad, w.grad, v.grad)]

    want = run(lambda a, b: a @ b.t(), "cpu", torch.float64)
    AP.set_precision("bf16x3")
    got = run(lambda a, b: AP.linear(a, b), "cuda", torch.float32)
    for name, gt, wt in zip(("y", "dy/dx", "x.grad", "w.grad", "v.grad"), got, want):
        assert rel_l2(gt, wt) < 1e-4, (name, rel_l2(gt, wt))
    AP.set_precision("bf16")
    try:
        got = run(lambda a, b: AP.linear(a, b), "cuda", torch.float32)
    finally:
        AP.set_precision("bf16x3")
    for name, gt, wt in zip(("y", "dy/dx", "x.grad", "w.grad", "v.grad"), got, want):
        assert rel_l2(gt, wt) < 2e-2, (name, rel_l2(gt, wt))    # bf16 operand rounding (2^-9 per product)


@pytest.mark.parametrize("M,N", [(1, 1), (7, 5), (300, 130), (4096, 512), (32768, 2048), (1000, 13312)])
def test_colsum_matches_fp64_sum_and_is_deterministic(M, N):
    """aid_colsum = grad.sum(0) of the Linear biases (two-stage, fixed order), incl. row-strided views."""
    from active_inference_diffusion_b200 import _lib, autograd_path as AP
    g = gen(M + N)
    x = torch.randn(M, N, generator=g).cuda()
    want = x.double().sum(0)
    got = _lib.colsum(x)
    assert float((got.double() - want).abs().max()) <= 1e-5 * (float(want.abs().max()) + M ** 0.5)
    assert torch.equal(got, _lib.colsum(x))
    if N >= 8:                                   # a column slice: row stride > N, unaligned start
        v = x[:, 3:N - 1]
        assert torch.allclose(_lib.colsum(v).double(), v.double().sum(0), rtol=1e-5, atol=1e-5 * M ** 0.5)
    # autograd: first derivative is a broadcast, and it composes under create_graph
    y = x[: min(M, 64)].clone().requires_grad_(True)
    s = AP.ColSum.apply(y * y)
    (gy,) = torch.autograd.grad(s.sum(), y, create_graph=True)
    assert torch.allclose(gy, 2 * y)
    gy.sum().backward()
    assert torch.allclose(y.grad, torch.full_like(y, 2.0))


def test_gelu_function_first_and_second_derivatives():
    """autograd_path.Gelu: value, gradient and the double-backward terms (aid_gelu_double_backward) vs
    torch's own exact GELU in fp64 under create_graph."""
    from active_inference_diffusion_b200 import autograd_path as AP
    g = gen(8)
    x0 = torch.randn(257, 131, generator=g) * 2
    v0 = torch.randn(257, 131, generator=g)

    def run(fn, dev, dt):
        x = x0.to(dev, dt).requires_grad_(True)
        v = v0.to(dev, dt).requires_grad_(True)
        y = fn(x) * v
        (gx,) = torch.autograd.grad(y.sum(), x, create_graph=True)
        ((gx ** 2).sum() + y.sum()).backward()
        return [t.detach().double().cpu() for t in (y, gx, x.grad, v.grad)]

    want = run(torch.nn.functional.gelu, "cpu", torch.float64)
    got = run(AP.Gelu.apply, "cuda", torch.float32)
    for name, a, b in zip(("y", "dy/dx", "x.grad", "v.grad"), got, want):
        assert rel_l2(a, b) < 1e-5, (name, rel_l2(a, b))


def test_pair_kernel_variant_subprocess():
    """The cta_group::2 kernels are opt-in via AID_PAIRS=1 (read once per process)."""
    import os, subprocess, sys
    code = ("import torch; from active_inference_diffusion_b200 import _lib; torch.manual_seed(0);"
            "x=torch.randn(700,512,device='cuda'); w=torch.randn(1024,512,device='cuda')/22; b=torch.randn(1024,device='cuda');"
            "y=_lib.linear(x,w,b,act=3,via_packed=True);"
            "r=torch.nn.functional.gelu(x.bfloat16().double()@w.bfloat16().double().T+b.double()).float();"
            "e=float((y-r).abs().max()/r.abs().max()); print(e); assert e<6e-3")
    env = dict(os.environ, AID_PAIRS="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], env=env, cwd=root, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]


@pytest.mark.parametrize("n", [1, 37, 4096])
def test_time_importance_update_is_bit_exact(n):
    """core/active_inference.py:750-771 on the device: same sequential double arithmetic with fp32
    storage as the reference's .item() loop -> identical bits; bin indices exact (integer work)."""
    from active_inference_diffusion_b200 import _lib
    from oracle import restatement as R
    g = gen(n)
    t = torch.rand(n, generator=g)
    t[0] = 0.0
    if n > 2:
        t[1], t[2] = 0.999999, 1.0          # top bin / clamp
    loss = torch.randn(n, generator=g).abs() * 3
    w0 = torch.rand(100, generator=g) + 0.5
    want = R.update_time_importance(w0, t, loss)
    w = w0.clone().cuda()
    bins = _lib.time_importance_update(t.cuda(), loss.cuda(), w, want_bins=True)
    assert torch.equal(w.cpu(), want)
    assert torch.equal(bins.cpu(), R.time_importance_bins(t))


def test_lambda_returns_bit_exact_vs_golden_and_oracle():
    """aid_lambda_returns: bit-identical to the reference's Python loop on the golden
    cases and to the oracle on a 4,099-long batch with long n_steps."""
    import os
    from active_inference_diffusion_b200 import _lib
    from oracle import restatement as R
    fx = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lambda_returns.pt"),
                    weights_only=False)
    for c in fx["cases"]:
        got = _lib.lambda_returns(c["rewards"].cuda(), c["next_values"].cuda(), c["dones"].cuda(), c["gamma"], c["lam"],
                                  c["n_steps"], c["exclude"])
        assert torch.equal(got.cpu(), c["out"]), (c["n_steps"], c["exclude"])
    g = torch.Generator().manual_seed(1)
    B = 4099
    r, nv = torch.randn(B, generator=g), torch.randn(B, generator=g)
    for dones in (torch.rand(B, generator=g) < 0.1, (torch.rand(B, generator=g) < 0.3).float()):
        for n_steps, excl in ((5, False), (17, True), (64, False)):
            want = R.lambda_returns(r, nv, dones, 0.97, 0.9, n_steps, excl)
            got = _lib.lambda_returns(r.cuda(), nv.cuda(), dones.cuda(), 0.97, 0.9, n_steps, excl)
            assert torch.equal(got.cpu(), want), (n_steps, excl, float((got.cpu() - want).abs().max()))
    with pytest.raises(RuntimeError):
        _lib.lambda_returns(r.cuda(), nv.cuda(), dones.cuda(), 0.97, 0.9, 65, False)
