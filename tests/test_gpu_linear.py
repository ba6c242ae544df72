"""tcgen05 GEMM primitive vs torch on bf16-rounded operands (products exact in fp32, so only the
accumulation order differs): validates the packed (K-major, no-swizzle) tile layout, UMMA descriptors, TMEM
epilogue addressing, ring/accumulator pipelines and every tiling mode."""
import pytest
import torch

from tests.util import gen, max_rel

pytestmark = pytest.mark.gpu

ACTS = {0: lambda x: x, 1: torch.nn.functional.silu, 2: torch.relu,
        3: lambda x: torch.nn.functional.gelu(x)}

SHAPES = [
    (128, 128, 64),      # one tile, one k-block
    (200, 96, 100),      # ragged M, N, K
    (1, 8, 3),           # degenerate
    (1000, 512, 512),    # resident A, 4 n-tiles, 8 row tiles
    (300, 512, 2048),    # streamed A, two 256-column units per row tile
    (257, 256, 1024),    # streamed A, one unit per row tile
    (129, 128, 640),     # streamed A, N=128 MMAs (odd n-tile count)
    (3000, 2048, 512),   # many units per CTA
    (40000, 1024, 512),  # > 148 row tiles: persistent loop over several row tiles per CTA
]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("act", [0, 3])
def test_linear_f32_epilogue(M, N, K, act):
    from active_inference_diffusion_b200 import _lib
    g = gen(M * 7 + N * 3 + K)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    y = _lib.linear(x.cuda(), w.cuda(), b.cuda(), act=act).cpu()
    ref = ACTS[act](x.bfloat16().double() @ w.bfloat16().double().T + b.double()).float()
    assert torch.isfinite(y).all()
    # GELU is evaluated with the tanh form on the MUFU unit: |err| <= 5e-4 abs (precision)
    tol = 2e-5 if act != 3 else 1e-3
    assert max_rel(y, ref) < tol, max_rel(y, ref)


@pytest.mark.parametrize("M,N,K", [(200, 96, 100), (1000, 512, 512), (300, 512, 2048)])
@pytest.mark.parametrize("act", [1, 2, 3])
def test_linear_packed_epilogue(M, N, K, act):
    from active_inference_diffusion_b200 import _lib
    g = gen(M + N + K + act)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    y = _lib.linear(x.cuda(), w.cuda(), b.cuda(), act=act, via_packed=True).cpu()
    ref = ACTS[act](x.bfloat16().double() @ w.bfloat16().double().T + b.double()).float()
    # output rounded to bf16 by the packed epilogue
    assert max_rel(y, ref) < 6e-3, max_rel(y, ref)


def test_linear_no_bias():
    from active_inference_diffusion_b200 import _lib
    g = gen(5)
    x = torch.randn(64, 256, generator=g)
    w = torch.randn(128, 256, generator=g) / 16
    y = _lib.linear(x.cuda(), w.cuda(), None).cpu()
    ref = (x.bfloat16().double() @ w.bfloat16().double().T).float()
    assert max_rel(y, ref) < 2e-5
