"""3x3 convolution of the DrQ-v2 encoder's training graph on the library's tcgen05 GEMMs
(csrc/conv_train.inc): `conv3x3(x, w, stride, precision)` is `F.conv2d(x, w, None, stride, padding=1)` with
its forward, input gradient and weight gradient computed by `aid_conv3x3_forward`, `aid_conv3x3_wgrad` and
(first layer, only when pixel gradients are requested) `aid_conv3x3_dgrad_direct` -- no library
convolution in the autograd graph (reference encoder/visual_encoders.py:56-76,166-176).

Cotangents are multiplied by a power of two chosen on the device from their largest magnitude before they
become tensor-core operands (exact: removed again from the result), so small gradients stay inside the
fp16 range.  The weight gradient always runs on the fp16-operand library (the MN-major weight-gradient
GEMM has no hi/lo split; fp16's 11-bit significand keeps it at ~3e-4)."""
from __future__ import annotations

import torch

from . import _lib


# Images per library call: rows are independent, so larger batches are processed in chunks (forward /
# input gradient: concatenated; weight gradient: summed) and the workspace stays bounded
# (512 images of 42x42x256 fp32 rows = 1 GB).
MAX_IMAGES_PER_CALL = 512


def _workspace(l, dev, n, cin, cout, H, W, stride, prec):
    need = l.aid_conv3x3_workspace_bytes(n, cin, cout, H, W, stride, prec)
    if need == 0:
        _lib.check(-1, "aid_conv3x3_workspace_bytes", l)
    return torch.empty(need, dtype=torch.uint8, device=dev)


def conv3x3_forward(x: torch.Tensor, w: torch.Tensor, stride: int = 1, precision: str = "bf16x3") -> torch.Tensor:
    dev = _lib.require_cuda(x, w)
    if x.shape[0] > MAX_IMAGES_PER_CALL:
        return torch.cat([conv3x3_forward(c, w, stride, precision) for c in x.split(MAX_IMAGES_PER_CALL)], dim=0)
    op, prec = _lib.PRECISIONS[precision]
    x, w = _lib.f32c(x), _lib.f32c(w)
    n, cin, H, W = x.shape
    cout = w.shape[0]
    if tuple(w.shape) != (cout, cin, 3, 3):
        raise ValueError(f"conv3x3: weight {tuple(w.shape)} does not match input channels {cin}")
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    y = torch.empty(n, cout, Ho, Wo, dtype=torch.float32, device=dev)
    if n == 0:
        return y
    l = _lib.lib(op)
    ws = _workspace(l, dev, n, cin, cout, H, W, stride, prec)
    with torch.cuda.device(dev):
        _lib.check(l.aid_conv3x3_forward(x.data_ptr(), w.data_ptr(), None, n, cin, cout, H, W, stride, prec, y.data_ptr(),
                                         ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)), "aid_conv3x3_forward", l)
    return y


def _pow2_scale(t: torch.Tensor) -> torch.Tensor:
    """2^k with max|t| * 2^k in [128, 256): device tensor [1], no host read."""
    amax = t.detach().abs().amax().clamp_min(1e-30)
    return torch.exp2(torch.floor(8.0 - torch.log2(amax))).reshape(1).float()


def conv3x3_wgrad(x: torch.Tensor, dy: torch.Tensor, stride: int = 1, operand: str = "f16") -> torch.Tensor:
    dev = _lib.require_cuda(x, dy)
    if x.shape[0] > MAX_IMAGES_PER_CALL:
        parts = [conv3x3_wgrad(a, b, stride, operand) for a, b in zip(x.split(MAX_IMAGES_PER_CALL), dy.split(MAX_IMAGES_PER_CALL))]
        return torch.stack(parts).sum(dim=0)
    x, dy = _lib.f32c(x), _lib.f32c(dy)
    n, cin, H, W = x.shape
    cout = dy.shape[1]
    dw = torch.empty(cout, cin, 3, 3, dtype=torch.float32, device=dev)
    if n == 0:
        return dw.zero_()
    l = _lib.lib(operand)
    ws = _workspace(l, dev, n, cin, cout, H, W, stride, 0)
    scale = _pow2_scale(dy)
    with torch.cuda.device(dev):
        _lib.check(l.aid_conv3x3_wgrad(x.data_ptr(), dy.data_ptr(), scale.data_ptr(), n, cin, cout, H, W, stride,
                                       dw.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)),
                   "aid_conv3x3_wgrad", l)
    return dw / scale


def conv3x3_dgrad(dy: torch.Tensor, w: torch.Tensor, x_shape, stride: int = 1, precision: str = "bf16x3") -> torch.Tensor:
    if stride == 1:
        # full correlation with the flipped kernel, channels swapped: the same forward kernel
        scale = _pow2_scale(dy)
        wt = w.detach().flip(2, 3).transpose(0, 1).contiguous()
        return conv3x3_forward(dy * scale, wt, 1, precision) / scale
    dev = _lib.require_cuda(dy, w)
    dy, w = _lib.f32c(dy), _lib.f32c(w)
    n, cin, H, W = x_shape
    dx = torch.empty(n, cin, H, W, dtype=torch.float32, device=dev)
    if n == 0:
        return dx
    l = _lib.lib()
    with torch.cuda.device(dev):
        _lib.check(l.aid_conv3x3_dgrad_direct(dy.data_ptr(), w.data_ptr(), n, cin, w.shape[0], H, W, stride, dx.data_ptr(),
                                              _lib.stream_ptr(dev)), "aid_conv3x3_dgrad_direct", l)
    return dx


class _Conv3x3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, stride, precision):
        ctx.save_for_backward(x, w)
        ctx.stride, ctx.precision = stride, precision
        return conv3x3_forward(x, w, stride, precision)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx = conv3x3_dgrad(dy, w, x.shape, ctx.stride, ctx.precision) if ctx.needs_input_grad[0] else None
        dw = conv3x3_wgrad(x, dy, ctx.stride) if ctx.needs_input_grad[1] else None
        return dx, dw, None, None


def conv3x3(x: torch.Tensor, w: torch.Tensor, stride: int = 1, precision: str = "bf16x3") -> torch.Tensor:
    return _Conv3x3.apply(x, w, stride, precision)
