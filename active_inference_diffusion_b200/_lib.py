"""ctypes binding of libaid_sm100.so (C ABI declared in include/aid_b200.h).

There is no CPU fallback: `lib()` raises if the shared library is missing, and every compute
wrapper raises if the tensors are not CUDA tensors.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_size_t, c_void_p
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# One source tree, one shared library per tensor-core operand type (csrc/gemm.cuh, pack_op16x2):
#   "bf16": bf16 operands (8-bit significand), stated bf16 bounds (precision)
#   "f16" : IEEE fp16 operands (11-bit significand = the TF32 significand) at the same tensor-pipe
#           rate: the rel-1e-3 ("fp32/TF32") contract of the sampled latents, EFE, losses, gradients
LIB_PATHS = {"bf16": os.path.join(_HERE, "libaid_sm100.so"), "f16": os.path.join(_HERE, "libaid_sm100_f16.so")}
LIB_PATH = LIB_PATHS["bf16"]
ABI_VERSION = 4
OPERAND_TYPES = tuple(LIB_PATHS)
_operand = os.environ.get("AID_PRECISION", "bf16")
if _operand not in LIB_PATHS:
    raise RuntimeError(f"AID_PRECISION={_operand!r}: expected one of {sorted(LIB_PATHS)}")


def operand_type() -> str:
    """Operand type the fused inference calls (score forward, sampler, EFE rollout, heads, encoder)
    use on this thread of control: "bf16" or "f16"."""
    return _operand


def set_operand_type(name: str) -> str:
    """Select the tensor-core operand type of the fused calls; returns the previous one."""
    global _operand
    if name not in LIB_PATHS:
        raise ValueError(f"unknown operand type {name!r}; expected one of {sorted(LIB_PATHS)}")
    prev, _operand = _operand, name
    return prev


class operand:
    """`with _lib.operand("f16"):` — operand type for a region."""

    def __init__(self, name: str):
        if name not in LIB_PATHS:
            raise ValueError(f"unknown operand type {name!r}; expected one of {sorted(LIB_PATHS)}")
        self.name = name

    def __enter__(self):
        self.prev = set_operand_type(self.name)

    def __exit__(self, *exc):
        set_operand_type(self.prev)


class AidScoreDims(ctypes.Structure):
    _fields_ = [("latent_dim", c_int32), ("obs_dim", c_int32), ("hidden_dim", c_int32),
                ("time_embed_dim", c_int32), ("num_blocks", c_int32)]


class AidHeadsDims(ctypes.Structure):
    _fields_ = [("latent_dim", c_int32), ("action_dim", c_int32), ("hidden_dim", c_int32),
                ("time_embed_dim", c_int32)]


class AidEpistemicDims(ctypes.Structure):
    _fields_ = [("latent_dim", c_int32), ("hidden_dim", c_int32), ("observation_dim", c_int32),
                ("jacobian_dim", c_int32)]


class AidEncoderDims(ctypes.Structure):
    _fields_ = [("in_channels", c_int32), ("height", c_int32), ("width", c_int32),
                ("num_filters", c_int32), ("num_layers", c_int32), ("feature_dim", c_int32),
                ("use_attention", c_int32), ("precision", c_int32)]


class AidSampleNoise(ctypes.Structure):
    _fields_ = [("z_init", c_void_p), ("noise", c_void_p), ("philox", c_void_p), ("row_offset", c_int64),
                ("deterministic", c_int32)]


class AidEfeConfig(ctypes.Structure):
    _fields_ = [("epistemic_weight", c_float), ("pragmatic_weight", c_float),
                ("consistency_weight", c_float), ("discount_factor", c_float)]


# mirrors enum AidScoreParam / AidScoreBlockParam in include/aid_b200.h
SCORE_PARAM_KEYS = [
    "time_scale", "output_multiplier", "time_embed.0.freq_scale",
    "time_embed.1.weight", "time_embed.1.bias", "time_embed.3.weight", "time_embed.3.bias",
    "obs_encoder.0.weight", "obs_encoder.0.bias", "obs_encoder.1.weight", "obs_encoder.1.bias",
    "obs_encoder.4.weight", "obs_encoder.4.bias", "obs_encoder.5.weight", "obs_encoder.5.bias",
    "obs_encoder.7.weight", "obs_encoder.7.bias", "obs_encoder.8.weight", "obs_encoder.8.bias",
    "continuous_time_embed.0.weight", "continuous_time_embed.0.bias",
    "continuous_time_embed.2.weight", "continuous_time_embed.2.bias",
    "continuous_time_embed.4.weight", "continuous_time_embed.4.bias",
    "latent_proj.weight", "latent_proj.bias",
    "norm_final.adaLN_modulation.1.weight", "norm_final.adaLN_modulation.1.bias",
    "output_proj.0.weight", "output_proj.0.bias", "output_proj.2.weight",
]
SCORE_BLOCK_KEYS = [
    "norm1.adaLN_modulation.1.weight", "norm1.adaLN_modulation.1.bias",
    "norm2.adaLN_modulation.1.weight", "norm2.adaLN_modulation.1.bias",
    "attention.in_proj_weight", "attention.in_proj_bias",
    "attention.out_proj.weight", "attention.out_proj.bias",
    "mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias",
]

_libs: dict = {}


def _declare(l: ctypes.CDLL) -> None:
    P = POINTER
    l.aid_abi_version.restype = c_int32
    l.aid_last_error.restype = c_char_p
    l.aid_device_count.restype = c_int32
    l.aid_launch_count.restype = c_int64
    l.aid_reset_launch_count.restype = None
    l.aid_profile_select.restype = c_int32
    l.aid_profile_select.argtypes = [c_int32, c_int32, c_int32]
    l.aid_profile_collect.restype = c_int32
    l.aid_profile_collect.argtypes = [P(ctypes.c_double), P(c_int64)]
    l.aid_score_packed_bytes.restype = c_size_t
    l.aid_score_packed_bytes.argtypes = [P(AidScoreDims)]
    l.aid_score_num_params.restype = c_int32
    l.aid_score_num_params.argtypes = [P(AidScoreDims)]
    l.aid_score_pack.restype = c_int32
    l.aid_score_pack.argtypes = [P(AidScoreDims), P(c_void_p), c_int32, c_void_p, c_size_t, c_void_p]
    l.aid_score_workspace_bytes.restype = c_size_t
    l.aid_score_workspace_bytes.argtypes = [P(AidScoreDims), c_int32, c_int32]
    l.aid_score_forward.restype = c_int32
    l.aid_score_forward.argtypes = [P(AidScoreDims), c_void_p, c_void_p, c_size_t, c_void_p, c_void_p,
                                    c_void_p, c_int32, c_int32, c_void_p, c_void_p]
    l.aid_sample.restype = c_int32
    l.aid_sample.argtypes = [P(AidScoreDims), c_void_p, c_void_p, c_size_t, c_int32, c_int32,
                             P(c_float), P(c_int32), P(c_float), c_int32, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p]
    l.aid_heads_packed_bytes.restype = c_size_t
    l.aid_heads_packed_bytes.argtypes = [P(AidHeadsDims)]
    l.aid_heads_pack.restype = c_int32
    l.aid_heads_pack.argtypes = [P(AidHeadsDims), P(c_void_p), c_int32, c_void_p, c_size_t, c_void_p]
    l.aid_heads_workspace_bytes.restype = c_size_t
    l.aid_heads_workspace_bytes.argtypes = [P(AidHeadsDims), c_int32]
    l.aid_efe_rollout.restype = c_int32
    l.aid_efe_rollout.argtypes = [P(AidHeadsDims), c_void_p, c_void_p, c_size_t, c_int32, c_int32, c_int32,
                                  P(AidEfeConfig), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    l.aid_head_forward.restype = c_int32
    l.aid_head_forward.argtypes = [P(AidHeadsDims), c_void_p, c_void_p, c_size_t, c_int32, c_int32,
                                   c_void_p, c_void_p, c_void_p, c_void_p]
    l.aid_fp_belief_update.restype = c_int32
    l.aid_fp_belief_update.argtypes = [c_void_p] * 5 + [c_int32, c_int32] + [ctypes.c_double] * 6 + [c_void_p] * 4
    l.aid_linear_workspace_bytes.restype = c_size_t
    l.aid_linear_workspace_bytes.argtypes = [c_int32, c_int32, c_int32]
    l.aid_linear.restype = c_int32
    l.aid_linear.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                             c_int32, c_int32, c_void_p, c_size_t, c_void_p]


def _declare_train(l: ctypes.CDLL) -> None:
    l.aid_gemm_nt_workspace_bytes.restype = c_size_t
    l.aid_gemm_nt_workspace_bytes.argtypes = [c_int32, c_int32, c_int32, c_int32]
    l.aid_gemm_nt.restype = c_int32
    l.aid_gemm_nt.argtypes = [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p,
                              c_int32, c_int32, c_int32, c_int32, c_void_p, c_size_t, c_void_p]


def _declare_misc(l: ctypes.CDLL) -> None:
    l.aid_time_importance_update.restype = c_int32
    l.aid_time_importance_update.argtypes = [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_void_p]
    l.aid_lambda_returns.restype = c_int32
    l.aid_lambda_returns.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, ctypes.c_double, ctypes.c_double,
                                     c_int32, c_int32, c_void_p, c_void_p]


def _declare_colsum(l: ctypes.CDLL) -> None:
    l.aid_colsum_workspace_bytes.restype = c_size_t
    l.aid_colsum_workspace_bytes.argtypes = [c_int32, c_int32]
    l.aid_colsum.restype = c_int32
    l.aid_colsum.argtypes = [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_size_t, c_void_p]
    l.aid_gelu_double_backward.restype = c_int32
    l.aid_gelu_double_backward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]


def _declare_r2(l: ctypes.CDLL) -> None:
    """Round-2 entry points (include/aid_b200.h)."""
    l.aid_operand_type.restype = c_int32
    l.aid_profile_select_slot.restype = c_int32
    l.aid_profile_select_slot.argtypes = [c_int32, c_int32, c_int32, c_int32]
    l.aid_profile_collect_slot.restype = c_int32
    l.aid_profile_collect_slot.argtypes = [c_int32, POINTER(ctypes.c_double), POINTER(c_int64)]
    l.aid_sample_ex.restype = c_int32
    l.aid_sample_ex.argtypes = [POINTER(AidScoreDims), c_void_p, c_void_p, c_size_t, c_int32, c_int32,
                                POINTER(c_float), POINTER(c_int32), POINTER(c_float), c_int32, c_void_p,
                                POINTER(AidSampleNoise), c_void_p, c_void_p, c_void_p]
    D = POINTER(AidScoreDims)
    l.aid_train_packed_bytes.restype = c_size_t
    l.aid_train_packed_bytes.argtypes = [D]
    l.aid_train_num_params.restype = c_int32
    l.aid_train_num_params.argtypes = [D]
    l.aid_train_pack.restype = c_int32
    l.aid_train_pack.argtypes = [D, POINTER(c_void_p), c_int32, c_void_p, c_size_t, c_void_p]
    l.aid_train_workspace_bytes.restype = c_size_t
    l.aid_train_workspace_bytes.argtypes = [D, c_int32]
    l.aid_train_debug_offset.restype = c_int64
    l.aid_train_debug_offset.argtypes = [D, c_int32, c_char_p, c_int32]
    l.aid_dsm_forward.restype = c_int32
    l.aid_dsm_forward.argtypes = [D, c_void_p, c_void_p, c_size_t, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p]
    l.aid_gp_forward_backward.restype = c_int32
    l.aid_gp_forward_backward.argtypes = [D, c_void_p, c_void_p, c_size_t, c_int32, c_int32, c_void_p, c_void_p,
                                          c_void_p, c_void_p, POINTER(c_void_p), c_void_p]
    l.aid_dsm_backward.restype = c_int32
    l.aid_dsm_backward.argtypes = [D, c_void_p, c_void_p, c_size_t, c_int32, c_void_p, c_void_p, c_void_p, c_int32,
                                   POINTER(c_void_p), c_void_p, c_void_p, c_int32, c_int32, c_void_p]
    l.aid_wgrad_workspace_bytes.restype = c_size_t
    l.aid_wgrad_workspace_bytes.argtypes = [c_int32, c_int32, c_int32]
    l.aid_wgrad.restype = c_int32
    l.aid_wgrad.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_size_t, c_void_p]
    l.aid_philox_normal.restype = c_int32
    l.aid_philox_normal.argtypes = [c_void_p, ctypes.c_uint32, c_int64, c_void_p, c_int32, c_int32, c_void_p]


def _declare_epistemic(l: ctypes.CDLL) -> None:
    D = POINTER(AidEpistemicDims)
    l.aid_epistemic_packed_bytes.restype = c_size_t
    l.aid_epistemic_packed_bytes.argtypes = [D]
    l.aid_epistemic_pack.restype = c_int32
    l.aid_epistemic_pack.argtypes = [D, POINTER(c_void_p), c_int32, c_void_p, c_size_t, c_void_p]
    l.aid_epistemic_workspace_bytes.restype = c_size_t
    l.aid_epistemic_workspace_bytes.argtypes = [D, c_int32, c_int32]
    l.aid_epistemic_forward.restype = c_int32
    l.aid_epistemic_forward.argtypes = [D, c_void_p, c_void_p, c_size_t, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p]
    l.aid_epistemic_forward_grouped.restype = c_int32
    l.aid_epistemic_forward_grouped.argtypes = [D, c_void_p, c_void_p, c_size_t, c_int32, c_int32, c_int32, c_void_p,
                                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    l.aid_ema_sequence.restype = c_int32
    l.aid_ema_sequence.argtypes = [c_void_p, c_int32, c_float, c_void_p, c_void_p]


def _declare_conv(l: ctypes.CDLL) -> None:
    l.aid_conv3x3_workspace_bytes.restype = c_size_t
    l.aid_conv3x3_workspace_bytes.argtypes = [c_int32] * 7
    l.aid_conv3x3_forward.restype = c_int32
    l.aid_conv3x3_forward.argtypes = [c_void_p, c_void_p, c_void_p] + [c_int32] * 7 + [c_void_p, c_void_p, c_size_t, c_void_p]
    l.aid_conv3x3_wgrad.restype = c_int32
    l.aid_conv3x3_wgrad.argtypes = [c_void_p, c_void_p, c_void_p] + [c_int32] * 6 + [c_void_p, c_void_p, c_size_t, c_void_p]
    l.aid_conv3x3_dgrad_direct.restype = c_int32
    l.aid_conv3x3_dgrad_direct.argtypes = [c_void_p, c_void_p] + [c_int32] * 6 + [c_void_p, c_void_p]


def _declare_encoder(l: ctypes.CDLL) -> None:
    D = POINTER(AidEncoderDims)
    l.aid_encoder_packed_bytes.restype = c_size_t
    l.aid_encoder_packed_bytes.argtypes = [D]
    l.aid_encoder_num_params.restype = c_int32
    l.aid_encoder_num_params.argtypes = [D]
    l.aid_encoder_pack.restype = c_int32
    l.aid_encoder_pack.argtypes = [D, POINTER(c_void_p), c_int32, c_void_p, c_size_t, c_void_p]
    l.aid_encoder_workspace_bytes.restype = c_size_t
    l.aid_encoder_workspace_bytes.argtypes = [D, c_int32]
    l.aid_encoder_forward.restype = c_int32
    l.aid_encoder_forward.argtypes = [D, c_void_p, c_void_p, c_size_t, c_int32, c_void_p, c_int32, c_void_p,
                                      c_void_p]


def lib(operand_type: Optional[str] = None) -> ctypes.CDLL:
    """Load the CUDA extension of the given (default: current) operand type; fail loudly when it
    has not been built."""
    name = operand_type or _operand
    l = _libs.get(name)
    if l is None:
        path = LIB_PATHS[name]
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  There is no CPU fallback.")
        l = ctypes.CDLL(path)
        _declare(l)
        _declare_train(l)
        _declare_misc(l)
        _declare_encoder(l)
        _declare_colsum(l)
        _declare_r2(l)
        _declare_epistemic(l)
        _declare_conv(l)
        if l.aid_abi_version() != ABI_VERSION:
            raise RuntimeError(f"{path}: ABI version {l.aid_abi_version()} != {ABI_VERSION}; rebuild")
        _libs[name] = l
    return l


def check(rc: int, what: str, l: Optional[ctypes.CDLL] = None) -> None:
    if rc != 0:
        msg = (l or lib()).aid_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed: {msg}")


def require_cuda(*tensors: Optional[torch.Tensor]) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("active_inference_diffusion_b200 runs on CUDA (sm_100a) only: got a "
                               f"{t.device} tensor; there is no CPU fallback")
        dev = t.device if dev is None else dev
    if dev is None:
        raise RuntimeError("no tensor given")
    return dev


class PackedCache:
    """Derived caches of a parameter-holding module: packed tensor-core operand tiles (one entry per
    operand type and device) and scratch workspaces (one per device and stream, so a collector
    thread and the trainer never share scratch memory).

    A pack is rebuilt when the identity or the autograd version counter of any parameter changes
    (optimizer steps, `load_state_dict`, `.to()` all bump one of them).  In-place edits through
    `.data` (`p.data.copy_()`, `p.data.mul_()` — the EMA / target-sync idiom) do NOT bump the
    version counter: call `invalidate()` (modules expose it as `invalidate_packed()`) after them.
    `verify=True` (or AID_VERIFY_PACKED=1) additionally compares a device-side checksum of the
    parameters with the one taken at pack time; it costs a host read per call and is meant for
    debugging and tests."""

    VERIFY = os.environ.get("AID_VERIFY_PACKED", "0") not in ("", "0")

    def __init__(self):
        self.entries: dict = {}
        self.scratch: dict = {}

    def invalidate(self) -> None:
        self.entries.clear()

    @staticmethod
    def _checksum(params) -> torch.Tensor:
        live = [p.detach().double().sum() for p in params if p is not None]
        return torch.stack(live).sum() if live else torch.zeros((), dtype=torch.float64)

    def get(self, tag, params, build, verify: bool = False) -> torch.Tensor:
        """`build()` -> packed tensor; cached under `tag` (operand type, precision ...)."""
        live = [p for p in params if p is not None]
        dev = require_cuda(*live)
        key = tuple((p.data_ptr(), p._version) for p in live)
        full = (tag, dev)
        hit = self.entries.get(full)
        check = verify or self.VERIFY
        if hit is not None and hit[0] == key:
            if not check or bool(self._checksum(live) == hit[2]):
                return hit[1]
        packed = build()
        self.entries[full] = (key, packed, self._checksum(live) if check else None)
        return packed

    def workspace(self, need: int, device: torch.device) -> torch.Tensor:
        sk = (device, torch.cuda.current_stream(device).cuda_stream)
        ws = self.scratch.get(sk)
        if ws is None or ws.numel() < need:
            ws = self.scratch[sk] = torch.empty(need, dtype=torch.uint8, device=device)
        return ws


def f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def launch_count() -> int:
    return int(lib().aid_launch_count())


def reset_launch_count() -> None:
    lib().aid_reset_launch_count()


def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor] = None, act: int = 0,
           via_packed: bool = False) -> torch.Tensor:
    """y = act(x W^T + b) through the tcgen05 GEMM (test primitive)."""
    dev = require_cuda(x, w, b)
    x, w, b = f32c(x), f32c(w), f32c(b)
    M, K = x.shape
    N = w.shape[0]
    l = lib()
    ws_bytes = l.aid_linear_workspace_bytes(M, N, K)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    y = torch.empty(M, N, dtype=torch.float32, device=dev)
    check(l.aid_linear(ptr(x), ptr(w), ptr(b), ptr(y), M, N, K, act, int(via_packed), ptr(ws), ws_bytes,
                       stream_ptr(dev)), "aid_linear")
    return y


def _strided_2d(t: torch.Tensor) -> torch.Tensor:
    """fp32 2-D tensor with a unit stride in one dimension (transposed views are kept as views)."""
    if t.dtype != torch.float32:
        t = t.float()
    if t.stride(1) == 1 or t.stride(0) == 1:
        return t
    return t.contiguous()


# training-graph GEMM precisions: (operand type = which library, hi/lo split along K)
PRECISIONS = {"bf16": ("bf16", 0), "bf16x3": ("bf16", 1), "f16": ("f16", 0)}


def gemm_nt(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None,
            precision: str = "bf16") -> torch.Tensor:
    """out[M,N] = a[M,K] @ b[N,K]^T (+ bias) on the tcgen05 path; a and b may be transposed views.
    precision "bf16" rounds the operands to bf16; "bf16x3" uses the hi/lo split (fp32-class products)."""
    op, prec = PRECISIONS[precision]
    dev = require_cuda(a, b, bias)
    a, b, bias = _strided_2d(a), _strided_2d(b), f32c(bias)
    M, K = a.shape
    N, K2 = b.shape
    if K != K2:
        raise ValueError(f"gemm_nt: inner dimensions differ ({K} vs {K2})")
    out = torch.empty(M, N, dtype=torch.float32, device=dev)
    if M == 0 or N == 0:
        return out
    if K == 0:
        return out.zero_() if bias is None else out.copy_(bias.expand(M, N))
    l = lib(op)
    ws_bytes = l.aid_gemm_nt_workspace_bytes(M, N, K, prec)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(l.aid_gemm_nt(ptr(a), a.stride(0), a.stride(1), ptr(b), b.stride(0), b.stride(1), ptr(bias), ptr(out),
                            M, N, K, prec, ptr(ws), ws_bytes, stream_ptr(dev)), "aid_gemm_nt", l)
    return out


def wgrad(dy: torch.Tensor, x: torch.Tensor, operand: Optional[str] = None) -> torch.Tensor:
    """dy[rows,N]^T @ x[rows,K] -> [N,K] through the MN-major weight-gradient kernel (aid_wgrad)."""
    dev = require_cuda(dy, x)
    dy, x = f32c(dy), f32c(x)
    rows, N = dy.shape
    K = x.shape[1]
    if x.shape[0] != rows:
        raise ValueError("wgrad: row counts differ")
    out = torch.empty(N, K, dtype=torch.float32, device=dev)
    l = lib(operand)
    nbytes = l.aid_wgrad_workspace_bytes(rows, N, K)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(l.aid_wgrad(ptr(dy), ptr(x), ptr(out), rows, N, K, ptr(ws), nbytes, stream_ptr(dev)), "aid_wgrad", l)
    return out


def colsum(x: torch.Tensor) -> torch.Tensor:
    """x.sum(0) of a 2-D fp32 tensor with unit column stride (aid_colsum, deterministic)."""
    dev = require_cuda(x)
    if x.dtype != torch.float32:
        x = x.float()
    if x.dim() != 2:
        raise ValueError("colsum expects a 2-D tensor")
    if x.stride(1) != 1 or x.stride(0) < x.shape[1]:
        x = x.contiguous()
    M, N = x.shape
    out = torch.empty(N, dtype=torch.float32, device=dev)
    if M == 0 or N == 0:
        return out.zero_()
    l = lib()
    ws_bytes = l.aid_colsum_workspace_bytes(M, N)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(l.aid_colsum(ptr(x), x.stride(0), M, N, ptr(out), ptr(ws), ws_bytes, stream_ptr(dev)), "aid_colsum")
    return out


def gelu_double_backward(gg: torch.Tensor, g: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """gg * g * phi(x) * (2 - x^2), element-wise (aid_gelu_double_backward)."""
    dev = require_cuda(gg, g, x)
    gg, g, x = f32c(gg), f32c(g), f32c(x)
    if not (gg.shape == g.shape == x.shape):
        raise ValueError("gelu_double_backward: shapes differ")
    out = torch.empty_like(x)
    check(lib().aid_gelu_double_backward(ptr(gg), ptr(g), ptr(x), ptr(out), x.numel(), stream_ptr(dev)),
          "aid_gelu_double_backward")
    return out


def time_importance_update(t: torch.Tensor, loss: torch.Tensor, weights: torch.Tensor,
                           want_bins: bool = False) -> Optional[torch.Tensor]:
    """In-place sequential EMA of `weights` [n_bins] over the batch (aid_time_importance_update)."""
    dev = require_cuda(t, loss, weights)
    t, loss = f32c(t.detach().reshape(-1)), f32c(loss.detach().reshape(-1))
    if weights.dtype != torch.float32 or not weights.is_contiguous():
        raise ValueError("weights must be a contiguous float32 tensor")
    bins = torch.empty(t.numel(), dtype=torch.int64, device=dev) if want_bins else None
    check(lib().aid_time_importance_update(ptr(t), ptr(loss), t.numel(), ptr(weights), weights.numel(), ptr(bins),
                                           stream_ptr(dev)), "aid_time_importance_update")
    return bins


def lambda_returns(rewards: torch.Tensor, next_values: torch.Tensor, dones: torch.Tensor, discount_factor: float,
                   lambda_: float = 0.95, n_steps: int = 5, exclude_immediate_rewards: bool = False) -> torch.Tensor:
    """lambda-returns over the batch axis (aid_lambda_returns); all inputs [B] on the device."""
    dev = require_cuda(rewards, next_values, dones)
    r, nv = f32c(rewards.detach().reshape(-1)), f32c(next_values.detach().reshape(-1))
    d = (dones.reshape(-1) != 0).to(torch.uint8).contiguous()
    if not (r.numel() == nv.numel() == d.numel()):
        raise ValueError("rewards, next_values and dones must have the same length")
    out = torch.empty_like(r)
    check(lib().aid_lambda_returns(ptr(r), ptr(nv), ptr(d), r.numel(), float(discount_factor), float(lambda_),
                                   int(n_steps), int(bool(exclude_immediate_rewards)), ptr(out), stream_ptr(dev)),
          "aid_lambda_returns")
    return out.reshape(rewards.shape)


def profile_select(epi: int, k: int = 0, n: int = 0, slot: Optional[int] = None) -> None:
    """Time every tcgen05 GEMM launch of the class (epilogue kind, K, N) with CUDA events on the launching
    stream; `slot` (0..3) keeps several classes at once, None = slot 0 only (others cleared)."""
    if slot is None:
        check(lib().aid_profile_select(epi, k, n), "aid_profile_select")
    else:
        check(lib().aid_profile_select_slot(slot, epi, k, n), "aid_profile_select_slot")


def profile_collect(slot: int = 0):
    """(total device ms, launches) of the selected GEMM class since the last collect."""
    ms, n = ctypes.c_double(0.0), c_int64(0)
    check(lib().aid_profile_collect_slot(slot, ctypes.byref(ms), ctypes.byref(n)), "aid_profile_collect_slot")
    return ms.value, int(n.value)
