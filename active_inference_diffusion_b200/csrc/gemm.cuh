// Persistent, warp-specialised tcgen05 GEMM with fused epilogues (sm_100a).
//
//   D[128 x 128] (TMEM, fp32) = A[128 x K] (16-bit operands, K-major, no-swizzle tiles) * B[128 x K]^T
//
// Data layout in HBM (all produced by this library, never by the caller):
//   * "packed" bf16 operand  : [row_tile][k_block] tiles of R rows x 64 cols (R = 128 for
//     activations, 128 or 256 for weights), each tile R*128 bytes contiguous and already the
//     shared-memory image of the K-major NO-swizzle UMMA layout: [16-byte K chunk (8)][row (R)][8 bf16].
//     One cp.async.bulk (TMA engine) moves a tile and the UMMA descriptor reads it directly.  With
//     one thread per row, a warp writing one K chunk stores 512 contiguous bytes: epilogues and pack
//     kernels write packed operands with fully coalesced 16-byte stores (the 128-byte-swizzled
//     image needs 32 different 128-byte lines per warp store and was L1TEX-wavefront bound), and
//     the tensor pipe reads it at the same rate (scripts/micro/mma_rate.cu: 128 cyc per N=256 MMA).
//   * "tiled" fp32 activation: [row_tile][col/4][128 rows] float4, so that the epilogue's
//     one-thread-per-row accesses are fully coalesced.
//   * LayerNorm partials     : [row_tile][n_tile][128 rows] float2 (mean, M2) per 128 columns.
//
// Work decomposition: unit = (row_tile, group of G consecutive 128-col n-tiles).  Each CTA
// (one per SM) takes a contiguous range of units so the A row tile stays resident in shared
// memory while the weight tiles stream through a ring of 16 KiB stages.
//
// Roles (384 threads): warp0 lane0 = bulk-copy producer, warp1 lane0 = MMA issuer,
// warp2 = TMEM allocator, warps 4..11 = two epilogue groups (tile sequence number parity).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "ptx.cuh"
#include "philox.cuh"

namespace aid {

constexpr int TILE_M = 128;
constexpr int TILE_N = 128;
constexpr int TILE_K = 64;
constexpr int TILE_BYTES = TILE_M * TILE_K * 2;  // 16384
constexpr int TILE_ELEMS = TILE_M * TILE_K;      // 8192
constexpr int MAX_RES_KB = 8;                    // A resident up to K = 512
constexpr int MAX_RING = 14;                     // ring stages (16 KiB each)
constexpr int GEMM_THREADS = 384;
constexpr int SMEM_LIMIT = 232448;               // 227 KiB opt-in max per CTA
constexpr int SMEM_CTRL = 2048;                  // barriers + tmem pointer + bias stages

enum Act : int { ACT_NONE = 0, ACT_SILU = 1, ACT_RELU = 2, ACT_GELU = 3 };

enum EpiKind : int {
  EPI_PACK = 0,   // y = act(acc + bias)                -> packed bf16 (next GEMM's A)
  EPI_F32 = 1,    // y = acc + bias (+ resid)           -> tiled fp32 and/or row-major fp32 (+ LN partials)
  EPI_MODLN = 2,  // xn = LN(h)*(1+scale)+shift         -> packed bf16   (acc = [scale | shift])
  EPI_SCORE = 3,  // clamp/scale score, optional reverse-diffusion step -> z (fp32 row-major + packed)
  EPI_LNACT = 4,  // y = act(LayerNorm(acc + bias) * gamma + beta) (+ resid) -> packed bf16; the whole
                  // 512-column row is normalised from TMEM (CTA-pair kernel only, gemm2.cuh)
  EPI_DACT = 5,   // training (train.inc): activation forward with the pre-activation saved, and the
                  // products with act'(pre) / act''(pre) of the three backward passes
};
// EPI_DACT modes (ACT = GELU or SiLU; pre = saved pre-activation, tiled fp32)
enum DactMode : int {
  DACT_FWD = 1,   // y = acc + bias: pre <- y (saved), out_packed <- act(y)
  DACT_VJP = 2,   // aux_out <- acc (cotangent before the activation), out_packed <- acc * act'(pre)
  DACT_HAT = 3,   // out_packed <- acc * act'(pre); aux_out <- acc * aux_in * act''(pre) * aux_scale
  DACT_BWD = 4,   // out_packed <- acc * act'(pre) (+ aux_in on the rows of the first src_rt row tiles)
};

// Implicit im2col of a 3x3 / stride-1 / padding-1 convolution (encoder.inc): the A operand is the
// "halo" activation [Cin/8 planes][guard + virtual rows + guard][8 bf16] and the K chunk (tap, 8
// channels) of 128 consecutive output rows is one contiguous 2 KiB run of it, so the producer
// fetches a k-block as eight 2 KiB bulk copies instead of one packed 16 KiB tile.  cin8 == 0: off.
struct ConvA {
  const uint8_t* hi;
  const uint8_t* lo;        // low halves (bf16x3 operand split), regions [hi | hi | lo] along K
  long long plane_bytes;    // bytes per 8-channel plane
  int guard;                // zero guard rows in front of every plane (>= 128: also the zero K padding)
  int cin8;                 // Cin / 8
  int vw;                   // virtual row width (Wo + 2)
  int kreg8;                // 16-byte chunks per region when split, else 0
};

struct GemmArgs {
  ConvA conv;
  const uint8_t* A;  // packed
  const uint8_t* B;  // packed
  int row_tiles;
  int kb;       // K / 64 processed per unit (per split)
  int n_tiles;  // N_pad / 128
  int kb_stride;  // k-blocks per row tile in memory (= kb * splits)
  int splits;     // split-K: unit = (split, row tile, column group); split s reduces k-blocks
                  // [s*kb, (s+1)*kb) and stores to row block s of the (row-major) output
  int* err;
  int reverse;  // walk the units in descending order (see launch_gemm: L2 reuse between kernels)
  int debug;    // developer knobs (AID_DEBUG env): 1 = skip epilogue math, 2 = skip B loads
};

struct EpiArgs {
  int debug;           // developer knobs (AID_DEBUG): 32 = no epilogue stores, 64 = no activation,
                       // 128 = no epilogue global loads (residual / LayerNorm rows)
  const float* bias;   // [n_tiles*128] (padded), may be null
  int act;
  int n_valid;         // number of real output columns
  int rows_valid;      // number of real rows (batch)
  // EPI_PACK / EPI_MODLN / EPI_SCORE(packed z)
  __nv_bfloat16* out_packed;
  int out_kb;          // k-blocks per row tile of the packed output
  // EPI_F32
  float4* out_tiled;         // may be null
  const float4* resid_tiled; // may be null
  int ld4;                   // float4 columns per row of tiled buffers (N_pad/4)
  float2* stats_out;         // may be null; [rt][n_tiles][128]
  float* out_rm;             // optional row-major output [rows_valid, ld_rm]
  int ld_rm;
  int split_rt;              // row tiles per split (= row_tiles); out_rm row = rt*128 + r over all splits
  // EPI_LNACT (resid_tiled / ld4 above: residual added AFTER the activation)
  const float* ln_gamma;     // [n_valid]
  const float* ln_beta;      // [n_valid]
  // EPI_MODLN
  const float4* h_tiled;     // [rt][h_ld4][128]
  int h_ld4;
  const float2* stats_in;    // [rt][stats_nt][128]
  int stats_nt;
  int h_dim;                 // real hidden width (LayerNorm length)
  // EPI_SCORE
  const float* out_mult;     // device scalar (output_multiplier)
  const float* tw_rows;      // per-row time weight or null
  float tw_scalar;           // used when tw_rows == null
  int do_step;               // 0: write score; 1: reverse-diffusion update
  const float* z_in;         // [rows_valid, n_valid] row-major
  const float* eps;          // [rows_valid, n_valid] or null (deterministic / t == 0 / Philox mode)
  const PhiloxState* philox; // device {seed, call offset}: draw eps in the epilogue (no noise tensor), or null
  unsigned int philox_draw;  // draw index of this step (philox.cuh)
  long long row_offset;      // global index of row 0 (sharded batches draw the unsharded stream)
  float c_s1, c_ra, c_c1, c_c2, c_sigma;
  float* z_out;              // row-major fp32 (score or new z)
  float* r_out;              // optional: the pre-clamp output (training keeps it for the clamp mask)
  // EPI_DACT
  int dact_mode;
  float4* pre_tiled;         // [src_rt][pre_ld4][128]: written in DACT_FWD, read otherwise
  int pre_ld4;
  int src_rt;                // row tiles of the saved tensors; stacked cotangent rows use rt % src_rt
  const __nv_bfloat16* aux_in;
  __nv_bfloat16* aux_out;
  const float* aux_scale;    // device scalar (DACT_HAT) or null
};

// The source address of K chunk g for row tile 0 depends only on g; it is tabulated in shared
// memory once per CTA (the first version recomputed it with integer divisions in the single
// producer thread, ~3000 cycles of dependent instructions per k-block: the implicit GEMMs ran 2.5x
// slower than GEMMs on packed tiles).  CONV_TAB_BYTES of dynamic shared memory follow the ring.
constexpr int CONV_TAB_BYTES = 16384;
__device__ __forceinline__ void conv_fill_table(const ConvA& c, unsigned long long* tab, int chunks) {
  for (int i = threadIdx.x; i < chunks; i += blockDim.x) {
    int g = i;
    bool want_lo = false;
    if (c.kreg8) {
      const int region = g / c.kreg8;
      g -= region * c.kreg8;
      want_lo = region == 2;
    }
    // K padding: the zero guard rows of plane 0, flagged in bit 0 (does not move with the row tile)
    unsigned long long entry = reinterpret_cast<unsigned long long>(c.hi) | 1ull;
    if (g < 9 * c.cin8) {
      const int tap = g / c.cin8, c8 = g - tap * c.cin8;
      const long long row = (long long)c.guard + (tap / 3 - 1) * c.vw + (tap % 3 - 1);
      entry = reinterpret_cast<unsigned long long>((want_lo ? c.lo : c.hi) + (long long)c8 * c.plane_bytes + row * 16);
    }
    tab[i] = entry;
  }
}
// one 2 KiB chunk (j = 0..7) of k-block kb: the warp-wide producers give one chunk to each of 8 lanes
__device__ __forceinline__ void conv_load_chunk(const unsigned long long* tab, uint32_t dst, int rt, int kb, int j,
                                                uint32_t bar, uint64_t policy) {
  const unsigned long long t = tab[kb * 8 + j];
  const uint8_t* src = reinterpret_cast<const uint8_t*>(
      (t & 1ull) ? t - 1ull : t + (unsigned long long)rt * (TILE_M * 16));
  bulk_g2s_hint(dst + j * (TILE_M * 16), src, TILE_M * 16, bar, policy);
}
__device__ __forceinline__ void conv_load_a(const ConvA& c, const unsigned long long* tab, uint32_t dst, int rt,
                                            int kb, uint32_t bar, uint64_t policy) {
  const unsigned long long row_off = (unsigned long long)rt * (TILE_M * 16);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const unsigned long long t = tab[kb * 8 + j];
    const uint8_t* src = reinterpret_cast<const uint8_t*>((t & 1ull) ? t - 1ull : t + row_off);
    bulk_g2s_hint(dst + j * (TILE_M * 16), src, TILE_M * 16, bar, policy);
  }
}

// element offset of (row r, col c) inside an R-row x 64-col packed tile
__device__ __forceinline__ int packed_off(int r, int c, int R = TILE_M) {
  return (c >> 3) * (R * 8) + r * 8 + (c & 7);
}

// bare MUFU operations (exp2f / __frcp_rn expand to multi-instruction sequences with range fix-ups)
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float act_silu(float x) { return x / (1.0f + __expf(-x)); }
// GELU.  The reference uses the exact erf form (nn.GELU(), models/score_networks.py:199).  The
// epilogue budget is ~16 issue slots per output element (K=512), erff alone costs ~40, so the hot
// path evaluates x*Phi(x) with Phi(x) ~= 0.5(1+tanh(sqrt(2/pi)(x+0.044715x^3))) on the MUFU tanh
// unit: |gelu_tanh - gelu_erf| <= 5e-4 absolute (worst near |x|~2.5), i.e. below the bf16 rounding
// (2^-9 relative) applied to this output right after.  AID_EXACT_GELU restores erff.
// fp16-operand builds (-DAID_F16, the rel-1e-3 mode) use erf in the Abramowitz-Stegun 7.1.26 form
// (|erf error| <= 1.5e-7): with q = (a1 t + ... + a5 t^5) exp(-y^2), t = 1/(1 + p y), y = |x|/sqrt2,
// 1 + erf(x/sqrt2) is q for x < 0 and 2 - q for x >= 0 -- no cancellation in the negative tail; one
// MUFU reciprocal and one MUFU exponential per element.
__device__ __forceinline__ float act_gelu_as(float x) {
  const float y = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, y, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float q = p * t * __expf(-y * y);
  return 0.5f * x * (x < 0.f ? q : 2.0f - q);
}
__device__ __forceinline__ float act_gelu(float x) {
#ifdef AID_EXACT_GELU
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
#elif defined(AID_F16)
  return act_gelu_as(x);
#else
  const float u = x * fmaf(0.0356774081f, x * x, 0.7978845608f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
#endif
}
__device__ __forceinline__ float act_apply(float x, int act) {
  switch (act) {
    case ACT_SILU: return act_silu(x);
    case ACT_RELU: return fmaxf(x, 0.0f);
    case ACT_GELU: return act_gelu(x);
    default: return x;
  }
}
// `act` is warp-uniform: branch once per 32-column chunk, not once per element.
__device__ __forceinline__ void act_apply32(float (&y)[32], int act) {
  if (act == ACT_SILU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) y[j] = act_silu(y[j]);
  } else if (act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.f);
  } else if (act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) y[j] = act_gelu(y[j]);
  }
}

// Tensor-core operand element type, fixed per build of the library:
//   default     : bf16 (8-bit significand)  -> libaid_sm100.so      ("bf16" precision)
//   -DAID_F16   : IEEE fp16 (11-bit significand, the TF32 significand) -> libaid_sm100_f16.so ("f16")
// Both are kind::f16 UMMAs at the same tensor-pipe rate with fp32 accumulation; fp16 operands give
// TF32-class products (rounding 2^-11 instead of 2^-8) and are what meets the rel-1e-3 contract.
// fp16 has a 5-bit exponent: conversions saturate to +-65504 instead of producing inf, activations
// are LayerNorm-scaled, and the training path scales its cotangent streams (train.inc).
// pack_op16x2(a, b): a in the low half, b in the high half, round to nearest even.
__device__ __forceinline__ uint32_t pack_op16x2(float a, float b) {
#ifdef AID_F16
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
#else
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
#endif
}
// value of x after rounding to the operand type (hi part of the x3 split) and one stored element back to fp32
__device__ __forceinline__ float op16_round(float x) {
#ifdef AID_F16
  return __half2float(__float2half_rn(x));
#else
  return __bfloat162float(__float2bfloat16_rn(x));
#endif
}
__device__ __forceinline__ float op16_to_float(const __nv_bfloat16* p) {
#ifdef AID_F16
  return __half2float(*reinterpret_cast<const __half*>(p));
#else
  return __bfloat162float(*p);
#endif
}

// Write 32 consecutive columns [c0, c0+32) of row r (c0 % 32 == 0) into a packed tile row.
__device__ __forceinline__ void store_packed32(__nv_bfloat16* tile_base, int r, int c0_in_tile,
                                               const float (&y)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 v;
    v.x = pack_op16x2(y[q * 8 + 0], y[q * 8 + 1]);
    v.y = pack_op16x2(y[q * 8 + 2], y[q * 8 + 3]);
    v.z = pack_op16x2(y[q * 8 + 4], y[q * 8 + 5]);
    v.w = pack_op16x2(y[q * 8 + 6], y[q * 8 + 7]);
    int chunk = (c0_in_tile >> 3) + q;
    *reinterpret_cast<uint4*>(tile_base + chunk * (TILE_M * 8) + r * 8) = v;
  }
}

// Chan/Welford merge of (n_b, mean_b, M2_b) into (n, mean, M2).
__device__ __forceinline__ void stats_merge(float& n, float& mean, float& m2, float nb, float mb,
                                            float m2b) {
  if (nb <= 0.f) return;
  float nt = n + nb;
  float d = mb - mean;
  mean += d * (nb / nt);
  m2 += m2b + d * d * (n * nb / nt);
  n = nt;
}

#include "epilogue.cuh"

// ------------------------------------------------------------------------------------------
// Shared-memory control block
struct alignas(8) GemmCtrl {
  uint64_t ring_full[MAX_RING];
  uint64_t ring_empty[MAX_RING];
  uint64_t a_full[MAX_RES_KB];
  uint64_t a_empty[MAX_RES_KB];
  uint64_t acc_full[4];
  uint64_t acc_empty[4];
  uint32_t tmem_base;
  uint32_t pad_[(1024 - (2 * MAX_RING + 2 * MAX_RES_KB + 8) * 8 - 4) / 4];
  float bias_stage[2][TILE_N];   // per epilogue group; 1024-byte offset
  uint8_t pad2_[SMEM_CTRL - 1024 - 2 * TILE_N * 4];
};
static_assert(sizeof(GemmCtrl) == SMEM_CTRL, "control block must be exactly SMEM_CTRL bytes");

// NW  : 128-col n-tiles covered by ONE tcgen05.mma (1 -> N=128, 2 -> N=256).  N=256 is the
//       instruction shape that reaches the tensor-pipe rate on one CTA (measured on B200,
//       scripts/micro/mma_rate.cu: N=128 86 cyc/MMA vs 64 ideal; N=256 128 cyc = ideal; and
//       switching the destination accumulator between consecutive MMAs costs ~250 cyc).
// G   : MMA units accumulated concurrently (they share each streamed A k-block).
// RES : A row tile resident in shared memory (kb <= MAX_RES_KB) vs streamed through the ring.
// A "unit" is TU = NW*G consecutive n-tiles of one row tile; TMEM holds 4/TU units in flight.
// ACT : activation of EPI_PACK / EPI_F32 resolved at compile time (the others take ACT_NONE).
template <int EPI, int NW, int G, bool RES, int ACT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const GemmArgs ga, const EpiArgs ea, const int ring_stages) {
  constexpr int TU = NW * G;
  constexpr int SLOT_BYTES = NW * TILE_BYTES;
  static_assert(TU <= 4, "unit does not fit TMEM");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  GemmCtrl* ctrl = reinterpret_cast<GemmCtrl*>(smem);
  const uint32_t a_smem = base + SMEM_CTRL;                                  // RES: kb tiles
  const uint32_t ring_smem = a_smem + (RES ? ga.kb * TILE_BYTES : 0);        // ring_stages slots

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int groups = ga.n_tiles / TU;  // units per row tile
  const int num_units = ga.splits * ga.row_tiles * groups;
  // RES: a contiguous range of units per CTA, so the resident A row tile serves all its column
  // groups.  Streamed A: units are dealt round-robin, so the column groups of one row tile run at
  // the same time on neighbouring CTAs and the second read of the A tile hits L2 (with contiguous
  // ranges ncu showed mlp.2 pulling its 268 MB A operand from DRAM twice).
  const int u_lo = (int)((long long)blockIdx.x * num_units / gridDim.x);
  const int u_hi = (int)((long long)(blockIdx.x + 1) * num_units / gridDim.x);
  const int u_begin = 0;
  const int u_end = RES ? u_hi - u_lo
                        : (num_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto unit_at = [&](int i) {   // i-th unit of this CTA -> global unit index (before reversal)
    const int u = RES ? u_lo + i : (int)blockIdx.x + i * (int)gridDim.x;
    return ga.reverse ? num_units - 1 - u : u;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_RING; ++i) {
      mbar_init(smem_u32(&ctrl->ring_full[i]), 1);
      mbar_init(smem_u32(&ctrl->ring_empty[i]), 1);
    }
    for (int i = 0; i < MAX_RES_KB; ++i) {
      mbar_init(smem_u32(&ctrl->a_full[i]), 1);
      mbar_init(smem_u32(&ctrl->a_empty[i]), 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(smem_u32(&ctrl->acc_full[i]), 1);
      mbar_init(smem_u32(&ctrl->acc_empty[i]), 4);  // one arrive per epilogue warp of a group
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&ctrl->tmem_base), 512);
    tmem_relinquish();
  }
  // Programmatic dependent launch: the prologue above (barrier init, TMEM allocation) overlaps the
  // tail of the previous kernel of the chain; nothing below touches global memory before the wait.
  unsigned long long* conv_tab =
      reinterpret_cast<unsigned long long*>(smem + SMEM_CTRL + (RES ? ga.kb * TILE_BYTES : 0) + ring_stages * SLOT_BYTES);
  if (ga.conv.cin8) conv_fill_table(ga.conv, conv_tab, ga.kb * 8);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;
  pdl_wait();

  if (warp == 0) {
    // ===================== producer =====================
    if (!RES && ga.conv.cin8) {
      // Implicit-im2col mode, warp-wide: lane 0 owns the barriers, lanes 0-7 each issue one 2 KiB
      // chunk of the A k-block, lane 8 the weight tile.  (One thread issuing all nine copies of a
      // k-block was the bottleneck: 506 vs 334 us for conv4 against the same GEMM on packed tiles.)
      const uint64_t keep = policy_evict_last();
      int stage = 0;
      uint32_t phase = 0;
      for (int u = u_begin; u < u_end; ++u) {
        const int ue = unit_at(u);
        const int rt = ue / groups, ng = ue % groups;
        for (int kb = 0; kb < ga.kb; ++kb) {
          uint32_t fb = smem_u32(&ctrl->ring_full[stage]);
          if (lane == 0) {
            mbar_wait(smem_u32(&ctrl->ring_empty[stage]), phase ^ 1, ga.err, 2);
            mbar_arrive_expect_tx(fb, TILE_BYTES);
          }
          __syncwarp();
          if (lane < 8) conv_load_chunk(conv_tab, ring_smem + stage * SLOT_BYTES, rt, kb, lane, fb, keep);
          if (++stage == ring_stages) { stage = 0; phase ^= 1; }
#pragma unroll
          for (int g = 0; g < G; ++g) {
            fb = smem_u32(&ctrl->ring_full[stage]);
            if (lane == 8) {
              mbar_wait(smem_u32(&ctrl->ring_empty[stage]), phase ^ 1, ga.err, 3);
              mbar_arrive_expect_tx(fb, SLOT_BYTES);
              bulk_g2s_hint(ring_smem + stage * SLOT_BYTES,
                            ga.B + ((size_t)(ng * G + g) * ga.kb_stride + kb) * SLOT_BYTES, SLOT_BYTES, fb, keep);
            }
            if (++stage == ring_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (lane == 0) {
      // L2 policy: weight tiles are read by every CTA and a streamed A row tile is read again by
      // the next unit of the same CTA -> keep (evict_last); a resident A tile is read once ->
      // evict_first (ncu showed mlp.2 re-reading its 268 MB A operand from DRAM on the second pass).
      const uint64_t keep = policy_evict_last(), once = policy_evict_first();
      int stage = 0;
      uint32_t phase = 0;
      uint32_t a_par = 0;
      int prev_rt = -1;
      for (int u = u_begin; u < u_end; ++u) {
        const int ue = unit_at(u);
        const int rt = ue / groups, ng = ue % groups;   // rt counts (split, row tile) pairs
        const size_t a_row = (size_t)(rt % ga.row_tiles) * ga.kb_stride + (size_t)(rt / ga.row_tiles) * ga.kb;
        const size_t kb_split = (size_t)(rt / ga.row_tiles) * ga.kb;
        const bool new_rt = RES && (rt != prev_rt);
        for (int kb = 0; kb < ga.kb; ++kb) {
          if (RES) {
            if (new_rt) {
              const uint32_t fb = smem_u32(&ctrl->a_full[kb]);
              mbar_wait(smem_u32(&ctrl->a_empty[kb]), a_par ^ 1, ga.err, 1);
              mbar_arrive_expect_tx(fb, TILE_BYTES);
              bulk_g2s_hint(a_smem + kb * TILE_BYTES, ga.A + (a_row + kb) * TILE_BYTES, TILE_BYTES, fb, once);
            }
          } else {
            const uint32_t fb = smem_u32(&ctrl->ring_full[stage]);
            mbar_wait(smem_u32(&ctrl->ring_empty[stage]), phase ^ 1, ga.err, 2);
            mbar_arrive_expect_tx(fb, TILE_BYTES);
            if (ga.conv.cin8)
              conv_load_a(ga.conv, conv_tab, ring_smem + stage * SLOT_BYTES, rt, kb, fb, keep);
            else
              bulk_g2s_hint(ring_smem + stage * SLOT_BYTES, ga.A + (a_row + kb) * TILE_BYTES, TILE_BYTES, fb,
                            (ga.debug & 8) ? once : keep);
            if (++stage == ring_stages) { stage = 0; phase ^= 1; }
          }
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const uint32_t fb = smem_u32(&ctrl->ring_full[stage]);
            mbar_wait(smem_u32(&ctrl->ring_empty[stage]), phase ^ 1, ga.err, 3);
            if (ga.debug & 2) {
              mbar_arrive(fb);
            } else {
              mbar_arrive_expect_tx(fb, SLOT_BYTES);   // weight tiles are packed NW*128 rows tall
              bulk_g2s_hint(ring_smem + stage * SLOT_BYTES,
                            ga.B + ((size_t)(ng * G + g) * ga.kb_stride + kb_split + kb) * SLOT_BYTES,
                            SLOT_BYTES, fb, keep);
            }
            if (++stage == ring_stages) { stage = 0; phase ^= 1; }
          }
        }
        if (new_rt) { a_par ^= 1; prev_rt = rt; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop (uniform control flow, descriptors in uniform registers); one
    // elected lane issues the tcgen05 instructions.  With the loop under `if (lane == 0)` every MMA
    // paid ELECT + 5 R2UR moves and the single issue thread, not the tensor pipe, set the pace
    // (MMA-only kernels at 58 % of the pipe rate).
    {
      constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, TILE_N * NW);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t a_par = 0;
      int prev_rt = -1;
      int q = 0;  // 128-col tile sequence number (multiple of TU at unit start)
      for (int u = u_begin; u < u_end; ++u) {
        const int rt = unit_at(u) / groups;
        const bool new_rt = RES && (rt != prev_rt);
        const bool last_of_rt =
            RES && (u + 1 == u_end || unit_at(u + 1) / groups != rt);
        for (int kb = 0; kb < ga.kb; ++kb) {
          uint32_t a_tile;
          int a_stage = -1;
          if (RES) {
            if (new_rt) mbar_wait(smem_u32(&ctrl->a_full[kb]), a_par, ga.err, 5);
            a_tile = a_smem + kb * TILE_BYTES;
          } else {
            mbar_wait(smem_u32(&ctrl->ring_full[stage]), phase, ga.err, 6);
            a_tile = ring_smem + stage * SLOT_BYTES;
            a_stage = stage;
            if (++stage == ring_stages) { stage = 0; phase ^= 1; }
          }
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if (kb == 0) {  // this unit's accumulator slots must have been drained by the epilogue
#pragma unroll
              for (int j = 0; j < NW; ++j) {
                const int sl = q + g * NW + j;
                mbar_wait(smem_u32(&ctrl->acc_empty[sl & 3]), ((sl >> 2) & 1) ^ 1, ga.err, 4);
              }
            }
            mbar_wait(smem_u32(&ctrl->ring_full[stage]), phase, ga.err, 7);
            tc_fence_after();
            const uint32_t b_tile = ring_smem + stage * SLOT_BYTES;
            const uint32_t d = tmem_base + (uint32_t)(((q + g * NW) & 3) * TILE_N);
            const uint64_t ad = umma_desc_kmajor(a_tile, TILE_M * 16);
            const uint64_t bd = umma_desc_kmajor(b_tile, NW * TILE_M * 16);
            if (elect_one()) {
              if (!(ga.debug & 1024))   // 1024: no MMAs (accumulators keep their content): epilogue-only timing
#pragma unroll
              for (int k = 0; k < TILE_K / 16; ++k) {
                // advancing the 14-bit start-address field by a constant stays inside the field
                umma_bf16(d, ad + (uint64_t)(k * (2 * TILE_M * 16) >> 4),
                          bd + (uint64_t)(k * (2 * NW * TILE_M * 16) >> 4), idesc, (kb | k) ? 1u : 0u);
              }
              umma_commit(smem_u32(&ctrl->ring_empty[stage]));
            }
            __syncwarp();
            if (++stage == ring_stages) { stage = 0; phase ^= 1; }
          }
          if (elect_one()) {
            if (!RES) umma_commit(smem_u32(&ctrl->ring_empty[a_stage]));
            if (last_of_rt) umma_commit(smem_u32(&ctrl->a_empty[kb]));
          }
          __syncwarp();
        }
        if (elect_one()) {
#pragma unroll
          for (int t = 0; t < TU; ++t) umma_commit(smem_u32(&ctrl->acc_full[(q + t) & 3]));
        }
        __syncwarp();
        q += TU;
        if (new_rt) { a_par ^= 1; prev_rt = rt; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (two groups of 4 warps) =====================
    // Group eg owns the tiles with sequence number q = eg, eg+2, ... of this CTA (q counts 128-col
    // tiles in MMA order; accumulator slot q & 3, use count q >> 2).
    const int eg = (warp - 4) >> 2;
    const int lq = warp & 3;            // TMEM lane quadrant this warp may access
    const int r = lq * 32 + lane;
    const int n_local = (u_end - u_begin) * TU;
    float* sb = ctrl->bias_stage[eg];
    auto coords = [&](int q, int& rt, int& nt) {   // rt counts (split, row tile) pairs
      const int ue = unit_at(u_begin + q / TU);
      rt = ue / groups;
      nt = (ue % groups) * TU + q % TU;
    };
    EpiState<EPI> st;
    int rt = 0, nt = 0;
    if (eg < n_local && !(ga.debug & 1)) {
      coords(eg, rt, nt);
      epi_first<EPI>(ea, rt, nt, r, st);
    }
#pragma unroll 1
    for (int q = eg; q < n_local; q += 2) {
      coords(q, rt, nt);
      const bool has_next = q + 2 < n_local;
      int rt2 = rt, nt2 = nt;
      if (has_next) coords(q + 2, rt2, nt2);
      const int buf = q & 3, use = q >> 2;
      if (!(ga.debug & 1)) epi_stage_bias<EPI>(ea, sb, 1 + eg, r, st, has_next, nt2);
      mbar_wait(smem_u32(&ctrl->acc_full[buf]), use & 1, ga.err, 8);
      tc_fence_after();
      const uint32_t tm = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(buf * TILE_N);
      const AccRelease rel{smem_u32(&ctrl->acc_empty[buf]), 0};
      if (!(ga.debug & 1)) epi_finish<EPI, ACT>(ea, tm, rt, nt, ga.n_tiles, r, sb, st, has_next, rt2, nt2, rel);
      else acc_release(rel);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace aid
