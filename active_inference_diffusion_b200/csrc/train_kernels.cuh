// Device code of the native training path (train.inc): the weight-gradient GEMM on MN-major
// operands, the LayerNorm / adaLN row kernels of the four passes, and the small element-wise seeds.
// Specification: oracle/manual_score_grad.py (hand-derived first/second-order backward of the trunk
// of models/score_networks.py:151-171, checked against autograd in tests/test_manual_score_grad.py).
#pragma once
#include "gemm.cuh"

namespace aid {

// =================================================================================================
// Weight gradient  dW[n, k] = sum_b dY[b, n] * X[b, k]   (contraction over the batch rows)
//
// Both operands are the ROW-major packs the forward / input-gradient GEMMs already made
// ([row tile][64-feature block] tiles of [8-feature chunk][128 rows][8 elements]).  Read with MN =
// feature and K = row this is the MN-major no-swizzle canonical layout (LBO = 128 B: next 8 rows,
// SBO = 2 KiB: next 8-feature chunk; a K = 16 MMA advances the start address by 256 B; verified on
// B200 by scripts/micro/mn_major_probe.cu), so no transposed operand pack exists anywhere.
//   unit  = (split of the row tiles, 128 features of dY -> MMA M, up to 256 features of X -> MMA N)
//   stage = one 128-row tile: 32 KiB of dY (two adjacent feature blocks) + up to 64 KiB of X,
//           8 x (M128, N<=256, K16) MMAs; 2 stages of 96 KiB; accumulators double-buffered in TMEM.
// Output: fp32 partials [split][n][k] (row-major), reduced by k_wgrad_reduce.
constexpr int WG_STAGE_A = 2 * TILE_BYTES;
constexpr int WG_STAGE_B = 4 * TILE_BYTES;
constexpr int WG_STAGE = WG_STAGE_A + WG_STAGE_B;
constexpr int WG_STAGES = 2;
constexpr int WG_THREADS = 320;   // warp 0 producer, warp 1 MMA + TMEM alloc, warps 2-9 epilogue (2 groups)

struct WgradArgs {
  const uint8_t* dy;   // packed, kb_dy blocks per row tile (even)
  const uint8_t* x;    // packed, kb_x blocks per row tile
  int kb_dy, kb_x;
  int row_tiles;       // row tiles to reduce over (the first row_tiles of both packs)
  int splits;
  int n_blocks;        // 128-feature blocks of dY  (= ceil(N / 128))
  int k_groups;        // 256-feature groups of X   (= ceil(kb_x / 4))
  int N, K;            // logical output shape
  float* partial;      // [splits][N][K]
  int* err;
};

struct alignas(8) WgradCtrl {
  uint64_t full[WG_STAGES];
  uint64_t empty[WG_STAGES];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  uint32_t pad_[256 - 2 * (2 * WG_STAGES + 4) - 1];
};
static_assert(sizeof(WgradCtrl) == 1024, "wgrad control block");

__device__ __forceinline__ uint64_t umma_desc_mnmajor(uint32_t smem_addr) {
  // MN-major, no swizzle: LBO (bits 16-29) = 128 B (next 8 rows of K), SBO (bits 32-45) = 2048 B
  return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(128 >> 4) << 16) |
         (static_cast<uint64_t>(2048 >> 4) << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_kernel(const WgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  WgradCtrl* ctrl = reinterpret_cast<WgradCtrl*>(smem);
  const uint32_t ring = base + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int units = a.splits * a.n_blocks * a.k_groups;

  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_STAGES; ++i) {
      mbar_init(smem_u32(&ctrl->full[i]), 1);
      mbar_init(smem_u32(&ctrl->empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&ctrl->acc_full[i]), 1);
      mbar_init(smem_u32(&ctrl->acc_empty[i]), 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&ctrl->tmem_base), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;

  // unit -> (split, n block, k group); the row-tile range of a split
  auto decode = [&](int u, int& sp, int& nb, int& kg, int& rt0, int& rt1) {
    kg = u % a.k_groups;
    nb = (u / a.k_groups) % a.n_blocks;
    sp = u / (a.k_groups * a.n_blocks);
    rt0 = (int)((long long)sp * a.row_tiles / a.splits);
    rt1 = (int)((long long)(sp + 1) * a.row_tiles / a.splits);
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        int sp, nb, kg, rt0, rt1;
        decode(u, sp, nb, kg, rt0, rt1);
        const int nt = min(4, a.kb_x - kg * 4);            // 64-feature blocks of X in this group
        // dY packs with an odd number of feature blocks: the last 128-feature block has one tile; the
        // MMA then reads stale shared memory for features 64..127, which only reaches output rows >= N
        const int na = min(2, a.kb_dy - nb * 2);
        for (int rt = rt0; rt < rt1; ++rt) {
          mbar_wait(smem_u32(&ctrl->empty[stage]), phase ^ 1, a.err, 21);
          const uint32_t fb = smem_u32(&ctrl->full[stage]);
          mbar_arrive_expect_tx(fb, (na + nt) * TILE_BYTES);
          const uint32_t dst = ring + stage * WG_STAGE;
          bulk_g2s(dst, a.dy + ((size_t)rt * a.kb_dy + nb * 2) * TILE_BYTES, na * TILE_BYTES, fb);
          bulk_g2s(dst + WG_STAGE_A, a.x + ((size_t)rt * a.kb_x + kg * 4) * TILE_BYTES, nt * TILE_BYTES, fb);
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    int q = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++q) {
      int sp, nb, kg, rt0, rt1;
      decode(u, sp, nb, kg, rt0, rt1);
      const int nt = min(4, a.kb_x - kg * 4);
      const uint32_t idesc = umma_idesc_op16_mn(TILE_M, nt * TILE_K);
      mbar_wait(smem_u32(&ctrl->acc_empty[q & 1]), ((q >> 1) & 1) ^ 1, a.err, 22);
      const uint32_t d = tmem_base + (uint32_t)((q & 1) * 256);
      for (int rt = rt0; rt < rt1; ++rt) {
        mbar_wait(smem_u32(&ctrl->full[stage]), phase, a.err, 23);
        tc_fence_after();
        const uint32_t s_a = ring + stage * WG_STAGE;
        const uint64_t ad = umma_desc_mnmajor(s_a);
        const uint64_t bd = umma_desc_mnmajor(s_a + WG_STAGE_A);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < TILE_M / 16; ++k)
            umma_bf16(d, ad + (uint64_t)((k * 256) >> 4), bd + (uint64_t)((k * 256) >> 4), idesc,
                      (rt > rt0 || k) ? 1u : 0u);
          umma_commit(smem_u32(&ctrl->empty[stage]));
        }
        __syncwarp();
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(smem_u32(&ctrl->acc_full[q & 1]));
      __syncwarp();
    }
  } else {
    const int eg = (warp - 2) >> 2;     // epilogue group = accumulator parity
    const int lq = warp & 3;            // TMEM lane quadrant of this warp
    const int r = lq * 32 + lane;
    int q = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++q) {
      if ((q & 1) != eg) continue;
      int sp, nb, kg, rt0, rt1;
      decode(u, sp, nb, kg, rt0, rt1);
      const int nt = min(4, a.kb_x - kg * 4);
      mbar_wait(smem_u32(&ctrl->acc_full[q & 1]), (q >> 1) & 1, a.err, 24);
      tc_fence_after();
      const uint32_t tm = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)((q & 1) * 256);
      const int n = nb * TILE_M + r;
      float* orow = a.partial + ((size_t)sp * a.N + n) * a.K + kg * 256;
      const bool empty = rt1 <= rt0;    // more splits than row tiles: this split contributes zeros
      for (int c = 0; c < nt * 2; ++c) {
        uint32_t raw[32];
        tmem_ld32(tm + c * 32, raw);
        tmem_ld_wait();
        if (n < a.N) {
          const int k0 = kg * 256 + c * 32;
          if (k0 + 32 <= a.K && (a.K & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(orow + c * 32 + j * 4) =
                  empty ? make_float4(0.f, 0.f, 0.f, 0.f)
                        : make_float4(__uint_as_float(raw[j * 4]), __uint_as_float(raw[j * 4 + 1]),
                                      __uint_as_float(raw[j * 4 + 2]), __uint_as_float(raw[j * 4 + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (k0 + j < a.K) orow[c * 32 + j] = empty ? 0.f : __uint_as_float(raw[j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&ctrl->acc_empty[q & 1]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// out[i] (+)= inv_scale * sum_s partial[s][i];  inv_scale = scale_a * (*scale_dev) (scale_dev may be null)
__global__ void k_wgrad_reduce(const float* __restrict__ partial, int splits, size_t n, float* __restrict__ out,
                               float scale_a, const float* __restrict__ scale_dev, int accumulate) {
  const float sc = scale_a * (scale_dev ? __ldg(scale_dev) : 1.0f);
  const size_t n4 = n >> 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
      const float4 v = reinterpret_cast<const float4*>(partial + (size_t)s * n)[i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float4 o = make_float4(acc.x * sc, acc.y * sc, acc.z * sc, acc.w * sc);
    if (accumulate) {
      const float4 p = reinterpret_cast<float4*>(out)[i];
      o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
    }
    reinterpret_cast<float4*>(out)[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (size_t i = n4 * 4; i < n; ++i) {
      float acc = 0.f;
      for (int s = 0; s < splits; ++s) acc += partial[(size_t)s * n + i];
      out[i] = acc * sc + (accumulate ? out[i] : 0.f);
    }
  }
}

// =================================================================================================
// All weight packs of a step in ONE launch (the per-Linear launches were 134 tiny kernels, ~2 ms of a
// 33 ms step): blockIdx.y selects a job, the job table travels as a kernel parameter.
struct PackJob {
  const float* src;
  long long rs, cs;          // element (row, col) of the operand at src[row * rs + col * cs]
  int rows, cols;
  __nv_bfloat16* dst;
  int row_tiles, kb_total, nw;
};
constexpr int PACK_JOBS = 64;
struct PackJobs {
  PackJob job[PACK_JOBS];
};
__global__ void __launch_bounds__(256) k_pack_multi(const PackJobs jobs) {
  const PackJob j = jobs.job[blockIdx.y];
  const size_t total = (size_t)j.row_tiles * j.kb_total * 1024;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int rt, kb, r, ch;
    chunk_coords(idx, j.kb_total, rt, kb, r, ch);
    if (j.cs == 1) {            // row-major source: neighbouring lanes take neighbouring 32-byte pieces of one row
      ch = (int)(idx & 7);
      r = (int)((idx >> 3) & 127);
    }
    const int row = rt * TILE_M + r, c0 = kb * TILE_K + ch * 8;
    float v[8];
    if (j.cs == 1 && (j.rs & 3) == 0 && row < j.rows && c0 + 8 <= j.cols) {
      const float4* p = reinterpret_cast<const float4*>(j.src + (long long)row * j.rs + c0);
      const float4 a = __ldg(p), b = __ldg(p + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        v[i] = (row < j.rows && c0 + i < j.cols) ? __ldg(j.src + (long long)row * j.rs + (long long)(c0 + i) * j.cs) : 0.f;
    }
    store_chunk(j.dst, rt, kb, j.kb_total, r, ch, v, j.nw);
  }
}
struct BiasJob {
  const float* src;          // null: zeros
  int n;
  float* dst;
  int n_pad;
};
struct BiasJobs {
  BiasJob job[PACK_JOBS];
};
__global__ void __launch_bounds__(256) k_bias_multi(const BiasJobs jobs) {
  const BiasJob j = jobs.job[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < j.n_pad; i += gridDim.x * blockDim.x)
    j.dst[i] = (j.src && i < j.n) ? j.src[i] : 0.f;
}

// =================================================================================================
// Column sums of a packed operand (bias gradients): out[n] (+)= scale * sum_rows pack[row, n].
// grid = (feature blocks, row splits); block = 256 threads = 8 chunks x 32 row lanes.
__global__ void k_colsum_packed(const __nv_bfloat16* __restrict__ pk, int kb_total, int row_tiles, int N,
                                float* __restrict__ partial /* [gridDim.y][kb_total*64] */) {
  const int kb = blockIdx.x, ch = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rt0 = (int)((long long)blockIdx.y * row_tiles / gridDim.y);
  const int rt1 = (int)((long long)(blockIdx.y + 1) * row_tiles / gridDim.y);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int rt = rt0; rt < rt1; ++rt) {
    const __nv_bfloat16* tile = pk + ((size_t)rt * kb_total + kb) * TILE_ELEMS + ch * (TILE_M * 8);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 v = *reinterpret_cast<const uint4*>(tile + (i * 32 + lane) * 8);
      float f[8];
      unpack_op16x8(v, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = kb * 64 + ch * 8 + j;
      if (n < N) partial[(size_t)blockIdx.y * kb_total * 64 + n] = acc[j];
    }
  }
}
__global__ void k_colsum_finish(const float* __restrict__ partial, int parts, int stride, int N,
                                float* __restrict__ out, float scale_a, const float* __restrict__ scale_dev,
                                int accumulate) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float acc = 0.f;
  for (int p = 0; p < parts; ++p) acc += partial[(size_t)p * stride + n];
  acc *= scale_a * (scale_dev ? __ldg(scale_dev) : 1.0f);
  out[n] = acc + (accumulate ? out[n] : 0.f);
}

// =================================================================================================
// Row kernels.  block = RK_THREADS / CG rows of one 128-row tile x CG column groups (thread = (group, row);
// a warp = 32 consecutive rows of one group, so tiled fp32 accesses and packed 16-byte stores are coalesced).
// A thread walks the 8-column chunks ch = cg, cg + CG, ...  Row reductions go through shared memory.
// CG = 4: a block owns a whole row tile (grid = row tiles; fills the chip from ~150 row tiles on).
// CG = 16: a block owns 32 rows (grid = 4 x row tiles) -- for the few row tiles of a data-parallel shard
// (4,096 rows per rank = 32 row tiles: the CG = 4 launches ran on 32 of 148 SMs, 23-36 us each).
constexpr int RK_CG = 4;
constexpr int RK_CG_WIDE = 16;
constexpr int RK_THREADS = RK_CG * TILE_M;
__host__ __device__ constexpr int rk_rows(int cg) { return RK_THREADS / cg; }
static inline bool rk_wide(int row_tiles) { return row_tiles < 96; }

struct Tl {            // tiled fp32 tensor view [rt][ld4][128] float4 + float4-column offset
  const float4* p;
  int ld4;
  int off4;
};
__device__ __forceinline__ void tl_load8(const Tl& t, int rt, int ch, int r, float* v) {
  const float4 a = t.p[((size_t)rt * t.ld4 + t.off4 + ch * 2) * TILE_M + r];
  const float4 b = t.p[((size_t)rt * t.ld4 + t.off4 + ch * 2 + 1) * TILE_M + r];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void tl_store8(float4* p, int ld4, int off4, int rt, int ch, int r, const float* v) {
  p[((size_t)rt * ld4 + off4 + ch * 2) * TILE_M + r] = make_float4(v[0], v[1], v[2], v[3]);
  p[((size_t)rt * ld4 + off4 + ch * 2 + 1) * TILE_M + r] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void pk_store8(__nv_bfloat16* pk, int kb_total, int kb_off, int rt, int ch, int r,
                                          const float* v) {
  uint4 o;
  o.x = pack_op16x2(v[0], v[1]); o.y = pack_op16x2(v[2], v[3]);
  o.z = pack_op16x2(v[4], v[5]); o.w = pack_op16x2(v[6], v[7]);
  __nv_bfloat16* tile = pk + ((size_t)rt * kb_total + kb_off + (ch >> 3)) * TILE_ELEMS;
  *reinterpret_cast<uint4*>(tile + (ch & 7) * (TILE_M * 8) + r * 8) = o;
}
// sum of `v` over the RK_CG column groups of row r (all threads of the block call it)
template <int NV, int CG>
__device__ __forceinline__ void rk_reduce(float (&v)[NV], float* sm /* [NV][CG][rows of the block] */, int cg, int rl) {
  constexpr int ROWS = rk_rows(CG);
#pragma unroll
  for (int i = 0; i < NV; ++i) sm[(i * CG + cg) * ROWS + rl] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < CG; ++g) s += sm[(i * CG + g) * ROWS + rl];
    v[i] = s;
  }
  __syncthreads();
}
// block -> (row tile, row inside the tile, row inside the block, column group)
template <int CG>
__device__ __forceinline__ void rk_coords(int& rt, int& r, int& rl, int& cg) {
  constexpr int ROWS = rk_rows(CG), SUBS = TILE_M / ROWS;
  rt = blockIdx.x / SUBS;
  rl = threadIdx.x % ROWS;
  r = (blockIdx.x % SUBS) * ROWS + rl;
  cg = threadIdx.x / ROWS;
}

// ---- forward: adaLN(h) = LN(h) * (1 + scale) + shift -> packed operand; (mean, rstd) saved --------
struct AdaLnFwdArgs {
  Tl h;                       // [rt][H4][128]
  const float2* partials;     // LayerNorm partials of h from the producing GEMM [rt][stats_nt][128]
  int stats_nt;
  Tl scale, shift;            // modulation columns of this site
  int H;
  float2* stat_out;           // [rt*128 + r] (mean, rstd)
  __nv_bfloat16* out_packed;  // [rt][H/64]
};
template <int CG>
__global__ void __launch_bounds__(RK_THREADS) k_adaln_fwd(const AdaLnFwdArgs a) {
  int rt, r, rl, cg;
  rk_coords<CG>(rt, r, rl, cg);
  (void)rl;
  float sn = 0.f, mean = 0.f, m2 = 0.f;
  for (int p = 0; p < a.stats_nt; ++p) {
    const float2 s = a.partials[((size_t)rt * a.stats_nt + p) * TILE_M + r];
    stats_merge(sn, mean, m2, (float)min(TILE_N, a.H - p * TILE_N), s.x, s.y);
  }
  const float rstd = rsqrtf(m2 / (float)a.H + 1e-5f);
  if (cg == 0) a.stat_out[(size_t)rt * TILE_M + r] = make_float2(mean, rstd);
  const int nch = a.H >> 3, kb_total = (a.H + 63) >> 6;
  for (int ch = cg; ch < nch; ch += CG) {
    float x[8], sc[8], sh[8], y[8];
    tl_load8(a.h, rt, ch, r, x);
    tl_load8(a.scale, rt, ch, r, sc);
    tl_load8(a.shift, rt, ch, r, sh);
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = fmaf((x[i] - mean) * rstd, 1.0f + sc[i], sh[i]);
    pk_store8(a.out_packed, kb_total, 0, rt, ch, r, y);
  }
}

// ---- LayerNorm-modulate backward (VJP pass and final backward pass) -------------------------------
//   c_out = c_in + rstd * (a*s1 - mean(a*s1) - n * mean(a*s1*n)) (+ gp)        s1 = 1 + scale
// rows may be stacked ([M; D], 2*src_rt row tiles): saved forward tensors are read at rt % src_rt and
// the modulation gradients / second-order term only concern the first src_rt row tiles.
struct LnBwdArgs {
  Tl a;                       // cotangent of the adaLN output (GEMM result), [rows_rt][H4][128]
  Tl h;                       // forward input of the LayerNorm, [src_rt]...
  const float2* stat;         // [src_rt*128] (mean, rstd)
  Tl scale;                   // [src_rt]
  Tl c_in;                    // residual cotangent (p == null: zero)
  Tl gp;                      // second-order term added on the first src_rt tiles (p == null: none)
  int src_rt, H;
  float4* c_out;              // tiled [rows_rt][H4][128] (may alias c_in.p)
  __nv_bfloat16* c_packed;    // [rows_rt][H/64]
  // modulation gradients (first src_rt tiles): d scale = a * n (+ s1hat), d shift = a -> packed operand
  __nv_bfloat16* dmod_packed; // null: not wanted
  int dmod_kb_total, dmod_kb_scale, dmod_kb_shift;
  Tl s1hat;                   // p == null: none; scaled by s1hat_mul
  float s1hat_mul;
};
template <int CG>
__global__ void __launch_bounds__(RK_THREADS) k_ln_bwd(const LnBwdArgs a) {
  __shared__ float sm[2 * RK_THREADS];
  int rt, r, rl, cg;
  rk_coords<CG>(rt, r, rl, cg);
  const int srt = rt % a.src_rt;
  const bool first = rt < a.src_rt;
  const float2 st = a.stat[(size_t)srt * TILE_M + r];
  const float mean = st.x, rstd = st.y;
  const int nch = a.H >> 3, kb_total = (a.H + 63) >> 6;
  float red[2] = {0.f, 0.f};
  for (int ch = cg; ch < nch; ch += CG) {
    float av[8], sc[8], x[8];
    tl_load8(a.a, rt, ch, r, av);
    tl_load8(a.scale, srt, ch, r, sc);
    tl_load8(a.h, srt, ch, r, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float cn = av[i] * (1.0f + sc[i]);
      red[0] += cn;
      red[1] = fmaf(cn, (x[i] - mean) * rstd, red[1]);
    }
  }
  rk_reduce<2, CG>(red, sm, cg, rl);
  const float inv_h = 1.0f / (float)a.H;
  const float m1 = red[0] * inv_h, m2 = red[1] * inv_h;
  for (int ch = cg; ch < nch; ch += CG) {
    float av[8], sc[8], x[8], ci[8], y[8];
    tl_load8(a.a, rt, ch, r, av);
    tl_load8(a.scale, srt, ch, r, sc);
    tl_load8(a.h, srt, ch, r, x);
    if (a.c_in.p) tl_load8(a.c_in, rt, ch, r, ci);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) ci[i] = 0.f;
    }
    if (a.gp.p && first) {
      float g[8];
      tl_load8(a.gp, srt, ch, r, g);
#pragma unroll
      for (int i = 0; i < 8; ++i) ci[i] += g[i];
    }
    float n[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      n[i] = (x[i] - mean) * rstd;
      y[i] = fmaf(rstd, av[i] * (1.0f + sc[i]) - m1 - n[i] * m2, ci[i]);
    }
    tl_store8(a.c_out, a.a.ld4, 0, rt, ch, r, y);
    pk_store8(a.c_packed, kb_total, 0, rt, ch, r, y);
    if (a.dmod_packed && first) {
      float ds[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) ds[i] = av[i] * n[i];
      if (a.s1hat.p) {
        float sh[8];
        tl_load8(a.s1hat, srt, ch, r, sh);
#pragma unroll
        for (int i = 0; i < 8; ++i) ds[i] = fmaf(sh[i], a.s1hat_mul, ds[i]);
      }
      pk_store8(a.dmod_packed, a.dmod_kb_total, a.dmod_kb_scale, rt, ch, r, ds);
      pk_store8(a.dmod_packed, a.dmod_kb_total, a.dmod_kb_shift, rt, ch, r, av);
    }
  }
}

// ---- adjoint of the VJP at a LayerNorm-modulate node (pass 3) ------------------------------------
//   a  = adjoint of the VJP stream's residual cotangent after the node,  c = c_a * s1 * c_mul
//   an = ln_bwd(a);  out_packed = an * s1;  s1hat = an * c_a * c_mul;  gp = ln_second(a, c)
struct LnHatArgs {
  Tl a;                       // adjoint stream, [rt][H4][128]
  Tl ca;                      // saved VJP cotangent of the adaLN output (carries the c-stream scale)
  const float* c_mul;         // device: 1 / c-stream scale
  Tl h;
  const float2* stat;
  Tl scale;
  int H;
  __nv_bfloat16* out_packed;  // [rt][H/64]
  float4* s1hat;              // tiled [rt][H4][128]
  float4* gp;                 // tiled [rt][H4][128]
};
template <int CG>
__global__ void __launch_bounds__(RK_THREADS) k_ln_hat(const LnHatArgs a) {
  __shared__ float sm[5 * RK_THREADS];
  int rt, r, rl, cg;
  rk_coords<CG>(rt, r, rl, cg);
  const float2 st = a.stat[(size_t)rt * TILE_M + r];
  const float mean = st.x, rstd = st.y;
  const float c_mul = __ldg(a.c_mul);
  const int nch = a.H >> 3, kb_total = (a.H + 63) >> 6, H4 = a.a.ld4;
  float red[5] = {0.f, 0.f, 0.f, 0.f, 0.f};   // sum a, sum a n, sum c, sum c n, sum a c
  for (int ch = cg; ch < nch; ch += CG) {
    float av[8], cv[8], sc[8], x[8];
    tl_load8(a.a, rt, ch, r, av);
    tl_load8(a.ca, rt, ch, r, cv);
    tl_load8(a.scale, rt, ch, r, sc);
    tl_load8(a.h, rt, ch, r, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float n = (x[i] - mean) * rstd;
      const float c = cv[i] * (1.0f + sc[i]) * c_mul;
      red[0] += av[i];
      red[1] = fmaf(av[i], n, red[1]);
      red[2] += c;
      red[3] = fmaf(c, n, red[3]);
      red[4] = fmaf(av[i], c, red[4]);
    }
  }
  rk_reduce<5, CG>(red, sm, cg, rl);
  const float inv_h = 1.0f / (float)a.H;
  const float m_a = red[0] * inv_h, m_an = red[1] * inv_h, m_c = red[2] * inv_h, m_cn = red[3] * inv_h;
  const float phi0 = red[4] - (float)a.H * (m_a * m_c + m_an * m_cn);
  const float mk = -rstd * (m_a * m_cn + m_c * m_an);       // mean(k)
  const float mkn = -2.0f * rstd * m_an * m_cn;             // mean(k n)
  const float tail = phi0 * rstd * rstd * inv_h;
  for (int ch = cg; ch < nch; ch += CG) {
    float av[8], cv[8], sc[8], x[8], o[8], sh[8], g[8];
    tl_load8(a.a, rt, ch, r, av);
    tl_load8(a.ca, rt, ch, r, cv);
    tl_load8(a.scale, rt, ch, r, sc);
    tl_load8(a.h, rt, ch, r, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float n = (x[i] - mean) * rstd;
      const float s1 = 1.0f + sc[i];
      const float c = cv[i] * s1 * c_mul;
      const float an = rstd * (av[i] - m_a - n * m_an);
      o[i] = an * s1;
      sh[i] = an * cv[i] * c_mul;
      const float k = -rstd * (av[i] * m_cn + c * m_an);
      g[i] = rstd * (k - mk - n * mkn) - tail * n;
    }
    pk_store8(a.out_packed, kb_total, 0, rt, ch, r, o);
    tl_store8(a.s1hat, H4, 0, rt, ch, r, sh);
    tl_store8(a.gp, H4, 0, rt, ch, r, g);
  }
}

// =================================================================================================
// Element-wise seeds and outputs (row-major fp32 [B, n] on the caller's side)

// SiLU(cond) -> packed operand of the modulation GEMM
__global__ void k_silu_pack(const float* __restrict__ cond, int rows, int H, __nv_bfloat16* __restrict__ dst,
                            int row_tiles) {
  const int kb_total = (H + 63) >> 6;
  const size_t total = (size_t)row_tiles * kb_total * 1024;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int rt, kb, r, ch;
    chunk_coords_rowmajor(idx, kb_total, rt, kb, r, ch);
    const int row = rt * TILE_M + r, c0 = kb * 64 + ch * 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float x = (row < rows && c0 + i < H) ? __ldg(cond + (size_t)row * H + c0 + i) : 0.f;
      v[i] = (row < rows && c0 + i < H) ? x / (1.0f + __expf(-x)) : 0.f;
    }
    store_chunk(dst, rt, kb, kb_total, r, ch, v);
  }
}

// src[rows, cols] * mul (* *mul_dev) -> packed [copies][row_tiles][kb_alloc] (zero padded; `copies`
// stacked images of the same rows)
__global__ void k_pack_scaled(const float* __restrict__ src, int rows, int cols, float mul,
                              const float* __restrict__ mul_dev, __nv_bfloat16* __restrict__ dst, int row_tiles,
                              int kb_alloc, int copies) {
  const float m = mul * (mul_dev ? __ldg(mul_dev) : 1.0f);
  const size_t total = (size_t)row_tiles * kb_alloc * 1024;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int rt, kb, r, ch;
    chunk_coords_rowmajor(idx, kb_alloc, rt, kb, r, ch);
    const int row = rt * TILE_M + r, c0 = kb * 64 + ch * 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      v[i] = (row < rows && c0 + i < cols) ? __ldg(src + (size_t)row * cols + c0 + i) * m : 0.f;
    for (int c = 0; c < copies; ++c) store_chunk(dst, c * row_tiles + rt, kb, kb_alloc, r, ch, v);
  }
}

// Seeds at the clamp (models/score_networks.py:167-170):  s = clamp(r, +-10) * mult * tw.
//   mode 0 (VJP):      out = [|r| <= 10] * mult * tw * mul                       (cotangent 1 on every s)
//   mode 1 (backward): out = s_bar * [|r| <= 10] * mult * tw * mul * (*mul_dev), two stacked copies;
//                      block partial sums of s_bar * clamp(r) * tw -> part[blockIdx.x]   (d mult)
//   mode 2 (adjoint):  block partial sums of x * [|r| <= 10] * tw -> part[blockIdx.x]  (d mult, penalty)
__global__ void __launch_bounds__(256) k_clamp_seed(const float* __restrict__ rr, const float* __restrict__ tw,
                                                    const float* __restrict__ mult, const float* __restrict__ x,
                                                    int rows, int L, float mul, const float* __restrict__ mul_dev,
                                                    __nv_bfloat16* __restrict__ dst, int row_tiles, int kb_alloc,
                                                    int mode, float* __restrict__ part) {
  __shared__ float red[256];
  const float m = __ldg(mult) * mul * ((mul_dev && mode != 2) ? __ldg(mul_dev) : 1.0f);
  const size_t total = (size_t)row_tiles * kb_alloc * 1024;
  float acc = 0.f;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int rt, kb, r, ch;
    chunk_coords_rowmajor(idx, kb_alloc, rt, kb, r, ch);
    const int row = rt * TILE_M + r, c0 = kb * 64 + ch * 8;
    const float t = (row < rows && tw) ? __ldg(tw + row) : 1.0f;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = 0.f;
      if (row < rows && c0 + i < L) {
        const float rv = __ldg(rr + (size_t)row * L + c0 + i);
        const float mask = (rv >= -10.f && rv <= 10.f) ? 1.f : 0.f;
        if (mode == 0) {
          v[i] = mask * m * t;
        } else {
          const float xv = __ldg(x + (size_t)row * L + c0 + i);
          if (mode == 1) {
            v[i] = xv * mask * m * t;
            acc = fmaf(xv * fminf(fmaxf(rv, -10.f), 10.f), t, acc);
          } else {
            acc = fmaf(xv * mask, t, acc);
          }
        }
      }
    }
    if (mode == 0) store_chunk(dst, rt, kb, kb_alloc, r, ch, v);
    else if (mode == 1) {
      store_chunk(dst, rt, kb, kb_alloc, r, ch, v);
      store_chunk(dst, row_tiles + rt, kb, kb_alloc, r, ch, v);
    }
  }
  if (mode != 0) {
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = red[0];
  }
}
// d mult = sum(part_bwd) + sum(part_hat) * (*inv_scale_dev)
__global__ void k_mult_grad(const float* __restrict__ part_bwd, int n_bwd, const float* __restrict__ part_hat,
                            int n_hat, const float* __restrict__ inv_scale_dev, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float a = 0.f, b = 0.f;
  for (int i = 0; i < n_bwd; ++i) a += part_bwd[i];
  if (part_hat)
    for (int i = 0; i < n_hat; ++i) b += part_hat[i];
  out[0] = a + b * (part_hat ? __ldg(inv_scale_dev) : 0.f);
}

// Stream scales (fp16 operands have a 5-bit exponent: the cotangent streams are kept near 1; the
// scaling is by exact powers of two and is removed at the outputs).
//   scale[3] = S_c, scale[4] = 1/S_c : VJP ("c") stream, |mult| * S_c in [0.25, 0.5)  (seed = mask * mult * tw)
//   scale[0] = S_b, scale[1] = 1/S_b, scale[2] = 1/(S_b S_c) : backward streams,
//              max(|s_bar|_max * |mult|, |g_bar|_max) * S_b in [32, 64)
__device__ __forceinline__ float pow2_scale(float mx, int target_exp) {
  if (!(mx > 0.f) || !(mx < 3.0e38f)) return 1.0f;
  int e;
  frexpf(mx, &e);                 // mx = f * 2^e, f in [0.5, 1)
  e = target_exp - e;
  e = e > 100 ? 100 : (e < -100 ? -100 : e);
  return ldexpf(1.0f, e);
}
__global__ void k_cscale(const float* __restrict__ mult, float* __restrict__ scale) {
  if (threadIdx.x != 0) return;
  const float S = pow2_scale(fabsf(__ldg(mult)), -1);
  scale[3] = S;
  scale[4] = 1.0f / S;
}
// two stages: per-block maxima folded with atomicMax on the bit pattern (non-negative floats order like
// unsigned integers) into scale[8] (a), scale[9] (b); then one thread derives the scales
__global__ void __launch_bounds__(256) k_amax_partial(const float* __restrict__ a, size_t na, const float* __restrict__ b,
                                                      size_t nb, float* __restrict__ scale) {
  __shared__ float red[2][256];
  float ma = 0.f, mb = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x, i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = i0; i < (na >> 2); i += stride) {
    const float4 v = reinterpret_cast<const float4*>(a)[i];
    ma = fmaxf(ma, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  for (size_t i = (na & ~(size_t)3) + i0; i < na; i += stride) ma = fmaxf(ma, fabsf(a[i]));
  if (b) {
    for (size_t i = i0; i < (nb >> 2); i += stride) {
      const float4 v = reinterpret_cast<const float4*>(b)[i];
      mb = fmaxf(mb, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (size_t i = (nb & ~(size_t)3) + i0; i < nb; i += stride) mb = fmaxf(mb, fabsf(b[i]));
  }
  red[0][threadIdx.x] = ma;
  red[1][threadIdx.x] = mb;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      red[0][threadIdx.x] = fmaxf(red[0][threadIdx.x], red[0][threadIdx.x + o]);
      red[1][threadIdx.x] = fmaxf(red[1][threadIdx.x], red[1][threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    atomicMax(reinterpret_cast<unsigned int*>(scale + 8), __float_as_uint(red[0][0]));
    atomicMax(reinterpret_cast<unsigned int*>(scale + 9), __float_as_uint(red[1][0]));
  }
}
__global__ void k_amax_finish(const float* __restrict__ a_mul, float* __restrict__ scale) {
  if (threadIdx.x != 0) return;
  const float m = fmaxf(scale[8] * fabsf(__ldg(a_mul)), scale[9]);
  const float S = pow2_scale(m, 6);
  scale[0] = S;
  scale[1] = 1.0f / S;
  scale[2] = 1.0f / (S * scale[3]);
}

// out[i] = in[i] * mul * (*mul_dev)
__global__ void k_scale_copy(const float* in, size_t n, float mul, const float* __restrict__ mul_dev, float* out) {
  const float m = mul * (mul_dev ? __ldg(mul_dev) : 1.0f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = in[i] * m;
}
// d cond = d csilu * silu'(cond) * (*mul_dev)
__global__ void k_dcond(const float* __restrict__ dcs, const float* __restrict__ cond, size_t n,
                        const float* __restrict__ mul_dev, float* __restrict__ out) {
  const float m = __ldg(mul_dev);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = dcs[i] * dact_d1<ACT_SILU>(cond[i]) * m;
}

}  // namespace aid
