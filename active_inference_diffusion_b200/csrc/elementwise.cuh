// Memory-bound helper kernels around the tcgen05 GEMMs: operand packing (fp32 row-major ->
// bf16 swizzled tiles), weight packing/folding, time embeddings, conditioning, LayerNorm.
#pragma once
#include "gemm.cuh"

namespace aid {

// One thread per 16-byte chunk (8 bf16) of the packed output; consecutive threads take
// consecutive rows of the same K chunk, i.e. consecutive 16-byte slots of the packed tile.
// chunk id -> (tile, chunk_in_row, r); tile -> (rt, kb), rt counting 128-row tiles.
__device__ __forceinline__ void chunk_coords(size_t idx, int kb_total, int& rt, int& kb, int& r,
                                             int& ch) {
  r = (int)(idx & 127);
  ch = (int)((idx >> 7) & 7);
  size_t tile = idx >> 10;
  kb = (int)(tile % kb_total);
  rt = (int)(tile / kb_total);
}

// Same work item, but neighbouring threads take neighbouring 16-byte chunks of ONE row: the mapping
// for kernels that read row-major fp32 (8 lanes cover 256 contiguous bytes of a row; the packed
// stores then form 64-byte runs).
__device__ __forceinline__ void chunk_coords_rowmajor(size_t idx, int kb_total, int& rt, int& kb, int& r,
                                                      int& ch) {
  ch = (int)(idx & 7);
  r = (int)((idx >> 3) & 127);
  size_t tile = idx >> 10;
  kb = (int)(tile % kb_total);
  rt = (int)(tile / kb_total);
}

// nw: packed tiles are nw*128 rows tall (weights of N=256 MMAs: nw = 2); rt counts 128-row tiles.
__device__ __forceinline__ void store_chunk(__nv_bfloat16* dst, int rt, int kb, int kb_total, int r,
                                            int ch, const float (&v)[8], int nw = 1) {
  uint4 o;
  o.x = pack_op16x2(v[0], v[1]);
  o.y = pack_op16x2(v[2], v[3]);
  o.z = pack_op16x2(v[4], v[5]);
  o.w = pack_op16x2(v[6], v[7]);
  const int R = nw * TILE_M;
  __nv_bfloat16* tile = dst + ((size_t)(rt / nw) * kb_total + kb) * ((size_t)nw * TILE_ELEMS);
  *reinterpret_cast<uint4*>(tile + ch * (R * 8) + ((rt % nw) * TILE_M + r) * 8) = o;
}

// fp32 row-major [rows, cols] (leading dim ld) -> packed bf16 [row_tiles][kb]; zero padding.
// src_row_map (optional) gathers rows: packed row i reads source row src_row_map(i).
enum RowMap : int { MAP_PLAIN = 0, MAP_MODLN = 1 };
// MAP_MODLN: adaLN modulation weight [2H, H]; packed n-tile t holds [scale rows 64t..64t+63 |
// shift rows H+64t..H+64t+63] so one accumulator tile carries both halves for 64 hidden columns
// (models/score_networks.py:265-270: scale, shift = chunk(2)).
__device__ __forceinline__ int map_row(int i, int mode, int H) {
  if (mode == MAP_MODLN) {
    int t = i >> 7, j = i & 127;
    return (j < 64) ? (t * 64 + j) : (H + t * 64 + (j - 64));
  }
  return i;
}

__global__ void k_pack_rows(const float* __restrict__ src, int rows, int cols, int ld,
                            __nv_bfloat16* __restrict__ dst, int row_tiles, int kb_total, int mode,
                            int H, int nw) {
  size_t total = (size_t)row_tiles * kb_total * 1024;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int rt, kb, r, ch;
    chunk_coords_rowmajor(idx, kb_total, rt, kb, r, ch);
    int prow = rt * TILE_M + r;
    int c0 = kb * TILE_K + ch * 8;
    float v[8];
    int srow = -1;
    if (mode == MAP_MODLN) {
      // valid only when the mapped row exists (t*64+j < H)
      int t = prow >> 7, j = prow & 127;
      int hc = t * 64 + (j & 63);
      if (hc < H && prow < rows) srow = map_row(prow, mode, H);
    } else if (prow < rows) {
      srow = prow;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      v[i] = (srow >= 0 && c0 + i < cols) ? __ldg(src + (size_t)srow * ld + c0 + i) : 0.f;
    store_chunk(dst, rt, kb, kb_total, r, ch, v, nw);
  }
}

// bias [n] -> padded fp32 [n_pad] with the same row map.
__global__ void k_pack_bias(const float* __restrict__ src, int n, float* __restrict__ dst, int n_pad,
                            int mode, int H) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  float v = 0.f;
  if (mode == MAP_MODLN) {
    // scale rows (j < 64) carry 1 + bias so the epilogue needs no separate "1 + scale" add; padded
    // hidden columns get scale' = 0 and shift = 0 (their weight rows are zero too) -> output 0
    int t = i >> 7, j = i & 127;
    if (t * 64 + (j & 63) < H) v = src[map_row(i, mode, H)] + (j < 64 ? 1.0f : 0.0f);
  } else if (i < n && src) {
    v = src[i];
  }
  dst[i] = v;
}

// Folded single-token attention: nn.MultiheadAttention over sequence length 1 has softmax == 1,
// so attn(x) = W_o (W_v x + b_v) + b_o = (W_o W_v) x + (W_o b_v + b_o)
// (models/score_networks.py:189-194,224-227; W_v = in_proj_weight[2H:3H]).
// Computes the fp32 product and writes packed bf16 tiles directly.
__global__ void k_pack_folded_attn(const float* __restrict__ in_proj_w, const float* __restrict__ wo,
                                   int H, __nv_bfloat16* __restrict__ dst, int n_tiles, int kb_total,
                                   int nw) {
  size_t total = (size_t)n_tiles * kb_total * 1024;
  const float* wv = in_proj_w + (size_t)2 * H * H;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int nt, kb, r, ch;
    chunk_coords(idx, kb_total, nt, kb, r, ch);
    int n = nt * TILE_M + r;
    int k0 = kb * TILE_K + ch * 8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (n < H && k0 < H) {
      for (int j = 0; j < H; ++j) {
        float a = __ldg(wo + (size_t)n * H + j);
        const float* row = wv + (size_t)j * H + k0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (k0 + i < H) acc[i] = fmaf(a, __ldg(row + i), acc[i]);
      }
    }
    store_chunk(dst, nt, kb, kb_total, r, ch, acc, nw);
  }
}
__global__ void k_folded_attn_bias(const float* __restrict__ in_proj_b, const float* __restrict__ wo,
                                   const float* __restrict__ bo, int H, float* __restrict__ dst,
                                   int n_pad) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_pad) return;
  float acc = 0.f;
  if (n < H) {
    const float* bv = in_proj_b + 2 * H;
    for (int j = 0; j < H; ++j) acc = fmaf(__ldg(wo + (size_t)n * H + j), __ldg(bv + j), acc);
    acc += bo[n];
  }
  dst[n] = acc;
}

// Sinusoidal embedding rows -> packed bf16 [row_tiles][dim/64].
// models/score_networks.py:282-291: freq_i = exp(-i ln(1e4)/(half-1)) * freq_scale; [sin | cos].
__global__ void k_sincos_pack(const float* __restrict__ t_rows, int rows,
                              const float* __restrict__ freq_scale, int dim,
                              __nv_bfloat16* __restrict__ dst, int row_tiles) {
  const int kb_total = (dim + TILE_K - 1) / TILE_K;
  const int half = dim / 2;
  const float kf = logf(10000.0f) / (float)(half - 1);
  const float fs = __ldg(freq_scale);
  size_t total = (size_t)row_tiles * kb_total * 1024;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int rt, kb, r, ch;
    chunk_coords(idx, kb_total, rt, kb, r, ch);
    int row = rt * TILE_M + r;
    int c0 = kb * TILE_K + ch * 8;
    float v[8];
    float t = (row < rows) ? __ldg(t_rows + row) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int c = c0 + i;
      float o = 0.f;
      if (row < rows && c < 2 * half) {
        int fi = (c < half) ? c : c - half;
        float f = expf((float)fi * -kf) * fs;
        float a = t * f;
        o = (c < half) ? sinf(a) : cosf(a);
      }
      v[i] = o;
    }
    store_chunk(dst, rt, kb, kb_total, r, ch, v);
  }
}

// continuous_time_embed[0]: Linear(1 -> E) + SiLU -> packed (models/score_networks.py:60-62).
__global__ void k_cont0_pack(const float* __restrict__ tn_rows, int rows, const float* __restrict__ w,
                             const float* __restrict__ b, int E, __nv_bfloat16* __restrict__ dst,
                             int row_tiles) {
  const int kb_total = (E + TILE_K - 1) / TILE_K;
  size_t total = (size_t)row_tiles * kb_total * 1024;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int rt, kb, r, ch;
    chunk_coords(idx, kb_total, rt, kb, r, ch);
    int row = rt * TILE_M + r;
    int c0 = kb * TILE_K + ch * 8;
    float tn = (row < rows) ? __ldg(tn_rows + row) : 0.f;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int c = c0 + i;
      v[i] = (row < rows && c < E) ? act_silu(fmaf(__ldg(w + c), tn, __ldg(b + c))) : 0.f;
    }
    store_chunk(dst, rt, kb, kb_total, r, ch, v);
  }
}

// Conditioning: c = t_sin[row|srow] + cont_flag * time_scale * t_cont[...] + obs_emb[row]
// (models/score_networks.py:123-153).  Output SiLU(c) packed (adaLN input) or raw c tiled.
// t_* may be indexed by a fixed row `fixed_row` (sampler: the step's row of the step table)
// instead of the batch row.
struct CondArgs {
  const float4* t_sin;     // tiled [.. ][ld4][128]
  const float4* t_cont;    // tiled or null
  const float* cont_flag;  // per t-row flag (1 = continuous branch) or null (= use t_cont if given)
  const float* time_scale; // device scalar
  const float4* obs_emb;   // tiled [rt][ld4][128] or null
  int fixed_row;           // >= 0: read t_* at this row for every batch row
  int ld4;
  int H;
  int rows;
  int row_tiles;
  __nv_bfloat16* out_packed;  // SiLU(c) packed [rt][H/64]  (may be null)
  float4* out_tiled;          // c tiled (may be null)
};
__global__ void k_cond(const CondArgs a) {
  // one thread per (row, 8 columns); consecutive threads = consecutive rows, so the tiled fp32
  // reads/writes and the 16-byte packed stores are all coalesced
  const int kb_total = (a.H + TILE_K - 1) / TILE_K;
  const int nch = a.ld4 / 2;
  size_t total = (size_t)a.row_tiles * nch * TILE_M;
  const float ts = a.t_cont ? __ldg(a.time_scale) : 0.f;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx & 127);
    const int chn = (int)((idx >> 7) % nch);
    const int rt = (int)((idx >> 7) / nch);
    const int trow = (a.fixed_row >= 0) ? a.fixed_row : rt * TILE_M + r;
    float f = 0.f;
    if (a.t_cont) f = a.cont_flag ? __ldg(a.cont_flag + trow) : 1.f;
    float o8[8];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int c4 = chn * 2 + hh;
      const size_t toff = ((size_t)(trow >> 7) * a.ld4 + c4) * TILE_M + (trow & 127);
      const size_t boff = ((size_t)rt * a.ld4 + c4) * TILE_M + r;
      float4 v = a.t_sin[toff];
      if (f != 0.f) {
        float4 c = a.t_cont[toff];
        v.x = __fadd_rn(v.x, __fmul_rn(ts, c.x));
        v.y = __fadd_rn(v.y, __fmul_rn(ts, c.y));
        v.z = __fadd_rn(v.z, __fmul_rn(ts, c.z));
        v.w = __fadd_rn(v.w, __fmul_rn(ts, c.w));
      }
      if (a.obs_emb) {
        float4 o = a.obs_emb[boff];
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      if (a.out_tiled) a.out_tiled[boff] = v;
      o8[hh * 4 + 0] = act_silu(v.x); o8[hh * 4 + 1] = act_silu(v.y);
      o8[hh * 4 + 2] = act_silu(v.z); o8[hh * 4 + 3] = act_silu(v.w);
    }
    if (a.out_packed) {
      const int c = chn * 8;
      if (c < a.H) store_chunk(a.out_packed, rt, c >> 6, kb_total, r, (c >> 3) & 7, o8);
    }
  }
}

// LayerNorm (affine) + activation over tiled fp32 input with (mean, M2) partials.
// grid = (row_tiles, ceil(n/64)); block = 128 (thread = row).
struct LnArgs {
  const float4* x;       // tiled [rt][ld4][128]
  const float2* stats;   // [rt][stats_nt][128]
  int stats_nt;
  int ld4;
  int n;                 // LayerNorm width
  const float* gamma;
  const float* beta;
  int act;
  __nv_bfloat16* out_packed;  // [rt][ceil(n/64)] or null
  float4* out_tiled;          // [rt][ld4][128] or null
  const float4* resid;        // tiled [rt][ld4][128], added AFTER the activation, or null
};
__global__ void k_ln_act(const LnArgs a) {
  const int rt = blockIdx.x, kb = blockIdx.y, r = threadIdx.x;
  const int kb_total = (a.n + TILE_K - 1) / TILE_K;
  // this block's 64 gamma / beta values: staged once in shared memory and read back as float4
  // broadcasts (per-element __ldg was 128 uniform global loads per thread next to 16 data loads)
  __shared__ float4 s_gb[32];
  if (r < 32) {
    const int c = kb * 64 + (r & 15) * 4;
    const float* src = (r < 16) ? a.gamma : a.beta;
    float4 v;
    v.x = (c + 0 < a.n) ? __ldg(src + c + 0) : 0.f;
    v.y = (c + 1 < a.n) ? __ldg(src + c + 1) : 0.f;
    v.z = (c + 2 < a.n) ? __ldg(src + c + 2) : 0.f;
    v.w = (c + 3 < a.n) ? __ldg(src + c + 3) : 0.f;
    s_gb[r] = v;
  }
  float sn = 0.f, mean = 0.f, m2 = 0.f;
  for (int p = 0; p < a.stats_nt; ++p) {
    float2 s = a.stats[((size_t)rt * a.stats_nt + p) * TILE_M + r];
    stats_merge(sn, mean, m2, (float)min(TILE_N, a.n - p * TILE_N), s.x, s.y);
  }
  const float rstd = rsqrtf(m2 / (float)a.n + 1e-5f);
  __syncthreads();
#pragma unroll
  for (int half = 0; half < 2; ++half) {   // unrolled: all 16 float4 loads of the 64 columns in flight
    float y[32];
    const int c0 = kb * 64 + half * 32;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      int c = c0 + q * 4;
      float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < a.ld4 * 4) xv = a.x[((size_t)rt * a.ld4 + (c >> 2)) * TILE_M + r];
      float xs[4] = {xv.x, xv.y, xv.z, xv.w};
      float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.resid && c < a.ld4 * 4) rv = a.resid[((size_t)rt * a.ld4 + (c >> 2)) * TILE_M + r];
      const float rs[4] = {rv.x, rv.y, rv.z, rv.w};
      const float4 g4 = s_gb[half * 8 + q], b4 = s_gb[16 + half * 8 + q];
      const float gs[4] = {g4.x, g4.y, g4.z, g4.w}, bs[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = 0.f;
        if (c + i < a.n) {
          v = (xs[i] - mean) * rstd * gs[i] + bs[i];
          v = act_apply(v, a.act) + rs[i];
        }
        y[q * 4 + i] = v;
      }
      if (a.out_tiled && c < a.ld4 * 4)
        a.out_tiled[((size_t)rt * a.ld4 + (c >> 2)) * TILE_M + r] =
            make_float4(y[q * 4 + 0], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
    }
    if (a.out_packed) {
      __nv_bfloat16* tile = a.out_packed + ((size_t)rt * kb_total + kb) * TILE_ELEMS;
      store_packed32(tile, r, half * 32, y);
    }
  }
}

// tiled fp32 -> row-major fp32 (debug / public outputs)
__global__ void k_untile(const float4* __restrict__ x, int ld4, int rows, int cols,
                         float* __restrict__ out, int ld_out) {
  size_t total = (size_t)((rows + TILE_M - 1) / TILE_M) * ld4 * TILE_M;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(idx & 127);
    int c4 = (int)((idx >> 7) % ld4);
    int rt = (int)((idx >> 7) / ld4);
    int row = rt * TILE_M + r;
    if (row >= rows) continue;
    float4 v = x[idx];
    float vs[4] = {v.x, v.y, v.z, v.w};
    for (int i = 0; i < 4; ++i)
      if (c4 * 4 + i < cols) out[(size_t)row * ld_out + c4 * 4 + i] = vs[i];
  }
}

}  // namespace aid
