// Fused epilogues of the tcgen05 GEMM (included by gemm.cuh inside namespace aid).
//
// One 128x128 fp32 accumulator tile per call, one thread per row (TMEM lane).  Every operand an
// epilogue needs from global memory (bias, residual rows, LayerNorm input rows and statistics) is
// software-pipelined ACROSS tiles: the registers that held tile i's operands are refilled with
// tile i+1's as soon as they are consumed, so the loads fly during the rest of tile i's epilogue
// and tile i+1's MMAs, and nothing waits on HBM/L2 latency when the next accumulator is ready.
//   epi_first      : loads the operands of a group's first tile.
//   epi_stage_bias : puts the tile's 128 bias values in shared memory (broadcast reads).
//   epi_finish     : TMEM -> registers (next chunk's tcgen05.ld in flight while this chunk is
//                    processed), math, stores, refill for the group's next tile (rt2, nt2).
// The 32-column chunk loops are deliberately NOT fully unrolled: the fully unrolled kernels were
// 150-250 KB of SASS and stalled on instruction fetch (ncu: stall_no_instruction on top).
// rt/nt: row tile / n-tile indices; r: row inside the tile.

// ---- activation derivatives for the training epilogues (EPI_DACT) ------------------------------
// GELU (exact, models/score_networks.py:199) through the Abramowitz-Stegun erf of act_gelu_as:
// with e = exp(-x^2/2), q = poly(t) e:  Phi(x) = x < 0 ? q/2 : 1 - q/2,  phi(x) = e / sqrt(2 pi),
// gelu = x Phi, gelu' = Phi + x phi, gelu'' = phi (2 - x^2).
__device__ __forceinline__ void gelu_parts(float x, float& Phi, float& phi) {
  const float y = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, y, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = __expf(-y * y);
  const float hq = 0.5f * p * t * e;
  Phi = x < 0.f ? hq : 1.0f - hq;
  phi = e * 0.3989422804014327f;
}
template <int ACT>
__device__ __forceinline__ float dact_f(float x) {
  if constexpr (ACT == ACT_GELU) {
    float Phi, phi;
    gelu_parts(x, Phi, phi);
    return x * Phi;
  } else if constexpr (ACT == ACT_SILU) {
    return x / (1.0f + __expf(-x));
  } else {
    return x;
  }
}
template <int ACT>
__device__ __forceinline__ float dact_d1(float x) {
  if constexpr (ACT == ACT_GELU) {
    float Phi, phi;
    gelu_parts(x, Phi, phi);
    return fmaf(x, phi, Phi);
  } else if constexpr (ACT == ACT_SILU) {
    const float s = 1.0f / (1.0f + __expf(-x));
    return s * fmaf(x, 1.0f - s, 1.0f);
  } else {
    return 1.0f;
  }
}
template <int ACT>
__device__ __forceinline__ float dact_d2(float x) {
  if constexpr (ACT == ACT_GELU) {
    float Phi, phi;
    gelu_parts(x, Phi, phi);
    return phi * (2.0f - x * x);
  } else if constexpr (ACT == ACT_SILU) {
    const float s = 1.0f / (1.0f + __expf(-x));
    return s * (1.0f - s) * fmaf(x, 1.0f - 2.0f * s, 2.0f);
  } else {
    return 0.0f;
  }
}
// 8 packed 16-bit operand elements -> fp32
__device__ __forceinline__ void unpack_op16x8(const uint4 v, float* out) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#ifdef AID_F16
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
#else
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
#endif
    out[2 * i] = f.x;
    out[2 * i + 1] = f.y;
  }
}

// value, first and second derivative of the activation on a register pair
template <int ACT>
__device__ __forceinline__ void dact_eval2(const float2 x, float2& f, float2& d1, float2& d2) {
  if constexpr (ACT == ACT_GELU) {
    // erf in the Abramowitz-Stegun 7.1.26 form (see gelu_parts), all polynomial work as FFMA2
    const float2 y = __fmul2_rn(make_float2(fabsf(x.x), fabsf(x.y)), make_float2(0.70710678118654752f, 0.70710678118654752f));
    const float2 den = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), y, make_float2(1.0f, 1.0f));
    const float2 t = make_float2(rcp_approx(den.x), rcp_approx(den.y));
    // 0.5 * (a1 + a2 t + a3 t^2 + a4 t^3 + a5 t^4): the 0.5 of Phi = 1 - q/2 folded into the coefficients
    float2 p = __ffma2_rn(make_float2(0.5307027145f, 0.5307027145f), t, make_float2(-0.7265760135f, -0.7265760135f));
    p = __ffma2_rn(p, t, make_float2(0.7107068705f, 0.7107068705f));
    p = __ffma2_rn(p, t, make_float2(-0.142248368f, -0.142248368f));
    p = __ffma2_rn(p, t, make_float2(0.127414796f, 0.127414796f));
    const float2 y2 = __fmul2_rn(y, y);
    float2 ex;
    ex.x = ex2_approx(-1.4426950408889634f * y2.x);
    ex.y = ex2_approx(-1.4426950408889634f * y2.y);
    const float2 hq = __fmul2_rn(__fmul2_rn(p, t), ex);
    const float2 Phi = make_float2(x.x < 0.f ? hq.x : 1.0f - hq.x, x.y < 0.f ? hq.y : 1.0f - hq.y);
    const float2 phi = __fmul2_rn(ex, make_float2(0.3989422804014327f, 0.3989422804014327f));
    f = __fmul2_rn(x, Phi);
    d1 = __ffma2_rn(x, phi, Phi);
    d2 = __fmul2_rn(phi, __ffma2_rn(make_float2(-x.x, -x.y), x, make_float2(2.0f, 2.0f)));
  } else {
    f = make_float2(dact_f<ACT>(x.x), dact_f<ACT>(x.y));
    d1 = make_float2(dact_d1<ACT>(x.x), dact_d1<ACT>(x.y));
    d2 = make_float2(dact_d2<ACT>(x.x), dact_d2<ACT>(x.y));
  }
}

template <int EPI>
struct EpiState {
  float bias;        // bias[nt*128 + r] of the tile about to be processed
};
template <>
struct EpiState<EPI_F32> {
  float bias;
  float4 res[2][8];  // residual chunks 0 and 1 of the tile about to be processed
};
template <>
struct EpiState<EPI_MODLN> {
  float bias;
  float4 h[16];      // the 64 hidden columns the tile normalises
  float mean, rstd;
  int rt;            // row tile the statistics belong to
};

__device__ __forceinline__ void bias32_from_smem(const float* sb, int c0, float (&b)[32]) {
  const float4* p = reinterpret_cast<const float4*>(sb + c0);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float4 v = p[q];
    b[q * 4 + 0] = v.x; b[q * 4 + 1] = v.y; b[q * 4 + 2] = v.z; b[q * 4 + 3] = v.w;
  }
}

__device__ __forceinline__ void group_bar(int bar_id) {
  asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
}

// ---- optional shared-memory staging of output tiles (pair kernels) ---------------------------
// A packed bf16 k-block tile [8 chunks][128 rows][16 B] and a 32-column block of the tiled fp32
// layout [8 float4 columns][128 rows][16 B] are both 16 KiB contiguous in global memory and have
// the same shape, so an epilogue group can assemble one in shared memory (a conflict-free 16-byte
// st.shared per thread and chunk) and hand it to the TMA engine as ONE bulk store.  With one CTA
// per tile this lost (the shared-memory port is the bottleneck there); in the CTA-pair kernels the
// UMMA and the weight refills need 96 instead of 160 B/clk of that port, and taking the stores off
// the LSU path pays.  SMODE 0 = direct 16-byte global stores; 1 = assemble + TMA bulk store;
// 2 = assemble in shared memory only (the tile is the next GEMM's A operand, chain2.cuh: the
// caller owns the synchronisation).
struct Stage {
  uint32_t base;  // shared-memory address of this group's 16 KiB staging buffer (0: not staged)
  int bar_id;     // named barrier of the group
  bool leader;    // the thread that issues (and therefore tracks) the bulk stores
};
__device__ __forceinline__ void stage_acquire(const Stage& s) {
  if (s.leader) bulk_wait_read<0>();   // the previous bulk store has finished reading the buffer
  group_bar(s.bar_id);
}
__device__ __forceinline__ void stage_flush(const Stage& s, void* gdst) {
  fence_async_smem();
  group_bar(s.bar_id);
  if (s.leader) {
    bulk_s2g(gdst, s.base, TILE_BYTES);
    bulk_commit();
  }
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
// 8 consecutive bf16 columns (16-byte chunk `chunk`) of row r into a staged packed tile
__device__ __forceinline__ void sts_packed8(uint32_t tile_smem, int r, int chunk, const float* y) {
  sts_v4(tile_smem + chunk * (TILE_M * 16) + r * 16, pack_op16x2(y[0], y[1]), pack_op16x2(y[2], y[3]),
         pack_op16x2(y[4], y[5]), pack_op16x2(y[6], y[7]));
}

// Packed fp32x2 arithmetic (sm_100 FADD2 / FMUL2 / FFMA2: two IEEE-rn fp32 results per issue slot).
// The epilogues are bound by the FP32 pipe's issue rate, so the hot math works on register pairs.
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
  const float2 r = __fadd2_rn(make_float2(a0, a1), make_float2(b0, b1));
  a0 = r.x; a1 = r.y;
}
// GELU on two elements.  Default: the tanh form on the MUFU unit (see act_gelu).  -DAID_GELU_POLY:
// x*Phi(x) with erf(x/sqrt2) ~= xc*P(xc^2), xc = clamp(x, +-4), P of degree 6 (minimax fit,
// |gelu error| <= 1.9e-4 vs the exact erf form -- closer than the tanh form's 5e-4), packed FFMA2
// only; measured SLOWER on B200 (mlp.0 epilogue compute 92 us vs 72 us: the FMA pipe, not the XU
// pipe, becomes the bound), kept as an accuracy option.  -DAID_EXACT_GELU: libm erff.
__device__ __forceinline__ void act_gelu2(float& a, float& b) {
#if defined(AID_EXACT_GELU)
  a = act_gelu(a); b = act_gelu(b);
#elif defined(AID_F16)
  // fp16-operand builds (rel-1e-3 mode): Phi from erfc(t) = 2^(t Q(t)), t = |x|/sqrt2 <= 4.3, Q of degree 5
  // (weighted least-squares fit of log2 erfc; |erfc error| <= 5e-7 evaluated in fp32): ONE MUFU
  // operation per element like the tanh form of the bf16 build, erf-accurate.  For x < 0 the
  // result is x * erfc/2 directly (no cancellation in the tail).
  const float2 x = make_float2(a, b);
  const float2 t = make_float2(fminf(fabsf(a) * 0.70710678118654752f, 4.3f), fminf(fabsf(b) * 0.70710678118654752f, 4.3f));
#define AID_C2(v) make_float2(v, v)
  float2 q = __ffma2_rn(t, AID_C2(0.00024000316524137267f), AID_C2(-0.004126049044898346f));
  q = __ffma2_rn(q, t, AID_C2(0.031667067247158676f));
  q = __ffma2_rn(q, t, AID_C2(-0.15025340151855682f));
  q = __ffma2_rn(q, t, AID_C2(-0.9180013625978741f));
  q = __ffma2_rn(q, t, AID_C2(-1.6279399503718077f));
#undef AID_C2
  const float2 pw = __fmul2_rn(q, t);
  const float h0 = 0.5f * ex2_approx(pw.x), h1 = 0.5f * ex2_approx(pw.y);
  const float2 Phi = make_float2(a < 0.f ? h0 : 1.0f - h0, b < 0.f ? h1 : 1.0f - h1);
  const float2 r = __fmul2_rn(x, Phi);
  a = r.x; b = r.y;
#elif !defined(AID_GELU_POLY)
  const float2 x = make_float2(a, b);
  const float2 p = __ffma2_rn(__fmul2_rn(x, x), make_float2(0.0356774081f, 0.0356774081f),
                              make_float2(0.7978845608f, 0.7978845608f));
  const float2 u = __fmul2_rn(x, p);
  float t0, t1;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(u.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(u.y));
  const float2 hx = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  const float2 r = __ffma2_rn(hx, make_float2(t0, t1), hx);
  a = r.x; b = r.y;
#else
  const float2 x = make_float2(a, b);
  const float2 xc = make_float2(fminf(fmaxf(a, -4.0f), 4.0f), fminf(fmaxf(b, -4.0f), 4.0f));
  const float2 t = __fmul2_rn(xc, xc);
#define AID_C2(v) make_float2(v, v)
  float2 p = __ffma2_rn(t, AID_C2(4.556275002e-08f), AID_C2(-3.197196975e-06f));
  p = __ffma2_rn(p, t, AID_C2(9.591087291e-05f));
  p = __ffma2_rn(p, t, AID_C2(-1.628029160e-03f));
  p = __ffma2_rn(p, t, AID_C2(1.754476689e-02f));
  p = __ffma2_rn(p, t, AID_C2(-1.291462034e-01f));
  p = __ffma2_rn(p, t, AID_C2(7.957667112e-01f));
#undef AID_C2
  float2 e = __fmul2_rn(p, xc);                                   // ~erf(x / sqrt 2), in (-1.0002, 1.0002)
  e = make_float2(fminf(fmaxf(e.x, -1.0f), 1.0f), fminf(fmaxf(e.y, -1.0f), 1.0f));
  const float2 hx = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  const float2 r = __ffma2_rn(hx, e, hx);
  a = r.x; b = r.y;
#endif
}

template <int ACT>
__device__ __forceinline__ void act_apply32_ct(float (&y)[32]) {
  if constexpr (ACT == ACT_SILU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) y[j] = act_silu(y[j]);
  } else if constexpr (ACT == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.f);
  } else if constexpr (ACT == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) act_gelu2(y[j], y[j + 1]);
  }
}

__device__ __forceinline__ void modln_stats(const EpiArgs& e, int rt, int r, float& mean, float& rstd) {
  float sn = 0.f, m = 0.f, m2 = 0.f;
  for (int p = 0; p < e.stats_nt; ++p) {
    float2 s = e.stats_in[((size_t)rt * e.stats_nt + p) * TILE_M + r];
    stats_merge(sn, m, m2, (float)min(TILE_N, e.h_dim - p * TILE_N), s.x, s.y);
  }
  mean = m;
  rstd = rsqrtf(m2 / (float)e.h_dim + 1e-5f);
}

template <int EPI>
__device__ __forceinline__ void epi_first(const EpiArgs& e, int rt, int nt, int r, EpiState<EPI>& st) {
  st.bias = 0.f;
  if constexpr (EPI != EPI_SCORE) st.bias = e.bias ? __ldg(e.bias + nt * TILE_N + r) : 0.f;
  if constexpr (EPI == EPI_F32) {
    if (e.resid_tiled) {
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int q = 0; q < 8; ++q)
          st.res[c][q] =
              e.resid_tiled[((size_t)rt * e.ld4 + ((nt * TILE_N + c * 32) >> 2) + q) * TILE_M + r];
    }
  } else if constexpr (EPI == EPI_MODLN) {
#pragma unroll
    for (int q = 0; q < 16; ++q)
      st.h[q] = e.h_tiled[((size_t)rt * e.h_ld4 + nt * 16 + q) * TILE_M + r];
    modln_stats(e, rt, r, st.mean, st.rstd);
    st.rt = rt;
  }
}

// Stage this tile's bias in shared memory and start the load of the next tile's.
template <int EPI>
__device__ __forceinline__ void epi_stage_bias(const EpiArgs& e, float* sb, int bar_id, int r,
                                               EpiState<EPI>& st, bool has_next, int nt2) {
  if constexpr (EPI != EPI_SCORE) {
    group_bar(bar_id);  // every reader of the previous tile's stage is done
    sb[r] = st.bias;
    st.bias = (has_next && e.bias) ? __ldg(e.bias + nt2 * TILE_N + r) : 0.f;
    group_bar(bar_id);  // stage visible to the whole group
  }
}

// Early release of the accumulator slot: once the LAST tcgen05.ld of a tile has completed the
// values are in registers and the MMA warp may overwrite the slot, while this warp still does the
// math and the stores of the last chunk.  bar = shared::cta address of acc_empty[slot] (remote = 0)
// or its shared::cluster address in the pair's leader CTA (remote = 1, relaxed arrive: the reads
// are complete and nothing this warp wrote is consumed through the barrier).
struct AccRelease {
  uint32_t bar;
  int remote;
};
__device__ __forceinline__ void acc_release(const AccRelease& rel) {
  tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    if (rel.remote)
      asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(rel.bar) : "memory");
    else
      mbar_arrive(rel.bar);
  }
}

template <int EPI, int ACT, int SMODE = 0>
__device__ __forceinline__ void epi_finish(const EpiArgs& e, uint32_t tmem_tile, int rt, int nt,
                                           int n_tiles, int r, const float* sb, EpiState<EPI>& st,
                                           bool has_next, int rt2, int nt2, const AccRelease rel,
                                           const Stage stg = Stage{0, 0, false}) {
  const int row = rt * TILE_M + r;

  if constexpr (EPI == EPI_PACK) {
    uint32_t raw[32];
    tmem_ld32(tmem_tile, raw);
#pragma unroll 1
    for (int cc = 0; cc < 4; cc += 2) {
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        const int c = cc + hb;
        const int n0 = nt * TILE_N + c * 32;
        float y[32];
        bias32_from_smem(sb, c * 32, y);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 2)
          add2(y[j], y[j + 1], __uint_as_float(raw[j]), __uint_as_float(raw[j + 1]));
        if (c + 1 < 4) tmem_ld32(tmem_tile + (c + 1) * 32, raw);   // in flight during this chunk's math
        else acc_release(rel);                                      // whole tile is in registers
        if (!(e.debug & 64)) act_apply32_ct<ACT>(y);
        if (n0 + 32 > e.n_valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = (n0 + j < e.n_valid) ? y[j] : 0.f;
        }
        const int kb_out = n0 >> 6;
        if (kb_out < e.out_kb && !(e.debug & 32)) {   // uniform over the group
          __nv_bfloat16* tile = e.out_packed + (size_t)(rt * e.out_kb + kb_out) * TILE_ELEMS;
          if constexpr (SMODE == 1) {
            if (hb == 0) stage_acquire(stg);
#pragma unroll
            for (int q = 0; q < 4; ++q) sts_packed8(stg.base, r, hb * 4 + q, y + q * 8);
            if (hb == 1) stage_flush(stg, tile);
          } else {
            store_packed32(tile, r, n0 & 63, y);
          }
        }
      }
    }
  } else if constexpr (EPI == EPI_F32) {
    float sn = 0.f, smean = 0.f, sm2 = 0.f;
    uint32_t raw[32];
    tmem_ld32(tmem_tile, raw);
#pragma unroll 1
    for (int cc = 0; cc < 4; cc += 2) {
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        const int c = cc + hb;
        const int n0 = nt * TILE_N + c * 32;
        float y[32];
        bias32_from_smem(sb, c * 32, y);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) y[j] += __uint_as_float(raw[j]);   // scalar: packed pairs spill here
        if (c + 1 < 4) tmem_ld32(tmem_tile + (c + 1) * 32, raw);   // in flight during this chunk's math
        else acc_release(rel);                                      // whole tile is in registers
        if (e.resid_tiled && !(e.debug & 128)) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 hv = st.res[hb][q];
            y[q * 4 + 0] += hv.x; y[q * 4 + 1] += hv.y; y[q * 4 + 2] += hv.z; y[q * 4 + 3] += hv.w;
          }
          // refill this buffer: chunk c+2 of this tile, or chunk c-2 of the group's next tile
          if (c < 2) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              st.res[hb][q] = e.resid_tiled[((size_t)rt * e.ld4 + ((n0 + 64) >> 2) + q) * TILE_M + r];
          } else if (has_next) {
            const int m0 = nt2 * TILE_N + (c - 2) * 32;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              st.res[hb][q] = e.resid_tiled[((size_t)rt2 * e.ld4 + (m0 >> 2) + q) * TILE_M + r];
          }
        }
        act_apply32_ct<ACT>(y);
        if (e.out_tiled && !(e.debug & 32)) {         // uniform over the group
          if constexpr (SMODE == 1) {   // 32 columns = 8 float4 columns x 128 rows: one contiguous 16 KiB block
            stage_acquire(stg);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              sts_v4(stg.base + q * (TILE_M * 16) + r * 16, __float_as_uint(y[q * 4 + 0]), __float_as_uint(y[q * 4 + 1]),
                     __float_as_uint(y[q * 4 + 2]), __float_as_uint(y[q * 4 + 3]));
            stage_flush(stg, e.out_tiled + ((size_t)rt * e.ld4 + (n0 >> 2)) * TILE_M);
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              e.out_tiled[((size_t)rt * e.ld4 + (n0 >> 2) + q) * TILE_M + r] =
                  make_float4(y[q * 4 + 0], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
          }
        }
        if (e.out_packed && (n0 >> 6) < e.out_kb) {
          float yp[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) yp[j] = (n0 + j < e.n_valid) ? y[j] : 0.f;
          __nv_bfloat16* tile = e.out_packed + (size_t)(rt * e.out_kb + (n0 >> 6)) * TILE_ELEMS;
          store_packed32(tile, r, n0 & 63, yp);
        }
        if (e.out_rm && (rt % e.split_rt) * TILE_M + r < e.rows_valid) {
          // split-K: rt = split * split_rt + row tile; every split owns a block of split_rt*128 rows
          float* orow = e.out_rm + (size_t)row * e.ld_rm + n0;
          if (n0 + 32 <= e.n_valid && (e.ld_rm & 3) == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(orow + q * 4) =
                  make_float4(y[q * 4 + 0], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < e.n_valid) orow[j] = y[j];
          }
        }
        if (e.stats_out) {
          const int nv = min(32, max(0, e.n_valid - n0));
          if (nv == 32) {                       // full chunk (every hot layer): no per-element masks
            float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 32; j += 2) s2 = __fadd2_rn(s2, make_float2(y[j], y[j + 1]));
            const float m = (s2.x + s2.y) * (1.0f / 32.0f);
            const float2 nm = make_float2(-m, -m);
            float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float2 d = __fadd2_rn(make_float2(y[j], y[j + 1]), nm);
              q2 = __ffma2_rn(d, d, q2);
            }
            stats_merge(sn, smean, sm2, 32.f, m, q2.x + q2.y);
          } else if (nv > 0) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) s += (j < nv) ? y[j] : 0.f;
            float m = s / (float)nv;
            float q2 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float d = (j < nv) ? (y[j] - m) : 0.f;
              q2 += d * d;
            }
            stats_merge(sn, smean, sm2, (float)nv, m, q2);
          }
        }
      }
    }
    if (e.stats_out) e.stats_out[((size_t)rt * n_tiles + nt) * TILE_M + r] = make_float2(smean, sm2);
  } else if constexpr (EPI == EPI_MODLN) {
    // y = LN(h) * (1 + scale) + shift with hn = h*rstd - mean*rstd; the packed bias of the scale
    // rows already holds 1 + b (k_pack_bias), so one output costs 2 FADD + 2 FFMA.
    const float rstd = st.rstd, nmr = -st.mean * st.rstd;
    __nv_bfloat16* tile = e.out_packed + (size_t)(rt * e.out_kb + nt) * TILE_ELEMS;
    // 8 chunks of 8 hidden columns, TMEM loads double-buffered one chunk ahead: the math of a chunk
    // is short (8 packed instructions), so a single-buffered load -> wait -> math chain left the
    // warps waiting on tcgen05.ld most of the time.
    uint32_t rs[2][8], rh[2][8];
    tmem_ld8(tmem_tile, rs[0]);                   // scale cols
    tmem_ld8(tmem_tile + 64, rh[0]);              // shift cols
    // LayerNorm statistics change with the row tile only; when the next tile starts a new row tile
    // its partials are fetched NOW and merged after this tile's work (ncu: the un-prefetched loads
    // were ~18 % of the adaLN epilogue's stall samples).
    const bool new_stats = has_next && rt2 != st.rt;
    float2 sp[4];
    if (new_stats) {
#pragma unroll
      for (int p = 0; p < 4; ++p)
        sp[p] = (p < e.stats_nt) ? e.stats_in[((size_t)rt2 * e.stats_nt + p) * TILE_M + r] : make_float2(0.f, 0.f);
    }
    if constexpr (SMODE == 1) {
      if (!(e.debug & 32)) stage_acquire(stg);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {                 // 8 hidden columns = one 16-byte output chunk
      const float4* bsp = reinterpret_cast<const float4*>(sb + c * 8);        // 1 + scale biases
      const float4* bhp = reinterpret_cast<const float4*>(sb + 64 + c * 8);   // shift biases
      float y[8];
      tmem_ld_wait();                             // chunk c (and nothing else) is outstanding
      if (c + 1 < 8) {
        tmem_ld8(tmem_tile + (c + 1) * 8, rs[(c + 1) & 1]);
        tmem_ld8(tmem_tile + 64 + (c + 1) * 8, rh[(c + 1) & 1]);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 hv = st.h[c * 2 + q];
        const float4 b1 = bsp[q], b2 = bhp[q];
#pragma unroll
        for (int i = 0; i < 2; ++i) {             // two hidden columns per packed instruction
          const int j = q * 4 + i * 2;
          const float2 hx = i ? make_float2(hv.z, hv.w) : make_float2(hv.x, hv.y);
          const float2 bs = i ? make_float2(b1.z, b1.w) : make_float2(b1.x, b1.y);
          const float2 bh = i ? make_float2(b2.z, b2.w) : make_float2(b2.x, b2.y);
          const float2 scale1 = __fadd2_rn(make_float2(__uint_as_float(rs[c & 1][j]), __uint_as_float(rs[c & 1][j + 1])), bs);
          const float2 shift = __fadd2_rn(make_float2(__uint_as_float(rh[c & 1][j]), __uint_as_float(rh[c & 1][j + 1])), bh);
          const float2 hn = __ffma2_rn(hx, make_float2(rstd, rstd), make_float2(nmr, nmr));
          const float2 o = __ffma2_rn(hn, scale1, shift);
          y[j] = o.x; y[j + 1] = o.y;
        }
      }
      if (c == 6) {
        // chunk 7's loads were issued above; wait for them here so the slot can be released one
        // chunk early (its values are then in registers)
        tmem_ld_wait();
        acc_release(rel);
      }
      if (has_next && !(e.debug & 128)) {         // refill the consumed registers for the next tile
#pragma unroll
        for (int q = 0; q < 2; ++q)
          st.h[c * 2 + q] = e.h_tiled[((size_t)rt2 * e.h_ld4 + nt2 * 16 + c * 2 + q) * TILE_M + r];
      }
      if (!(e.debug & 32)) {
        if constexpr (SMODE != 0) {
          sts_packed8(stg.base, r, c, y);
        } else {
          uint4 v;
          v.x = pack_op16x2(y[0], y[1]);
          v.y = pack_op16x2(y[2], y[3]);
          v.z = pack_op16x2(y[4], y[5]);
          v.w = pack_op16x2(y[6], y[7]);
          *reinterpret_cast<uint4*>(tile + c * (TILE_M * 8) + r * 8) = v;
        }
      }
    }
    if constexpr (SMODE == 1) {
      if (!(e.debug & 32)) stage_flush(stg, tile);
    }
    if (new_stats) {
      float sn = 0.f, m = 0.f, m2 = 0.f;
#pragma unroll
      for (int p = 0; p < 4; ++p)
        if (p < e.stats_nt) stats_merge(sn, m, m2, (float)min(TILE_N, e.h_dim - p * TILE_N), sp[p].x, sp[p].y);
      for (int p = 4; p < e.stats_nt; ++p) {      // hidden widths beyond 512: not prefetched
        const float2 s2 = e.stats_in[((size_t)rt2 * e.stats_nt + p) * TILE_M + r];
        stats_merge(sn, m, m2, (float)min(TILE_N, e.h_dim - p * TILE_N), s2.x, s2.y);
      }
      st.mean = m;
      st.rstd = rsqrtf(m2 / (float)e.h_dim + 1e-5f);
      st.rt = rt2;
    }
  } else if constexpr (EPI == EPI_DACT) {
    // Training epilogues (train.inc).  One thread = one row.  Per 32-column accumulator chunk the saved
    // pre-activation (8 float4) and the auxiliary packed operand (4 x 16 B) come from HBM: they are
    // fetched one chunk AHEAD into a second register set (the first version loaded them right before
    // use and the backward modes ran 2.5x slower than the forward mode: exposed load latency with only
    // eight epilogue warps per SM).  The math consumes a chunk in four 8-column pieces (= one 16-byte
    // packed store each); activation derivatives are evaluated on register pairs (dact_eval2).
    const int srt = rt % e.src_rt;                       // row tile of the saved forward tensors
    const bool first_half = rt < e.src_rt;
    const float aux_scale = e.aux_scale ? __ldg(e.aux_scale) : 1.0f;
    const int mode = e.dact_mode;
    const bool use_aux = mode == DACT_HAT || (mode == DACT_BWD && first_half && e.aux_in != nullptr);
    const bool rd_pre = mode != DACT_FWD;
    float4 pre[2][8];
    uint4 aux[2][4];
    auto fetch = [&](int c, int buf) {
      const int n0 = nt * TILE_N + c * 32;
      if ((n0 >> 6) >= e.out_kb) return;
      if (rd_pre) {
        const float4* pp = e.pre_tiled + ((size_t)srt * e.pre_ld4 + (n0 >> 2)) * TILE_M + r;
#pragma unroll
        for (int q = 0; q < 8; ++q) pre[buf][q] = pp[(size_t)q * TILE_M];
      }
      if (use_aux) {
        const __nv_bfloat16* at = e.aux_in + (size_t)(srt * e.out_kb + (n0 >> 6)) * TILE_ELEMS + (size_t)((n0 & 63) >> 3) * (TILE_M * 8) + r * 8;
#pragma unroll
        for (int q = 0; q < 4; ++q) aux[buf][q] = *reinterpret_cast<const uint4*>(at + (size_t)q * (TILE_M * 8));
      }
    };
    fetch(0, 0);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int n0 = nt * TILE_N + c * 32;
      const int kb_out = n0 >> 6;
      const bool in_range = kb_out < e.out_kb;           // uniform over the group
      uint32_t raw[32];
      tmem_ld32(tmem_tile + c * 32, raw);
      if (c + 1 < 4) fetch(c + 1, (c + 1) & 1);          // next chunk's operands fly during this chunk's math
      float b32[32];
      if (mode == DACT_FWD) bias32_from_smem(sb, c * 32, b32);
      tmem_ld_wait();
      if (c == 3) acc_release(rel);                      // the whole tile is in registers
      if (!in_range) continue;
      __nv_bfloat16* tile = e.out_packed + (size_t)(rt * e.out_kb + kb_out) * TILE_ELEMS;
      const int chunk0 = (n0 & 63) >> 3;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = n0 + q * 8;
        float y[8], o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = __uint_as_float(raw[q * 8 + i]);
        const size_t packed_at = (size_t)(chunk0 + q) * (TILE_M * 8) + r * 8;
        if (mode == DACT_FWD) {
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = (col + i < e.n_valid) ? y[i] + b32[q * 8 + i] : 0.f;
          float4* pp = e.pre_tiled + ((size_t)rt * e.pre_ld4 + (col >> 2)) * TILE_M + r;
          pp[0] = make_float4(y[0], y[1], y[2], y[3]);
          pp[TILE_M] = make_float4(y[4], y[5], y[6], y[7]);
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            float2 f, d1, d2;
            dact_eval2<ACT>(make_float2(y[i], y[i + 1]), f, d1, d2);
            o[i] = f.x; o[i + 1] = f.y;
          }
        } else {
          const float4 p0 = pre[c & 1][q * 2], p1 = pre[c & 1][q * 2 + 1];
          const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
          float ax[8];
          if (use_aux) unpack_op16x8(aux[c & 1][q], ax);
          if (mode == DACT_VJP && e.aux_out) {
            uint4 v;
            v.x = pack_op16x2(y[0], y[1]); v.y = pack_op16x2(y[2], y[3]);
            v.z = pack_op16x2(y[4], y[5]); v.w = pack_op16x2(y[6], y[7]);
            *reinterpret_cast<uint4*>(e.aux_out + (size_t)(rt * e.out_kb + kb_out) * TILE_ELEMS + packed_at) = v;
          }
          float g2[8];
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            float2 f, d1, d2;
            dact_eval2<ACT>(make_float2(pv[i], pv[i + 1]), f, d1, d2);
            const float2 acc = make_float2(y[i], y[i + 1]);
            float2 out = __fmul2_rn(acc, d1);
            if (mode == DACT_HAT) {
              const float2 t2 = __fmul2_rn(__fmul2_rn(acc, make_float2(ax[i], ax[i + 1])),
                                           __fmul2_rn(d2, make_float2(aux_scale, aux_scale)));
              g2[i] = (col + i < e.n_valid) ? t2.x : 0.f;
              g2[i + 1] = (col + i + 1 < e.n_valid) ? t2.y : 0.f;
            } else if (mode == DACT_BWD && use_aux) {
              out = __fadd2_rn(out, make_float2(ax[i], ax[i + 1]));
            }
            o[i] = (col + i < e.n_valid) ? out.x : 0.f;
            o[i + 1] = (col + i + 1 < e.n_valid) ? out.y : 0.f;
          }
          if (mode == DACT_HAT) {
            uint4 v;
            v.x = pack_op16x2(g2[0], g2[1]); v.y = pack_op16x2(g2[2], g2[3]);
            v.z = pack_op16x2(g2[4], g2[5]); v.w = pack_op16x2(g2[6], g2[7]);
            *reinterpret_cast<uint4*>(e.aux_out + (size_t)(rt * e.out_kb + kb_out) * TILE_ELEMS + packed_at) = v;
          }
        }
        uint4 v;
        v.x = pack_op16x2(o[0], o[1]); v.y = pack_op16x2(o[2], o[3]);
        v.z = pack_op16x2(o[4], o[5]); v.w = pack_op16x2(o[6], o[7]);
        *reinterpret_cast<uint4*>(tile + packed_at) = v;
      }
    }
  } else {  // EPI_SCORE
    uint32_t raw[32];
    const float mult = __ldg(e.out_mult);
    const float tw = e.tw_rows ? ((row < e.rows_valid) ? __ldg(e.tw_rows + row) : 0.f) : e.tw_scalar;
    PhiloxState ps{0ull, 0ull};
    if (e.philox) ps = *e.philox;
    const bool has_eps = e.eps != nullptr || e.philox != nullptr;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const int n0 = nt * TILE_N + c * 32;
      if (n0 >= e.n_valid && !e.out_packed) break;
      tmem_ld32(tmem_tile + c * 32, raw);
      const bool live = row < e.rows_valid;
      // the step's operands do not depend on the accumulator: load them under the TMEM latency
      float4 zv[8], ev[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int col = n0 + q * 4;
        zv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        ev[q] = zv[q];
        if (e.do_step && live && (col + 3 < e.n_valid)) {
          zv[q] = *reinterpret_cast<const float4*>(e.z_in + (size_t)row * e.n_valid + col);
          if (e.eps) ev[q] = *reinterpret_cast<const float4*>(e.eps + (size_t)row * e.n_valid + col);
          else if (e.philox)
            ev[q] = philox_normal4(ps, e.philox_draw, (unsigned long long)(e.row_offset + row), (uint32_t)(col >> 2));
        }
      }
      tmem_ld_wait();
      if (e.r_out && live) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int col = n0 + q * 4;
          if (col + 3 < e.n_valid)
            *reinterpret_cast<float4*>(e.r_out + (size_t)row * e.n_valid + col) =
                make_float4(__uint_as_float(raw[q * 4 + 0]), __uint_as_float(raw[q * 4 + 1]),
                            __uint_as_float(raw[q * 4 + 2]), __uint_as_float(raw[q * 4 + 3]));
        }
      }
      float y[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float s = fminf(fmaxf(__uint_as_float(raw[j]), -10.f), 10.f);
        s = __fmul_rn(s, mult);
        if (e.tw_rows || e.tw_scalar != 1.0f) s = __fmul_rn(s, tw);
        y[j] = s;
      }
      if (e.do_step) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int col = n0 + q * 4;
          const bool ok = live && (col + 3 < e.n_valid);
          float zz[4] = {zv[q].x, zv[q].y, zv[q].z, zv[q].w}, ee[4] = {ev[q].x, ev[q].y, ev[q].z, ev[q].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            // (z + s1*score) * ra ; c1*pred + c2*z ; + sigma*eps   (rounding order of the reference)
            float pred = __fmul_rn(__fadd_rn(zz[i], __fmul_rn(e.c_s1, y[q * 4 + i])), e.c_ra);
            float mu = __fadd_rn(__fmul_rn(e.c_c1, pred), __fmul_rn(e.c_c2, zz[i]));
            if (has_eps) mu = __fadd_rn(mu, __fmul_rn(e.c_sigma, ee[i]));
            y[q * 4 + i] = ok ? mu : 0.f;
          }
          if (ok)
            *reinterpret_cast<float4*>(e.z_out + (size_t)row * e.n_valid + col) =
                make_float4(y[q * 4 + 0], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
        }
        if (e.out_packed && (n0 >> 6) < e.out_kb) {
          __nv_bfloat16* tile = e.out_packed + (size_t)(rt * e.out_kb + (n0 >> 6)) * TILE_ELEMS;
          store_packed32(tile, r, n0 & 63, y);
        }
      } else if (live) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int col = n0 + q * 4;
          if (col + 3 < e.n_valid)
            *reinterpret_cast<float4*>(e.z_out + (size_t)row * e.n_valid + col) =
                make_float4(y[q * 4 + 0], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
        }
      }
    }
    acc_release(rel);
  }
}
