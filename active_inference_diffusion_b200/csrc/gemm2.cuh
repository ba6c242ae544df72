// CTA-pair (cta_group::2) variant of the persistent tcgen05 GEMM.
//
// Why: with one CTA per 128-row tile every SM streams the FULL weight matrix of a layer through
// its L2->shared-memory port: 64 B/clk/SM at the tensor-pipe rate, 10-15 TB/s over the chip.
// Measured on B200 (scripts/perf_kernels.py) every hot GEMM of the score network ran at
// (weight + activation bytes) / ~10.4 TB/s whether or not its epilogue did anything -- the L2
// fabric, not the tensor pipe, set the time.  Pairing two SMs on one 256-row x 256-col UMMA
// (tcgen05.mma.cta_group::2) lets each CTA fetch only HALF of every weight tile (its 128 of the
// 256 weight rows); the pair's tensor cores read both halves.  Weight traffic per SM halves.
//
// Layout per CTA (rank c of the pair):
//   A  : its own 128-row activation tile (row tile 2*rp + c), resident (K <= 512) or streamed.
//   B  : 128 x 64 half tiles (n-tile 2*ng + c of the 256-column unit) in a ring of stages.
//   D  : TMEM accumulators of its own 128 rows x 256 columns, two units in flight.
// Protocol (all mbarriers in each CTA's control block at identical offsets):
//   ring_full[s]          : the CTA's own bulk copies (complete_tx); the LEADER's copy additionally
//                           takes a remote arrive from rank 1's "forwarder" thread (which waits on
//                           rank 1's a_full/ring_full), so one wait tells the leader's MMA warp that
//                           both halves (and rank 1's A k-block) are in shared memory.
//   ring_empty[s], a_empty[kb], acc_full[slot] : signalled in BOTH CTAs by the leader's
//                           tcgen05.commit ... multicast::cluster (mask 0b11).
//   acc_empty[slot] leader : 8 arrivals = 4 epilogue warps of each CTA (rank 1 arrives remotely).
// Only rank 0 issues MMAs.  Epilogues are the per-CTA ones of epilogue.cuh, unchanged.
#pragma once
#include "gemm.cuh"

namespace aid {

constexpr int MAX_RING2 = 12;

struct alignas(8) Gemm2Ctrl {
  uint64_t ring_full[MAX_RING2];
  uint64_t ring_empty[MAX_RING2];
  uint64_t peer_full[MAX_RING2];
  uint64_t a_full[MAX_RES_KB];
  uint64_t a_empty[MAX_RES_KB];
  uint64_t acc_full[4];
  uint64_t acc_empty[4];
  uint32_t tmem_base;
  uint32_t pad_[(1024 - (3 * MAX_RING2 + 2 * MAX_RES_KB + 8) * 8 - 4) / 4];
  float bias_stage[2][TILE_N];   // per epilogue group; 1024-byte offset
  uint8_t pad2_[SMEM_CTRL - 1024 - 2 * TILE_N * 4];
};
static_assert(sizeof(Gemm2Ctrl) == SMEM_CTRL, "control block must be exactly SMEM_CTRL bytes");

// ---- cluster / cta_group::2 PTX -----------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t local_bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// Same, without the release fence: for the forwarder, whose signal only relays the completion of
// TMA writes (async proxy, already complete when its local barrier flipped) -- the thread itself
// wrote nothing that the leader reads.  The release form costs a cluster-scope fence per stage and
// made the forwarder loop, not the tensor pipe, set the pace of the pair kernel.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  return remote;
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem, 256 rows over the pair] (+)= A * B^T, issued by the leader CTA only
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (same offset in both CTAs of the pair) once all tcgen05 ops issued so far completed
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(mask)
      : "memory");
}

// RES : A row tile resident in shared memory (kb <= MAX_RES_KB) vs streamed with the B halves.
// A "pair unit" = (row-tile pair rp, 256-column group ng); rank c owns row tile 2*rp + c.
// Shared memory reserved for staging output tiles (one 16 KiB buffer per epilogue group) when the
// build defines AID_STAGE_STORES.  Off by default: measured on B200 it is neutral for the adaLN and
// attention kernels and slower for the MLP kernels (mlp.0 114 -> 125 us), because the two buffers
// cost two of the six operand stages.  The GPU suite passes with it on.
// EPI_LNACT keeps bias / gamma / beta of the 512 columns (6 KiB) and the double-buffered exchange of
// the two epilogue groups' LayerNorm partials (4 KiB) in the same region.
constexpr int LN_COLS = 4 * TILE_N;
constexpr int LN_SMEM_BYTES = 3 * LN_COLS * 4 + 2 * 2 * TILE_M * 8;
__host__ __device__ constexpr int gemm2_stage_bytes(int epi) {
  if (epi == EPI_LNACT) return LN_SMEM_BYTES;
#ifdef AID_STAGE_STORES
  return (epi == EPI_SCORE) ? 0 : 2 * TILE_BYTES;
#else
  (void)epi;
  return 0;
#endif
}

template <int EPI, bool RES, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_kernel(const GemmArgs ga, const EpiArgs ea, const int ring_stages) {
  constexpr int TU = 2;                                       // 128-col tiles per unit
  constexpr int STAGE_BYTES2 = (RES ? 1 : 2) * TILE_BYTES;    // [A k-block |] B half tile
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;          // same offset in both CTAs
  uint8_t* smem = smem_raw + (base - raw_addr);
  Gemm2Ctrl* ctrl = reinterpret_cast<Gemm2Ctrl*>(smem);
  constexpr int STAGE_OUT = gemm2_stage_bytes(EPI);
  constexpr bool STAGED = STAGE_OUT != 0 && EPI != EPI_LNACT;
  const uint32_t out_stage = base + SMEM_CTRL;                 // 2 x 16 KiB (one per epilogue group)
  const uint32_t a_smem = out_stage + STAGE_OUT;
  const uint32_t ring_smem = a_smem + (RES ? ga.kb * TILE_BYTES : 0);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int groups = ga.n_tiles / TU;                         // units per row-tile pair
  const int rps = (ga.row_tiles + 1) >> 1;
  const int num_units = rps * groups;
  // Resident A: a contiguous range of units per pair, so the resident row tiles serve all their
  // column groups.  Streamed A: units are dealt round-robin, so the column groups of one row pair run
  // at the same time on neighbouring pairs and the second read of the A tiles hits L2 (with contiguous
  // ranges ncu showed mlp.2 reading 620 MB from DRAM for 404 MB of operands: the pair came back to
  // the same 1 MB of A ~35 us later, after ~200 MB of other traffic had gone through the 126 MB L2).
  // EPI_LNACT normalises whole rows: a pair always takes BOTH column groups of a row-tile pair.
  const int u_lo = EPI == EPI_LNACT ? groups * (int)((long long)pair * rps / num_pairs)
                                    : (int)((long long)pair * num_units / num_pairs);
  const int u_hi = EPI == EPI_LNACT ? groups * (int)((long long)(pair + 1) * rps / num_pairs)
                                    : (int)((long long)(pair + 1) * num_units / num_pairs);
  const int u_begin = 0;
  const int u_end = RES ? u_hi - u_lo : (num_units - pair + num_pairs - 1) / num_pairs;

  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_RING2; ++i) {
      // leader: own producer's expect_tx arrive + rank 1's forwarder (one wait per stage for the MMA warp)
      mbar_init(smem_u32(&ctrl->ring_full[i]), rank == 0 ? 2 : 1);
      mbar_init(smem_u32(&ctrl->ring_empty[i]), 1);
      mbar_init(smem_u32(&ctrl->peer_full[i]), 1);
    }
    for (int i = 0; i < MAX_RES_KB; ++i) {
      mbar_init(smem_u32(&ctrl->a_full[i]), 1);
      mbar_init(smem_u32(&ctrl->a_empty[i]), 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(smem_u32(&ctrl->acc_full[i]), 1);
      mbar_init(smem_u32(&ctrl->acc_empty[i]), 8);  // 4 epilogue warps of each CTA
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc2(smem_u32(&ctrl->tmem_base), 512);
    tmem_relinquish2();
  }
  unsigned long long* conv_tab = reinterpret_cast<unsigned long long*>(
      smem + SMEM_CTRL + STAGE_OUT + (RES ? ga.kb * TILE_BYTES : 0) + ring_stages * STAGE_BYTES2);
  if (ga.conv.cin8) conv_fill_table(ga.conv, conv_tab, ga.kb * 8);
  pdl_launch_dependents();                                    // see gemm.cuh: prologue overlaps the previous kernel's tail
  tc_fence_before();
  cluster_sync_all();                                         // barriers of BOTH CTAs are live
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;
  pdl_wait();                                                 // nothing above touched global memory

  auto unit_coords = [&](int i, int& rp, int& ng) {   // i-th unit of this pair
    const int u = RES ? u_lo + i : pair + i * num_pairs;
    const int ue = ga.reverse ? num_units - 1 - u : u;
    rp = ue / groups;
    ng = ue % groups;
  };

  if (warp == 0) {
    // ===================== producer (both CTAs) =====================
    if (!RES && ga.conv.cin8) {
      // implicit-im2col mode, warp-wide (see gemm.cuh): lanes 0-7 one A chunk each, lane 8 the weights
      const uint64_t conv_policy = policy_evict_last();   // halo rows are re-read by the other eight taps
      int stage = 0;
      uint32_t phase = 0;
      for (int u = u_begin; u < u_end; ++u) {
        int rp, ng;
        unit_coords(u, rp, ng);
        const int rt = min(2 * rp + (int)rank, ga.row_tiles - 1);
        const int nt = ng * TU + (int)rank;
        for (int kb = 0; kb < ga.kb; ++kb) {
          const uint32_t fb = smem_u32(&ctrl->ring_full[stage]);
          if (lane == 0) {
            mbar_wait(smem_u32(&ctrl->ring_empty[stage]), phase ^ 1, ga.err, 2);
            mbar_arrive_expect_tx(fb, STAGE_BYTES2);
          }
          __syncwarp();
          const uint32_t dst = ring_smem + stage * STAGE_BYTES2;
          if (lane < 8) conv_load_chunk(conv_tab, dst, rt, kb, lane, fb, conv_policy);
          else if (lane == 8) bulk_g2s(dst + TILE_BYTES, ga.B + ((size_t)nt * ga.kb + kb) * TILE_BYTES, TILE_BYTES, fb);
          if (++stage == ring_stages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, a_par = 0;
      int prev_rp = -1;
      const uint64_t conv_policy = policy_evict_last();   // halo rows are re-read by the other eight taps
      for (int u = u_begin; u < u_end; ++u) {
        int rp, ng;
        unit_coords(u, rp, ng);
        const int rt = min(2 * rp + (int)rank, ga.row_tiles - 1);   // odd tail: reload a valid tile
        const int nt = ng * TU + (int)rank;                         // my half of the weight rows
        const bool new_rp = RES && (rp != prev_rp);
        for (int kb = 0; kb < ga.kb; ++kb) {
          if (RES && new_rp) {
            const uint32_t fb = smem_u32(&ctrl->a_full[kb]);
            mbar_wait(smem_u32(&ctrl->a_empty[kb]), a_par ^ 1, ga.err, 1);
            mbar_arrive_expect_tx(fb, TILE_BYTES);
            bulk_g2s(a_smem + kb * TILE_BYTES, ga.A + ((size_t)rt * ga.kb + kb) * TILE_BYTES,
                     TILE_BYTES, fb);
          }
          const uint32_t fb = smem_u32(&ctrl->ring_full[stage]);
          mbar_wait(smem_u32(&ctrl->ring_empty[stage]), phase ^ 1, ga.err, 2);
          if (ga.debug & 512) {
            mbar_arrive(fb);
          } else {
            mbar_arrive_expect_tx(fb, STAGE_BYTES2);
            uint32_t dst = ring_smem + stage * STAGE_BYTES2;
            if (!RES) {
              if (ga.conv.cin8)
                conv_load_a(ga.conv, conv_tab, dst, rt, kb, fb, conv_policy);
              else
                bulk_g2s(dst, ga.A + ((size_t)rt * ga.kb + kb) * TILE_BYTES, TILE_BYTES, fb);
              dst += TILE_BYTES;
            }
            bulk_g2s(dst, ga.B + ((size_t)nt * ga.kb + kb) * TILE_BYTES, TILE_BYTES, fb);
          }
          if (++stage == ring_stages) { stage = 0; phase ^= 1; }
        }
        if (new_rp) { a_par ^= 1; prev_rp = rp; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 1) {
      // ===================== forwarder (rank 1): my operands are in -> tell the leader ======
      const uint32_t leader_full = map_to_rank(smem_u32(&ctrl->ring_full[0]), 0);
      int stage = 0;
      uint32_t phase = 0, a_par = 0;
      int prev_rp = -1;
      for (int u = u_begin; u < u_end; ++u) {
        int rp, ng;
        unit_coords(u, rp, ng);
        const bool new_rp = RES && (rp != prev_rp);
        for (int kb = 0; kb < ga.kb; ++kb) {
          if (RES && new_rp) mbar_wait(smem_u32(&ctrl->a_full[kb]), a_par, ga.err, 9);
          mbar_wait(smem_u32(&ctrl->ring_full[stage]), phase, ga.err, 10);
          mbar_arrive_cluster_relaxed(leader_full + stage * 8);
          if (++stage == ring_stages) { stage = 0; phase ^= 1; }
        }
        if (new_rp) { a_par ^= 1; prev_rp = rp; }
      }
    } else if (rank == 0) {
      // ===================== MMA issuer (leader): whole warp, one elected lane issues ==========
      constexpr uint32_t idesc = umma_idesc_bf16(2 * TILE_M, 2 * TILE_N);
      int stage = 0;
      uint32_t phase = 0, a_par = 0;
      int prev_rp = -1;
      int q = 0;  // 128-col tile sequence number (multiple of TU at unit start)
      for (int u = u_begin; u < u_end; ++u) {
        int rp, ng;
        unit_coords(u, rp, ng);
        const bool new_rp = RES && (rp != prev_rp);
        bool last_of_rp = false;
        if (RES) {
          last_of_rp = true;
          if (u + 1 < u_end) {
            int rp2, ng2;
            unit_coords(u + 1, rp2, ng2);
            last_of_rp = rp2 != rp;
          }
        }
#pragma unroll
        for (int t = 0; t < TU; ++t)   // this unit's accumulator slots drained by both epilogues
          mbar_wait(smem_u32(&ctrl->acc_empty[(q + t) & 3]), (((q + t) >> 2) & 1) ^ 1, ga.err, 4);
        const uint32_t d = tmem_base + (uint32_t)((q & 3) * TILE_N);
        for (int kb = 0; kb < ga.kb; ++kb) {
          if (RES && new_rp) mbar_wait(smem_u32(&ctrl->a_full[kb]), a_par, ga.err, 5);
          mbar_wait(smem_u32(&ctrl->ring_full[stage]), phase, ga.err, 7);   // both halves are in
          tc_fence_after();
          const uint32_t s_base = ring_smem + stage * STAGE_BYTES2;
          const uint32_t a_tile = RES ? a_smem + kb * TILE_BYTES : s_base;
          const uint32_t b_tile = RES ? s_base : s_base + TILE_BYTES;
          const uint64_t ad = umma_desc_kmajor(a_tile, TILE_M * 16);
          const uint64_t bd = umma_desc_kmajor(b_tile, TILE_M * 16);
          if (elect_one()) {
            if (!(ga.debug & 1024))   // 1024: no MMAs: epilogue-only timing
#pragma unroll
            for (int k = 0; k < TILE_K / 16; ++k) {
              umma2_bf16(d, ad + (uint64_t)(k * (2 * TILE_M * 16) >> 4), bd + (uint64_t)(k * (2 * TILE_M * 16) >> 4),
                         idesc, (kb | k) ? 1u : 0u);
            }
            umma2_commit_both(smem_u32(&ctrl->ring_empty[stage]));
            if (last_of_rp) umma2_commit_both(smem_u32(&ctrl->a_empty[kb]));
          }
          __syncwarp();
          if (++stage == ring_stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) {
#pragma unroll
          for (int t = 0; t < TU; ++t) umma2_commit_both(smem_u32(&ctrl->acc_full[(q + t) & 3]));
        }
        __syncwarp();
        q += TU;
        if (new_rp) { a_par ^= 1; prev_rp = rp; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (two groups of 4 warps, both CTAs) =====================
    const int eg = (warp - 4) >> 2;
    const int lq = warp & 3;            // TMEM lane quadrant this warp may access
    const int r = lq * 32 + lane;
    const int n_local = (u_end - u_begin) * TU;
    float* sb = ctrl->bias_stage[eg];
    auto coords = [&](int q, int& rt, int& nt) {
      int rp, ng;
      unit_coords(u_begin + q / TU, rp, ng);
      rt = 2 * rp + (int)rank;
      nt = ng * TU + q % TU;
    };
    const uint32_t leader_acc_empty = map_to_rank(smem_u32(&ctrl->acc_empty[0]), 0);
    if constexpr (EPI == EPI_LNACT) {
      // ---- Linear -> LayerNorm -> activation (+ residual) -> packed bf16, whole rows from TMEM ----
      // n_tiles == 4: the two 256-column units of a row-tile pair are consecutive units of this
      // pair, so after their MMAs all 512 columns of this CTA's 128 rows sit in the four TMEM
      // slots.  Group eg owns tiles q = 4j + eg and 4j + 2 + eg.  Pass 1 reads both tiles and
      // accumulates (mean, M2); the groups swap partials through shared memory; pass 2 re-reads the
      // accumulators (TMEM reads are cheap: no HBM round trip of the fp32 pre-activation), normalises,
      // applies the activation and writes the next layer's packed operand.  A slot is released as
      // soon as its second read has completed, so the next row pair's MMAs overlap pass 2.
      float* s_bias = reinterpret_cast<float*>(smem + SMEM_CTRL);
      float* s_gamma = s_bias + LN_COLS;
      float* s_beta = s_gamma + LN_COLS;
      float2* xch = reinterpret_cast<float2*>(s_beta + LN_COLS);   // [parity][group][row]
      for (int i = (int)threadIdx.x - 128; i < LN_COLS; i += 256) {
        s_bias[i] = ea.bias ? __ldg(ea.bias + i) : 0.f;
        s_gamma[i] = __ldg(ea.ln_gamma + i);
        s_beta[i] = __ldg(ea.ln_beta + i);
      }
      asm volatile("bar.sync 3, 256;" ::: "memory");
      const int n_rowpairs = (u_end - u_begin) / TU;
      const float inv_n = 1.0f / (float)LN_COLS;
#pragma unroll 1
      for (int j = 0; j < n_rowpairs; ++j) {
        int rt = 0, nt[2];
        uint32_t tm[2];
        AccRelease rel[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int q = 4 * j + 2 * h + eg;
          coords(q, rt, nt[h]);
          const int buf = q & 3;
          tm[h] = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(buf * TILE_N);
          rel[h] = AccRelease{rank == 0 ? smem_u32(&ctrl->acc_empty[buf]) : leader_acc_empty + buf * 8, rank != 0};
        }
        const bool valid = rt < ga.row_tiles;                  // odd tail: rank 1 has no tile
        // pass 1: statistics of this group's 256 columns (the first unit's tile is read while the
        // second unit's MMAs are still running)
        float sn = 0.f, smean = 0.f, sm2 = 0.f;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          {
            const int q = 4 * j + 2 * h + eg;
            mbar_wait(smem_u32(&ctrl->acc_full[q & 3]), (q >> 2) & 1, ga.err, 8);
            tc_fence_after();
          }
          uint32_t raw[32];
          tmem_ld32(tm[h], raw);
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            float y[32];
            bias32_from_smem(s_bias + nt[h] * TILE_N, c * 32, y);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 2) add2(y[i], y[i + 1], __uint_as_float(raw[i]), __uint_as_float(raw[i + 1]));
            if (c + 1 < 4) tmem_ld32(tm[h] + (c + 1) * 32, raw);
            float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 32; i += 2) s2 = __fadd2_rn(s2, make_float2(y[i], y[i + 1]));
            const float m = (s2.x + s2.y) * (1.0f / 32.0f);
            const float2 nm = make_float2(-m, -m);
            float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float2 d = __fadd2_rn(make_float2(y[i], y[i + 1]), nm);
              q2 = __ffma2_rn(d, d, q2);
            }
            stats_merge(sn, smean, sm2, 32.f, m, q2.x + q2.y);
          }
        }
        float2* x = xch + (size_t)(j & 1) * 2 * TILE_M;
        x[eg * TILE_M + r] = make_float2(smean, sm2);
        asm volatile("bar.sync 3, 256;" ::: "memory");
        {   // merge in a fixed order (group 0 then group 1) so both groups get identical statistics
          const float2 a = x[r], b = x[TILE_M + r];
          sn = 0.f; smean = 0.f; sm2 = 0.f;
          stats_merge(sn, smean, sm2, 256.f, a.x, a.y);
          stats_merge(sn, smean, sm2, 256.f, b.x, b.y);
        }
        const float rstd = rsqrtf(sm2 * inv_n + 1e-5f);
        const float nmr = -smean * rstd;
        // pass 2: normalise, activation, residual, packed store
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          uint32_t raw[32];
          tmem_ld32(tm[h], raw);
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            const int n0 = nt[h] * TILE_N + c * 32;
            float y[32];
            bias32_from_smem(s_bias, n0, y);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 2) add2(y[i], y[i + 1], __uint_as_float(raw[i]), __uint_as_float(raw[i + 1]));
            if (c + 1 < 4) tmem_ld32(tm[h] + (c + 1) * 32, raw);
            else acc_release(rel[h]);                            // the tile's last read is complete
            float4 rv[8];
            const bool has_res = ea.resid_tiled != nullptr && valid;
            if (has_res) {
#pragma unroll
              for (int qq = 0; qq < 8; ++qq)
                rv[qq] = ea.resid_tiled[((size_t)rt * ea.ld4 + (n0 >> 2) + qq) * TILE_M + r];
            }
            const float4* gp = reinterpret_cast<const float4*>(s_gamma + n0);
            const float4* bp = reinterpret_cast<const float4*>(s_beta + n0);
            const float2 rs2 = make_float2(rstd, rstd), nm2 = make_float2(nmr, nmr);
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
              const float4 g4 = gp[qq], b4 = bp[qq];
              // ((y - mean) * rstd) * gamma + beta on register pairs (FFMA2)
              const float2 t0 = __ffma2_rn(make_float2(y[qq * 4 + 0], y[qq * 4 + 1]), rs2, nm2);
              const float2 t1 = __ffma2_rn(make_float2(y[qq * 4 + 2], y[qq * 4 + 3]), rs2, nm2);
              const float2 o0 = __ffma2_rn(t0, make_float2(g4.x, g4.y), make_float2(b4.x, b4.y));
              const float2 o1 = __ffma2_rn(t1, make_float2(g4.z, g4.w), make_float2(b4.z, b4.w));
              y[qq * 4 + 0] = o0.x; y[qq * 4 + 1] = o0.y; y[qq * 4 + 2] = o1.x; y[qq * 4 + 3] = o1.y;
            }
            act_apply32_ct<ACT>(y);
            if (has_res) {
#pragma unroll
              for (int qq = 0; qq < 8; ++qq) {
                y[qq * 4 + 0] += rv[qq].x; y[qq * 4 + 1] += rv[qq].y;
                y[qq * 4 + 2] += rv[qq].z; y[qq * 4 + 3] += rv[qq].w;
              }
            }
            if (valid) {
              __nv_bfloat16* tile = ea.out_packed + (size_t)(rt * ea.out_kb + (n0 >> 6)) * TILE_ELEMS;
              store_packed32(tile, r, n0 & 63, y);
            }
          }
        }
      }
    } else {
    const Stage stg{STAGED ? out_stage + eg * TILE_BYTES : 0u, 1 + eg, r == 0};
    EpiState<EPI> st;
    int rt = 0, nt = 0;
    const bool skip = (ga.debug & 1) != 0;
    if (eg < n_local && !skip) {
      coords(eg, rt, nt);
      if (rt < ga.row_tiles) epi_first<EPI>(ea, rt, nt, r, st);
    }
#pragma unroll 1
    for (int q = eg; q < n_local; q += 2) {
      coords(q, rt, nt);
      const bool valid = rt < ga.row_tiles;                    // odd tail: rank 1 has no tile
      int rt2 = rt, nt2 = nt;
      bool valid2 = false;
      if (q + 2 < n_local) {
        coords(q + 2, rt2, nt2);
        valid2 = rt2 < ga.row_tiles;
      }
      const int buf = q & 3, use = q >> 2;
      if (valid && !skip) epi_stage_bias<EPI>(ea, sb, 1 + eg, r, st, valid2, nt2);
      mbar_wait(smem_u32(&ctrl->acc_full[buf]), use & 1, ga.err, 8);
      tc_fence_after();
      const uint32_t tm = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(buf * TILE_N);
      const AccRelease rel{rank == 0 ? smem_u32(&ctrl->acc_empty[buf]) : leader_acc_empty + buf * 8, rank != 0};
      if (!skip && valid) {
        epi_finish<EPI, ACT, STAGED ? 1 : 0>(ea, tm, rt, nt, ga.n_tiles, r, sb, st, valid2, rt2, nt2, rel, stg);
      } else {
        acc_release(rel);
        if (!skip && valid2) epi_first<EPI>(ea, rt2, nt2, r, st);
      }
    }
    if (STAGED && stg.leader) bulk_wait_all<0>();   // every staged tile has reached global memory
    }
  }

  tc_fence_before();
  cluster_sync_all();   // no CTA leaves (or frees TMEM) while its peer can still signal it
  if (warp == 2) tmem_dealloc2(tmem_base, 512);
}

}  // namespace aid
