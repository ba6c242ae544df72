// Two-GEMM chain in one CTA-pair kernel: adaLN modulation -> next layer.
//
//   phase 1:  [scale|shift] = SiLU(cond) x Wmod^T ; xn = LN(h)*(1+scale)+shift   (EPI_MODLN)
//   phase 2:  acc = xn x W2^T ; second epilogue EPI2 (attention: h += acc + b, mlp.0: GELU -> packed,
//             output_proj.0: SiLU -> packed)
//
// The normalised tile xn (128 rows x H bf16 per CTA, 128 KiB at H=512) never leaves the SM: the
// phase-1 epilogue writes it straight into the shared-memory A operand of phase 2 in the packed
// K-major layout, instead of 67 MB out to HBM and 67 MB back per layer at 65,536 rows (15 % of the
// DRAM bytes of a denoise step), and one launch replaces two.
//
// Same CTA-pair scheme as gemm2.cuh (rank c owns row tile 2*rp + c and half of every weight tile;
// only rank 0 issues tcgen05.mma.cta_group::2; relaxed remote arrives).  Shared memory per CTA:
//   [control 2 KiB][xn: kb x 16 KiB][ring of 16 KiB granules]
// Phase 1 streams two granules per k-block (SiLU(cond) tile, Wmod half tile), phase 2 one (W2 half
// tile).  Extra barriers: xn_full[kb] (phase-1 epilogue -> MMA warp; rank 1's arrivals are relayed
// by its forwarder thread), xn_empty[kb] (last phase-2 MMAs of a row-tile pair -> phase-1 epilogue
// of the next pair, multicast commit).
#pragma once
#include "gemm2.cuh"

namespace aid {

constexpr int CHAIN_RING = 6;

struct alignas(8) ChainCtrl {
  uint64_t ring_full[CHAIN_RING];
  uint64_t ring_empty[CHAIN_RING];
  uint64_t xn_full[MAX_RES_KB];
  uint64_t xn_empty[MAX_RES_KB];
  uint64_t acc_full[4];
  uint64_t acc_empty[4];
  uint32_t tmem_base;
  uint32_t pad_[(512 - (2 * CHAIN_RING + 2 * MAX_RES_KB + 8) * 8 - 4) / 4];
  float bias_stage[2][TILE_N];   // per epilogue group; 512-byte offset
  uint8_t pad2_[SMEM_CTRL - 512 - 2 * TILE_N * 4];
};
static_assert(sizeof(ChainCtrl) == SMEM_CTRL, "control block must be exactly SMEM_CTRL bytes");

struct ChainArgs {
  const uint8_t* A;    // packed SiLU(cond) [row_tiles][kb]
  const uint8_t* B1;   // packed adaLN modulation weight (MODLN row map), 128-row tiles [kb n-tiles][kb]
  const uint8_t* B2;   // packed second weight, 128-row tiles [n_tiles2][kb]
  int row_tiles;
  int kb;              // H / 64 (<= MAX_RES_KB); phase 1 has kb accumulator tiles per row tile
  int n_tiles2;        // even
  int* err;
  int reverse;
  int debug;
};

template <int EPI2, int ACT2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
chain2_kernel(const ChainArgs ca, const EpiArgs e1, const EpiArgs e2) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;          // same offset in both CTAs
  uint8_t* smem = smem_raw + (base - raw_addr);
  ChainCtrl* ctrl = reinterpret_cast<ChainCtrl*>(smem);
  const uint32_t xn_smem = base + SMEM_CTRL;
  const uint32_t ring_smem = xn_smem + ca.kb * TILE_BYTES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int KB = ca.kb;
  const int U1 = KB / 2;                 // phase-1 units (256 accumulator columns = 128 hidden columns)
  const int U2 = ca.n_tiles2 / 2;        // phase-2 units
  const int rps = (ca.row_tiles + 1) >> 1;
  const int rp_begin = (int)((long long)pair * rps / num_pairs);
  const int rp_end = (int)((long long)(pair + 1) * rps / num_pairs);
  auto rp_at = [&](int i) { return ca.reverse ? rps - 1 - i : i; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < CHAIN_RING; ++i) {
      mbar_init(smem_u32(&ctrl->ring_full[i]), rank == 0 ? 2 : 1);   // leader: + rank 1's forwarder
      mbar_init(smem_u32(&ctrl->ring_empty[i]), 1);
    }
    for (int i = 0; i < MAX_RES_KB; ++i) {
      mbar_init(smem_u32(&ctrl->xn_full[i]), rank == 0 ? 5 : 4);     // 4 epilogue warps (+ forwarder)
      mbar_init(smem_u32(&ctrl->xn_empty[i]), 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(smem_u32(&ctrl->acc_full[i]), 1);
      mbar_init(smem_u32(&ctrl->acc_empty[i]), 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc2(smem_u32(&ctrl->tmem_base), 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;

  if (warp == 0) {
    // ===================== producer (both CTAs) =====================
    if (lane == 0) {
      int g = 0;
      uint32_t ph = 0;
      auto load = [&](const uint8_t* src) {
        const uint32_t fb = smem_u32(&ctrl->ring_full[g]);
        mbar_wait(smem_u32(&ctrl->ring_empty[g]), ph ^ 1, ca.err, 2);
        mbar_arrive_expect_tx(fb, TILE_BYTES);
        bulk_g2s(ring_smem + g * TILE_BYTES, src, TILE_BYTES, fb);
        if (++g == CHAIN_RING) { g = 0; ph ^= 1; }
      };
      for (int i = rp_begin; i < rp_end; ++i) {
        const int rp = rp_at(i);
        const int rt = min(2 * rp + (int)rank, ca.row_tiles - 1);   // odd tail: reload a valid tile
        for (int u = 0; u < U1; ++u)
          for (int kb = 0; kb < KB; ++kb) {
            load(ca.A + ((size_t)rt * KB + kb) * TILE_BYTES);
            load(ca.B1 + ((size_t)(2 * u + (int)rank) * KB + kb) * TILE_BYTES);
          }
        for (int u = 0; u < U2; ++u)
          for (int kb = 0; kb < KB; ++kb) load(ca.B2 + ((size_t)(2 * u + (int)rank) * KB + kb) * TILE_BYTES);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 1) {
      // ===================== forwarder (rank 1) =====================
      const uint32_t leader_full = map_to_rank(smem_u32(&ctrl->ring_full[0]), 0);
      const uint32_t leader_xn = map_to_rank(smem_u32(&ctrl->xn_full[0]), 0);
      int g = 0;
      uint32_t ph = 0, xpar = 0;
      auto fwd = [&]() {
        mbar_wait(smem_u32(&ctrl->ring_full[g]), ph, ca.err, 10);
        mbar_arrive_cluster_relaxed(leader_full + g * 8);
        if (++g == CHAIN_RING) { g = 0; ph ^= 1; }
      };
      for (int i = rp_begin; i < rp_end; ++i) {
        for (int u = 0; u < U1; ++u)
          for (int kb = 0; kb < KB; ++kb) { fwd(); fwd(); }
        for (int u = 0; u < U2; ++u)
          for (int kb = 0; kb < KB; ++kb) {
            if (u == 0) {   // my CTA's xn k-block is complete -> tell the leader's MMA warp
              mbar_wait(smem_u32(&ctrl->xn_full[kb]), xpar, ca.err, 12);
              mbar_arrive_cluster_relaxed(leader_xn + kb * 8);
            }
            fwd();
          }
        xpar ^= 1;
      }
    } else if (rank == 0) {
      // ===================== MMA issuer (leader): whole warp, one elected lane issues ==========
      constexpr uint32_t idesc = umma_idesc_bf16(2 * TILE_M, 2 * TILE_N);
      int g = 0;
      uint32_t ph = 0, xpar = 0;
      int q = 0;
      auto next = [&]() { if (++g == CHAIN_RING) { g = 0; ph ^= 1; } };
      for (int i = rp_begin; i < rp_end; ++i) {
        for (int u = 0; u < U1 + U2; ++u) {
          const bool p1 = u < U1;
#pragma unroll
          for (int t = 0; t < 2; ++t)
            mbar_wait(smem_u32(&ctrl->acc_empty[(q + t) & 3]), (((q + t) >> 2) & 1) ^ 1, ca.err, 4);
          const uint32_t d = tmem_base + (uint32_t)((q & 3) * TILE_N);
          for (int kb = 0; kb < KB; ++kb) {
            uint32_t a_tile, b_tile;
            int ga_ = -1;
            if (p1) {
              mbar_wait(smem_u32(&ctrl->ring_full[g]), ph, ca.err, 6);
              a_tile = ring_smem + g * TILE_BYTES;
              ga_ = g;
              next();
            } else {
              if (u == U1) mbar_wait(smem_u32(&ctrl->xn_full[kb]), xpar, ca.err, 13);
              a_tile = xn_smem + kb * TILE_BYTES;
            }
            mbar_wait(smem_u32(&ctrl->ring_full[g]), ph, ca.err, 7);
            b_tile = ring_smem + g * TILE_BYTES;
            tc_fence_after();
            const uint64_t ad = umma_desc_kmajor(a_tile, TILE_M * 16);
            const uint64_t bd = umma_desc_kmajor(b_tile, TILE_M * 16);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < TILE_K / 16; ++k)
                umma2_bf16(d, ad + (uint64_t)(k * (2 * TILE_M * 16) >> 4), bd + (uint64_t)(k * (2 * TILE_M * 16) >> 4),
                           idesc, (kb | k) ? 1u : 0u);
              if (p1) umma2_commit_both(smem_u32(&ctrl->ring_empty[ga_]));
              umma2_commit_both(smem_u32(&ctrl->ring_empty[g]));
              if (u == U1 + U2 - 1) umma2_commit_both(smem_u32(&ctrl->xn_empty[kb]));
            }
            __syncwarp();
            next();
          }
          if (elect_one()) {
            umma2_commit_both(smem_u32(&ctrl->acc_full[q & 3]));
            umma2_commit_both(smem_u32(&ctrl->acc_full[(q + 1) & 3]));
          }
          __syncwarp();
          q += 2;
        }
        xpar ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (two groups of 4 warps, both CTAs) =====================
    const int eg = (warp - 4) >> 2;
    const int lq = warp & 3;
    const int r = lq * 32 + lane;
    float* sb = ctrl->bias_stage[eg];
    const uint32_t leader_acc_empty = map_to_rank(smem_u32(&ctrl->acc_empty[0]), 0);
    const bool skip = (ca.debug & 1) != 0;
    int qbase = 0;
    uint32_t xpar = 0;
    for (int i = rp_begin; i < rp_end; ++i) {
      const int rp = rp_at(i);
      const int rt = 2 * rp + (int)rank;
      const bool valid = rt < ca.row_tiles && !skip;            // odd tail: rank 1 has no tile
      auto slot_wait = [&](int q) {
        mbar_wait(smem_u32(&ctrl->acc_full[q & 3]), (q >> 2) & 1, ca.err, 8);
        tc_fence_after();
        return tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)((q & 3) * TILE_N);
      };
      auto release_of = [&](int q) {
        return AccRelease{rank == 0 ? smem_u32(&ctrl->acc_empty[q & 3]) : leader_acc_empty + (q & 3) * 8, rank != 0};
      };
      // ---- phase 1: adaLN tiles nt = eg, eg+2, ... -> xn k-block nt in shared memory
      {
        EpiState<EPI_MODLN> st;
        if (valid) epi_first<EPI_MODLN>(e1, rt, eg, r, st);
        for (int nt = eg; nt < KB; nt += 2) {
          const int q = qbase + nt;
          const bool has_next = nt + 2 < KB;
          mbar_wait(smem_u32(&ctrl->xn_empty[nt]), xpar ^ 1, ca.err, 14);   // previous pair's phase 2 done
          if (valid) epi_stage_bias<EPI_MODLN>(e1, sb, 1 + eg, r, st, has_next, nt + 2);
          const uint32_t tm = slot_wait(q);
          const AccRelease rel = release_of(q);
          if (valid) {
            epi_finish<EPI_MODLN, ACT_NONE, 2>(e1, tm, rt, nt, KB, r, sb, st, has_next, rt, nt + 2, rel,
                                               Stage{xn_smem + (uint32_t)nt * TILE_BYTES, 1 + eg, false});
          } else {
            acc_release(rel);
          }
          fence_async_smem();            // my st.shared writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&ctrl->xn_full[nt]));
        }
      }
      // ---- phase 2: tiles t = eg, eg+2, ... of the second GEMM
      {
        EpiState<EPI2> st;
        if (valid) epi_first<EPI2>(e2, rt, eg, r, st);
        for (int t = eg; t < ca.n_tiles2; t += 2) {
          const int q = qbase + KB + t;
          const bool has_next = t + 2 < ca.n_tiles2;
          if (valid) epi_stage_bias<EPI2>(e2, sb, 1 + eg, r, st, has_next, t + 2);
          const uint32_t tm = slot_wait(q);
          const AccRelease rel = release_of(q);
          if (valid) epi_finish<EPI2, ACT2, 0>(e2, tm, rt, t, ca.n_tiles2, r, sb, st, has_next, rt, t + 2, rel);
          else acc_release(rel);
        }
      }
      qbase += KB + ca.n_tiles2;
      xpar ^= 1;
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc2(tmem_base, 512);
}

}  // namespace aid
