// libaid_sm100.so — host orchestration and C ABI (see include/aid_b200.h).
// Everything numeric runs in the kernels of gemm.cuh / elementwise.cuh; this file only lays
// out buffers and enqueues launches on the caller's stream.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/aid_b200.h"
#include "elementwise.cuh"
#include "gemm.cuh"
#include "gemm2.cuh"
#include "chain2.cuh"

using namespace aid;

// ------------------------------------------------------------------------------------------
// errors / bookkeeping
static thread_local std::string g_err;
static thread_local long long g_launches = 0;

static int fail(const std::string& msg) {
  g_err = msg;
  return -1;
}
#define AID_CHECK(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return fail(std::string(#expr) + ": " + cudaGetErrorString(_e));                    \
  } while (0)
// AID_SEGV_BT=1: print a native backtrace on SIGSEGV (developer aid; resolve with addr2line on the .so)
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void aid_segv_handler(int) {
  void* bt[64];
  const int n = backtrace(bt, 64);
  backtrace_symbols_fd(bt, n, 2);
  _exit(139);
}
static const bool g_segv_bt = [] {
  if (getenv("AID_SEGV_BT") && atoi(getenv("AID_SEGV_BT")) != 0) signal(SIGSEGV, aid_segv_handler);
  return true;
}();
// AID_TRACE=1: print every launch site to stderr (developer aid for locating host-side faults)
static const bool g_trace = getenv("AID_TRACE") && atoi(getenv("AID_TRACE")) != 0;
#define AID_LAUNCH_CHECK(name)                                                            \
  do {                                                                                    \
    if (g_trace) { fprintf(stderr, "[aid] %s\n", name); fflush(stderr); }                 \
    ++g_launches;                                                                         \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) return fail(std::string(name) + ": " + cudaGetErrorString(_e)); \
  } while (0)
#define AID_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != 0) return _r;      \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

static int num_sms() {
  static int sms[64] = {0};   // per device: one process may drive several GPUs
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms[dev] = n;
  }
  return sms[dev];
}

static int ew_grid(size_t work_items, int block = 256) {
  size_t g = (work_items + block - 1) / block;
  size_t cap = (size_t)num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ------------------------------------------------------------------------------------------
// bump allocator over a caller-owned buffer (also used in "measure" mode with base == null)
struct Arena {
  uint8_t* base;
  size_t off = 0;
  explicit Arena(void* b) : base(static_cast<uint8_t*>(b)) {}
  template <class T>
  T* take(size_t bytes) {
    off = align_up(off, 1024);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += bytes;
    return p;
  }
};

static size_t packed_tiles_bytes(int rows, int cols) {
  return (size_t)ceil_div(rows, TILE_M) * ceil_div(cols, TILE_K) * TILE_BYTES;
}
static size_t tiled_bytes(int rows, int cols) {
  return (size_t)ceil_div(rows, TILE_M) * (ceil_div(cols, TILE_N) * TILE_N) * TILE_M * 4;
}
static size_t stats_bytes(int rows, int cols) {
  return (size_t)ceil_div(rows, TILE_M) * ceil_div(cols, TILE_N) * TILE_M * sizeof(float2);
}

// ------------------------------------------------------------------------------------------
// packed Linear: weight tiles + padded fp32 bias
struct PLin {
  const uint8_t* w = nullptr;
  const float* b = nullptr;
  int n = 0, k = 0;        // logical [n, k]
  int n_tiles = 0, kb = 0; // padded tile counts (128-row n-tiles, 64-col k-blocks)
  int nw = 1;              // packed tile height in n-tiles: 2 -> 256-row tiles for N=256 MMAs
};

// CTA-pair kernels (gemm2.cuh, tcgen05.mma.cta_group::2) are the default for every layer whose padded
// width is a multiple of 256: each CTA fetches half of every weight tile and the UMMA reads 8 KiB
// instead of 12 KiB of operands per K=16 step, which takes the shared-memory port out of the
// critical path (MMA stream with loads: 1.83 PFLOP/s vs 1.60 for the single-CTA kernel; mlp.0
// 145 -> 114 us).  AID_PAIRS=0 selects the single-CTA kernels (weights then packed 256 rows tall).
static bool use_pairs() {
  static const bool on = !(getenv("AID_PAIRS") && atoi(getenv("AID_PAIRS")) == 0);
  return on;
}

static void plin_shape(PLin& p, int n, int k, bool modln = false) {
  p.n = n;
  p.k = k;
  p.kb = ceil_div(k, TILE_K);
  p.n_tiles = modln ? ceil_div(n / 2, 64) : ceil_div(n, TILE_N);
  static const bool narrow = getenv("AID_DEBUG") && (atoi(getenv("AID_DEBUG")) & 4);   // knob: N=128 units
  p.nw = (!use_pairs() && !narrow && p.n_tiles % 2 == 0) ? 2 : 1;
}
static size_t plin_w_bytes(const PLin& p) { return (size_t)p.n_tiles * p.kb * TILE_BYTES; }
static size_t plin_b_bytes(const PLin& p) { return (size_t)p.n_tiles * TILE_N * sizeof(float); }

static int pack_linear(const PLin& p, const float* w, const float* b, int mode, int H, cudaStream_t st) {
  size_t chunks = (size_t)p.n_tiles * p.kb * 1024;
  k_pack_rows<<<ew_grid(chunks), 256, 0, st>>>(
      w, p.n, p.k, p.k, reinterpret_cast<__nv_bfloat16*>(const_cast<uint8_t*>(p.w)), p.n_tiles, p.kb,
      mode, H, p.nw);
  AID_LAUNCH_CHECK("k_pack_rows(weight)");
  int n_pad = p.n_tiles * TILE_N;
  k_pack_bias<<<ceil_div(n_pad, 256), 256, 0, st>>>(b, p.n, const_cast<float*>(p.b), n_pad, mode, H);
  AID_LAUNCH_CHECK("k_pack_bias");
  return 0;
}

// ------------------------------------------------------------------------------------------
// Optional device timing of one GEMM class (bench.py roofline): CUDA events are recorded on the
// launching stream around every launch whose (epilogue, K blocks, N tiles) match the selection.
struct GemmProfile {
  bool on = false;
  int epi = -1, kb = -1, n_tiles = -1;
  std::vector<cudaEvent_t> ev;   // start/stop pairs
  size_t used = 0;
};
constexpr int PROF_SLOTS = 4;
static GemmProfile g_prof[PROF_SLOTS];

static int prof_set(int slot, int32_t epi, int32_t k, int32_t n) {
  GemmProfile& p = g_prof[slot];
  p.on = epi >= 0;
  p.epi = epi;
  p.kb = (k + TILE_K - 1) / TILE_K;
  p.n_tiles = (n + TILE_N - 1) / TILE_N;
  p.used = 0;
  return 0;
}
// slot 0 only (and clears the other slots): the round-1 interface
extern "C" int32_t aid_profile_select(int32_t epi, int32_t k, int32_t n) {
  for (int i = 1; i < PROF_SLOTS; ++i) prof_set(i, -1, 0, 0);
  return prof_set(0, epi, k, n);
}
extern "C" int32_t aid_profile_select_slot(int32_t slot, int32_t epi, int32_t k, int32_t n) {
  if (slot < 0 || slot >= PROF_SLOTS) return fail("aid_profile_select_slot: slot out of range");
  return prof_set(slot, epi, k, n);
}
static int prof_collect(int slot, double* total_ms, int64_t* launches) {
  GemmProfile& p = g_prof[slot];
  double tot = 0;
  for (size_t i = 0; i + 1 < p.used; i += 2) {
    AID_CHECK(cudaEventSynchronize(p.ev[i + 1]));
    float ms = 0;
    AID_CHECK(cudaEventElapsedTime(&ms, p.ev[i], p.ev[i + 1]));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = (int64_t)(p.used / 2);
  p.used = 0;
  return 0;
}
extern "C" int32_t aid_profile_collect(double* total_ms, int64_t* launches) { return prof_collect(0, total_ms, launches); }
extern "C" int32_t aid_profile_collect_slot(int32_t slot, double* total_ms, int64_t* launches) {
  if (slot < 0 || slot >= PROF_SLOTS) return fail("aid_profile_collect_slot: slot out of range");
  return prof_collect(slot, total_ms, launches);
}
// slot whose selection matches this launch, or -1
static int prof_match(int epi, int kb, int n_tiles) {
  for (int i = 0; i < PROF_SLOTS; ++i) {
    const GemmProfile& p = g_prof[i];
    if (p.on && p.epi == epi && p.kb == kb && p.n_tiles == n_tiles && p.used < 200000) return i;
  }
  return -1;
}
static cudaEvent_t prof_event(int slot) {
  GemmProfile& p = g_prof[slot];
  if (p.used == p.ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    p.ev.push_back(e);
  }
  return p.ev[p.used++];
}

// ------------------------------------------------------------------------------------------
// GEMM launch
template <int EPI, int NW, int G, bool RES, int ACT = ACT_NONE>
static int launch_gemm_inst(const GemmArgs& ga, const EpiArgs& ea, cudaStream_t st) {
  static bool configured = false;
  auto kern = gemm_kernel<EPI, NW, G, RES, ACT>;
  if (!configured) {
    AID_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    configured = true;
  }
  const int a_bytes = RES ? ga.kb * TILE_BYTES : 0;
  const int slot = NW * TILE_BYTES;
  const int tab_bytes = ga.conv.cin8 ? CONV_TAB_BYTES : 0;   // implicit-im2col address table after the ring
  if (ga.kb * 8 * 8 > CONV_TAB_BYTES && tab_bytes) return fail("gemm: conv K too long for the address table");
  int ring = (SMEM_LIMIT - 1024 - SMEM_CTRL - a_bytes - tab_bytes) / slot;
  if (ring > MAX_RING) ring = MAX_RING;
  if (ring < (RES ? 2 : G + 2)) return fail("gemm: not enough shared memory for the ring");
  const size_t smem = 1024 + SMEM_CTRL + a_bytes + (size_t)ring * slot + tab_bytes;
  const int units = (ga.splits > 1 ? ga.splits : 1) * ga.row_tiles * (ga.n_tiles / (NW * G));
  int grid = units < num_sms() ? units : num_sms();
  if (grid < 1) return 0;
  const int pslot = prof_match(EPI, ga.kb, ga.n_tiles);
  const bool prof = pslot >= 0;
  if (prof) cudaEventRecord(prof_event(pslot), st);
  {
    // programmatic stream serialization: this grid may begin (prologue only, see pdl_wait in the
    // kernel) while the previous kernel of the stream drains
    static const bool pdl = !(getenv("AID_DEBUG") && (atoi(getenv("AID_DEBUG")) & 2048));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && !prof) ? 1 : 0;
    AID_CHECK(cudaLaunchKernelEx(&cfg, kern, ga, ea, ring));
  }
  if (prof) cudaEventRecord(prof_event(pslot), st);
  AID_LAUNCH_CHECK("gemm_kernel");
  return 0;
}

template <int EPI, bool RES, int ACT>
static int launch_gemm2_inst(const GemmArgs& ga, const EpiArgs& ea, cudaStream_t st) {
  static bool configured = false;
  auto kern = gemm2_kernel<EPI, RES, ACT>;
  if (!configured) {
    AID_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    configured = true;
  }
  const int a_bytes = RES ? ga.kb * TILE_BYTES : 0;
  const int slot = (RES ? 1 : 2) * TILE_BYTES;
  constexpr int stage_out = gemm2_stage_bytes(EPI);
  const int tab_bytes = ga.conv.cin8 ? CONV_TAB_BYTES : 0;
  if (ga.kb * 8 * 8 > CONV_TAB_BYTES && tab_bytes) return fail("gemm2: conv K too long for the address table");
  int ring = (SMEM_LIMIT - 1024 - SMEM_CTRL - stage_out - a_bytes - tab_bytes) / slot;
  if (ring > MAX_RING2) ring = MAX_RING2;
  if (ring < 2) return fail("gemm2: not enough shared memory for the ring");
  const size_t smem = 1024 + SMEM_CTRL + stage_out + a_bytes + (size_t)ring * slot + tab_bytes;
  const int units = ((ga.row_tiles + 1) / 2) * (ga.n_tiles / 2);
  const int max_pairs = num_sms() / 2;
  const int pairs = units < max_pairs ? units : max_pairs;
  if (pairs < 1) return 0;
  const int pslot = prof_match(EPI, ga.kb, ga.n_tiles);
  const bool prof = pslot >= 0;
  if (prof) cudaEventRecord(prof_event(pslot), st);
  {
    static const bool pdl = !(getenv("AID_DEBUG") && (atoi(getenv("AID_DEBUG")) & 2048));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);        // cluster dims (2,1,1) come from the kernel's __cluster_dims__
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && !prof) ? 1 : 0;
    AID_CHECK(cudaLaunchKernelEx(&cfg, kern, ga, ea, ring));
  }
  if (prof) cudaEventRecord(prof_event(pslot), st);
  AID_LAUNCH_CHECK("gemm2_kernel");
  return 0;
}

template <int EPI, int ACT>
static int launch_gemm_shape(const GemmArgs& ga, const EpiArgs& ea, cudaStream_t st, bool res, bool wide) {
  if (use_pairs() && ga.n_tiles % 2 == 0 && ga.splits == 1) {
    if (res) return launch_gemm2_inst<EPI, true, ACT>(ga, ea, st);
    return launch_gemm2_inst<EPI, false, ACT>(ga, ea, st);
  }
  if (res) {
    if (wide) return launch_gemm_inst<EPI, 2, 1, true, ACT>(ga, ea, st);
    return launch_gemm_inst<EPI, 1, 1, true, ACT>(ga, ea, st);
  }
  if (wide) return launch_gemm_inst<EPI, 2, 1, false, ACT>(ga, ea, st);
  return launch_gemm_inst<EPI, 1, 1, false, ACT>(ga, ea, st);
}

template <int EPI>
static int launch_gemm(const uint8_t* A, int row_tiles, const PLin& w, EpiArgs ea, cudaStream_t st,
                       int* err_flag, int splits = 1, const ConvA* conv = nullptr) {
  GemmArgs ga;
  memset(&ga.conv, 0, sizeof(ga.conv));
  if (conv) ga.conv = *conv;
  ga.A = A;
  ga.B = w.w;
  ga.row_tiles = row_tiles;
  ga.kb = w.kb / splits;       // split-K: w.kb must be a multiple of splits
  ga.kb_stride = w.kb;
  ga.splits = splits;
  ga.n_tiles = w.n_tiles;
  ga.err = err_flag;
  ea.split_rt = row_tiles;
  static const int dbg = getenv("AID_DEBUG") ? atoi(getenv("AID_DEBUG")) : 0;
  ga.debug = dbg;
  ea.debug = dbg;
  // Consecutive GEMMs of a chain walk the row tiles in opposite directions: the tiles the
  // previous kernel wrote LAST are the ones this kernel reads FIRST, so they are still in the
  // 126 MB L2 instead of coming back from HBM (activations of a 65k-row batch exceed L2).
  static thread_local unsigned flip = 0;
  ga.reverse = (dbg & 8) ? 0 : (int)(flip++ & 1);
  if (!ea.bias) ea.bias = w.b;
  const bool res = ga.kb <= MAX_RES_KB && !conv;   // the implicit-im2col producer streams A
  const bool wide = w.nw == 2;   // N=256 MMAs whenever the tile count allows (fixed at pack time)
  // A resident in shared memory (K <= 512) or streamed through the ring.  Streamed A: one N=256
  // unit per pass so TMEM holds two units and the epilogue of unit u overlaps the MMAs of unit u+1
  // (a 512-column unit fills TMEM and serialises them: 248 us vs 103 us of MMA time for mlp.2 at
  // 65,536 rows).  The second pass over the same A row tile hits L2.
  if constexpr (EPI == EPI_DACT) {
    if (ea.act == ACT_GELU) return launch_gemm_shape<EPI, ACT_GELU>(ga, ea, st, res, wide);
    if (ea.act == ACT_SILU) return launch_gemm_shape<EPI, ACT_SILU>(ga, ea, st, res, wide);
    return fail("launch_gemm<EPI_DACT>: activation must be GELU or SiLU");
  } else if constexpr (EPI == EPI_PACK || EPI == EPI_F32) {
    switch (ea.act) {
      case ACT_SILU: return launch_gemm_shape<EPI, ACT_SILU>(ga, ea, st, res, wide);
      case ACT_RELU: return launch_gemm_shape<EPI, ACT_RELU>(ga, ea, st, res, wide);
      case ACT_GELU: return launch_gemm_shape<EPI, ACT_GELU>(ga, ea, st, res, wide);
      default: return launch_gemm_shape<EPI, ACT_NONE>(ga, ea, st, res, wide);
    }
  } else {
    return launch_gemm_shape<EPI, ACT_NONE>(ga, ea, st, res, wide);
  }
}

// Linear -> LayerNorm -> ReLU (+ residual) -> packed bf16 as ONE CTA-pair kernel (EPI_LNACT,
// gemm2.cuh): layers of width exactly 512 (the EFE heads at the default hidden size).  The fp32
// pre-activation never leaves TMEM; the two-kernel form (EPI_F32 + k_ln_act) wrote and re-read it
// through HBM (268 MB per layer at 65,536 rows).  AID_FUSED_LN=0 selects the two-kernel form.
static bool lnact_fusable(const PLin& w, int act) {
  static const bool off = getenv("AID_FUSED_LN") && atoi(getenv("AID_FUSED_LN")) == 0;
  return !off && use_pairs() && w.n == LN_COLS && w.n_tiles == 4 && w.nw == 1 && w.kb <= MAX_RES_KB &&
         act == ACT_RELU;
}
static int launch_gemm_lnact(const uint8_t* A, int row_tiles, const PLin& w, EpiArgs ea, cudaStream_t st,
                             int* err_flag) {
  GemmArgs ga;
  memset(&ga.conv, 0, sizeof(ga.conv));
  ga.A = A;
  ga.B = w.w;
  ga.row_tiles = row_tiles;
  ga.kb = w.kb;
  ga.kb_stride = w.kb;
  ga.splits = 1;
  ga.n_tiles = w.n_tiles;
  ga.err = err_flag;
  ea.split_rt = row_tiles;
  static const int dbg = getenv("AID_DEBUG") ? atoi(getenv("AID_DEBUG")) : 0;
  ga.debug = dbg & ~(1 | 1024);     // the epilogue-off / MMA-off knobs do not apply to this kernel
  ea.debug = 0;
  static thread_local unsigned flip = 0;
  ga.reverse = (dbg & 8) ? 0 : (int)(flip++ & 1);
  if (!ea.bias) ea.bias = w.b;
  return launch_gemm2_inst<EPI_LNACT, true, ACT_RELU>(ga, ea, st);
}

// adaLN modulation -> next layer in one CTA-pair kernel (chain2.cuh), opt-in with AID_CHAIN=1.
// Correct (the GPU suite passes with it) and it removes the 134 MB xn round trip per layer pair, but
// in its first form it is slower than the two separate kernels (3.65 vs 3.15 ms per denoise step at
// 65,536 rows): the two epilogues in one kernel spill at 168 registers and phase 1 has only three
// k-blocks of operands in flight.  Kept as the starting point of the cross-layer fusion work.
// AID_CHAIN=1 opts in.  Measured (round 2) also for the launch-latency-bound small batches, where
// removing 13 of a step's 47 launches looked attractive: B <= 4,096 rows took 34.7 ms per 50-step call with
// the chain kernel against 19.5 ms without -- a chain unit is a whole row-tile pair, so one CTA pair walks
// all column groups of both layers serially, while the separate kernels spread them over the SMs.
static bool use_chain(int row_tiles) {
  (void)row_tiles;
  static const bool on = use_pairs() && getenv("AID_CHAIN") && atoi(getenv("AID_CHAIN")) != 0;
  return on;
}
static unsigned g_chain_flip = 0;

template <int EPI2, int ACT2>
static int launch_chain_inst(const ChainArgs& ca, const EpiArgs& e1, const EpiArgs& e2, cudaStream_t st) {
  static bool configured = false;
  auto kern = chain2_kernel<EPI2, ACT2>;
  if (!configured) {
    AID_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    configured = true;
  }
  const size_t smem = 1024 + SMEM_CTRL + (size_t)ca.kb * TILE_BYTES + (size_t)CHAIN_RING * TILE_BYTES;
  const int rps = (ca.row_tiles + 1) / 2;
  const int max_pairs = num_sms() / 2;
  const int pairs = rps < max_pairs ? rps : max_pairs;
  if (pairs < 1) return 0;
  kern<<<2 * pairs, GEMM_THREADS, smem, st>>>(ca, e1, e2);
  AID_LAUNCH_CHECK("chain2_kernel");
  return 0;
}

// mod: adaLN modulation weight (MODLN row map), w2: the layer that consumes the normalised tile
static bool chain_ok(const PLin& mod, const PLin& w2, int row_tiles) {
  return use_chain(row_tiles) && mod.nw == 1 && w2.nw == 1 && mod.kb <= MAX_RES_KB && mod.kb % 2 == 0 &&
         mod.n_tiles == mod.kb && w2.kb == mod.kb && w2.n_tiles % 2 == 0;
}

template <int EPI2>
static int launch_chain(const uint8_t* csilu, int row_tiles, const PLin& mod, const PLin& w2, EpiArgs e1,
                        EpiArgs e2, cudaStream_t st, int* err_flag) {
  static const int dbg = getenv("AID_DEBUG") ? atoi(getenv("AID_DEBUG")) : 0;
  ChainArgs ca;
  ca.A = csilu; ca.B1 = mod.w; ca.B2 = w2.w;
  ca.row_tiles = row_tiles; ca.kb = mod.kb; ca.n_tiles2 = w2.n_tiles;
  ca.err = err_flag; ca.debug = dbg;
  ca.reverse = (dbg & 8) ? 0 : (int)(g_chain_flip++ & 1);
  e1.debug = dbg; e2.debug = dbg;
  e1.split_rt = row_tiles; e2.split_rt = row_tiles;
  if (!e1.bias) e1.bias = mod.b;
  if (!e2.bias) e2.bias = w2.b;
  if constexpr (EPI2 == EPI_PACK) {
    switch (e2.act) {
      case ACT_SILU: return launch_chain_inst<EPI2, ACT_SILU>(ca, e1, e2, st);
      case ACT_GELU: return launch_chain_inst<EPI2, ACT_GELU>(ca, e1, e2, st);
      default: return fail("launch_chain: unsupported activation");
    }
  } else {
    return launch_chain_inst<EPI2, ACT_NONE>(ca, e1, e2, st);
  }
}

static EpiArgs epi_zero() {
  EpiArgs e;
  memset(&e, 0, sizeof(e));
  e.tw_scalar = 1.0f;
  return e;
}

// ------------------------------------------------------------------------------------------
// Score network: packed layout
struct ScoreBlockW {
  PLin mod1, attn, mod2, fc1, fc2;
};
struct ScoreW {
  int L, O, H, E, NB;
  PLin te1, te3, oe0, oe4, oe7, ce2, ce4, lp, nf, out0, out2;
  std::vector<ScoreBlockW> blk;
  // fp32 vectors / scalars copied verbatim
  const float *oe1_g, *oe1_b, *oe5_g, *oe5_b, *oe8_g, *oe8_b, *ce0_w, *ce0_b;
  const float *time_scale, *out_mult, *freq_scale;
  size_t total = 0;
};

static int score_dims_ok(const AidScoreDims* d) {
  if (!d) return fail("dims is null");
  if (d->hidden_dim <= 0 || d->hidden_dim % 64) return fail("hidden_dim must be a positive multiple of 64");
  if (d->time_embed_dim <= 0 || d->time_embed_dim % 64) return fail("time_embed_dim must be a multiple of 64");
  if (d->latent_dim <= 0 || d->latent_dim % 4) return fail("latent_dim must be a positive multiple of 4");
  if (d->obs_dim <= 0) return fail("obs_dim must be positive");
  if (d->num_blocks < 0 || d->num_blocks > 64) return fail("num_blocks out of range");
  return 0;
}

static void score_layout(const AidScoreDims* d, void* packed, ScoreW& s) {
  s.L = d->latent_dim; s.O = d->obs_dim; s.H = d->hidden_dim; s.E = d->time_embed_dim; s.NB = d->num_blocks;
  const int H = s.H;
  Arena a(packed);
  auto place = [&](PLin& p, int n, int k, bool modln = false) {
    plin_shape(p, n, k, modln);
    p.w = a.take<uint8_t>(plin_w_bytes(p));
    p.b = a.take<float>(plin_b_bytes(p));
  };
  place(s.te1, 2 * H, s.E);
  place(s.te3, H, 2 * H);
  place(s.oe0, H, s.O);
  place(s.oe4, H, H);
  place(s.oe7, H, H);
  place(s.ce2, s.E, s.E);
  place(s.ce4, H, s.E);
  place(s.lp, H, s.L);
  place(s.nf, 2 * H, H, true);
  place(s.out0, H / 2, H);
  place(s.out2, s.L, H / 2);
  s.blk.resize(s.NB);
  for (auto& b : s.blk) {
    place(b.mod1, 2 * H, H, true);
    place(b.attn, H, H);
    place(b.mod2, 2 * H, H, true);
    place(b.fc1, 4 * H, H);
    place(b.fc2, H, 4 * H);
  }
  auto vec = [&](int n) { return a.take<float>((size_t)n * sizeof(float)); };
  s.oe1_g = vec(H); s.oe1_b = vec(H); s.oe5_g = vec(H); s.oe5_b = vec(H); s.oe8_g = vec(H); s.oe8_b = vec(H);
  s.ce0_w = vec(s.E); s.ce0_b = vec(s.E);
  s.time_scale = vec(1); s.out_mult = vec(1); s.freq_scale = vec(1);
  s.total = align_up(a.off, 1024);
}

extern "C" size_t aid_score_packed_bytes(const AidScoreDims* dims) {
  if (score_dims_ok(dims)) return 0;
  ScoreW s;
  score_layout(dims, nullptr, s);
  return s.total;
}

extern "C" int32_t aid_score_num_params(const AidScoreDims* dims) {
  if (score_dims_ok(dims)) return -1;
  return AID_SP_BLOCK0 + dims->num_blocks * AID_SP_BLOCK_STRIDE;
}

extern "C" int32_t aid_score_pack(const AidScoreDims* dims, const float* const* P, int32_t num_params,
                                  void* packed, size_t packed_bytes, void* stream) {
  AID_TRY(score_dims_ok(dims));
  if (!P || !packed) return fail("aid_score_pack: null pointer");
  if (num_params != aid_score_num_params(dims)) return fail("aid_score_pack: wrong parameter count");
  ScoreW s;
  score_layout(dims, packed, s);
  if (packed_bytes < s.total) return fail("aid_score_pack: packed buffer too small");
  for (int i = 0; i < num_params; ++i)
    if (!P[i]) return fail("aid_score_pack: null parameter pointer at index " + std::to_string(i));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int H = s.H;
  AID_TRY(pack_linear(s.te1, P[AID_SP_TE1_W], P[AID_SP_TE1_B], MAP_PLAIN, H, st));
  AID_TRY(pack_linear(s.te3, P[AID_SP_TE3_W], P[AID_SP_TE3_B], MAP_PLAIN, H, st));
  AID_TRY(pack_linear(s.oe0, P[AID_SP_OE0_W], P[AID_SP_OE0_B], MAP_PLAIN, H, st));
  AID_TRY(pack_linear(s.oe4, P[AID_SP_OE4_W], P[AID_SP_OE4_B], MAP_PLAIN, H, st));
  AID_TRY(pack_linear(s.oe7, P[AID_SP_OE7_W], P[AID_SP_OE7_B], MAP_PLAIN, H, st));
  AID_TRY(pack_linear(s.ce2, P[AID_SP_CE2_W], P[AID_SP_CE2_B], MAP_PLAIN, H, st));
  AID_TRY(pack_linear(s.ce4, P[AID_SP_CE4_W], P[AID_SP_CE4_B], MAP_PLAIN, H, st));
  AID_TRY(pack_linear(s.lp, P[AID_SP_LP_W], P[AID_SP_LP_B], MAP_PLAIN, H, st));
  AID_TRY(pack_linear(s.nf, P[AID_SP_NF_W], P[AID_SP_NF_B], MAP_MODLN, H, st));
  AID_TRY(pack_linear(s.out0, P[AID_SP_OUT0_W], P[AID_SP_OUT0_B], MAP_PLAIN, H, st));
  AID_TRY(pack_linear(s.out2, P[AID_SP_OUT2_W], nullptr, MAP_PLAIN, H, st));
  for (int i = 0; i < s.NB; ++i) {
    const float* const* B = P + AID_SP_BLOCK0 + i * AID_SP_BLOCK_STRIDE;
    ScoreBlockW& b = s.blk[i];
    AID_TRY(pack_linear(b.mod1, B[AID_SPB_N1_W], B[AID_SPB_N1_B], MAP_MODLN, H, st));
    AID_TRY(pack_linear(b.mod2, B[AID_SPB_N2_W], B[AID_SPB_N2_B], MAP_MODLN, H, st));
    size_t chunks = (size_t)b.attn.n_tiles * b.attn.kb * 1024;
    k_pack_folded_attn<<<ew_grid(chunks, 128), 128, 0, st>>>(
        B[AID_SPB_INPROJ_W], B[AID_SPB_OUTPROJ_W], H,
        reinterpret_cast<__nv_bfloat16*>(const_cast<uint8_t*>(b.attn.w)), b.attn.n_tiles, b.attn.kb,
        b.attn.nw);
    AID_LAUNCH_CHECK("k_pack_folded_attn");
    int n_pad = b.attn.n_tiles * TILE_N;
    k_folded_attn_bias<<<ceil_div(n_pad, 128), 128, 0, st>>>(
        B[AID_SPB_INPROJ_B], B[AID_SPB_OUTPROJ_W], B[AID_SPB_OUTPROJ_B], H, const_cast<float*>(b.attn.b), n_pad);
    AID_LAUNCH_CHECK("k_folded_attn_bias");
    AID_TRY(pack_linear(b.fc1, B[AID_SPB_FC1_W], B[AID_SPB_FC1_B], MAP_PLAIN, H, st));
    AID_TRY(pack_linear(b.fc2, B[AID_SPB_FC2_W], B[AID_SPB_FC2_B], MAP_PLAIN, H, st));
  }
  auto copyv = [&](const float* dst, const float* src, int n) {
    return cudaMemcpyAsync(const_cast<float*>(dst), src, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st);
  };
  AID_CHECK(copyv(s.oe1_g, P[AID_SP_OE1_G], H)); AID_CHECK(copyv(s.oe1_b, P[AID_SP_OE1_B], H));
  AID_CHECK(copyv(s.oe5_g, P[AID_SP_OE5_G], H)); AID_CHECK(copyv(s.oe5_b, P[AID_SP_OE5_B], H));
  AID_CHECK(copyv(s.oe8_g, P[AID_SP_OE8_G], H)); AID_CHECK(copyv(s.oe8_b, P[AID_SP_OE8_B], H));
  AID_CHECK(copyv(s.ce0_w, P[AID_SP_CE0_W], s.E)); AID_CHECK(copyv(s.ce0_b, P[AID_SP_CE0_B], s.E));
  AID_CHECK(copyv(s.time_scale, P[AID_SP_TIME_SCALE], 1));
  AID_CHECK(copyv(s.out_mult, P[AID_SP_OUTPUT_MULTIPLIER], 1));
  AID_CHECK(copyv(s.freq_scale, P[AID_SP_FREQ_SCALE], 1));
  return 0;
}

// ------------------------------------------------------------------------------------------
// Score network: workspace + forward pieces
struct ScoreWS {
  int B, RT;          // batch rows / row tiles
  int TR, TRT;        // time-table rows / row tiles (per-row times: TR == B)
  __nv_bfloat16 *zp, *obsp, *xn, *act, *o1, *csilu, *obs_h;   // batch-sized packed operands
  float4 *h, *obs_emb, *obs_y;                                 // batch-sized tiled fp32
  float2 *h_stats, *obs_stats;
  __nv_bfloat16 *te_e, *te_h, *ce_a, *ce_b;                    // time-table packed operands
  float4 *tsin, *tcont;                                        // time-table tiled fp32
  float *t_sin_arg, *t_norm, *t_flag, *t_w;                    // time-table per-row scalars
  // small-batch persistent sampler (small.inc), laid out when batch <= SM_MAX_BATCH
  float *sm_h0, *sm_h1, *sm_mod, *sm_tsin, *sm_tcont, *sm_obs;
  __nv_bfloat16 *sm_u, *sm_v;
  void* sm_steps_raw;
  unsigned int* sm_bar;
  int* err;
  size_t total;
};

static void score_ws_layout(const ScoreW& s, int batch, int table_rows, void* ws, ScoreWS& w) {
  const int H = s.H;
  w.B = batch; w.RT = ceil_div(batch, TILE_M);
  w.TR = table_rows; w.TRT = ceil_div(table_rows, TILE_M);
  Arena a(ws);
  w.err = a.take<int>(1024);
  w.zp = a.take<__nv_bfloat16>(packed_tiles_bytes(batch, s.L));
  w.obsp = a.take<__nv_bfloat16>(packed_tiles_bytes(batch, s.O));
  w.xn = a.take<__nv_bfloat16>(packed_tiles_bytes(batch, H));
  w.act = a.take<__nv_bfloat16>(packed_tiles_bytes(batch, 4 * H));
  w.o1 = a.take<__nv_bfloat16>(packed_tiles_bytes(batch, H / 2));
  w.csilu = a.take<__nv_bfloat16>(packed_tiles_bytes(batch, H));
  w.obs_h = a.take<__nv_bfloat16>(packed_tiles_bytes(batch, H));
  w.h = a.take<float4>(tiled_bytes(batch, H));
  w.obs_emb = a.take<float4>(tiled_bytes(batch, H));
  w.obs_y = a.take<float4>(tiled_bytes(batch, H));
  w.h_stats = a.take<float2>(stats_bytes(batch, H));
  w.obs_stats = a.take<float2>(stats_bytes(batch, H));
  w.te_e = a.take<__nv_bfloat16>(packed_tiles_bytes(table_rows, s.E));
  w.te_h = a.take<__nv_bfloat16>(packed_tiles_bytes(table_rows, 2 * H));
  w.ce_a = a.take<__nv_bfloat16>(packed_tiles_bytes(table_rows, s.E));
  w.ce_b = a.take<__nv_bfloat16>(packed_tiles_bytes(table_rows, s.E));
  w.tsin = a.take<float4>(tiled_bytes(table_rows, H));
  w.tcont = a.take<float4>(tiled_bytes(table_rows, H));
  w.t_sin_arg = a.take<float>((size_t)table_rows * 4);
  w.t_norm = a.take<float>((size_t)table_rows * 4);
  w.t_flag = a.take<float>((size_t)table_rows * 4);
  w.t_w = a.take<float>((size_t)table_rows * 4);
  w.sm_h0 = w.sm_h1 = w.sm_mod = w.sm_tsin = w.sm_tcont = w.sm_obs = nullptr;
  w.sm_u = w.sm_v = nullptr;
  w.sm_steps_raw = nullptr;
  w.sm_bar = nullptr;
  if (batch <= 256) {   // == SM_MAX_BATCH (small.inc)
    w.sm_h0 = a.take<float>((size_t)batch * H * 4);
    w.sm_h1 = a.take<float>((size_t)batch * H * 4);
    w.sm_mod = a.take<float>((size_t)batch * (2 * s.NB + 1) * 2 * H * 4);
    w.sm_tsin = a.take<float>((size_t)table_rows * H * 4);
    w.sm_tcont = a.take<float>((size_t)table_rows * H * 4);
    w.sm_obs = a.take<float>((size_t)batch * H * 4);
    w.sm_u = a.take<__nv_bfloat16>((size_t)batch * 4 * H * 2);
    w.sm_v = a.take<__nv_bfloat16>((size_t)batch * (H / 2) * 2);
    w.sm_steps_raw = a.take<uint8_t>((size_t)table_rows * 32);
    w.sm_bar = a.take<unsigned int>(16384);   // barrier counters (SM_BAR_BYTES)
  }
  w.total = align_up(a.off, 1024);
}

extern "C" size_t aid_score_workspace_bytes(const AidScoreDims* dims, int32_t batch, int32_t table_rows) {
  if (score_dims_ok(dims) || batch <= 0) return 0;
  if (table_rows <= 0) table_rows = batch;
  ScoreW s;
  score_layout(dims, nullptr, s);
  ScoreWS w;
  score_ws_layout(s, batch, table_rows, nullptr, w);
  return w.total;
}

// per-row time arguments for the score net's two branches (models/score_networks.py:121-141)
__global__ void k_time_args(const float* __restrict__ t, int rows, int continuous,
                            float* __restrict__ sin_arg, float* __restrict__ t_norm,
                            float* __restrict__ flag, float* __restrict__ tw) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  float ti = t[i];
  if (continuous) {
    sin_arg[i] = __fmul_rn(ti, 999.0f);
    t_norm[i] = __fsub_rn(__fmul_rn(2.0f, ti), 1.0f);
    flag[i] = 1.f;
    tw[i] = __fsqrt_rn(__fdiv_rn(1.0f, __fadd_rn(1e-5f, ti)));
  } else {
    sin_arg[i] = ti;
    t_norm[i] = 0.f;
    flag[i] = 0.f;
    tw[i] = 1.f;
  }
}

// time_embed (+ continuous_time_embed) over `w.TR` rows whose args are already in w.t_*.
static int run_time_table(const ScoreW& s, ScoreWS& w, bool any_continuous, cudaStream_t st) {
  const int H = s.H;
  k_sincos_pack<<<ew_grid((size_t)w.TRT * (s.E / 64) * 1024), 256, 0, st>>>(
      w.t_sin_arg, w.TR, s.freq_scale, s.E, w.te_e, w.TRT);
  AID_LAUNCH_CHECK("k_sincos_pack");
  EpiArgs e = epi_zero();
  e.act = ACT_SILU; e.n_valid = 2 * H; e.rows_valid = w.TR; e.out_packed = w.te_h; e.out_kb = 2 * H / 64;
  AID_TRY(launch_gemm<EPI_PACK>(reinterpret_cast<uint8_t*>(w.te_e), w.TRT, s.te1, e, st, w.err));
  e = epi_zero();
  e.n_valid = H; e.rows_valid = w.TR; e.out_tiled = w.tsin; e.ld4 = s.te3.n_tiles * TILE_N / 4;
  AID_TRY(launch_gemm<EPI_F32>(reinterpret_cast<uint8_t*>(w.te_h), w.TRT, s.te3, e, st, w.err));
  if (any_continuous) {
    k_cont0_pack<<<ew_grid((size_t)w.TRT * (s.E / 64) * 1024), 256, 0, st>>>(
        w.t_norm, w.TR, s.ce0_w, s.ce0_b, s.E, w.ce_a, w.TRT);
    AID_LAUNCH_CHECK("k_cont0_pack");
    e = epi_zero();
    e.act = ACT_SILU; e.n_valid = s.E; e.rows_valid = w.TR; e.out_packed = w.ce_b; e.out_kb = s.E / 64;
    AID_TRY(launch_gemm<EPI_PACK>(reinterpret_cast<uint8_t*>(w.ce_a), w.TRT, s.ce2, e, st, w.err));
    e = epi_zero();
    e.n_valid = H; e.rows_valid = w.TR; e.out_tiled = w.tcont; e.ld4 = s.ce4.n_tiles * TILE_N / 4;
    AID_TRY(launch_gemm<EPI_F32>(reinterpret_cast<uint8_t*>(w.ce_b), w.TRT, s.ce4, e, st, w.err));
  }
  return 0;
}

// obs_encoder (models/score_networks.py:49-59, eval mode) -> w.obs_emb (tiled fp32)
static int run_obs_encoder(const ScoreW& s, ScoreWS& w, const float* obs, cudaStream_t st) {
  const int H = s.H;
  const int ld4 = ceil_div(H, TILE_N) * TILE_N / 4;
  const int nt = ceil_div(H, TILE_N);
  if (!obs) {
    AID_CHECK(cudaMemsetAsync(w.obs_emb, 0, tiled_bytes(w.B, H), st));
    return 0;
  }
  const int kbo = ceil_div(s.O, TILE_K);
  k_pack_rows<<<ew_grid((size_t)w.RT * kbo * 1024), 256, 0, st>>>(obs, w.B, s.O, s.O, w.obsp, w.RT, kbo,
                                                                   MAP_PLAIN, 0, 1);
  AID_LAUNCH_CHECK("k_pack_rows(obs)");
  const PLin* lin[3] = {&s.oe0, &s.oe4, &s.oe7};
  const float* g[3] = {s.oe1_g, s.oe5_g, s.oe8_g};
  const float* b[3] = {s.oe1_b, s.oe5_b, s.oe8_b};
  const uint8_t* a_in = reinterpret_cast<uint8_t*>(w.obsp);
  for (int i = 0; i < 3; ++i) {
    EpiArgs e = epi_zero();
    e.n_valid = H; e.rows_valid = w.B; e.out_tiled = w.obs_y; e.ld4 = ld4; e.stats_out = w.obs_stats;
    AID_TRY(launch_gemm<EPI_F32>(a_in, w.RT, *lin[i], e, st, w.err));
    LnArgs l;
    l.x = w.obs_y; l.stats = w.obs_stats; l.stats_nt = nt; l.ld4 = ld4; l.n = H;
    l.gamma = g[i]; l.beta = b[i];
    l.act = (i < 2) ? ACT_SILU : ACT_NONE;
    l.out_packed = (i < 2) ? w.obs_h : nullptr;
    l.out_tiled = (i < 2) ? nullptr : w.obs_emb;
    l.resid = nullptr;
    k_ln_act<<<dim3(w.RT, ceil_div(H, 64)), 128, 0, st>>>(l);
    AID_LAUNCH_CHECK("k_ln_act");
    a_in = reinterpret_cast<uint8_t*>(w.obs_h);
  }
  return 0;
}

struct StepOut {
  // EPI_SCORE configuration for the last GEMM
  int do_step = 0;
  const float* tw_rows = nullptr;
  float tw_scalar = 1.0f;
  const float* z_in = nullptr;
  const float* eps = nullptr;
  const PhiloxState* philox = nullptr;
  unsigned philox_draw = 0;
  long long row_offset = 0;
  float c_s1 = 0, c_ra = 0, c_c1 = 0, c_c2 = 0, c_sigma = 0;
  float* z_out = nullptr;
  __nv_bfloat16* z_packed_out = nullptr;
};

// Trunk of the score net: w.zp (packed z) and w.csilu (packed SiLU(conditioning)) are ready.
// models/score_networks.py:156-170 with the attention folded (see k_pack_folded_attn).
static int run_score_trunk(const ScoreW& s, ScoreWS& w, const StepOut& o, cudaStream_t st) {
  const int H = s.H;
  const int ld4 = ceil_div(H, TILE_N) * TILE_N / 4;
  const int nt_h = ceil_div(H, TILE_N);
  const uint8_t* zp = reinterpret_cast<uint8_t*>(w.zp);
  const uint8_t* cs = reinterpret_cast<uint8_t*>(w.csilu);
  const uint8_t* xn = reinterpret_cast<uint8_t*>(w.xn);

  auto f32_out = [&](const PLin& lin, const uint8_t* a, bool resid) {
    EpiArgs e = epi_zero();
    e.n_valid = H; e.rows_valid = w.B; e.out_tiled = w.h; e.ld4 = ld4; e.stats_out = w.h_stats;
    e.resid_tiled = resid ? w.h : nullptr;
    return launch_gemm<EPI_F32>(a, w.RT, lin, e, st, w.err);
  };
  auto modln = [&](const PLin& lin) {
    EpiArgs e = epi_zero();
    e.rows_valid = w.B; e.n_valid = 2 * H; e.out_packed = w.xn; e.out_kb = H / 64;
    e.h_tiled = w.h; e.h_ld4 = ld4; e.stats_in = w.h_stats; e.stats_nt = nt_h; e.h_dim = H;
    return launch_gemm<EPI_MODLN>(cs, w.RT, lin, e, st, w.err);
  };

  auto modln_args = [&]() {
    EpiArgs e = epi_zero();
    e.rows_valid = w.B; e.n_valid = 2 * H; e.out_packed = w.xn; e.out_kb = H / 64;
    e.h_tiled = w.h; e.h_ld4 = ld4; e.stats_in = w.h_stats; e.stats_nt = nt_h; e.h_dim = H;
    return e;
  };
  auto f32_args = [&](bool resid) {
    EpiArgs e = epi_zero();
    e.n_valid = H; e.rows_valid = w.B; e.out_tiled = w.h; e.ld4 = ld4; e.stats_out = w.h_stats;
    e.resid_tiled = resid ? w.h : nullptr;
    return e;
  };
  AID_TRY(f32_out(s.lp, zp, false));
  for (int i = 0; i < s.NB; ++i) {
    const ScoreBlockW& b = s.blk[i];
    if (chain_ok(b.mod1, b.attn, w.RT)) {
      AID_TRY(launch_chain<EPI_F32>(cs, w.RT, b.mod1, b.attn, modln_args(), f32_args(true), st, w.err));
    } else {
      AID_TRY(modln(b.mod1));
      AID_TRY(f32_out(b.attn, xn, true));
    }
    EpiArgs e = epi_zero();
    e.act = ACT_GELU; e.n_valid = 4 * H; e.rows_valid = w.B; e.out_packed = w.act; e.out_kb = 4 * H / 64;
    if (chain_ok(b.mod2, b.fc1, w.RT)) {
      AID_TRY(launch_chain<EPI_PACK>(cs, w.RT, b.mod2, b.fc1, modln_args(), e, st, w.err));
    } else {
      AID_TRY(modln(b.mod2));
      AID_TRY(launch_gemm<EPI_PACK>(xn, w.RT, b.fc1, e, st, w.err));
    }
    AID_TRY(f32_out(b.fc2, reinterpret_cast<uint8_t*>(w.act), true));
  }
  EpiArgs e = epi_zero();
  e.act = ACT_SILU; e.n_valid = H / 2; e.rows_valid = w.B; e.out_packed = w.o1; e.out_kb = ceil_div(H / 2, 64);
  if (chain_ok(s.nf, s.out0, w.RT)) {
    AID_TRY(launch_chain<EPI_PACK>(cs, w.RT, s.nf, s.out0, modln_args(), e, st, w.err));
  } else {
    AID_TRY(modln(s.nf));
    AID_TRY(launch_gemm<EPI_PACK>(xn, w.RT, s.out0, e, st, w.err));
  }
  e = epi_zero();
  e.bias = nullptr;  // output_proj.2 has no bias; launch_gemm substitutes the zero-padded vector
  e.n_valid = s.L; e.rows_valid = w.B;
  e.out_mult = s.out_mult; e.tw_rows = o.tw_rows; e.tw_scalar = o.tw_scalar;
  e.do_step = o.do_step; e.z_in = o.z_in; e.eps = o.eps;
  e.philox = o.philox; e.philox_draw = o.philox_draw; e.row_offset = o.row_offset;
  e.c_s1 = o.c_s1; e.c_ra = o.c_ra; e.c_c1 = o.c_c1; e.c_c2 = o.c_c2; e.c_sigma = o.c_sigma;
  e.z_out = o.z_out; e.out_packed = o.z_packed_out; e.out_kb = ceil_div(s.L, 64);
  AID_TRY(launch_gemm<EPI_SCORE>(reinterpret_cast<uint8_t*>(w.o1), w.RT, s.out2, e, st, w.err));
  return 0;
}

static int run_cond(const ScoreW& s, ScoreWS& w, bool use_cont, int fixed_row, cudaStream_t st) {
  CondArgs c;
  c.t_sin = w.tsin;
  c.t_cont = use_cont ? w.tcont : nullptr;
  c.cont_flag = use_cont ? w.t_flag : nullptr;
  c.time_scale = s.time_scale;
  c.obs_emb = w.obs_emb;
  c.fixed_row = fixed_row;
  c.ld4 = ceil_div(s.H, TILE_N) * TILE_N / 4;
  c.H = s.H;
  c.rows = w.B;
  c.row_tiles = w.RT;
  c.out_packed = w.csilu;
  c.out_tiled = nullptr;
  k_cond<<<ew_grid((size_t)w.RT * c.ld4 * TILE_M), 256, 0, st>>>(c);
  AID_LAUNCH_CHECK("k_cond");
  return 0;
}

extern "C" int32_t aid_score_forward(const AidScoreDims* dims, const void* packed, void* workspace,
                                     size_t workspace_bytes, const float* z_t, const float* time,
                                     const float* observation, int32_t batch, int32_t continuous,
                                     float* score_out, void* stream) {
  AID_TRY(score_dims_ok(dims));
  if (!packed || !workspace || !z_t || !time || !score_out) return fail("aid_score_forward: null pointer");
  if (batch <= 0) return fail("aid_score_forward: batch must be positive");
  ScoreW s;
  score_layout(dims, const_cast<void*>(packed), s);
  ScoreWS w;
  score_ws_layout(s, batch, batch, workspace, w);
  if (workspace_bytes < w.total) return fail("aid_score_forward: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AID_CHECK(cudaMemsetAsync(w.err, 0, sizeof(int), st));
  k_time_args<<<ceil_div(batch, 256), 256, 0, st>>>(time, batch, continuous, w.t_sin_arg, w.t_norm, w.t_flag, w.t_w);
  AID_LAUNCH_CHECK("k_time_args");
  AID_TRY(run_time_table(s, w, continuous != 0, st));
  AID_TRY(run_obs_encoder(s, w, observation, st));
  AID_TRY(run_cond(s, w, continuous != 0, -1, st));
  const int kbl = ceil_div(s.L, TILE_K);
  k_pack_rows<<<ew_grid((size_t)w.RT * kbl * 1024), 256, 0, st>>>(z_t, batch, s.L, s.L, w.zp, w.RT, kbl, MAP_PLAIN, 0, 1);
  AID_LAUNCH_CHECK("k_pack_rows(z)");
  StepOut o;
  o.do_step = 0;
  o.tw_rows = continuous ? w.t_w : nullptr;
  o.tw_scalar = 1.0f;
  o.z_out = score_out;
  return run_score_trunk(s, w, o, st);
}

// ------------------------------------------------------------------------------------------
// Reverse diffusion
// Per-step time arguments of the score net's two branches (models/score_networks.py:121-137), same
// fp32 operations as the reference's tensor ops.  The step times travel as kernel parameters (no
// host-memory copy), which keeps aid_sample capturable in a CUDA graph.
struct StepTimes {
  float t[128];
};
__global__ void k_step_time_args(const StepTimes st, int n, int base, float* __restrict__ sin_arg,
                                 float* __restrict__ t_norm, float* __restrict__ flag) {
  const int i = threadIdx.x;
  if (i >= n) return;
  const float ti = st.t[i];
  const bool cont = (ti <= 1.0f) && (ti >= 0.0f);
  sin_arg[base + i] = cont ? __fmul_rn(ti, 999.0f) : ti;
  t_norm[base + i] = cont ? __fsub_rn(__fmul_rn(2.0f, ti), 1.0f) : 0.f;
  flag[base + i] = cont ? 1.f : 0.f;
}

// z_T (or any [rows, cols] block of standard normals) from the Philox stream of philox.cuh
__global__ void k_philox_normal(const PhiloxState* __restrict__ state, unsigned int draw, long long row_offset,
                                float* __restrict__ out, int rows, int cols4) {
  const PhiloxState ps = *state;
  const size_t total = (size_t)rows * cols4;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(idx / cols4), c4 = (int)(idx % cols4);
    reinterpret_cast<float4*>(out)[idx] =
        philox_normal4(ps, draw, (unsigned long long)(row_offset + row), (uint32_t)c4);
  }
}

extern "C" int32_t aid_philox_normal(const void* philox, uint32_t draw, int64_t row_offset, float* out,
                                     int32_t rows, int32_t cols, void* stream) {
  if (!philox || !out) return fail("aid_philox_normal: null pointer");
  if (rows <= 0 || cols <= 0 || (cols & 3)) return fail("aid_philox_normal: rows > 0 and cols a positive multiple of 4");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k_philox_normal<<<ew_grid((size_t)rows * (cols / 4)), 256, 0, st>>>(static_cast<const PhiloxState*>(philox), draw,
                                                                     row_offset, out, rows, cols / 4);
  AID_LAUNCH_CHECK("k_philox_normal");
  return 0;
}

#include "small.inc"

extern "C" int32_t aid_sample_ex(const AidScoreDims* dims, const void* packed, void* workspace,
                                 size_t workspace_bytes, int32_t batch, int32_t n_steps,
                                 const float* step_time_host, const int32_t* step_index_host,
                                 const float* coef_host, int32_t T, const float* observation,
                                 const AidSampleNoise* nz, float* z_out, float* traj_out, void* stream) {
  AID_TRY(score_dims_ok(dims));
  if (!packed || !workspace || !nz || !z_out || !step_time_host || !step_index_host || !coef_host)
    return fail("aid_sample: null pointer");
  if (batch <= 0 || n_steps <= 0 || T <= 0) return fail("aid_sample: batch, n_steps and T must be positive");
  if (!nz->z_init && !nz->philox) return fail("aid_sample: z_init is null and no Philox state was given");
  const PhiloxState* philox = static_cast<const PhiloxState*>(nz->philox);
  ScoreW s;
  score_layout(dims, const_cast<void*>(packed), s);
  ScoreWS w;
  score_ws_layout(s, batch, n_steps, workspace, w);
  if (workspace_bytes < w.total) return fail("aid_sample: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AID_CHECK(cudaMemsetAsync(w.err, 0, sizeof(int), st));

  // Per-step time arguments: the branch is batch-global and the batch is time-constant, so it is
  // resolved per step without a device sync; tw (a by-value kernel parameter of the
  // step's last GEMM) is evaluated on the host in fp32 exactly as the reference's tensor ops would.
  std::vector<float> flag(n_steps), tw(n_steps);
  bool any_cont = false;
  for (int i = 0; i < n_steps; ++i) {
    const float t = step_time_host[i];
    const bool cont = (t <= 1.0f) && (t >= 0.0f);
    if (step_index_host[i] < 0 || step_index_host[i] >= T) return fail("aid_sample: step index out of range");
    if (cont) {
      volatile float d = 1e-5f + t;
      volatile float e = 1.0f / d;
      flag[i] = 1.f; tw[i] = sqrtf(e);
      any_cont = true;
    } else {
      flag[i] = 0.f; tw[i] = 1.f;
    }
  }
  for (int base = 0; base < n_steps; base += 128) {
    StepTimes stt;
    const int n = n_steps - base < 128 ? n_steps - base : 128;
    for (int i = 0; i < 128; ++i) stt.t[i] = i < n ? step_time_host[base + i] : 0.f;
    k_step_time_args<<<1, 128, 0, st>>>(stt, n, base, w.t_sin_arg, w.t_norm, w.t_flag);
    AID_LAUNCH_CHECK("k_step_time_args");
  }
  AID_TRY(run_time_table(s, w, any_cont, st));
  AID_TRY(run_obs_encoder(s, w, observation, st));

  const size_t zl = (size_t)batch * s.L;
  const float* z_init = nz->z_init;
  if (!z_init) {
    // z_T ~ N(0, I) from the Philox stream (draw 0), written where the first step expects z
    float* z0 = traj_out ? traj_out : z_out;
    k_philox_normal<<<ew_grid(zl / 4), 256, 0, st>>>(philox, 0u, nz->row_offset, z0, batch, s.L / 4);
    AID_LAUNCH_CHECK("k_philox_normal");
    z_init = z0;
  } else if (traj_out) {
    AID_CHECK(cudaMemcpyAsync(traj_out, z_init, zl * 4, cudaMemcpyDeviceToDevice, st));
  }
  if (small_path_ok(s, batch))
    return sample_small(s, w, batch, n_steps, flag, tw, step_index_host, coef_host, T, observation != nullptr,
                        any_cont, z_init, nz, z_out, traj_out, st);
  const int kbl = ceil_div(s.L, TILE_K);
  k_pack_rows<<<ew_grid((size_t)w.RT * kbl * 1024), 256, 0, st>>>(z_init, batch, s.L, s.L, w.zp, w.RT, kbl, MAP_PLAIN, 0, 1);
  AID_LAUNCH_CHECK("k_pack_rows(z)");

  // Without a trajectory buffer every step after the first updates z_out in place: a thread of the
  // step epilogue reads its own (row, 32 columns) of z_in before it writes the same elements.
  const float* z_cur = z_init;
  int draw = 0;
  for (int i = 0; i < n_steps; ++i) {
    const int ti = step_index_host[i];
    AID_TRY(run_cond(s, w, flag[i] != 0.f, i, st));
    StepOut o;
    o.do_step = 1;
    o.tw_rows = nullptr;
    o.tw_scalar = tw[i];
    o.z_in = z_cur;
    const bool noisy = ti != 0 && !nz->deterministic;
    o.eps = (noisy && nz->noise) ? nz->noise + (size_t)draw * zl : nullptr;
    o.philox = (noisy && !nz->noise) ? philox : nullptr;
    o.philox_draw = 1u + (unsigned)i;
    o.row_offset = nz->row_offset;
    if (noisy) ++draw;
    o.c_s1 = coef_host[0 * T + ti];
    o.c_ra = coef_host[1 * T + ti];
    o.c_c1 = coef_host[2 * T + ti];
    o.c_c2 = coef_host[3 * T + ti];
    o.c_sigma = coef_host[4 * T + ti];
    o.z_out = traj_out ? traj_out + (size_t)(i + 1) * zl : z_out;
    o.z_packed_out = w.zp;
    AID_TRY(run_score_trunk(s, w, o, st));
    z_cur = o.z_out;
  }
  if (traj_out) AID_CHECK(cudaMemcpyAsync(z_out, z_cur, zl * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

extern "C" int32_t aid_sample(const AidScoreDims* dims, const void* packed, void* workspace,
                              size_t workspace_bytes, int32_t batch, int32_t n_steps,
                              const float* step_time_host, const int32_t* step_index_host,
                              const float* coef_host, int32_t T, const float* observation,
                              const float* z_init, const float* noise, float* z_out, float* traj_out,
                              void* stream) {
  if (!z_init) return fail("aid_sample: null pointer");
  AidSampleNoise nz;
  nz.z_init = z_init;
  nz.noise = noise;
  nz.philox = nullptr;
  nz.row_offset = 0;
  nz.deterministic = noise ? 0 : 1;
  return aid_sample_ex(dims, packed, workspace, workspace_bytes, batch, n_steps, step_time_host, step_index_host,
                       coef_host, T, observation, &nz, z_out, traj_out, stream);
}

// ------------------------------------------------------------------------------------------
// Primitive for tests
__global__ void k_unpack_rows(const __nv_bfloat16* __restrict__ src, int rows, int cols, int kb_total,
                              float* __restrict__ out, int ld) {
  size_t total = (size_t)rows * cols;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int row = (int)(idx / cols), c = (int)(idx % cols);
    int rt = row >> 7, r = row & 127, kb = c >> 6, cc = c & 63;
    const __nv_bfloat16* tile = src + ((size_t)rt * kb_total + kb) * TILE_ELEMS;
    out[(size_t)row * ld + c] = op16_to_float(tile + packed_off(r, cc));
  }
}

extern "C" size_t aid_linear_workspace_bytes(int32_t M, int32_t N, int32_t K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  Arena a(nullptr);
  PLin p;
  plin_shape(p, N, K);
  a.take<int>(1024);
  a.take<uint8_t>(packed_tiles_bytes(M, K));
  a.take<uint8_t>(plin_w_bytes(p));
  a.take<uint8_t>(plin_b_bytes(p));
  a.take<uint8_t>(packed_tiles_bytes(M, p.n_tiles * TILE_N));
  return align_up(a.off, 1024);
}

extern "C" int32_t aid_linear(const float* x, const float* wt, const float* bias, float* y, int32_t M,
                              int32_t N, int32_t K, int32_t act, int32_t via_packed, void* workspace,
                              size_t workspace_bytes, void* stream) {
  if (!x || !wt || !y || !workspace) return fail("aid_linear: null pointer");
  if (M <= 0 || N <= 0 || K <= 0) return fail("aid_linear: bad shape");
  if (workspace_bytes < aid_linear_workspace_bytes(M, N, K)) return fail("aid_linear: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Arena a(workspace);
  PLin p;
  plin_shape(p, N, K);
  int* err = a.take<int>(1024);
  __nv_bfloat16* xp = a.take<__nv_bfloat16>(packed_tiles_bytes(M, K));
  p.w = a.take<uint8_t>(plin_w_bytes(p));
  p.b = a.take<float>(plin_b_bytes(p));
  __nv_bfloat16* yp = a.take<__nv_bfloat16>(packed_tiles_bytes(M, p.n_tiles * TILE_N));
  AID_CHECK(cudaMemsetAsync(err, 0, sizeof(int), st));
  const int rt = ceil_div(M, TILE_M);
  k_pack_rows<<<ew_grid((size_t)rt * p.kb * 1024), 256, 0, st>>>(x, M, K, K, xp, rt, p.kb, MAP_PLAIN, 0, 1);
  AID_LAUNCH_CHECK("k_pack_rows(x)");
  AID_TRY(pack_linear(p, wt, bias, MAP_PLAIN, 0, st));
  EpiArgs e = epi_zero();
  e.act = act; e.n_valid = N; e.rows_valid = M;
  if (via_packed) {
    e.out_packed = yp; e.out_kb = p.n_tiles * 2;
    AID_TRY(launch_gemm<EPI_PACK>(reinterpret_cast<uint8_t*>(xp), rt, p, e, st, err));
    k_unpack_rows<<<ew_grid((size_t)M * N), 256, 0, st>>>(yp, M, N, p.n_tiles * 2, y, N);
    AID_LAUNCH_CHECK("k_unpack_rows");
  } else {
    e.out_rm = y; e.ld_rm = N;
    AID_TRY(launch_gemm<EPI_F32>(reinterpret_cast<uint8_t*>(xp), rt, p, e, st, err));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// General strided GEMM for the training graph: out[M,N] = A[M,K] * B[N,K]^T (+ bias[N]).
// Element (i,k) of A is a[i*a_rs + k*a_cs], likewise B, so forward (x W^T), dgrad (dY W) and
// wgrad (dY^T X) are the same call with different strides.  Operands are rounded to bf16 when
// packed; accumulation is fp32 in TMEM.
// `triple` != 0 packs the bf16x3 split of the operand along K (three regions of kreg columns,
// kreg = K rounded up to 64):  A side (1): [hi | hi | lo],  B side (2): [hi | lo | hi], with
// hi = bf16(x), lo = bf16(x - hi).  One GEMM over 3*kreg columns then evaluates
// hi*hi + hi*lo + lo*hi: the fp32 product to ~2^-16 relative, on the bf16 tensor pipe.
__global__ void k_pack_strided(const float* __restrict__ src, long long rs, long long cs, int rows, int cols,
                               __nv_bfloat16* __restrict__ dst, int row_tiles, int kb_total, int nw,
                               int triple, int kreg) {
  size_t total = (size_t)row_tiles * kb_total * 1024;
  const bool vec = cs == 1 && (rs & 3) == 0 && (reinterpret_cast<unsigned long long>(src) & 15ull) == 0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int rt, kb, r, ch;
    chunk_coords(idx, kb_total, rt, kb, r, ch);
    if (cs == 1) {   // row-major source: neighbouring lanes take neighbouring 32-byte pieces of one row
      ch = (int)(idx & 7);
      r = (int)((idx >> 3) & 127);
    }
    const int row = rt * TILE_M + r;
    int c0 = kb * TILE_K + ch * 8;
    int want_lo = 0;
    if (triple) {
      const int region = c0 / kreg;
      c0 -= region * kreg;
      want_lo = (triple == 1) ? (region == 2) : (region == 1);
    }
    float v[8];
    if (vec && row < rows && c0 + 8 <= cols) {   // row-major, 16-byte aligned rows: two 128-bit loads
      const float4* p = reinterpret_cast<const float4*>(src + (long long)row * rs + c0);
      const float4 a = __ldg(p), b = __ldg(p + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
      v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        v[i] = (row < rows && c0 + i < cols) ? __ldg(src + (long long)row * rs + (long long)(c0 + i) * cs) : 0.f;
    }
    if (want_lo) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] -= op16_round(v[i]);
    }
    store_chunk(dst, rt, kb, kb_total, r, ch, v, nw);
  }
}

// out[m, n] = sum_s partial[s][m][n] (+ bias[n]); partial rows are padded to m_pad per split
__global__ void k_sum_splits(const float* __restrict__ partial, int splits, int m_pad, int M, int N,
                             const float* __restrict__ bias, float* __restrict__ out) {
  const size_t total = (size_t)M * N;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(idx / N), n = (int)(idx % N);
    float acc = bias ? __ldg(bias + n) : 0.f;
    for (int s = 0; s < splits; ++s) acc += partial[((size_t)s * m_pad + m) * N + n];
    out[idx] = acc;
  }
}

static int gemm_nt_splits(int M, int N, int K) {
  // long reductions with few output tiles (weight gradients): spread the K range over the SMs
  const int units = ceil_div(M, TILE_M) * ceil_div(ceil_div(N, TILE_N), 2);
  const int kb = ceil_div(K, TILE_K);
  int s = 1;
  while (units * s * 2 <= num_sms() + num_sms() / 2 && kb % (s * 2) == 0 && kb / (s * 2) >= 16) s *= 2;
  return s;
}

struct GemmNtLayout {
  PLin p;
  int rt, splits;
  int* err;
  __nv_bfloat16* ap;
  float* partial;
  size_t total;
};
static void gemm_nt_layout(int M, int N, int K, int precision, void* ws, GemmNtLayout& g) {
  Arena a(ws);
  const int kreg = ceil_div(K, TILE_K) * TILE_K;
  const int keff = precision ? 3 * kreg : K;
  plin_shape(g.p, N, keff);
  g.rt = ceil_div(M, TILE_M);
  g.splits = gemm_nt_splits(M, N, keff);
  g.err = a.take<int>(1024);
  g.ap = a.take<__nv_bfloat16>(packed_tiles_bytes(M, keff));
  g.p.w = a.take<uint8_t>(plin_w_bytes(g.p));
  g.p.b = a.take<float>(plin_b_bytes(g.p));
  g.partial = g.splits > 1 ? a.take<float>((size_t)g.splits * g.rt * TILE_M * N * sizeof(float)) : nullptr;
  g.total = align_up(a.off, 1024);
}

extern "C" size_t aid_gemm_nt_workspace_bytes(int32_t M, int32_t N, int32_t K, int32_t precision) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  GemmNtLayout g;
  gemm_nt_layout(M, N, K, precision, nullptr, g);
  return g.total;
}

extern "C" int32_t aid_gemm_nt(const float* a, int64_t a_rs, int64_t a_cs, const float* b, int64_t b_rs,
                               int64_t b_cs, const float* bias, float* out, int32_t M, int32_t N, int32_t K,
                               int32_t precision, void* workspace, size_t workspace_bytes, void* stream) {
  if (!a || !b || !out || !workspace) return fail("aid_gemm_nt: null pointer");
  if (M <= 0 || N <= 0 || K <= 0) return fail("aid_gemm_nt: bad shape");
  GemmNtLayout g;
  gemm_nt_layout(M, N, K, precision, workspace, g);
  if (workspace_bytes < g.total) return fail("aid_gemm_nt: workspace too small");
  const int kreg = ceil_div(K, TILE_K) * TILE_K;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AID_CHECK(cudaMemsetAsync(g.err, 0, sizeof(int), st));
  k_pack_strided<<<ew_grid((size_t)g.rt * g.p.kb * 1024), 256, 0, st>>>(a, a_rs, a_cs, M, K, g.ap, g.rt, g.p.kb, 1,
                                                                         precision ? 1 : 0, kreg);
  AID_LAUNCH_CHECK("k_pack_strided(a)");
  k_pack_strided<<<ew_grid((size_t)g.p.n_tiles * g.p.kb * 1024), 256, 0, st>>>(
      b, b_rs, b_cs, N, K, reinterpret_cast<__nv_bfloat16*>(const_cast<uint8_t*>(g.p.w)), g.p.n_tiles, g.p.kb, g.p.nw,
      precision ? 2 : 0, kreg);
  AID_LAUNCH_CHECK("k_pack_strided(b)");
  const int n_pad = g.p.n_tiles * TILE_N;
  const bool direct = g.splits == 1;
  k_pack_bias<<<ceil_div(n_pad, 256), 256, 0, st>>>(direct ? bias : nullptr, direct && bias ? N : 0,
                                                   const_cast<float*>(g.p.b), n_pad, MAP_PLAIN, 0);
  AID_LAUNCH_CHECK("k_pack_bias");
  EpiArgs e = epi_zero();
  e.n_valid = N; e.rows_valid = M; e.ld_rm = N;
  e.out_rm = direct ? out : g.partial;
  AID_TRY(launch_gemm<EPI_F32>(reinterpret_cast<uint8_t*>(g.ap), g.rt, g.p, e, st, g.err, g.splits));
  if (!direct) {
    k_sum_splits<<<ew_grid((size_t)M * N), 256, 0, st>>>(g.partial, g.splits, g.rt * TILE_M, M, N, bias, out);
    AID_LAUNCH_CHECK("k_sum_splits");
  }
  return 0;
}

#include "heads.inc"
#include "epistemic.inc"
#include "train.inc"
#include "belief.inc"
#include "encoder.inc"
#include "conv_train.inc"

// ------------------------------------------------------------------------------------------
extern "C" int32_t aid_abi_version(void) { return AID_ABI_VERSION; }
extern "C" int32_t aid_operand_type(void) {
#ifdef AID_F16
  return 1;
#else
  return 0;
#endif
}
extern "C" const char* aid_last_error(void) { return g_err.c_str(); }
extern "C" int32_t aid_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
extern "C" int64_t aid_launch_count(void) { return g_launches; }
extern "C" void aid_reset_launch_count(void) { g_launches = 0; }
