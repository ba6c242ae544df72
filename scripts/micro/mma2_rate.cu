// Microbenchmark: tcgen05.mma.cta_group::2 (M=256 over a CTA pair) issue rate, K-major no-swizzle.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../active_inference_diffusion_b200/csrc/gemm2.cuh"
using namespace aid;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k2(int n_mma, int N, long long* out, int commit_mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[8];
  __shared__ uint32_t tbase;
  uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bar2[i]), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc2(smem_u32(&tbase), 512); tmem_relinquish2(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); cluster_sync_all(); tc_fence_after();
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0 && rank == 0) {
    uint32_t idesc = umma_idesc_bf16(256, N);
    uint64_t ad[4], bd[4];
    for (int j = 0; j < 4; ++j) {
      ad[j] = umma_desc_kmajor(base + j * 4096, 2048);
      bd[j] = umma_desc_kmajor(base + 32768 + j * (N / 2 * 32), N / 2 * 16);
    }
    long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
      for (int j = 0; j < 4; ++j) umma2_bf16(tbase, ad[j], bd[j], idesc, 1);
      if (commit_mode == 1) umma2_commit_both(smem_u32(&bar2[(i >> 2) & 7]));
      if (commit_mode == 2) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[(i >> 2) & 7])) : "memory");
#pragma unroll
      for (int j = 0; j < 4; ++j) umma2_bf16(tbase + (N % 512), ad[j], bd[j], idesc, 1);
      if (commit_mode == 1) umma2_commit_both(smem_u32(&bar2[((i >> 2) + 1) & 7]));
      if (commit_mode == 2) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[((i >> 2) + 1) & 7])) : "memory");
    }
    long long t1 = clock64();
    umma2_commit_both(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, nullptr, 0);
    long long t2 = clock64();
    out[(blockIdx.x >> 1) * 2] = t1 - t0;
    out[(blockIdx.x >> 1) * 2 + 1] = t2 - t0;
  }
  tc_fence_before(); cluster_sync_all();
  if (threadIdx.x < 32) tmem_dealloc2(tbase, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 2 * 8);
  cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  int n = 4096;
  for (int commit_mode : {0, 1, 2})
  for (int N : {256})
    for (int grid : {2, 148}) {
      for (int rep = 0; rep < 2; ++rep) {
        k2<<<grid, 128, 100 * 1024>>>(n, N, d, commit_mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < grid / 2; ++i) mx = h[2 * i + 1] > mx ? h[2 * i + 1] : mx;
      printf("commit_mode=%d cta_group::2 M=256 N=%3d grid=%3d: issue %.1f cyc/mma, complete %.1f (max %.1f); 1-CTA-rate ideal %d\n", commit_mode, N, grid,
             (double)h[0] / n, (double)h[1] / n, (double)mx / n, N / 2);
    }
  return 0;
}
