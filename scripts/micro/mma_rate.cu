// Microbenchmark: tcgen05.mma issue/throughput for cta_group::1, M=128, N in {128,256},
// same accumulator vs rotating accumulators.  Operands are whatever is in shared memory.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../active_inference_diffusion_b200/csrc/ptx.cuh"
using namespace aid;

// K-major, no swizzle ("interleaved"): core matrix = 8 rows x 16 B contiguous (128 B); LBO = byte
// stride between core matrices adjacent in K, SBO = between 8-row groups.
__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__global__ void __launch_bounds__(128, 1) k(int n_mma, int N, int rot, int kdep, long long* out, int nosw) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tbase), 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) {
    uint32_t idesc = umma_idesc_bf16(128, N);
    uint64_t ad[4], bd[4];
    for (int j = 0; j < 4; ++j) {
      if (nosw) {
        ad[j] = desc_nosw(base + (kdep ? j * 4096 : 0), 2048, 128);
        bd[j] = desc_nosw(base + 32768 + (kdep ? j * (N * 32) : 0), N * 16, 128);
      } else {
        ad[j] = umma_desc_sw128(base + (kdep ? j * 32 : 0));
        bd[j] = umma_desc_sw128(base + 32768 + (kdep ? j * 32 : 0));
      }
    }
    const uint32_t d0 = tbase, d1 = tbase + (rot ? N % 512 : 0);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
      for (int j = 0; j < 4; ++j) umma_bf16(d0, ad[j], bd[j], idesc, 1);
#pragma unroll
      for (int j = 0; j < 4; ++j) umma_bf16(d1, ad[j], bd[j], idesc, 1);
    }
    long long t1 = clock64();
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, nullptr, 0);
    long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tbase, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 2 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  int n = 4096;
  struct { int N, rot, kdep, grid; } cfg[] = {{128, 0, 1, 1}, {128, 4, 1, 1}, {256, 0, 1, 1}, {256, 2, 1, 1},
                                             {64, 0, 1, 1}, {128, 0, 1, 148}, {256, 0, 1, 148}, {256, 2, 1, 148}, {128, 4, 1, 148}};
  for (int nosw = 0; nosw < 2; ++nosw)
  for (auto c : cfg) {
    for (int rep = 0; rep < 2; ++rep) {
      k<<<c.grid, 128, 100 * 1024>>>(n, c.N, c.rot, c.kdep, d, nosw);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    }
    long long h[296]; cudaMemcpy(h, d, c.grid * 16, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < c.grid; ++i) mx = h[2 * i + 1] > mx ? h[2 * i + 1] : mx;
    printf("nosw=%d N=%3d rot=%d grid=%3d: issue %.1f cyc/mma, complete %.1f cyc/mma (max over CTAs %.1f); ideal %d\n", nosw, c.N, c.rot, c.grid,
           (double)h[0] / n, (double)h[1] / n, (double)mx / n, c.N / 2);
  }
  return 0;
}
