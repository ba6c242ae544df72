// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput and latency per SM, 32x32b.x32
// (4 KiB per warp instruction) and .x16, for 4 / 8 / 16 resident warps.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../active_inference_diffusion_b200/csrc/ptx.cuh"
using namespace aid;

template <int X>
__global__ void __launch_bounds__(512, 1) k(int iters, int depth, long long* out) {
  __shared__ uint32_t tbase;
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tbase), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const int warp = threadIdx.x >> 5;
  const uint32_t t0a = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (X == 32) {
      uint32_t r[32];
      tmem_ld32(t0a + ((i * 32) & 511), r);
      if (depth == 2) { uint32_t r2[32]; tmem_ld32(t0a + ((i * 32 + 256) & 511), r2); tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc += r2[j]; }
      else tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += r[j];
    } else {
      uint32_t r[16];
      tmem_ld16(t0a + ((i * 16) & 511), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) acc += r[j];
    }
  }
  long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
  if (acc == 0x12345) out[0] = 0;
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tbase, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 16 * 8);
  const int iters = 2048;
  for (int warps : {1, 4, 8, 16}) {
    for (int mode = 0; mode < 3; ++mode) {   // 0: x32 depth1, 1: x32 depth2, 2: x16
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 2) k<16><<<1, warps * 32>>>(iters, 1, d);
        else k<32><<<1, warps * 32>>>(iters, mode + 1, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bytes = (double)iters * warps * (mode == 2 ? 2048 : 4096) * (mode == 1 ? 2 : 1);
      printf("warps=%2d %s: %.1f cyc/iter/warp, %.1f B/cyc/SM\n", warps,
             mode == 0 ? "x32 depth1" : mode == 1 ? "x32 depth2" : "x16 depth1", (double)mx / iters, bytes / mx);
    }
  }
  return 0;
}
