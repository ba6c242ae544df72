// Latency of a dependent mma.sync.m16n8k16 (bf16, fp32 accumulate) chain on B200, and of two / four
// independent chains interleaved (what the persistent small-batch sampler's K loop can overlap).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/hmma_latency scripts/micro/hmma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float (&d)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int CH>
__global__ void k(long long* out, float* sink, unsigned seed) {
  float acc[CH][4];
  for (int c = 0; c < CH; ++c) for (int i = 0; i < 4; ++i) acc[c][i] = 0.f;
  unsigned a = seed + threadIdx.x, b = seed * 3 + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 256; ++it) {
#pragma unroll
    for (int c = 0; c < CH; ++c) mma(acc[c], a, a ^ 1, a ^ 2, a ^ 3, b, b ^ 5);
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int c = 0; c < CH; ++c) for (int i = 0; i < 4; ++i) s += acc[c][i];
  sink[threadIdx.x] = s;
  if (threadIdx.x == 0) out[0] = t1 - t0;
}
// a burst of 8 dependent MMAs after `gap` idle cycles (the persistent sampler issues a handful of MMAs per
// phase, then waits ~9k cycles at barriers): does the burst pay a wake-up cost?
__global__ void kgap(long long* out, float* sink, unsigned seed, int gap, int fma_instead) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  unsigned a = seed + threadIdx.x, b = seed * 3 + threadIdx.x;
  long long total = 0;
  for (int it = 0; it < 64; ++it) {
    long long t0 = clock64();
    while (clock64() - t0 < gap) { }
    __syncwarp();
    long long t1 = clock64();
    if (fma_instead) {
#pragma unroll
      for (int i = 0; i < 64; ++i) acc[i & 3] = fmaf(acc[i & 3], 1.0001f, __uint_as_float((a + i) & 0x3f800000u));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) mma(acc, a, a ^ 1, a ^ 2, a ^ 3, b, b ^ 5);
    }
    float s = acc[0] + acc[1] + acc[2] + acc[3];
    if (s == 123.456f) sink[0] = s;
    long long t2 = clock64();
    total += t2 - t1;
  }
  sink[threadIdx.x] = acc[0];
  if (threadIdx.x == 0) out[0] = total;
}
int main() {
  long long* out; float* sink; long long h;
  cudaMalloc(&out, 8); cudaMalloc(&sink, 4096);
  k<1><<<1, 32>>>(out, sink, 1); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("1 chain : %.1f cycles per dependent mma\n", h / 256.0);
  k<2><<<1, 32>>>(out, sink, 1); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("2 chains: %.1f cycles per pair\n", h / 256.0);
  k<4><<<1, 32>>>(out, sink, 1); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("4 chains: %.1f cycles per four\n", h / 256.0);
  k<8><<<1, 32>>>(out, sink, 1); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("8 chains: %.1f cycles per eight\n", h / 256.0);
  for (int gap : {0, 500, 2000, 8000, 30000}) {
    kgap<<<1, 32>>>(out, sink, 1, gap, 0); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("8 dependent mma after %5d idle cycles: %.0f cycles per burst\n", gap, h / 64.0);
  }
  kgap<<<1, 32>>>(out, sink, 1, 8000, 1); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("64 FFMA (16 deep x 4) after 8000 idle cycles: %.0f cycles per burst\n", h / 64.0);
  return 0;
}
