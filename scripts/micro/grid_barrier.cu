// Cost of a grid-wide barrier among 128 co-resident CTAs on B200, and of the alternatives the
// persistent small-batch sampler (csrc/small.inc) could use instead.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/grid_barrier scripts/micro/grid_barrier.cu
// Variants (cycles per round, CTA 0 / last CTA):
//   0  counter: bar.sync, thread 0 {red.release.gpu, relaxed poll, fence.acq_rel.gpu}, bar.sync
//   1  counter without the trailing fence (consumer loads bypass L1)
//   2  counter, data store + __threadfence by every thread before (what a real phase does)
//   3  flag-in-data ("LL"): every CTA stores {value, epoch} pairs, every CTA polls all pairs (no fence at all)
//   4  pure L2 round trip: dependent ld.cg chain (latency reference)
//   5  MEMBAR.ALL.GPU alone (thread 0), no traffic
//   6  variant 2 with the arrivals spread over 8 counters on different L2 slices (what small.inc uses)
//   7  flag-in-data done right: every CTA stores 4 {value, epoch} pairs, every thread of every CTA polls ONE pair
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512, 1) k(int variant, int rounds, unsigned* counter, uint2* ll, float* data,
                                             long long* out, const int* chain) {
  extern __shared__ unsigned char smem[];
  const int G = gridDim.x;
  long long t0 = clock64();
  unsigned acc = 0;
  if (variant <= 2) {
    for (int r = 1; r <= rounds; ++r) {
      if (variant == 2) {
        data[(size_t)blockIdx.x * 512 + threadIdx.x] = (float)r;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned v;
        do {
          asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while ((int)(v - (unsigned)r * G) < 0);
        if (variant != 1) asm volatile("fence.acq_rel.gpu;" ::: "memory");
      }
      __syncthreads();
      if (variant == 2) acc += (unsigned)__ldcg(data + (size_t)((blockIdx.x + 1) % G) * 512 + threadIdx.x);
    }
  } else if (variant == 6) {
    for (int r = 1; r <= rounds; ++r) {
      data[(size_t)blockIdx.x * 512 + threadIdx.x] = (float)r;
      __syncthreads();
      if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter + (blockIdx.x & 7) * 320) : "memory");
        for (;;) {
          unsigned v[8], sum = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v[i]) : "l"(counter + i * 320) : "memory");
#pragma unroll
          for (int i = 0; i < 8; ++i) sum += v[i];
          if ((int)(sum - (unsigned)r * G) >= 0) break;
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
      }
      __syncthreads();
      acc += (unsigned)__ldcg(data + (size_t)((blockIdx.x + 1) % G) * 512 + threadIdx.x);
    }
  } else if (variant == 7) {
    for (int r = 1; r <= rounds; ++r) {
      if (threadIdx.x < 4) {
        asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(ll + (size_t)blockIdx.x * 4 + threadIdx.x), "r"(r * 7 + threadIdx.x), "r"((unsigned)r) : "memory");
      }
      uint2 v;
      int spins = 0;
      do {   // epochs are monotonic: a later one also ends the wait (the writer may be a round ahead)
        asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(ll + threadIdx.x) : "memory");
      } while ((int)(v.y - (unsigned)r) < 0 && ++spins < (1 << 20));
      acc += v.x;
      __syncthreads();
    }
  } else if (variant == 3) {
    // every CTA publishes 64 pairs; every CTA reads all G * 64 pairs (512 threads: G*64/512 pairs each)
    for (int r = 1; r <= rounds; ++r) {
      if (threadIdx.x < 64) {
        uint2 v = make_uint2(r * 7 + threadIdx.x, (unsigned)r);
        asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(ll + (size_t)blockIdx.x * 64 + threadIdx.x), "r"(v.x), "r"(v.y) : "memory");
      }
      for (int i = threadIdx.x; i < G * 64; i += 512) {
        uint2 v;
        int spins = 0;
        do {
          asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(ll + i) : "memory");
        } while ((int)(v.y - (unsigned)r) < 0 && ++spins < (1 << 20));
        acc += v.x;
      }
      __syncthreads();   // the slab is complete in this CTA (the real kernel syncs before its MMAs too)
    }
  } else if (variant == 4) {
    int idx = blockIdx.x;
    if (threadIdx.x == 0)
      for (int r = 0; r < rounds; ++r) idx = __ldcg(chain + idx);
    acc = idx;
  } else {
    for (int r = 0; r < rounds; ++r) {
      if (threadIdx.x == 0) asm volatile("fence.acq_rel.gpu;" ::: "memory");
      __syncthreads();
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0xdeadbeef) out[G] = acc;
}

int main() {
  const int G = 128, rounds = 2000;
  unsigned* counter; uint2* ll; float* data; long long* out; int* chain;
  cudaMalloc(&counter, 16384); cudaMalloc(&ll, G * 64 * sizeof(uint2)); cudaMalloc(&data, G * 512 * 4);
  cudaMalloc(&out, (G + 1) * 8); cudaMalloc(&chain, 1 << 22);
  int* hc = (int*)malloc(1 << 22);
  for (int i = 0; i < (1 << 20); ++i) hc[i] = (int)(((long long)i * 40503 + 12345) % (1 << 20));
  cudaMemcpy(chain, hc, 1 << 22, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  for (int variant = 0; variant <= 7; ++variant) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(counter, 0, 16384); cudaMemset(ll, 0, G * 64 * sizeof(uint2));
      k<<<G, 512, 160 * 1024>>>(variant, rounds, counter, ll, data, out, chain);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d: %s\n", variant, cudaGetErrorString(e)); return 1; }
    }
    long long h[G + 1];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("variant %d: %.0f / %.0f cycles per round (CTA 0 / CTA %d)\n", variant, (double)h[0] / rounds,
           (double)h[G - 1] / rounds, G - 1);
  }
  return 0;
}
