// Micro-benchmark: throughput of the legacy warp-level int8 MMA (mma.sync.m16n8k32.s8)
// on B200, register-only.  Question: is an Ozaki-style int8-slice emulation of the FP64
// DHT contraction worth building on mma.sync, or only on tcgen05?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp/imma_peak tools/exp/imma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void imma(int (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3,
                                     unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(512) k(int* out, int iters) {
  unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  int c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = i + j;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) imma(c[i], a0, a1, a2, a3, b0, b1);
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int* out;
  cudaMalloc(&out, 148 * 4 * 512 * sizeof(int));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int threads : {128, 256, 512}) {
    for (int ctas : {1, 2, 4}) {
      if (threads * ctas > 2048) continue;
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k<<<148 * ctas, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
      }
      const double warps = 148.0 * ctas * threads / 32;
      const double ops = warps * iters * 8.0 * (16.0 * 8 * 32 * 2);
      printf("threads %d ctas/SM %d: %.3f ms  %.1f TOPS (int8 mma.sync m16n8k32)\n", threads, ctas, ms,
             ops / ms / 1e9);
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
