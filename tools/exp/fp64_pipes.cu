// Micro-benchmark: can the FP64 vector pipe (DFMA) and the FP64 tensor sub-pipe
// (DMMA.8x8x4) of a B200 SM run concurrently?  Three kernels on registers only:
//   mode 0: every warp issues DMMA    mode 1: every warp issues DFMA
//   mode 2: even warps DMMA, odd warps DFMA (same warp counts as modes 0/1 combined)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp/fp64_pipes tools/exp/fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>
__global__ void __launch_bounds__(512) k(double* out, int iters, int dmma_warps_of_4) {
  const int warp = threadIdx.x >> 5;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
  bool do_dmma = MODE == 0 || (MODE == 2 && (warp & 3) < dmma_warps_of_4);
  if (do_dmma) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) dmma(c[i], c[i + 1], a, b);
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) c[i] = fma(a, c[i], b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  double* out;
  cudaMalloc(&out, 148 * 4 * 512 * sizeof(double));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int threads : {256, 512}) {
    for (int ctas : {1, 2}) {
      for (int mode = 0; mode < 5; ++mode) {
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0);
          if (mode == 0) k<0><<<148 * ctas, threads>>>(out, iters, 0);
          else if (mode == 1) k<1><<<148 * ctas, threads>>>(out, iters, 0);
          else k<2><<<148 * ctas, threads>>>(out, iters, mode - 1);   // 1,2,3 of 4 warps DMMA
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          cudaEventElapsedTime(&ms, e0, e1);
        }
        const double warps = 148.0 * ctas * threads / 32;
        double fl_dmma = 0, fl_dfma = 0;
        // DMMA m8n8k4: 8*8*4*2 = 512 flop per warp instr, 8 per iter; DFMA: 32*2 flop, 16 per iter
        double frac = mode == 0 ? 1.0 : mode == 1 ? 0.0 : (mode - 1) / 4.0;
        fl_dmma = warps * frac * iters * 8.0 * 512.0;
        fl_dfma = warps * (1 - frac) * iters * 16.0 * 64.0;
        printf("threads %d ctas/SM %d mode %d (dmma warp frac %.2f): %.3f ms  DMMA %.2f TF/s  DFMA %.2f TF/s  total %.2f\n",
               threads, ctas, mode, frac, ms, fl_dmma / ms / 1e9, fl_dfma / ms / 1e9,
               (fl_dmma + fl_dfma) / ms / 1e9);
      }
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
