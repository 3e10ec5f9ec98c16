// chimera-b200: shared device helpers for the sm_100a kernels.
//
// Arithmetic contract (see DESIGN.md "Numerics"): everything that decides an
// INTEGER result (cell indices) or feeds the shape factors is evaluated with
// explicit round-to-nearest, non-contracted operations, i.e. exactly like the
// reference kernels compiled without FMA contraction (oracle/_ref).  The .cu
// files holding per-particle physics are additionally built with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define CHB_OK 0
#define CHB_ERR_ARG (-1)
#define CHB_ERR_WORKSPACE (-2)

#define CHB_RETURN_LAST_ERROR()                     \
  do {                                              \
    cudaError_t e__ = cudaGetLastError();           \
    return e__ == cudaSuccess ? CHB_OK : (int)e__;  \
  } while (0)

namespace chb {

constexpr int kSMs = 148;  // B200

struct GridGeom {
  const double* __restrict__ xmin;    // DataDev['Xmin']  (device scalar, moves with the window)
  const double* __restrict__ dx_inv;  // DataDev['dx_inv']
  const double* __restrict__ rmin;    // DataDev['Rmin']
  const double* __restrict__ dr_inv;  // DataDev['dr_inv']
  uint32_t Nx, Nr;
};

struct GridVals {
  double xmin, dx_inv, rmin, dr_inv;
  int Nx, Nr;
};

__device__ __forceinline__ GridVals load_geom(const GridGeom& g) {
  GridVals v;
  v.xmin = __ldg(g.xmin);
  v.dx_inv = __ldg(g.dx_inv);
  v.rmin = __ldg(g.rmin);
  v.dr_inv = __ldg(g.dr_inv);
  v.Nx = (int)g.Nx;
  v.Nr = (int)g.Nr;
  return v;
}

// (int)floor(v) with the x86 cvttsd2si convention for NaN (-> INT_MIN), so that
// NaN coordinates land in the trash bin exactly as in the host-compiled reference.
__device__ __forceinline__ int floor_to_int(double v) {
  return (v == v) ? (int)floor(v) : (int)0x80000000;
}

// r, ix, ir of particles_generic.cl:102-107 / grid_deposit_m1.cl:380-384.
__device__ __forceinline__ void cell_coords(double x, double y, double z,
                                            const GridVals& g, double& r,
                                            int& ix, int& ir) {
  r = __dsqrt_rn(__dadd_rn(__dmul_rn(y, y), __dmul_rn(z, z)));
  ix = floor_to_int(__dmul_rn(__dsub_rn(x, g.xmin), g.dx_inv));
  ir = floor_to_int(__dmul_rn(__dsub_rn(r, g.rmin), g.dr_inv));
}

// Cell id with the trash bin (particles_generic.cl:109-124).
__device__ __forceinline__ uint32_t cell_index(double x, double y, double z,
                                               const GridVals& g) {
  double r;
  int ix, ir;
  cell_coords(x, y, z, g, r, ix, ir);
  const int Nx_loc = g.Nx - 1, Nr_loc = g.Nr - 1;
  if (ix > 0 && ix < Nx_loc - 1 && ir < Nr_loc - 1 && ir >= 0)
    return (uint32_t)(ix + ir * Nx_loc);
  return (uint32_t)(Nr_loc * Nx_loc);
}

// Run-length aggregation inside a warp: lanes holding the same key as their
// left neighbour form a run.  Returns the lane of the run head, the rank of this
// lane inside the run and the run length.  Inactive lanes (valid=false) are
// their own runs and must be ignored by the caller.
__device__ __forceinline__ void warp_runs(uint32_t key, bool valid, int lane,
                                          int& head_lane, int& rank, int& len) {
  const unsigned full = 0xffffffffu;
  uint32_t prev = __shfl_up_sync(full, key, 1);
  bool pvalid = __shfl_up_sync(full, (int)valid, 1);
  bool head = (lane == 0) || (key != prev) || !valid || !pvalid;
  unsigned heads = __ballot_sync(full, head);
  unsigned below = heads & (0xffffffffu >> (31 - lane));  // heads at lanes <= lane
  head_lane = 31 - __clz(below);
  rank = lane - head_lane;
  unsigned above = (lane == 31) ? 0u : (heads >> (lane + 1));  // heads at lanes > lane
  int next = above ? (lane + 1 + (__ffs(above) - 1)) : 32;
  len = next - head_lane;
}

// One L2 atomic per run of equal cells inside a warp (storage is kept nearly
// cell-sorted, so a warp of 32 particles spans only a few cells).
__device__ __forceinline__ void histogram_add(uint32_t cell, bool valid,
                                              uint32_t* __restrict__ sum_in_cell) {
  const int lane = threadIdx.x & 31;
  int head, rank, len;
  warp_runs(cell, valid, lane, head, rank, len);
  if (valid && rank == 0) atomicAdd(&sum_in_cell[cell], (uint32_t)len);
}

// 8-byte asynchronous global -> shared copy (LDGSTS); the copy lands without
// occupying a register, so many particles' attributes can be in flight per thread.
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

__device__ __forceinline__ void red_add_f64(double* addr, double v) {
  atomicAdd(addr, v);  // result unused -> RED.E.ADD.F64
}

}  // namespace chb
