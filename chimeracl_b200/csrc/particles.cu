// chimera-b200 particle kernels: coordinate push, cell index + histogram,
// exclusive scan of the histogram, stable counting-sort scatter, particle align.
//
// Replaces (behaviour, not code) the reference's
//   chimeraCL/kernels/particles_generic.cl: push_xyz :129-153,
//   index_and_sum_in_cell :88-126, sort :186-201, data_align_dbl :156-169
//   + pyopencl.array.cumsum (methods/particles_methods_cl.py:303-311).
//
// All kernels are HBM-streaming integer/FP64 work (no tensor-core shape):
// coalesced 8-byte SoA accesses, 4 particles in flight per thread, warp-level
// run aggregation so that a warp issues one L2 atomic per distinct cell run
// instead of one per particle, grids sized in multiples of the 148 SMs.
#include "common.cuh"
#include "../../include/chimera_b200.h"

namespace chb {

constexpr int kBlock = 256;
constexpr int kIlp = 4;

static inline int stream_grid(uint64_t n, int per_block, int ctas_per_sm) {
  uint64_t need = (n + per_block - 1) / per_block;
  uint64_t cap = (uint64_t)kSMs * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ------------------------------------------------------------------ push_xyz
// x += (dt*g_inv)*px, non-contracted (bit-exact with the reference arithmetic).
__device__ __forceinline__ void push_one(double& x, double& y, double& z,
                                         double px, double py, double pz,
                                         double gi, double dt) {
  double dt_g = __dmul_rn(dt, gi);
  x = __dadd_rn(x, __dmul_rn(px, dt_g));
  y = __dadd_rn(y, __dmul_rn(py, dt_g));
  z = __dadd_rn(z, __dmul_rn(pz, dt_g));
}

__global__ void __launch_bounds__(kBlock)
push_xyz_kernel(double* __restrict__ x, double* __restrict__ y, double* __restrict__ z,
                const double* __restrict__ px, const double* __restrict__ py,
                const double* __restrict__ pz, const double* __restrict__ g_inv,
                const double* __restrict__ dt_p, uint32_t n) {
  const double dt = __ldg(dt_p);
  const uint32_t stride = gridDim.x * kBlock * kIlp;
  for (uint32_t base = blockIdx.x * kBlock * kIlp + threadIdx.x; base < n; base += stride) {
    double xv[kIlp], yv[kIlp], zv[kIlp], a[kIlp], b[kIlp], c[kIlp], gi[kIlp];
#pragma unroll
    for (int k = 0; k < kIlp; ++k) {
      uint32_t i = base + k * kBlock;
      if (i < n) {
        xv[k] = x[i]; yv[k] = y[i]; zv[k] = z[i];
        a[k] = px[i]; b[k] = py[i]; c[k] = pz[i]; gi[k] = g_inv[i];
      }
    }
#pragma unroll
    for (int k = 0; k < kIlp; ++k) {
      uint32_t i = base + k * kBlock;
      if (i < n) {
        push_one(xv[k], yv[k], zv[k], a[k], b[k], c[k], gi[k], dt);
        x[i] = xv[k]; y[i] = yv[k]; z[i] = zv[k];
      }
    }
  }
}

// ------------------------------------------------------------------ index + histogram
template <bool PUSH>
__global__ void __launch_bounds__(kBlock)
index_kernel(double* __restrict__ x, double* __restrict__ y, double* __restrict__ z,
             const double* __restrict__ px, const double* __restrict__ py,
             const double* __restrict__ pz, const double* __restrict__ g_inv,
             const double* __restrict__ dt_p, uint32_t* __restrict__ indx_in_cell,
             uint32_t* __restrict__ sum_in_cell, uint32_t n, GridGeom geom) {
  const GridVals g = load_geom(geom);
  const double dt = PUSH ? __ldg(dt_p) : 0.0;
  const uint32_t stride = gridDim.x * kBlock * kIlp;
  // uniform trip count per warp (warp_runs uses full-mask shuffles)
  const uint32_t n_round = ((n + 31u) / 32u) * 32u;
  for (uint32_t base = blockIdx.x * kBlock * kIlp + threadIdx.x;
       base - (threadIdx.x & 31) < n_round; base += stride) {
    double xv[kIlp], yv[kIlp], zv[kIlp];
    double a[kIlp], b[kIlp], c[kIlp], gi[kIlp];
#pragma unroll
    for (int k = 0; k < kIlp; ++k) {
      uint32_t i = base + k * kBlock;
      if (i < n) {
        xv[k] = x[i]; yv[k] = y[i]; zv[k] = z[i];
        if (PUSH) { a[k] = px[i]; b[k] = py[i]; c[k] = pz[i]; gi[k] = g_inv[i]; }
      }
    }
#pragma unroll
    for (int k = 0; k < kIlp; ++k) {
      uint32_t i = base + k * kBlock;
      bool valid = i < n;
      uint32_t cell = 0xffffffffu;
      if (valid) {
        if (PUSH) {
          push_one(xv[k], yv[k], zv[k], a[k], b[k], c[k], gi[k], dt);
          x[i] = xv[k]; y[i] = yv[k]; z[i] = zv[k];
        }
        cell = cell_index(xv[k], yv[k], zv[k], g);
        indx_in_cell[i] = cell;
      }
      // warp-uniform: the whole warp is either inside [0, n_round) or not
      if ((i - (threadIdx.x & 31)) < n_round) histogram_add(cell, valid, sum_in_cell);
    }
  }
}

// ------------------------------------------------------------------ scan
// cell_offset[0] = 0, cell_offset[i+1] = sum_{j<=i} sum_in_cell[j]; the same
// exclusive offsets are also written to `cursor` (scatter's running slots).
constexpr int kScanBlock = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanBlock * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* smem,
                                                         uint32_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = (lane < (int)(blockDim.x >> 5)) ? smem[lane] : 0u;
    uint32_t winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += t;
    }
    smem[32 + lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) smem[64] = winc;
  }
  __syncthreads();
  total = smem[64];
  uint32_t r = inc - v + smem[32 + warp];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kScanBlock)
scan_partials_kernel(const uint32_t* __restrict__ in, uint32_t n,
                     uint32_t* __restrict__ partials) {
  __shared__ uint32_t smem[65];
  uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (base + k < n) s += in[base + k];
  uint32_t total;
  block_exclusive_scan(s, smem, total);
  if (threadIdx.x == 0) partials[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanBlock)
scan_spine_kernel(uint32_t* __restrict__ partials, uint32_t nparts) {
  __shared__ uint32_t smem[65];
  uint32_t carry = 0;
  for (uint32_t b = 0; b < nparts; b += kScanBlock) {
    uint32_t i = b + threadIdx.x;
    uint32_t v = i < nparts ? partials[i] : 0u;
    uint32_t total;
    uint32_t ex = block_exclusive_scan(v, smem, total);
    if (i < nparts) partials[i] = ex + carry;
    carry += total;
  }
}

__global__ void __launch_bounds__(kScanBlock)
scan_final_kernel(const uint32_t* __restrict__ in, uint32_t n,
                  const uint32_t* __restrict__ partials,
                  uint32_t* __restrict__ cell_offset, uint32_t* __restrict__ cursor,
                  uint32_t* __restrict__ np_stay_out) {
  __shared__ uint32_t smem[65];
  uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0u;
    s += v[k];
  }
  uint32_t total;
  uint32_t ex = block_exclusive_scan(s, smem, total) + partials[blockIdx.x];
  if (blockIdx.x == 0 && threadIdx.x == 0) cell_offset[0] = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    uint32_t i = base + k;
    if (i < n) {
      if (cursor) cursor[i] = ex;
      ex += v[k];
      cell_offset[i + 1] = ex;
      // Np_stay = cell_offset[-2] = cell_offset[n-1] = particles in real cells
      if (i + 1 == n - 1 && np_stay_out) *np_stay_out = ex;
    }
  }
  if (n == 1 && blockIdx.x == 0 && threadIdx.x == 0 && np_stay_out) *np_stay_out = 0;
}

// ------------------------------------------------------------------ scatter
// slot = cursor[cell]++ claimed once per warp run; lanes of a run are in
// ascending storage order, runs of different warps land in arrival order and are
// put into the stable order by sort_fixup below.
__global__ void __launch_bounds__(kBlock)
sort_scatter_kernel(const uint32_t* __restrict__ indx_in_cell,
                    uint32_t* __restrict__ cursor, uint32_t* __restrict__ sort_indx,
                    uint32_t n) {
  const int lane = threadIdx.x & 31;
  const uint32_t n_round = ((n + 31u) / 32u) * 32u;
  const uint32_t stride = gridDim.x * kBlock * kIlp;
  // kIlp independent particles per thread: the slot-claiming atomics of all of them are
  // in flight before the first result is needed.  A warp owns kIlp CONSECUTIVE 32-particle
  // chunks and claims them in order, so a cell whose particles straddle chunks of the same
  // warp still receives them in ascending storage order (fewer segments for the fix-up
  // to repair: 0.279 -> 0.266 ms for scatter + fix-up; correctness never depends on it).
  constexpr uint32_t kStep = 32;
  const uint32_t first = (threadIdx.x >> 5) * 32 * kIlp + lane;
  for (uint32_t base = blockIdx.x * kBlock * kIlp + first; base - lane < n_round;
       base += stride) {
    uint32_t cell[kIlp], basev[kIlp];
    int head[kIlp], rank[kIlp];
#pragma unroll
    for (int k = 0; k < kIlp; ++k) {
      const uint32_t i = base + k * kStep;
      cell[k] = i < n ? indx_in_cell[i] : 0xffffffffu;
    }
#pragma unroll
    for (int k = 0; k < kIlp; ++k) {
      const uint32_t i = base + k * kStep;
      basev[k] = 0;
      head[k] = rank[k] = 0;
      if (i - lane < n_round) {            // warp-uniform
        int len;
        warp_runs(cell[k], i < n, lane, head[k], rank[k], len);
        if (i < n && rank[k] == 0) basev[k] = atomicAdd(&cursor[cell[k]], (uint32_t)len);
      }
    }
#pragma unroll
    for (int k = 0; k < kIlp; ++k) {
      const uint32_t i = base + k * kStep;
      if (i - lane < n_round) {
        const uint32_t b = __shfl_sync(0xffffffffu, basev[k], head[k]);
        if (i < n) sort_indx[b + rank[k]] = i;
      }
    }
  }
}

// ------------------------------------------------------------------ fix-up
// Every cell segment sort_indx[cell_offset[c] .. cell_offset[c+1]) already holds
// the right SET of storage indices; the stable counting sort of the reference
// (serial execution of particles_generic.cl:186-201) is that set in ascending
// order.  Segments are sorted in shared memory:
//   n <= kSmall      : per-thread insertion sort (input is a few ascending runs)
//   n <= kFixCap     : CTA-wide bitonic network (arbitrary n)
//   n >  kFixCap     : queued for sort_giant_kernel (single-CTA LSD radix sort)
#ifndef CHB_FIX_BLOCK
#define CHB_FIX_BLOCK 256
#endif
constexpr int kFixBlock = CHB_FIX_BLOCK;   // cells per CTA, one thread per cell
constexpr int kFixCap = 8192;    // staged entries (32 KiB)
constexpr int kSmall = 48;

// Shared staging buffer is addressed through SP(): one pad word every 16 entries, so
// that threads walking their own ~16-entry segments (stride ~16 words between lanes)
// hit different banks instead of the same one.
#define SP(i) ((i) + ((i) >> 4))
constexpr int kFixCapPadded = kFixCap + kFixCap / 16 + 1;

__device__ void bitonic_sort_smem(uint32_t* a, int base, int n) {
  int P = 1;
  while (P < n) P <<= 1;
  for (int k = 2; k <= P; k <<= 1) {
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
      int l = i ^ (k - 1);
      if (l > i && l < n) {
        uint32_t u = a[SP(base + i)], v = a[SP(base + l)];
        if (u > v) { a[SP(base + i)] = v; a[SP(base + l)] = u; }
      }
    }
    __syncthreads();
    for (int j = k >> 2; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        int l = i ^ j;
        if (l > i && l < n) {
          uint32_t u = a[SP(base + i)], v = a[SP(base + l)];
          if (u > v) { a[SP(base + i)] = v; a[SP(base + l)] = u; }
        }
      }
      __syncthreads();
    }
  }
}

// insertion sort of a[base .. base+n) addressed through SP() (shared staging)
__device__ __forceinline__ bool insertion_sort_sp(uint32_t* a, int base, int n) {
  bool changed = false;
  uint32_t prev = a[SP(base)];
  for (int i = 1; i < n; ++i) {
    uint32_t v = a[SP(base + i)];
    if (prev <= v) { prev = v; continue; }
    changed = true;
    int j = i - 1;
    while (j >= 0 && a[SP(base + j)] > v) { a[SP(base + j + 1)] = a[SP(base + j)]; --j; }
    a[SP(base + j + 1)] = v;
    // prev stays the maximum so far (a[base+i] after the shift)
  }
  return changed;
}

__device__ __forceinline__ bool insertion_sort(uint32_t* a, int n) {
  bool changed = false;
  for (int i = 1; i < n; ++i) {
    uint32_t v = a[i];
    int j = i - 1;
    if (a[j] <= v) continue;
    changed = true;
    while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; --j; }
    a[j + 1] = v;
  }
  return changed;
}

__global__ void __launch_bounds__(kFixBlock)
sort_fixup_kernel(const uint32_t* __restrict__ cell_offset, uint32_t nbins,
                  uint32_t* __restrict__ sort_indx, uint32_t* __restrict__ giant_count,
                  uint32_t* __restrict__ giant_list, uint32_t giant_cap) {
  __shared__ uint32_t stage[kFixCapPadded];
  __shared__ uint32_t big_list[kFixBlock];
  __shared__ uint32_t big_n;
  __shared__ int any_changed;

  for (uint32_t c0 = blockIdx.x * kFixBlock; c0 < nbins; c0 += gridDim.x * kFixBlock) {
    uint32_t c1 = min(c0 + kFixBlock, nbins);
    const uint32_t lo = cell_offset[c0], hi = cell_offset[c1];
    const uint32_t c = c0 + threadIdx.x;
    uint32_t s = 0, e = 0;
    if (c < c1) { s = cell_offset[c]; e = cell_offset[c + 1]; }
    const uint32_t nseg = e - s;
    if (threadIdx.x == 0) { big_n = 0; any_changed = 0; }
    __syncthreads();
    if (hi == lo) continue;

    if (hi - lo <= (uint32_t)kFixCap) {
      // stage the whole range (coalesced), sort segments in shared memory
      for (uint32_t i = threadIdx.x; i < hi - lo; i += kFixBlock) stage[SP(i)] = sort_indx[lo + i];
      __syncthreads();
      if (nseg > 1) {
        if (nseg <= (uint32_t)kSmall) {
          if (insertion_sort_sp(stage, (int)(s - lo), (int)nseg)) any_changed = 1;
        } else {
          big_list[atomicAdd(&big_n, 1u)] = threadIdx.x;
        }
      }
      __syncthreads();
      const uint32_t nb = big_n;
      for (uint32_t b = 0; b < nb; ++b) {
        uint32_t t = big_list[b];
        uint32_t bs = cell_offset[c0 + t], be = cell_offset[c0 + t + 1];
        bitonic_sort_smem(stage, (int)(bs - lo), (int)(be - bs));
      }
      if (nb) any_changed = 1;
      __syncthreads();
      if (any_changed)
        for (uint32_t i = threadIdx.x; i < hi - lo; i += kFixBlock) sort_indx[lo + i] = stage[SP(i)];
      __syncthreads();
    } else {
      // crowded range: small segments in place in global memory (each thread
      // touches only its own segment), larger ones staged one at a time
      if (nseg > 1) {
        if (nseg <= (uint32_t)kSmall) {
          insertion_sort(sort_indx + s, (int)nseg);
        } else {
          big_list[atomicAdd(&big_n, 1u)] = threadIdx.x;
        }
      }
      __syncthreads();
      const uint32_t nb = big_n;
      for (uint32_t b = 0; b < nb; ++b) {
        uint32_t t = big_list[b];
        uint32_t bs = cell_offset[c0 + t], be = cell_offset[c0 + t + 1];
        uint32_t bn = be - bs;
        if (bn <= (uint32_t)kFixCap) {
          for (uint32_t i = threadIdx.x; i < bn; i += kFixBlock) stage[SP(i)] = sort_indx[bs + i];
          __syncthreads();
          bitonic_sort_smem(stage, 0, (int)bn);
          for (uint32_t i = threadIdx.x; i < bn; i += kFixBlock) sort_indx[bs + i] = stage[SP(i)];
          __syncthreads();
        } else if (threadIdx.x == 0) {
          uint32_t k = atomicAdd(giant_count, 1u);
          if (k < giant_cap) giant_list[k] = c0 + t;
        }
      }
      __syncthreads();
    }
  }
}

// Single-CTA stable LSD radix sort (8-bit digits) of one giant segment, keys =
// storage indices < 2^key_bits.  Ping-pongs between the segment and `tmp`.
constexpr int kRadixBlock = 1024;

__global__ void __launch_bounds__(kRadixBlock)
sort_giant_kernel(const uint32_t* __restrict__ cell_offset,
                  const uint32_t* __restrict__ giant_count,
                  const uint32_t* __restrict__ giant_list, uint32_t giant_cap,
                  uint32_t* __restrict__ sort_indx, uint32_t* __restrict__ tmp,
                  int key_bits) {
  __shared__ uint32_t digit_base[256];
  __shared__ uint16_t warp_cnt[32][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t ng = min(*giant_count, giant_cap);
  for (uint32_t gi = blockIdx.x; gi < ng; gi += gridDim.x) {
    const uint32_t cell = giant_list[gi];
    const uint32_t s = cell_offset[cell], n = cell_offset[cell + 1] - s;
    uint32_t* src = sort_indx + s;
    uint32_t* dst = tmp + s;
    int npass = (key_bits + 7) / 8;
    if (npass & 1) ++npass;  // even number of passes: result ends in sort_indx
    for (int pass = 0; pass < npass; ++pass) {
      const int shift = pass * 8;
      // digit histogram
      if (threadIdx.x < 256) digit_base[threadIdx.x] = 0;
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < n; i += kRadixBlock)
        atomicAdd(&digit_base[(shift < 32 ? (src[i] >> shift) : 0u) & 255u], 1u);
      __syncthreads();
      // exclusive scan over the 256 digits (warp 0, 8 per lane)
      if (warp == 0) {
        uint32_t v[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[k] = digit_base[lane * 8 + k]; sum += v[k]; }
        uint32_t inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
          if (lane >= d) inc += t;
        }
        uint32_t ex = inc - sum;
#pragma unroll
        for (int k = 0; k < 8; ++k) { digit_base[lane * 8 + k] = ex; ex += v[k]; }
      }
      __syncthreads();
      // stable scatter, tile by tile in storage order
      for (uint32_t t0 = 0; t0 < n; t0 += kRadixBlock) {
        for (int k = threadIdx.x; k < 32 * 256; k += kRadixBlock) (&warp_cnt[0][0])[k] = 0;
        __syncthreads();
        uint32_t i = t0 + threadIdx.x;
        bool valid = i < n;
        uint32_t key = valid ? src[i] : 0u;
        uint32_t d = valid ? ((shift < 32 ? (key >> shift) : 0u) & 255u) : 256u + lane;
        unsigned m = __match_any_sync(0xffffffffu, d);
        int rank = __popc(m & ((1u << lane) - 1u));
        if (valid && rank == 0) warp_cnt[warp][d] = (uint16_t)__popc(m);
        __syncthreads();
        uint32_t below = 0;
        if (valid && rank == 0)
          for (int w = 0; w < warp; ++w) below += warp_cnt[w][d];
        below = __shfl_sync(0xffffffffu, below, __ffs(m) - 1);
        if (valid) dst[digit_base[d] + below + rank] = key;
        __syncthreads();
        // advance the digit bases by this tile's totals
        if (threadIdx.x < 256) {
          uint32_t tot = 0;
          for (int w = 0; w < 32; ++w) tot += warp_cnt[w][threadIdx.x];
          digit_base[threadIdx.x] += tot;
        }
        __syncthreads();
      }
      uint32_t* sw = src; src = dst; dst = sw;
    }
  }
}

// ------------------------------------------------------------------ align
struct AlignArgs {
  const double* src[CHB_MAX_ATTRS];
  double* dst[CHB_MAX_ATTRS];
  int nattr;
};

__global__ void __launch_bounds__(kBlock)
align_kernel(AlignArgs a, const uint32_t* __restrict__ sort_indx, uint32_t n) {
  const uint32_t stride = gridDim.x * kBlock;
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
    uint32_t s = sort_indx[i];
#pragma unroll
    for (int k = 0; k < CHB_MAX_ATTRS; ++k)
      if (k < a.nattr) a.dst[k][i] = __ldg(a.src[k] + s);
  }
}

__global__ void iota_kernel(uint32_t* out, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = i;
}

}  // namespace chb

using namespace chb;

extern "C" {

int chb_push_xyz(double* x, double* y, double* z, const double* px, const double* py,
                 const double* pz, const double* g_inv, const double* dt_dev,
                 uint32_t np, void* stream) {
  if (np == 0) return CHB_OK;
  int grid = stream_grid(np, kBlock * kIlp, 8);
  push_xyz_kernel<<<grid, kBlock, 0, (cudaStream_t)stream>>>(x, y, z, px, py, pz, g_inv, dt_dev, np);
  CHB_RETURN_LAST_ERROR();
}

int chb_index_and_sum(const double* x, const double* y, const double* z,
                      uint32_t* indx_in_cell, uint32_t* sum_in_cell, uint32_t np,
                      uint32_t Nx, uint32_t Nr, const double* xmin, const double* dx_inv,
                      const double* rmin, const double* dr_inv, void* stream) {
  if (np == 0) return CHB_OK;
  GridGeom g{xmin, dx_inv, rmin, dr_inv, Nx, Nr};
  int grid = stream_grid(np, kBlock * kIlp, 8);
  index_kernel<false><<<grid, kBlock, 0, (cudaStream_t)stream>>>(
      (double*)x, (double*)y, (double*)z, nullptr, nullptr, nullptr, nullptr, nullptr,
      indx_in_cell, sum_in_cell, np, g);
  CHB_RETURN_LAST_ERROR();
}

int chb_push_index(double* x, double* y, double* z, const double* px, const double* py,
                   const double* pz, const double* g_inv, const double* dt_dev,
                   uint32_t* indx_in_cell, uint32_t* sum_in_cell, uint32_t np,
                   uint32_t Nx, uint32_t Nr, const double* xmin, const double* dx_inv,
                   const double* rmin, const double* dr_inv, void* stream) {
  if (np == 0) return CHB_OK;
  GridGeom g{xmin, dx_inv, rmin, dr_inv, Nx, Nr};
  int grid = stream_grid(np, kBlock * kIlp, 8);
  index_kernel<true><<<grid, kBlock, 0, (cudaStream_t)stream>>>(
      x, y, z, px, py, pz, g_inv, dt_dev, indx_in_cell, sum_in_cell, np, g);
  CHB_RETURN_LAST_ERROR();
}

size_t chb_cell_offsets_workspace_bytes(uint32_t nbins) {
  size_t nparts = (nbins + kScanTile - 1) / kScanTile;
  return (nparts + 1) * sizeof(uint32_t);
}

int chb_cell_offsets(const uint32_t* sum_in_cell, uint32_t nbins, uint32_t* cell_offset,
                     uint32_t* cursor, uint32_t* np_stay_dev, void* workspace,
                     size_t workspace_bytes, void* stream) {
  if (nbins == 0) return CHB_ERR_ARG;
  if (workspace_bytes < chb_cell_offsets_workspace_bytes(nbins)) return CHB_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t nparts = (nbins + kScanTile - 1) / kScanTile;
  uint32_t* partials = (uint32_t*)workspace;
  scan_partials_kernel<<<nparts, kScanBlock, 0, st>>>(sum_in_cell, nbins, partials);
  scan_spine_kernel<<<1, kScanBlock, 0, st>>>(partials, nparts);
  scan_final_kernel<<<nparts, kScanBlock, 0, st>>>(sum_in_cell, nbins, partials, cell_offset,
                                                  cursor, np_stay_dev);
  CHB_RETURN_LAST_ERROR();
}

size_t chb_sort_workspace_bytes(uint32_t np, uint32_t nbins) {
  // cursor[nbins] + giant_count[1] + giant_list[CHB_GIANT_CAP] + tmp[np]
  return ((size_t)nbins + 1 + CHB_GIANT_CAP + np) * sizeof(uint32_t) + 256;
}

int chb_sort_scatter_stable(const uint32_t* indx_in_cell, const uint32_t* cell_offset,
                            uint32_t* cursor, uint32_t* sort_indx, uint32_t np,
                            uint32_t nbins, void* workspace, size_t workspace_bytes,
                            void* stream) {
  if (np == 0) return CHB_OK;
  size_t need = ((size_t)1 + CHB_GIANT_CAP + np) * sizeof(uint32_t);
  if (workspace_bytes < need) return CHB_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* giant_count = (uint32_t*)workspace;
  uint32_t* giant_list = giant_count + 1;
  uint32_t* tmp = giant_list + CHB_GIANT_CAP;
  cudaError_t e = cudaMemsetAsync(giant_count, 0, sizeof(uint32_t), st);
  if (e != cudaSuccess) return (int)e;
  sort_scatter_kernel<<<stream_grid(np, kBlock * kIlp, 8), kBlock, 0, st>>>(indx_in_cell, cursor,
                                                                      sort_indx, np);
  int fgrid = stream_grid(nbins, kFixBlock, 8);
  sort_fixup_kernel<<<fgrid, kFixBlock, 0, st>>>(cell_offset, nbins, sort_indx, giant_count,
                                                 giant_list, CHB_GIANT_CAP);
  int key_bits = 1;
  while (key_bits < 32 && (1ull << key_bits) < (uint64_t)np) ++key_bits;
  sort_giant_kernel<<<32, kRadixBlock, 0, st>>>(cell_offset, giant_count, giant_list,
                                                CHB_GIANT_CAP, sort_indx, tmp, key_bits);
  CHB_RETURN_LAST_ERROR();
}

int chb_align(const double* const* src, double* const* dst, int nattr,
              const uint32_t* sort_indx, uint32_t np_stay, uint32_t* sort_indx_out,
              void* stream) {
  if (nattr < 0 || nattr > CHB_MAX_ATTRS) return CHB_ERR_ARG;
  if (np_stay == 0) return CHB_OK;
  AlignArgs a;
  a.nattr = nattr;
  for (int k = 0; k < CHB_MAX_ATTRS; ++k) {
    a.src[k] = k < nattr ? src[k] : nullptr;
    a.dst[k] = k < nattr ? dst[k] : nullptr;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int grid = stream_grid(np_stay, kBlock, 16);
  align_kernel<<<grid, kBlock, 0, st>>>(a, sort_indx, np_stay);
  if (sort_indx_out) iota_kernel<<<grid, kBlock, 0, st>>>(sort_indx_out, np_stay);
  CHB_RETURN_LAST_ERROR();
}

}  // extern "C"
