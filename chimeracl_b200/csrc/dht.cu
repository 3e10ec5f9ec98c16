// chimera-b200 discrete Hankel transform: dense FP64 contraction on the tensor pipe.
//
// Replaces (behaviour, not code) the reference's Reikna MatrixMul calls `_ddot` /
// `_cdot` (methods/transformer_methods_cl.py:458-480; on CPU devices literally
// np.dot, :474-480), used by the Fourier-Bessel transforms (:296,:320,:355,:379)
// and the spectral grad/rot operators (:111,:125,:221,:228,:248,:255).
//
//   C[M x N] (op)= alpha * A[M x K] . B[K x N]      A real, B real or complex
//
// Because A is real, a complex right-hand side is the same real GEMM on the
// (re, im)-interleaved view with 2N columns: one kernel serves both.  FP64 does
// not exist on tcgen05, so the tensor path for this contraction is the FP64
// DMMA (`mma.sync.m8n8k4.f64`); operands are staged in padded shared memory
// (conflict-free fragment reads), next k-slab prefetched into registers while the
// current one is multiplied.  Optional fused epilogue: complex alpha and
// accumulate-into-C, which folds the reference's zpaxz/append_c2c passes
// (kernels/generic.cl:18-45) into the contraction.
#include <cstdlib>
#include "common.cuh"
#include "dht_epilogue.cuh"
#include "../../include/chimera_b200.h"

namespace chb {

constexpr int BM = 128, BN = 64, BK = 16;
constexpr int kGemmThreads = 256;       // 8 warps: 4 (M) x 2 (N), warp tile 32 x 32
constexpr int LDA_S = BK + 4;           // 20: (g*20 + t) mod 16 distinct over a half warp
constexpr int LDB_S = BN + 4;           // 68: (t*68 + g) mod 16 distinct over a half warp

__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

struct GemmArgs {
  const double* __restrict__ A;
  const double* Bv[CHB_MAX_FIELDS];   // batched right-hand sides (blockIdx.z)
  double* Cv[CHB_MAX_FIELDS];
  uint32_t lda, ldb, ldc;   // in doubles
  uint32_t M, N, K;         // N in doubles (2*Nx for complex data)
  uint32_t ka;              // K rounded up to even: A columns the 16-byte loads may touch
  double alpha_re, alpha_im;
  int complex_pairs;        // columns are (re, im) pairs -> complex alpha allowed
  int accumulate;           // C += ... instead of C = ...
  // optional second output of the same product: C2 (op)= alpha2 * A.B
  double* __restrict__ C2;
  uint32_t ldc2;
  double alpha2_re, alpha2_im;
  int accumulate2;
  // > 0 (complex data, wide kernel only): B is the x-spectrum of a REAL field, i.e.
  // B(:, Nx-k) = conj(B(:, k)); only the columns k <= Nx/2 are contracted (N is set
  // accordingly) and every result is also written, conjugated, to column Nx-k
  uint32_t mirror_nx;
};

constexpr int kSlabDoubles = BM * LDA_S + BK * LDB_S;
constexpr int kGemmSmem = 2 * kSlabDoubles * (int)sizeof(double);   // double buffered

// VEC: A rows and B rows are 16-byte aligned with even leading dimensions (the DHT
// matrices are stored with a padded leading dimension, zeros in the pad), so the
// global loads are 128-bit.
template <bool VEC>
__global__ void __launch_bounds__(kGemmThreads, 2)
dht_gemm_kernel(GemmArgs p) {
  extern __shared__ double gemm_smem[];
  const double* __restrict__ Bg = p.Bv[blockIdx.z];
  double* __restrict__ Cg = p.Cv[blockIdx.z];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
  const uint32_t m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  // global -> register staging assignment
  const int a_row = tid >> 1, a_k = (tid & 1) * 8;   // 8 consecutive k of one row
  const int b_row = tid >> 4, b_col = (tid & 15) * 4; // 4 consecutive columns of one k
  double ra[8], rb[4];

  auto load_slab = [&](uint32_t k0) {
    const uint32_t gr = m0 + a_row;
    const uint32_t gk = k0 + b_row;
    if (VEC) {
      // pairs (gk, gk+1): the pad column of A (index >= K, < lda) holds zeros
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const uint32_t ka = k0 + a_k + i;
        double2 v = make_double2(0.0, 0.0);
        if (gr < p.M && ka < p.K) v = __ldg(reinterpret_cast<const double2*>(p.A + (size_t)gr * p.lda + ka));
        ra[i] = v.x; ra[i + 1] = v.y;
      }
#pragma unroll
      for (int i = 0; i < 4; i += 2) {
        const uint32_t gc = n0 + b_col + i;
        double2 v = make_double2(0.0, 0.0);
        if (gk < p.K && gc < p.N) v = __ldg(reinterpret_cast<const double2*>(Bg + (size_t)gk * p.ldb + gc));
        rb[i] = v.x; rb[i + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t ka = k0 + a_k + i;
        ra[i] = (gr < p.M && ka < p.K) ? __ldg(p.A + (size_t)gr * p.lda + ka) : 0.0;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t gc = n0 + b_col + i;
        rb[i] = (gk < p.K && gc < p.N) ? __ldg(Bg + (size_t)gk * p.ldb + gc) : 0.0;
      }
    }
  };
  auto store_slab = [&](int buf) {
    double* As = gemm_smem + buf * kSlabDoubles;
    double* Bs = As + BM * LDA_S;
#pragma unroll
    for (int i = 0; i < 8; ++i) As[a_row * LDA_S + a_k + i] = ra[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) Bs[b_row * LDB_S + b_col + i] = rb[i];
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  load_slab(0);
  store_slab(0);
  __syncthreads();
  int cur = 0;
  for (uint32_t k0 = 0; k0 < p.K; k0 += BK) {
    const bool more = k0 + BK < p.K;
    if (more) load_slab(k0 + BK);   // global loads in flight while multiplying
    const double* As = gemm_smem + cur * kSlabDoubles;
    const double* Bs = As + BM * LDA_S;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[(wm * 32 + i * 8 + g) * LDA_S + kk + t];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[(kk + t) * LDB_S + wn * 32 + j * 8 + g];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    if (more) store_slab(cur ^ 1);  // the other buffer was consumed one iteration ago
    __syncthreads();
    cur ^= 1;
  }

  // epilogue: thread owns C(row g, cols 2t, 2t+1) of every 8x8 block
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t row = m0 + wm * 32 + i * 8 + g;
    if (row >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t col = n0 + wn * 32 + j * 8 + 2 * t;
      if (col >= p.N) continue;
      const bool pair = (col + 1 < p.N);
      gemm_store(Cg + (size_t)row * p.ldc + col, acc[i][j][0], acc[i][j][1], p.alpha_re,
                 p.alpha_im, p.complex_pairs, p.accumulate, pair);
      if (p.C2)
        gemm_store(p.C2 + (size_t)row * p.ldc2 + col, acc[i][j][0], acc[i][j][1], p.alpha2_re,
                   p.alpha2_im, p.complex_pairs, p.accumulate2, pair);
    }
  }
}


// ---------------------------------------------------------------------------------------
// Wide-tile variant: one 16-warp CTA per SM, 128 x (16*NT) tile, warp tile 16 x (8*NT)
// (a single warp cannot issue DMMAs back to back at the pipe's rate -- measured with
// tools/exp/fp64_pipes.cu: one warp per scheduler reaches half of it -- so four warps
// per scheduler keep the pipe fed while others load fragments or wait), operands
// brought in by a 3-stage 16-byte cp.async pipeline of 32-deep k-slabs (zero-filled at the
// K / N edges), one barrier per slab.  The tile width is chosen per launch so that the number of tiles is
// (just under) a multiple of the 148 SMs: for Nx = 4096 a 112-wide tile gives exactly 148
// (real) / 296 (complex) tiles per contraction, i.e. whole waves, where the 128 x 64 tiles
// above ran 1.73 waves.  Needs 16-byte aligned operands (see dht_launch).
#ifndef CHB_DHT_STAGES
#define CHB_DHT_STAGES 3
#endif
#ifndef CHB_DHT_WBK
#define CHB_DHT_WBK 32
#endif
constexpr int kWideStages = CHB_DHT_STAGES;
constexpr int kWideThreads = 512;      // 16 warps: 8 (M) x 2 (N)
constexpr int WBK = CHB_DHT_WBK;       // k-slab of the wide kernel
constexpr int WLDA = WBK + 4;          // 36: (g*36 + t) mod 16 distinct over a half warp

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, bool valid) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem_src), "r"(n)
               : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// WMW = warps along M: 8 -> 128-row tiles, 2 warps (16*NT columns) across; 4 -> 64-row
// tiles, 4 warps (32*NT columns) across, for results of <= 64 rows (the owned kr rows of
// a sharded forward / grad / rot contraction), which would fill half of a 128-row tile.
// The warp tile (16 x 8*NT), and with it the fragment-load : DMMA ratio, is the same.
template <int NT, int WMW = 8>
struct WideShape {
  static constexpr int kBM = 16 * WMW;
  static constexpr int kWN = 16 / WMW;
  static constexpr int kBN = kWN * 8 * NT;
  static constexpr int kLdb = kBN + 4;                 // (t*kLdb + g) mod 16 distinct (kBN % 16 == 0)
  static constexpr int kStageDoubles = kBM * WLDA + WBK * kLdb;
  static constexpr int kSmem = kWideStages * kStageDoubles * (int)sizeof(double);
};

template <int NT, int WMW = 8>
__global__ void __launch_bounds__(kWideThreads, 1)
dht_gemm_wide_kernel(const __grid_constant__ GemmArgs p) {
  using S = WideShape<NT, WMW>;
  constexpr int BNW = S::kBN;
  constexpr int BMW = S::kBM;
  extern __shared__ __align__(16) double gemm_smem[];
  const double* __restrict__ Bg = p.Bv[blockIdx.z];
  double* __restrict__ Cg = p.Cv[blockIdx.z];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp / S::kWN, wn = warp % S::kWN;  // wm: 0..WMW-1
  const uint32_t m0 = blockIdx.y * BMW, n0 = blockIdx.x * BNW;

  // cp.async assignment, fixed per thread for the whole k loop: the A slab is BMW rows x
  // 16 chunks of 16 bytes, the B slab 32 rows x BNW/2 chunks; only the k offset moves
  constexpr int kAPer = BMW * (WBK / 2) / kWideThreads;                    // 4 (64 rows: 2)
  constexpr int kBChunks = WBK * (BNW / 2);
  constexpr int kBPer = (kBChunks + kWideThreads - 1) / kWideThreads;      // 4 (NT = 7: 3.5)
  const double* a_src[kAPer];
  uint32_t a_dst[kAPer], a_k[kAPer];
  bool a_ok[kAPer];
#pragma unroll
  for (int i = 0; i < kAPer; ++i) {
    const int c = tid + i * kWideThreads;
    const int row = c / (WBK / 2), kc = (c % (WBK / 2)) * 2;
    a_ok[i] = m0 + row < p.M;
    a_src[i] = p.A + (size_t)(a_ok[i] ? m0 + row : 0) * p.lda + kc;
    a_dst[i] = row * WLDA + kc;
    a_k[i] = kc;
  }
  const double* b_src[kBPer];
  uint32_t b_dst[kBPer], b_k[kBPer];
  bool b_ok[kBPer];
#pragma unroll
  for (int i = 0; i < kBPer; ++i) {
    const int c = tid + i * kWideThreads;
    const int row = c / (BNW / 2), cc = (c - row * (BNW / 2)) * 2;
    b_ok[i] = c < kBChunks && n0 + cc < p.N;
    b_src[i] = Bg + (size_t)row * p.ldb + (b_ok[i] ? n0 + cc : 0);
    b_dst[i] = BMW * WLDA + row * S::kLdb + cc;
    b_k[i] = row;
  }
  const size_t b_step = (size_t)WBK * p.ldb;

  // slabs are issued strictly in order, so the sources just advance
  auto issue = [&](uint32_t k0, int stage) {
    double* st = gemm_smem + stage * S::kStageDoubles;
#pragma unroll
    for (int i = 0; i < kAPer; ++i) {
      // column K of an odd-K operand is a (finite) pad or neighbour element of the same row
      // and meets a zero-filled B row; nothing beyond it is touched, so A may be a column
      // block of a larger matrix (the kr-sharded backward transform)
      const bool ok = a_ok[i] && k0 + a_k[i] < p.ka;
      cp_async16_zfill(st + a_dst[i], ok ? a_src[i] : p.A, ok);
      a_src[i] += WBK;
    }
#pragma unroll
    for (int i = 0; i < kBPer; ++i) {
      if (kBChunks % kWideThreads != 0 && i == kBPer - 1 && tid + i * kWideThreads >= kBChunks)
        break;
      const bool ok = b_ok[i] && k0 + b_k[i] < p.K;
      cp_async16_zfill(st + b_dst[i], ok ? b_src[i] : Bg, ok);
      b_src[i] += b_step;
    }
  };

  double acc[2][NT][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const uint32_t nslab = (p.K + WBK - 1) / WBK;
#pragma unroll
  for (int s = 0; s < kWideStages - 1; ++s) {
    if ((uint32_t)s < nslab) issue(s * WBK, s);
    cp_async_commit();
  }
  int stage = 0;
  for (uint32_t it = 0; it < nslab; ++it) {
    cp_async_wait_group<kWideStages - 2>();   // slab `it` has landed (this thread's part)
    __syncthreads();                          // ... everybody's; slab it-1 fully consumed
    const double* As = gemm_smem + stage * S::kStageDoubles;
    const double* Bs = As + BMW * WLDA;
#pragma unroll
    for (int kk = 0; kk < WBK; kk += 4) {
      double a[2], b[NT];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = As[(wm * 16 + i * 8 + g) * WLDA + kk + t];
#pragma unroll
      for (int j = 0; j < NT; ++j) b[j] = Bs[(kk + t) * S::kLdb + wn * (8 * NT) + j * 8 + g];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma_884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      if (kk == 0) {
        // refill the buffer consumed in the previous iteration while this warp's first
        // DMMAs are queued in the pipe
        const uint32_t nxt = it + kWideStages - 1;
        int nstage = stage + kWideStages - 1;
        if (nstage >= kWideStages) nstage -= kWideStages;
        if (nxt < nslab) issue(nxt * WBK, nstage);
        cp_async_commit();
      }
    }
    if (++stage == kWideStages) stage = 0;
  }
  cp_async_wait_group<0>();

#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const uint32_t row = m0 + wm * 16 + i * 8 + g;
    if (row >= p.M) continue;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const uint32_t col = n0 + wn * (8 * NT) + j * 8 + 2 * t;
      if (col >= p.N) continue;
      const bool pair = (col + 1 < p.N);
      gemm_store(Cg + (size_t)row * p.ldc + col, acc[i][j][0], acc[i][j][1], p.alpha_re,
                 p.alpha_im, p.complex_pairs, p.accumulate, pair);
      if (p.C2)
        gemm_store(p.C2 + (size_t)row * p.ldc2 + col, acc[i][j][0], acc[i][j][1], p.alpha2_re,
                   p.alpha2_im, p.complex_pairs, p.accumulate2, pair);
      if (p.mirror_nx) {
        const uint32_t k = col >> 1;                     // complex column
        if (k > 0 && 2 * k < p.mirror_nx) {
          const uint32_t mcol = 2 * (p.mirror_nx - k);
          gemm_store(Cg + (size_t)row * p.ldc + mcol, acc[i][j][0], -acc[i][j][1], p.alpha_re,
                     p.alpha_im, 1, p.accumulate, true);
          if (p.C2)
            gemm_store(p.C2 + (size_t)row * p.ldc2 + mcol, acc[i][j][0], -acc[i][j][1],
                       p.alpha2_re, p.alpha2_im, 1, p.accumulate2, true);
        }
      }
    }
  }
}

template <int NT, int WMW = 8>
static cudaError_t launch_wide(const GemmArgs& p, int nbatch, cudaStream_t st) {
  using S = WideShape<NT, WMW>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(dht_gemm_wide_kernel<NT, WMW>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, S::kSmem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((p.N + S::kBN - 1) / S::kBN, (p.M + S::kBM - 1) / S::kBM, nbatch);
  dht_gemm_wide_kernel<NT, WMW><<<grid, kWideThreads, S::kSmem, st>>>(p);
  return cudaGetLastError();
}

// 64-row tiles: width 32*nt columns, nt = 2..7 (nt = 8 would not fit the shared memory).
// Narrow tiles (nt = 2, 3) exist for the single right-hand sides of an 8-way sharded solve:
// a 64 x 8192 result is 37-64 tiles with nt >= 4, i.e. a quarter to a half of the 148 SMs;
// their fragment-load : DMMA ratio is worse ((2 + nt) / 2nt: 1.0, 0.83 against 0.64 at
// nt = 7), which the cost accounts for.
static int pick_wide64_nt(uint32_t N, int nbatch) {
  int best = 0;
  double best_cost = 0;
  for (int nt = 2; nt <= 7; ++nt) {
    const uint64_t tiles = (uint64_t)((N + 32 * nt - 1) / (32 * nt)) * nbatch;
    const uint64_t waves = (tiles + kSMs - 1) / kSMs;
    const double eff = nt >= 4 ? 1.0 : (nt == 3 ? 0.9 : 0.8);
    const double cost = (double)waves * nt / eff;
    if (best == 0 || cost <= best_cost) { best = nt; best_cost = cost; }
  }
  return best;
}

// tile width (in units of 16 columns) minimising waves x width on 148 SMs
static int pick_wide_nt(uint32_t M, uint32_t N, int nbatch) {
  int best = 0;
  double best_cost = 0;
  for (int nt = 4; nt <= 8; ++nt) {
    const uint64_t tiles = (uint64_t)((N + 16 * nt - 1) / (16 * nt)) * ((M + BM - 1) / BM) * nbatch;
    const uint64_t waves = (tiles + kSMs - 1) / kSMs;
    const double cost = (double)waves * nt;
    if (best == 0 || cost <= best_cost) { best = nt; best_cost = cost; }
  }
  return best;
}

// Register-only DMMA loop: the FP64 tensor pipe's issue ceiling on this device (the
// roofline denominator bench.py reports for the contraction; MEASURED_PEAKS.json has no
// FP64 entry).  16 warps per SM, 8 independent accumulators per warp.
__global__ void __launch_bounds__(512) dmma_peak_kernel(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) dmma_884(c[i], c[i + 1], a, b);
  }
  double sum = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) sum += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

}  // namespace chb

using namespace chb;

extern "C" int chb_dmma_peak(double* scratch, size_t scratch_doubles, int iters,
                             double* flops_out, void* stream) {
  if (scratch_doubles < (size_t)kSMs * 512 || iters <= 0 || !flops_out) return CHB_ERR_ARG;
  dmma_peak_kernel<<<kSMs, 512, 0, (cudaStream_t)stream>>>(scratch, iters);
  // 16 warps x 8 DMMA.8x8x4 (512 flop each) per iteration and SM
  *flops_out = (double)kSMs * 16.0 * 8.0 * 512.0 * (double)iters;
  CHB_RETURN_LAST_ERROR();
}

static int dht_launch(const double* A, uint32_t lda, const double* const* Bv, int nbatch,
                      uint32_t ldb, double* const* Cv, uint32_t ldc, uint32_t M, uint32_t K,
                      uint32_t N,
                      int is_complex, double alpha_re, double alpha_im, int accumulate,
                      double* C2, uint32_t ldc2, double alpha2_re, double alpha2_im,
                      int accumulate2, void* stream, int hermitian = 0) {
  if (M == 0 || N == 0 || nbatch == 0) return CHB_OK;
  if (nbatch < 0 || nbatch > CHB_MAX_FIELDS || (C2 && nbatch != 1)) return CHB_ERR_ARG;
  if (!is_complex && (alpha_im != 0.0 || alpha2_im != 0.0)) return CHB_ERR_ARG;
  GemmArgs p;
  for (int k = 0; k < CHB_MAX_FIELDS; ++k) {
    p.Bv[k] = k < nbatch ? Bv[k] : nullptr;
    p.Cv[k] = k < nbatch ? Cv[k] : nullptr;
  }
  p.C2 = C2;
  p.ldc2 = is_complex ? 2 * ldc2 : ldc2;
  p.alpha2_re = alpha2_re; p.alpha2_im = alpha2_im;
  p.accumulate2 = accumulate2;
  p.A = A;
  p.lda = lda;
  p.ldb = is_complex ? 2 * ldb : ldb;
  p.ldc = is_complex ? 2 * ldc : ldc;
  p.M = M; p.K = K;
  p.ka = (K + 1) & ~1u;
  p.N = is_complex ? 2 * N : N;
  p.alpha_re = alpha_re; p.alpha_im = alpha_im;
  p.complex_pairs = is_complex;
  p.accumulate = accumulate;
  p.mirror_nx = 0;
  dim3 grid((p.N + BN - 1) / BN, (M + BM - 1) / BM, nbatch);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(dht_gemm_kernel<false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(dht_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               kGemmSmem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  // 128-bit loads need 16-byte aligned rows: even leading dimensions and aligned bases;
  // with the pairs (k, k+1) the A rows must have a (zero) pad element when K is odd
  bool vec = (p.lda % 2 == 0) && (p.ldb % 2 == 0) && (p.N % 2 == 0) &&
             (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (p.lda > p.K || p.K % 2 == 0);
  for (int k = 0; k < nbatch && vec; ++k) vec = reinterpret_cast<uintptr_t>(Bv[k]) % 16 == 0;
  const bool wide = vec && (p.ldc % 2 == 0) && (!C2 || p.ldc2 % 2 == 0) && !getenv("CHB_DHT_NARROW");
  if (wide && hermitian && is_complex && N >= 4 && N % 2 == 0) {
    p.mirror_nx = N;                 // complex columns of the full result
    p.N = 2 * (N / 2 + 1);           // contracted: k = 0 .. Nx/2
  }
  // results of <= 64 rows (kr-sharded forward / grad / rot contractions at 8 ranks): 64-row
  // tiles.  Measured with the 8 shards of cfg3 on one B200 (profiles/r1h_tile64_timing.txt):
  // batched launches 19 % faster; single right-hand sides were 5-25 % slower with the
  // 128-256 column tiles of round 1 (too few tiles) -- with the 64 / 96 column tiles they
  // fill the SMs as well, so 64-row tiles now serve every result of <= 64 rows.
  // CHB_DHT_TILE64=0: never; =2: the round-1 rule (>= 3 right-hand sides, nt >= 4).
  static const int tile64 = [] { const char* e = getenv("CHB_DHT_TILE64");
                                 return e ? atoi(e) : 1; }();
  if (wide && M <= 64 && (tile64 == 1 || (tile64 == 2 && nbatch >= 3))) {
    cudaError_t e;
    int nt64 = pick_wide64_nt(p.N, nbatch);
    if (tile64 == 2 && nt64 < 4) nt64 = 4;
    switch (nt64) {
      case 2: e = launch_wide<2, 4>(p, nbatch, (cudaStream_t)stream); break;
      case 3: e = launch_wide<3, 4>(p, nbatch, (cudaStream_t)stream); break;
      case 4: e = launch_wide<4, 4>(p, nbatch, (cudaStream_t)stream); break;
      case 5: e = launch_wide<5, 4>(p, nbatch, (cudaStream_t)stream); break;
      case 6: e = launch_wide<6, 4>(p, nbatch, (cudaStream_t)stream); break;
      default: e = launch_wide<7, 4>(p, nbatch, (cudaStream_t)stream); break;
    }
    return e == cudaSuccess ? CHB_OK : (int)e;
  }
  if (wide) {
    cudaError_t e;
    switch (pick_wide_nt(M, p.N, nbatch)) {
      case 4: e = launch_wide<4>(p, nbatch, (cudaStream_t)stream); break;
      case 5: e = launch_wide<5>(p, nbatch, (cudaStream_t)stream); break;
      case 6: e = launch_wide<6>(p, nbatch, (cudaStream_t)stream); break;
      case 7: e = launch_wide<7>(p, nbatch, (cudaStream_t)stream); break;
      default: e = launch_wide<8>(p, nbatch, (cudaStream_t)stream); break;
    }
    return e == cudaSuccess ? CHB_OK : (int)e;
  }
  if (vec)
    dht_gemm_kernel<true><<<grid, kGemmThreads, kGemmSmem, (cudaStream_t)stream>>>(p);
  else
    dht_gemm_kernel<false><<<grid, kGemmThreads, kGemmSmem, (cudaStream_t)stream>>>(p);
  CHB_RETURN_LAST_ERROR();
}

extern "C" int chb_dht_tile_columns(uint32_t M, uint32_t N_doubles, int nbatch) {
  return 16 * pick_wide_nt(M, N_doubles, nbatch < 1 ? 1 : nbatch);
}

extern "C" int chb_dht(const double* A, uint32_t lda, const double* B, uint32_t ldb,
                       double* C, uint32_t ldc, uint32_t M, uint32_t K, uint32_t N,
                       int is_complex, double alpha_re, double alpha_im, int accumulate,
                       void* stream) {
  return dht_launch(A, lda, &B, 1, ldb, &C, ldc, M, K, N, is_complex, alpha_re, alpha_im,
                    accumulate, nullptr, 0, 1.0, 0.0, 0, stream);
}

extern "C" int chb_dht_batched(const double* A, uint32_t lda, const double* const* B_host,
                               double* const* C_host, int nbatch, uint32_t ldb, uint32_t ldc,
                               uint32_t M, uint32_t K, uint32_t N, int is_complex, void* stream) {
  return dht_launch(A, lda, B_host, nbatch, ldb, C_host, ldc, M, K, N, is_complex, 1.0, 0.0, 0,
                    nullptr, 0, 1.0, 0.0, 0, stream);
}

extern "C" int chb_dht2_hermitian(const double* A, uint32_t lda, const double* B, uint32_t ldb,
                                  double* C1, double a1_re, double a1_im, int accumulate1,
                                  double* C2, double a2_re, double a2_im, int accumulate2,
                                  uint32_t ldc, uint32_t M, uint32_t K, uint32_t N, void* stream) {
  return dht_launch(A, lda, &B, 1, ldb, &C1, ldc, M, K, N, 1, a1_re, a1_im, accumulate1, C2, ldc,
                    a2_re, a2_im, accumulate2, stream, 1);
}

extern "C" int chb_dht2(const double* A, uint32_t lda, const double* B, uint32_t ldb,
                        double* C1, double a1_re, double a1_im, int accumulate1, double* C2,
                        double a2_re, double a2_im, int accumulate2, uint32_t ldc, uint32_t M,
                        uint32_t K, uint32_t N, int is_complex, void* stream) {
  return dht_launch(A, lda, &B, 1, ldb, &C1, ldc, M, K, N, is_complex, a1_re, a1_im,
                    accumulate1, C2, ldc, a2_re, a2_im, accumulate2, stream);
}
