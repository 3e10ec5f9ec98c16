// Experiment (round-2 groundwork): FP64-accurate contraction C = A.B (A: M x K real, B: K x N
// real) on the int8 tensor path of sm_100a by error-free slicing (Ozaki scheme I):
//   a_ik = 2^ea_i * sum_s qa_s[i,k] 2^-(7s+6),  b_kj = 2^eb_j * sum_t qb_t[k,j] 2^-(7t+6),
//   qa, qb in [-64, 64] (int8), s, t = 0..S-1; the int8 products are accumulated EXACTLY in
//   int32 in TMEM, one accumulator per diagonal d = s+t (d < S), and combined in FP64.
// Layout: both operands pre-tiled in global memory in the UMMA canonical K-major
// no-swizzle layout (see umma_i8.cu), so a k-block of all S slices is one bulk copy.
// Kernel: tile 128 x 64, S accumulators of 64 TMEM columns, warp 0 = bulk-copy producer,
// warp 1 = MMA issuer (one thread), warps 2-5 = epilogue.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Xcompiler -fopenmp -o tools/exp/ozaki_dht tools/exp/ozaki_dht.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

#ifndef S_SLICES
#define S_SLICES 7
#endif
#ifndef KBYTES
#define KBYTES 64
#endif
#ifndef ORDER
#define ORDER 0
#endif
#ifndef NSTAGES
#define NSTAGES 2
#endif
constexpr int S = S_SLICES;                        // slices per operand
constexpr int TM = 128, TN = 64, KB = KBYTES;
constexpr int A_BLK = TM * KB, B_BLK = TN * KB;    // bytes per slice and k-block
constexpr int A_STAGE = S * A_BLK, B_STAGE = S * B_BLK;
constexpr int STAGES = NSTAGES;
constexpr uint32_t A_LBO = (TM / 8) * 128, B_LBO = (TN / 8) * 128, SBO = 128;
constexpr int kThreads = 192;

__host__ __device__ inline int can_off(int row, int k, int rows) {
  return (k / 16) * (rows / 8) * 128 + (row / 8) * 128 + (row % 8) * 16 + (k % 16);
}

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// same, with the A-operand collector hint (fill on the first MMA that uses this A tile,
// use on the following ones, lastuse on the last)
template <int USE>
__device__ __forceinline__ void umma_i8_col(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  if (USE == 0)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8.collector::a::fill [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
  else if (USE == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8.collector::a::use [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8.collector::a::lastuse [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tLAB_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra LAB_DONE_%=;\n\tbra LAB_WAIT_%=;\n\tLAB_DONE_%=:\n\t}\n"
      :: "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
      :: "r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int32_t* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// ------------------------------------------------------------------ slicing of B
// B: K x N row-major (ldb).  One CTA per 64-column tile: column maxima -> power-of-two
// scales, then S int8 slices of every element written in the canonical tile layout
// Bs[ntile][kb][slice][4096].  The second read of the tile hits L2.
__device__ __forceinline__ void slice7(double x, int8_t* q) {
  // x in (-1, 1): x = sum_s q[s] 2^-(7s+6) + O(2^-(7S-1)), q[s] in [-64, 64]; every step exact
  double r = x * 64.0;
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const double qi = rint(r);
    q[s] = (int8_t)(int)qi;
    r = (r - qi) * 128.0;
  }
}

// pass 1: column maxima of |B| (positive doubles order like their bit patterns)
__global__ void __launch_bounds__(256) colmax_kernel(const double* __restrict__ B, int K, int N, int ldb,
                                                     unsigned long long* __restrict__ cmax) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  const int k0 = blockIdx.y * 64, k1 = min(k0 + 64, K);
  double m = 0.0;
  for (int k = k0; k < k1; ++k) m = fmax(m, fabs(B[(size_t)k * ldb + n]));
  atomicMax(cmax + n, (unsigned long long)__double_as_longlong(m));
}

// pass 2: one CTA per (64-column tile, 64-row super block): scales from the maxima, S int8
// slices of every element into the canonical tile layout Bs[ntile][kb][slice][B_BLK]
__global__ void __launch_bounds__(256) slice_b_kernel(const double* __restrict__ B, int K, int N, int ldb,
                                                      const unsigned long long* __restrict__ cmax,
                                                      int8_t* __restrict__ Bs, double* __restrict__ scale_b) {
  const int n0 = blockIdx.x * TN, tid = threadIdx.x, sb = blockIdx.y;
  const int c = tid & 63, g = tid >> 6;
  const int n = n0 + c;
  const int nkb = (K + KB - 1) / KB;
  int e = 0;
  if (n < N) {
    const double m = __longlong_as_double((long long)cmax[n]);
    if (m > 0.0) frexp(m, &e);                        // m = f * 2^e, f in [0.5, 1)  ->  |x| 2^-e < 1
    if (sb == 0 && g == 0) scale_b[n] = ldexp(1.0, e);
  }
  const double inv = ldexp(1.0, -e);
  const int kb = (sb * 64 + g * 16) / KB;
  const int gg = ((sb * 64 + g * 16) % KB) / 16;
  if (kb >= nkb) return;
  int8_t q[16][S];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int k = sb * 64 + g * 16 + i;
    const double x = (k < K && n < N) ? B[(size_t)k * ldb + n] * inv : 0.0;
    slice7(x, q[i]);
  }
  int8_t* blk = Bs + ((size_t)blockIdx.x * nkb + kb) * B_STAGE;
#pragma unroll
  for (int s = 0; s < S; ++s) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      w[j] = (uint32_t)(uint8_t)q[4 * j][s] | ((uint32_t)(uint8_t)q[4 * j + 1][s] << 8) |
             ((uint32_t)(uint8_t)q[4 * j + 2][s] << 16) | ((uint32_t)(uint8_t)q[4 * j + 3][s] << 24);
    *reinterpret_cast<uint4*>(blk + s * B_BLK + can_off(c, gg * 16, TN)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ------------------------------------------------------------------ the contraction
struct OzArgs {
  const int8_t* As;       // [mtile][kb][slice][A_BLK]
  const int8_t* Bs;       // [ntile][kb][slice][B_BLK]
  const double* scale_a;  // 2^ea_i  [mtiles*128]
  const double* scale_b;  // 2^eb_j  [ntiles*64]
  double* C;
  int M, N, ldc, nkb;
  int mode;               // probe: 1 = operands loaded once per tile, 2 = no MMAs issued
};

__global__ void __launch_bounds__(kThreads, 1) ozaki_gemm_kernel(const __grid_constant__ OzArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 1];
  __shared__ uint32_t tmem_base_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  const uint32_t accum_bar = bar0 + 8u * (2 * STAGES);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                 :: "r"((uint32_t)__cvta_generic_to_shared(&tmem_base_smem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  const int nt = blockIdx.x, mt = blockIdx.y, nkb = p.nkb;

  if (warp == 0) {
    if (lane == 0) {
      const int8_t* ga = p.As + (size_t)mt * nkb * A_STAGE;
      const int8_t* gb = p.Bs + (size_t)nt * nkb * B_STAGE;
      for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb % STAGES;
        if (kb >= STAGES) mbar_wait(empty_bar(st), ((kb / STAGES) - 1) & 1);
        if (p.mode == 1 && kb >= STAGES) {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(full_bar(st)) : "memory");
          continue;
        }
        mbar_expect_tx(full_bar(st), A_STAGE + B_STAGE);
        const uint32_t dst = smem_base + st * (A_STAGE + B_STAGE);
        bulk_g2s(dst, ga + (size_t)kb * A_STAGE, A_STAGE, full_bar(st));
        bulk_g2s(dst + A_STAGE, gb + (size_t)kb * B_STAGE, B_STAGE, full_bar(st));
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_i8(TM, TN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb % STAGES;
        mbar_wait(full_bar(st), (kb / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t sa = smem_base + st * (A_STAGE + B_STAGE), sb = sa + A_STAGE;
#if ORDER == 0
#pragma unroll
        for (int d = 0; d < (p.mode == 2 ? 0 : S); ++d) {
#pragma unroll
          for (int s = 0; s <= d; ++s) {
            const int t = d - s;
#pragma unroll
            for (int j = 0; j < KB / 32; ++j) {
              const uint64_t ad = smem_desc(sa + s * A_BLK + j * 2 * A_LBO, A_LBO, SBO);
              const uint64_t bd = smem_desc(sb + t * B_BLK + j * 2 * B_LBO, B_LBO, SBO);
              umma_i8(tmem_base + d * TN, ad, bd, idesc, (kb > 0 || s > 0 || j > 0) ? 1u : 0u);
            }
          }
        }
#else
        // A-stationary order: one A tile (slice s, k-step j) against every B slice it pairs with
#pragma unroll
        for (int s = 0; s < (p.mode == 2 ? 0 : S); ++s) {
#pragma unroll
          for (int j = 0; j < KB / 32; ++j) {
            const uint64_t ad = smem_desc(sa + s * A_BLK + j * 2 * A_LBO, A_LBO, SBO);
#pragma unroll
            for (int t = 0; t < S - s; ++t) {
              const uint64_t bd = smem_desc(sb + t * B_BLK + j * 2 * B_LBO, B_LBO, SBO);
              const uint32_t acc = (kb > 0 || s > 0 || j > 0) ? 1u : 0u;
#if ORDER == 1
              umma_i8(tmem_base + (s + t) * TN, ad, bd, idesc, acc);
#else
              if (S - s == 1) umma_i8(tmem_base + (s + t) * TN, ad, bd, idesc, acc);
              else if (t == 0) umma_i8_col<0>(tmem_base + (s + t) * TN, ad, bd, idesc, acc);
              else if (t == S - s - 1) umma_i8_col<2>(tmem_base + (s + t) * TN, ad, bd, idesc, acc);
              else umma_i8_col<1>(tmem_base + (s + t) * TN, ad, bd, idesc, acc);
#endif
            }
          }
        }
#endif
        umma_commit(empty_bar(st));
      }
      umma_commit(accum_bar);
    }
  } else {
    // epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 = tile rows
    mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const int q = warp & 3;
    const int row = mt * TM + q * 32 + lane;
    const double sa_row = p.scale_a[row] * (1.0 / 4096.0);     // 2^ea * 2^-12 (weights of slice 0)
    for (int c0 = 0; c0 < TN; c0 += 32) {
      double acc[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] = 0.0;
#pragma unroll
      for (int d = S - 1; d >= 0; --d) {
        int32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + d * TN + c0, v);
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = fma(acc[c], 1.0 / 128.0, (double)v[c]);
      }
      if (row < p.M) {
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const int col = nt * TN + c0 + c;
          if (col + 1 < p.N) {
            const double2 sb = *reinterpret_cast<const double2*>(p.scale_b + col);
            *reinterpret_cast<double2*>(p.C + (size_t)row * p.ldc + col) =
                make_double2(acc[c] * sa_row * sb.x, acc[c + 1] * sa_row * sb.y);
          } else if (col < p.N) {
            p.C[(size_t)row * p.ldc + col] = acc[c] * sa_row * p.scale_b[col];
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "r"(512u)
                 : "memory");
}

// ------------------------------------------------------------------ host
static void slice_a_host(const std::vector<double>& A, int M, int K, std::vector<int8_t>& As,
                         std::vector<double>& scale_a) {
  const int mt = (M + TM - 1) / TM, nkb = (K + KB - 1) / KB;
  As.assign((size_t)mt * nkb * A_STAGE, 0);
  scale_a.assign((size_t)mt * TM, 1.0);
  for (int i = 0; i < M; ++i) {
    double m = 0;
    for (int k = 0; k < K; ++k) m = std::fmax(m, std::fabs(A[(size_t)i * K + k]));
    int e = 0;
    if (m > 0) std::frexp(m, &e);
    scale_a[i] = std::ldexp(1.0, e);
    for (int k = 0; k < K; ++k) {
      double r = std::ldexp(A[(size_t)i * K + k], -e) * 64.0;
      for (int s = 0; s < S; ++s) {
        const double qi = std::rint(r);
        As[((size_t)(i / TM) * nkb + k / KB) * A_STAGE + (size_t)s * A_BLK + can_off(i % TM, k % KB, TM)] =
            (int8_t)(int)qi;
        r = (r - qi) * 128.0;
      }
    }
  }
}

int main(int argc, char** argv) {
  const int M = 511, K = 511, N = argc > 1 ? atoi(argv[1]) : 8192;
  std::vector<double> A((size_t)M * K), B((size_t)K * N), Cref((size_t)M * N);
  srand(3);
  auto rnd = [] { return (rand() / (double)RAND_MAX) * 2.0 - 1.0; };
  // Bessel-matrix-like dynamic range in A, columns of B spanning 12 orders of magnitude
  for (int i = 0; i < M; ++i)
    for (int k = 0; k < K; ++k) A[(size_t)i * K + k] = rnd() * std::exp(-3.0 * rnd() * rnd()) / (1.0 + 0.01 * i);
  for (int k = 0; k < K; ++k)
    for (int j = 0; j < N; ++j) B[(size_t)k * N + j] = rnd() * std::pow(10.0, -12.0 * (j % 97) / 96.0);
#pragma omp parallel for
  for (int i = 0; i < M; ++i) {
    std::vector<long double> row(N, 0.0L);
    for (int k = 0; k < K; ++k) {
      const long double a = A[(size_t)i * K + k];
      const double* b = &B[(size_t)k * N];
      for (int j = 0; j < N; ++j) row[j] += a * (long double)b[j];
    }
    for (int j = 0; j < N; ++j) Cref[(size_t)i * N + j] = (double)row[j];
  }

  std::vector<int8_t> As;
  std::vector<double> scale_a;
  slice_a_host(A, M, K, As, scale_a);
  const int mt = (M + TM - 1) / TM, ntl = (N + TN - 1) / TN, nkb = (K + KB - 1) / KB;
  int8_t *dAs, *dBs;
  double *dsa, *dsb, *dB, *dC;
  unsigned long long* dcmax;
  CK(cudaMalloc(&dAs, As.size()));
  CK(cudaMalloc(&dBs, (size_t)ntl * nkb * B_STAGE));
  CK(cudaMalloc(&dsa, scale_a.size() * 8));
  CK(cudaMalloc(&dsb, (size_t)ntl * TN * 8));
  CK(cudaMalloc(&dB, B.size() * 8));
  CK(cudaMalloc(&dcmax, (size_t)N * 8));
  CK(cudaMalloc(&dC, Cref.size() * 8));
  CK(cudaMemcpy(dAs, As.data(), As.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dsa, scale_a.data(), scale_a.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0, Cref.size() * 8));
  const int smem_bytes = STAGES * (A_STAGE + B_STAGE);
  CK(cudaFuncSetAttribute(ozaki_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  OzArgs p{dAs, dBs, dsa, dsb, dC, M, N, N, nkb, argc > 2 ? atoi(argv[2]) : 0};
  cudaEvent_t e0, e1, e2;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  float ms_slice = 0, ms_gemm = 0;
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(e0);
    cudaMemsetAsync(dcmax, 0, (size_t)N * 8);
    colmax_kernel<<<dim3((N + 255) / 256, (K + 63) / 64), 256>>>(dB, K, N, N, dcmax);
    slice_b_kernel<<<dim3(ntl, (K + 63) / 64), 256>>>(dB, K, N, N, dcmax, dBs, dsb);
    cudaEventRecord(e1);
    ozaki_gemm_kernel<<<dim3(ntl, mt), kThreads, smem_bytes>>>(p);
    cudaEventRecord(e2);
    CK(cudaEventSynchronize(e2));
    cudaEventElapsedTime(&ms_slice, e0, e1);
    cudaEventElapsedTime(&ms_gemm, e1, e2);
  }
  std::vector<double> C(Cref.size());
  CK(cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost));
  double cmax = 0, emax = 0, erel_col = 0;
  std::vector<double> colmax(N, 0.0);
  for (size_t i = 0; i < C.size(); ++i) {
    cmax = std::fmax(cmax, std::fabs(Cref[i]));
    colmax[i % N] = std::fmax(colmax[i % N], std::fabs(Cref[i]));
  }
  for (size_t i = 0; i < C.size(); ++i) {
    const double e = std::fabs(C[i] - Cref[i]);
    emax = std::fmax(emax, e);
    if (colmax[i % N] > 0) erel_col = std::fmax(erel_col, e / colmax[i % N]);
  }
  const double flop = 2.0 * M * K * N;
  printf("KB=%d stages=%d order=%d  ", KB, STAGES, ORDER);
  printf("M=%d K=%d N=%d S=%d: slice_b %.3f ms, gemm %.3f ms (%.1f TFLOP/s FP64-equivalent, %.0f TOPS int8)\n",
         M, K, N, S, ms_slice, ms_gemm, flop / ms_gemm / 1e9, flop * (S * (S + 1) / 2) / ms_gemm / 1e9);
  printf("max |err| / max|C| = %.3e   max over columns of |err| / max|C(:,j)| = %.3e\n", emax / cmax, erel_col);
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
