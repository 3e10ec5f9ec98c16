// chimera-b200 batched 1-D complex FFT along x (the contiguous axis), FP64.
//
// Replaces (behaviour, not code) the reference's Reikna FFT plan `_fft`
// (methods/transformer_methods_cl.py:482-509; numpy conventions: forward
// e^{-ikx}, normalised inverse) AND the element-wise passes the reference runs
// around it: the real->complex cast (:301), the x phase shift multiply_by_phase
// (kernels/transformer_generic.cl:58-80; :306-311, :344-348), the real-part
// extraction cast_array_c2d (:351) and the [1:] row-slice copies (:295, :358).
//
// One CTA transforms one row entirely in shared memory: a Stockham autosort FFT
// with radix-8 passes (plus one radix-4/2 pass), every thread holding 8 complex
// values in REGISTERS per pass (transform length is a template parameter, so all
// loops unroll and nothing spills to local memory) and one padded buffer
// (index i -> i + i/8, which makes the stride-8 stores of the first pass
// conflict-free).  Several arrays (components / modes of one fb_transform call)
// go into one launch (blockIdx.y), which removes the wave-quantisation loss of
// Nr-1 ~ 511 rows on 148 SMs.  HBM/L2 traffic is the compulsory 16 B read +
// 16 B write per point (8 B when one side is real).
// Lengths that are not powers of two (the reference's Nx=900 example) use
// Bluestein's chirp-z algorithm on top of the same in-smem power-of-two FFT.
#include "common.cuh"
#include "../../include/chimera_b200.h"

namespace chb {

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmulf(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }  // * (-i)
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }

// forward DFTs of size 2, 4, 8 (sign -), natural-order output
__device__ __forceinline__ void dft2(double2& a, double2& b) {
  double2 s = cadd(a, b);
  b = csub(a, b);
  a = s;
}
__device__ __forceinline__ void dft4(double2& y0, double2& y1, double2& y2, double2& y3) {
  double2 p0 = cadd(y0, y2), p1 = cadd(y1, y3);
  double2 q0 = csub(y0, y2), q1 = mul_mi(csub(y1, y3));
  y0 = cadd(p0, p1);
  y2 = csub(p0, p1);
  y1 = cadd(q0, q1);
  y3 = csub(q0, q1);
}
__device__ __forceinline__ void dft8(double2* v) {
  const double h = 0.70710678118654752440;
  double2 u0 = cadd(v[0], v[4]), u1 = cadd(v[1], v[5]), u2 = cadd(v[2], v[6]), u3 = cadd(v[3], v[7]);
  double2 d0 = csub(v[0], v[4]), d1 = csub(v[1], v[5]), d2 = csub(v[2], v[6]), d3 = csub(v[3], v[7]);
  d1 = make_double2(h * (d1.x + d1.y), h * (d1.y - d1.x));    // * W8^1 = (1 - i)/sqrt2
  d2 = mul_mi(d2);                                            // * W8^2 = -i
  d3 = make_double2(h * (d3.y - d3.x), -h * (d3.x + d3.y));   // * W8^3 = (-1 - i)/sqrt2
  dft4(u0, u1, u2, u3);
  dft4(d0, d1, d2, d3);
  v[0] = u0; v[2] = u1; v[4] = u2; v[6] = u3;
  v[1] = d0; v[3] = d1; v[5] = d2; v[7] = d3;
}

__device__ __forceinline__ int pad(int i) { return i + (i >> 3); }
__host__ __device__ constexpr int padded_len(int n) { return n + (n >> 3); }

struct FftArgs {
  const double* in[CHB_MAX_FIELDS];   // real or complex rows, one entry per batched array
  double* out[CHB_MAX_FIELDS];
  size_t in_stride;      // elements between rows (of the input element type)
  size_t out_stride;
  const double2* tw;     // N (or L for Bluestein) forward roots of unity (N/2 roots for split rows)
  const double2* tw2;    // split rows: W_N^i, i < N/2
  const double2* phase;  // Nx entries or nullptr
  const double2* chirp;  // Bluestein: exp(-i pi n^2 / N), n < N
  const double2* bfft;   // Bluestein: FFT_L(b)/L
  const double* filter;  // optional real (rows x Nx) factor on the output (spectral smoothing)
  int N;
  int inverse;           // normalised inverse transform
  int in_real, out_real; // element types
  int phase_in;          // multiply the input by phase[ix] (backward path) ...
  int phase_out;         // ... or the output (forward path)
  double scale;          // extra real factor on the output
};

__device__ __forceinline__ double2 load_in(const FftArgs& a, const double* row, int ix) {
  double2 v = a.in_real ? make_double2(row[ix], 0.0) : reinterpret_cast<const double2*>(row)[ix];
  if (a.phase_in) v = cmulf(v, __ldg(a.phase + ix));
  if (a.inverse) v = cconj(v);
  return v;
}
__device__ __forceinline__ void store_out(const FftArgs& a, double* row, int ix, double2 v) {
  if (a.inverse) v = cconj(v);
  v.x *= a.scale; v.y *= a.scale;
  if (a.phase_out) v = cmulf(v, __ldg(a.phase + ix));
  if (a.filter) {
    const double f = __ldg(a.filter + (size_t)blockIdx.x * a.N + ix);
    v.x = f * v.x; v.y = f * v.y;
  }
  if (a.out_real) row[ix] = v.x;
  else reinterpret_cast<double2*>(row)[ix] = v;
}

// One Stockham pass of radix 2^LR at sub-transform length NS over s[0..N).
// FIRST: the inputs come straight from global memory (with the fused prologue) instead
// of the shared buffer; LAST: the outputs go straight to global memory (with the fused
// epilogue).  Both accesses are coalesced: the first pass reads j + r*N/R, the last
// pass (NS = N/R) writes k + r*NS, consecutive in the thread index.
// REG (chained transforms, fft_damp_kernel): the first pass takes its inputs from / the
// last pass leaves its outputs in the register file instead, io[m] <-> element
// tid + m*N/8 -- the first pass reads and the last pass writes exactly that element set.
template <int LOGN, int LR, int NS, bool FIRST, bool LAST, bool REG = false>
__device__ __forceinline__ void stockham_pass(double2* s, const double2* __restrict__ tw, int tid,
                                              const FftArgs& a, const double* rin, double* rout,
                                              double2* io = nullptr) {
  constexpr int N = 1 << LOGN, T = N >> 3, R = 1 << LR, NB = 8 >> LR;
  double2 v[8];
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int j = tid + q * T;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (FIRST && REG) v[q * R + r] = io[q + r * (8 / R)];
      else if (FIRST) v[q * R + r] = load_in(a, rin, j + r * (N / R));
      else v[q * R + r] = s[pad(j + r * (N / R))];
    }
  }
  if (!FIRST) __syncthreads();
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int j = tid + q * T;
    const int k = j & (NS - 1);
    double2* b = v + q * R;
    if (NS > 1) {
      const double2 w1 = __ldg(tw + (size_t)k * (N / (NS * R)));
      double2 w = w1;
#pragma unroll
      for (int r = 1; r < R; ++r) {
        b[r] = cmulf(b[r], w);
        if (r + 1 < R) w = cmulf(w, w1);
      }
    }
    if (LR == 3) dft8(b);
    else if (LR == 2) dft4(b[0], b[1], b[2], b[3]);
    else dft2(b[0], b[1]);
    const int j0 = ((j - k) << LR) + k;   // (j / NS) * NS * R + k
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (LAST && REG) io[q + r * (8 / R)] = b[r];
      else if (LAST) store_out(a, rout, j0 + r * NS, b[r]);
      else s[pad(j0 + r * NS)] = b[r];
    }
  }
  if (!LAST) __syncthreads();
}

// GLOBAL_IO: first pass loads from / last pass stores to global memory (REG: registers)
template <int LOGN, int DONE, int NS, bool GLOBAL_IO, bool REG = false>
struct Passes {
  static __device__ __forceinline__ void run(double2* s, const double2* __restrict__ tw, int tid,
                                             const FftArgs& a, const double* rin, double* rout,
                                             double2* io = nullptr) {
    constexpr int REM = LOGN - DONE;
    constexpr int LR = REM >= 3 ? 3 : REM;
    constexpr bool FIRST = GLOBAL_IO && DONE == 0;
    constexpr bool LAST = GLOBAL_IO && (DONE + LR == LOGN);
    stockham_pass<LOGN, LR, NS, FIRST, LAST, REG>(s, tw, tid, a, rin, rout, io);
    Passes<LOGN, DONE + LR, (NS << LR), GLOBAL_IO, REG>::run(s, tw, tid, a, rin, rout, io);
  }
};
template <int LOGN, int NS, bool GLOBAL_IO, bool REG>
struct Passes<LOGN, LOGN, NS, GLOBAL_IO, REG> {
  static __device__ __forceinline__ void run(double2*, const double2* __restrict__, int,
                                             const FftArgs&, const double*, double*,
                                             double2* = nullptr) {}
};

// In-place forward FFT of the padded buffer s (natural order in, natural order
// out), executed by exactly N/8 threads; tw[j] = exp(-2 pi i j / N).
template <int LOGN>
__device__ __forceinline__ void fft_smem_forward(double2* s, const double2* __restrict__ tw, int tid) {
  FftArgs dummy;
  Passes<LOGN, 0, 1, false>::run(s, tw, tid, dummy, nullptr, nullptr);
}

template <int LOGN>
__global__ void __launch_bounds__((1 << LOGN) / 8 < 32 ? 32 : (1 << LOGN) / 8)
fft_pow2_kernel(FftArgs a) {
  extern __shared__ double2 s[];
  constexpr int N = 1 << LOGN, T = N >> 3;
  const int tid = threadIdx.x;
  const double* rin = a.in[blockIdx.y] + (size_t)blockIdx.x * a.in_stride * (a.in_real ? 1 : 2);
  double* rout = a.out[blockIdx.y] + (size_t)blockIdx.x * a.out_stride * (a.out_real ? 1 : 2);
  (void)T;
  // global -> registers -> (smem passes) -> registers -> global: the row never makes
  // an extra round trip through shared memory, and an in-place call is safe because
  // every input of the row is in registers/smem before the last pass writes
  Passes<LOGN, 0, 1, true>::run(s, a.tw, tid, a, rin, rout);
}

// damp_fields of the reference (solver.py:32-35) on one spectral row, in place:
//   half backward transform (x phase, inverse FFT, real part for m = 0)
//   -> profile_edges (solver_ms_pic.cl:5-55: Nf edge columns on each side times prof)
//   -> half forward transform (FFT, x phase)
// The row stays on chip between the two transforms (the last pass of the first one
// leaves in registers exactly the elements the first pass of the second one needs), so
// it is read and written once instead of three times; every element goes through the
// same operations as in the three separate calls (bit-identical results).
struct DampArgs {
  double* row[CHB_MAX_FIELDS];     // spectral arrays (complex rows), transformed in place
  int real_x[CHB_MAX_FIELDS];      // m = 0: the x-space half transform keeps the real part
  size_t stride;                   // complex elements between rows
  const double2* tw;
  const double2* phase_bwd;        // exp(+i kx Xmin) applied before the inverse FFT
  const double2* phase_fwd;        // exp(-i kx Xmin) applied after the forward FFT
  const double* prof;
  int N, Nf;
};

template <int LOGN>
__global__ void __launch_bounds__((1 << LOGN) / 8 < 32 ? 32 : (1 << LOGN) / 8)
fft_damp_kernel(DampArgs d) {
  extern __shared__ double2 s[];
  constexpr int N = 1 << LOGN, T = N >> 3;
  const int tid = threadIdx.x;
  double2* row = reinterpret_cast<double2*>(d.row[blockIdx.y]) + (size_t)blockIdx.x * d.stride;
  const bool real_x = d.real_x[blockIdx.y] != 0;
  FftArgs dummy;
  double2 io[8];
  // inverse transform as conj(FFT(conj(.))) / N, phase on the input
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int ix = tid + m * T;
    io[m] = cconj(cmulf(row[ix], __ldg(d.phase_bwd + ix)));
  }
  Passes<LOGN, 0, 1, true, true>::run(s, d.tw, tid, dummy, nullptr, nullptr, io);
  const double scale = 1.0 / (double)N;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int ix = tid + m * T;
    double2 v = cconj(io[m]);
    v.x *= scale; v.y *= scale;
    if (real_x) v.y = 0.0;
    if (ix < d.Nf) { const double f = __ldg(d.prof + ix); v.x *= f; if (!real_x) v.y *= f; }
    if (ix > N - d.Nf) { const double f = __ldg(d.prof + (N - ix)); v.x *= f; if (!real_x) v.y *= f; }
    io[m] = v;
  }
  __syncthreads();   // the shared buffer of the first transform is free again
  Passes<LOGN, 0, 1, true, true>::run(s, d.tw, tid, dummy, nullptr, nullptr, io);
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int ix = tid + m * T;
    row[ix] = cmulf(io[m], __ldg(d.phase_fwd + ix));
  }
}

template <int LOGN>
static int launch_damp(const DampArgs& d, uint32_t rows, int nbatch, cudaStream_t st) {
  constexpr int L = 1 << LOGN;
  const size_t smem = (size_t)padded_len(L) * sizeof(double2);
  cudaError_t e = cudaFuncSetAttribute(fft_damp_kernel<LOGN>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  fft_damp_kernel<LOGN><<<dim3(rows, nbatch), L / 8, smem, st>>>(d);
  CHB_RETURN_LAST_ERROR();
}

// Rows longer than one CTA's shared memory (N = 16384: 256 KiB): one radix-2
// decimation-in-frequency stage is folded into the load and the row is shared by TWO
// CTAs (blockIdx.z = p): CTA p transforms u_p[n] = (x[n] + (-1)^p x[n+N/2]) * W_N^{p n},
// n < N/2, and owns the outputs X[2k+p].  Both read the whole row (the second read
// hits L2), each writes every other element.
template <int LOGH>
__global__ void __launch_bounds__((1 << LOGH) / 8)
fft_split2_kernel(FftArgs a) {
  extern __shared__ double2 s[];
  constexpr int H = 1 << LOGH, T = H >> 3;
  const int tid = threadIdx.x, par = blockIdx.z;
  const double* rin = a.in[blockIdx.y] + (size_t)blockIdx.x * a.in_stride * (a.in_real ? 1 : 2);
  double* rout = a.out[blockIdx.y] + (size_t)blockIdx.x * a.out_stride * (a.out_real ? 1 : 2);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int i = tid + q * T;
    const double2 lo = load_in(a, rin, i), hi = load_in(a, rin, i + H);
    double2 u = par ? csub(lo, hi) : cadd(lo, hi);
    if (par) u = cmulf(u, __ldg(a.tw2 + i));       // W_N^i, N = 2H
    s[pad(i)] = u;
  }
  __syncthreads();
  fft_smem_forward<LOGH>(s, a.tw, tid);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int i = tid + q * T;
    store_out(a, rout, 2 * i + par, s[pad(i)]);
  }
}

template <int LOGL>
__global__ void __launch_bounds__((1 << LOGL) / 8 < 32 ? 32 : (1 << LOGL) / 8)
fft_bluestein_kernel(FftArgs a) {
  extern __shared__ double2 s[];
  constexpr int L = 1 << LOGL, T = L >> 3;
  const int N = a.N, tid = threadIdx.x;
  const double* rin = a.in[blockIdx.y] + (size_t)blockIdx.x * a.in_stride * (a.in_real ? 1 : 2);
  double* rout = a.out[blockIdx.y] + (size_t)blockIdx.x * a.out_stride * (a.out_real ? 1 : 2);
  for (int i = tid; i < L; i += T)
    s[pad(i)] = i < N ? cmulf(load_in(a, rin, i), __ldg(a.chirp + i)) : make_double2(0.0, 0.0);
  __syncthreads();
  fft_smem_forward<LOGL>(s, a.tw, tid);
  // pointwise product with FFT(b)/L, conjugated so that the second forward FFT
  // acts as the inverse: ifft(z) = conj(fft(conj(z)))
  for (int i = tid; i < L; i += T) s[pad(i)] = cconj(cmulf(s[pad(i)], __ldg(a.bfft + i)));
  __syncthreads();
  fft_smem_forward<LOGL>(s, a.tw, tid);
  for (int i = tid; i < N; i += T)
    store_out(a, rout, i, cmulf(cconj(s[pad(i)]), __ldg(a.chirp + i)));
}

template <int LOGL>
static int launch_fft(const FftArgs& a, bool pow2, uint32_t rows, int nbatch, cudaStream_t st) {
  constexpr int L = 1 << LOGL;
  const int threads = L / 8;   // >= 1; kernels index with tid < L/8 only
  const size_t smem = (size_t)padded_len(L) * sizeof(double2);
  cudaError_t e;
  dim3 grid(rows, nbatch);
  if (pow2) {
    e = cudaFuncSetAttribute(fft_pow2_kernel<LOGL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    fft_pow2_kernel<LOGL><<<grid, threads, smem, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(fft_bluestein_kernel<LOGL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    fft_bluestein_kernel<LOGL><<<grid, threads, smem, st>>>(a);
  }
  CHB_RETURN_LAST_ERROR();
}

}  // namespace chb

using namespace chb;

extern "C" {

int chb_fft_max_pow2(void) { return 16384; }

int chb_fft_x_batched(const double* const* in_host, double* const* out_host, int nbatch,
                      uint32_t rows, uint32_t Nx, size_t in_stride, size_t out_stride,
                      int inverse, int in_real, int out_real, const double* phase,
                      int phase_on_input, const double* twiddles, uint32_t L,
                      const double* chirp, const double* bfft, const double* out_filter,
                      void* stream) {
  if (rows == 0 || Nx == 0 || nbatch == 0) return CHB_OK;
  if (nbatch < 0 || nbatch > CHB_MAX_FIELDS) return CHB_ERR_ARG;
  if (L < 8 || (L & (L - 1)) || L > 16384) return CHB_ERR_ARG;
  const bool pow2 = (L == Nx);
  if (L > 8192 && !pow2) return CHB_ERR_ARG;   // Bluestein only up to L = 8192
  if (!pow2 && (L < 2 * Nx - 1 || !chirp || !bfft)) return CHB_ERR_ARG;
  FftArgs a;
  for (int k = 0; k < CHB_MAX_FIELDS; ++k) {
    a.in[k] = k < nbatch ? in_host[k] : nullptr;
    a.out[k] = k < nbatch ? out_host[k] : nullptr;
  }
  a.in_stride = in_stride; a.out_stride = out_stride;
  a.tw = (const double2*)twiddles;
  a.phase = (const double2*)phase;
  a.chirp = (const double2*)chirp;
  a.bfft = (const double2*)bfft;
  a.filter = out_filter;
  a.N = (int)Nx;
  a.inverse = inverse;
  a.in_real = in_real; a.out_real = out_real;
  a.phase_in = (phase && phase_on_input) ? 1 : 0;
  a.phase_out = (phase && !phase_on_input) ? 1 : 0;
  a.scale = inverse ? 1.0 / (double)Nx : 1.0;
  int logL = 0;
  while ((1u << logL) < L) ++logL;
  cudaStream_t st = (cudaStream_t)stream;
  a.tw2 = nullptr;
  if (L == 16384) {
    // twiddles = [8192 roots of the half-length transform | 8192 values W_16384^i]
    a.tw2 = a.tw + 8192;
    const size_t smem = (size_t)padded_len(8192) * sizeof(double2);
    cudaError_t e = cudaFuncSetAttribute(fft_split2_kernel<13>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    fft_split2_kernel<13><<<dim3(rows, nbatch, 2), 1024, smem, st>>>(a);
    CHB_RETURN_LAST_ERROR();
  }
  switch (logL) {
    case 3: return launch_fft<3>(a, pow2, rows, nbatch, st);
    case 4: return launch_fft<4>(a, pow2, rows, nbatch, st);
    case 5: return launch_fft<5>(a, pow2, rows, nbatch, st);
    case 6: return launch_fft<6>(a, pow2, rows, nbatch, st);
    case 7: return launch_fft<7>(a, pow2, rows, nbatch, st);
    case 8: return launch_fft<8>(a, pow2, rows, nbatch, st);
    case 9: return launch_fft<9>(a, pow2, rows, nbatch, st);
    case 10: return launch_fft<10>(a, pow2, rows, nbatch, st);
    case 11: return launch_fft<11>(a, pow2, rows, nbatch, st);
    case 12: return launch_fft<12>(a, pow2, rows, nbatch, st);
    case 13: return launch_fft<13>(a, pow2, rows, nbatch, st);
    default: return CHB_ERR_ARG;
  }
}

int chb_fft_damp_x_batched(double* const* spec_host, const int* real_x_host, int nbatch,
                           uint32_t rows, uint32_t Nx, size_t stride, const double* phase_bwd,
                           const double* phase_fwd, const double* prof, uint32_t Nf,
                           const double* twiddles, void* stream) {
  if (rows == 0 || nbatch == 0) return CHB_OK;
  if (nbatch < 0 || nbatch > CHB_MAX_FIELDS) return CHB_ERR_ARG;
  if (Nx < 256 || (Nx & (Nx - 1)) || Nx > 8192 || Nf > Nx || !phase_bwd || !phase_fwd || !prof)
    return CHB_ERR_ARG;                    // one CTA per row, 32..1024 threads
  DampArgs d;
  for (int k = 0; k < CHB_MAX_FIELDS; ++k) {
    d.row[k] = k < nbatch ? spec_host[k] : nullptr;
    d.real_x[k] = k < nbatch ? real_x_host[k] : 0;
  }
  d.stride = stride;
  d.tw = (const double2*)twiddles;
  d.phase_bwd = (const double2*)phase_bwd;
  d.phase_fwd = (const double2*)phase_fwd;
  d.prof = prof;
  d.N = (int)Nx; d.Nf = (int)Nf;
  cudaStream_t st = (cudaStream_t)stream;
  int logN = 0;
  while ((1u << logN) < Nx) ++logN;
  switch (logN) {
    case 8: return launch_damp<8>(d, rows, nbatch, st);
    case 9: return launch_damp<9>(d, rows, nbatch, st);
    case 10: return launch_damp<10>(d, rows, nbatch, st);
    case 11: return launch_damp<11>(d, rows, nbatch, st);
    case 12: return launch_damp<12>(d, rows, nbatch, st);
    case 13: return launch_damp<13>(d, rows, nbatch, st);
    default: return CHB_ERR_ARG;
  }
}

int chb_fft_x(const double* in, double* out, uint32_t rows, uint32_t Nx, size_t in_stride,
              size_t out_stride, int inverse, int in_real, int out_real,
              const double* phase, int phase_on_input, const double* twiddles,
              uint32_t L, const double* chirp, const double* bfft, void* stream) {
  return chb_fft_x_batched(&in, &out, 1, rows, Nx, in_stride, out_stride, inverse, in_real,
                           out_real, phase, phase_on_input, twiddles, L, chirp, bfft, nullptr,
                           stream);
}

}  // extern "C"
