// chimera-b200 batched 1-D complex FFT along x (the contiguous axis), FP64.
//
// Replaces (behaviour, not code) the reference's Reikna FFT plan `_fft`
// (methods/transformer_methods_cl.py:482-509; numpy conventions: forward
// e^{-ikx}, normalised inverse) AND the element-wise passes the reference runs
// around it: the real->complex cast (:301), the x phase shift multiply_by_phase
// (kernels/transformer_generic.cl:58-80; :306-311, :344-348), the real-part
// extraction cast_array_c2d (:351) and the [1:] row-slice copies (:295, :358).
//
// One CTA transforms one row entirely in shared memory (N*16 bytes): a Stockham
// autosort FFT with radix-8 passes (plus one radix-4/2 pass), every thread
// holding 8 complex values in registers per pass, so one buffer suffices.
// HBM traffic is the compulsory 16 B read + 16 B write per point (8 B when one
// side is real); the prologue/epilogue ops ride along for free.
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

// In-place forward FFT of s[0..N) in shared memory, N = 2^logN >= 8, executed by
// exactly N/8 threads (tid in [0, N/8)); other threads of the CTA only hit the
// barriers.  tw[j] = exp(-2 pi i j / N).
__device__ void fft_smem_forward(double2* s, int N, int logN, const double2* __restrict__ tw,
                                 int tid, bool active) {
  const int T = N >> 3;
  int Ns = 1;
  int done = 0;
  while (done < logN) {
    const int rem = logN - done;
    const int lr = rem >= 3 ? 3 : rem;   // log2 of this pass's radix
    const int R = 1 << lr;
    const int nb = 8 >> lr;              // butterflies per thread
    double2 v[8];
    if (active) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q < nb) {
          const int j = tid + q * T;
          for (int r = 0; r < R; ++r) v[q * R + r] = s[j + r * (N / R)];
        }
      }
    }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q < nb) {
          const int j = tid + q * T;
          const int k = j & (Ns - 1);
          double2* b = v + q * R;
          if (Ns > 1) {
            const double2 w1 = __ldg(tw + (size_t)k * (N / (Ns * R)));
            double2 w = w1;
            for (int r = 1; r < R; ++r) {
              b[r] = cmulf(b[r], w);
              if (r + 1 < R) w = cmulf(w, w1);
            }
          }
          if (lr == 3) dft8(b);
          else if (lr == 2) dft4(b[0], b[1], b[2], b[3]);
          else dft2(b[0], b[1]);
          const int j0 = ((j - k) << lr) + k;   // (j / Ns) * Ns * R + k
          for (int r = 0; r < R; ++r) s[j0 + r * Ns] = b[r];
        }
      }
    }
    __syncthreads();
    Ns <<= lr;
    done += lr;
  }
}

struct FftArgs {
  const double* in;      // real or complex rows
  double* out;
  size_t in_stride;      // elements between rows (of the input element type)
  size_t out_stride;
  const double2* tw;     // N (or L for Bluestein) forward roots of unity
  const double2* phase;  // Nx entries or nullptr
  const double2* chirp;  // Bluestein: exp(-i pi n^2 / N), n < N
  const double2* bfft;   // Bluestein: FFT_L(b)/L
  int N, L, logL;
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
  if (a.out_real) row[ix] = v.x;
  else reinterpret_cast<double2*>(row)[ix] = v;
}

__global__ void fft_pow2_kernel(FftArgs a) {
  extern __shared__ double2 s[];
  const int N = a.N, T = N >> 3, tid = threadIdx.x;
  const double* rin = a.in + (size_t)blockIdx.x * a.in_stride * (a.in_real ? 1 : 2);
  double* rout = a.out + (size_t)blockIdx.x * a.out_stride * (a.out_real ? 1 : 2);
  for (int i = tid; i < N; i += blockDim.x) s[i] = load_in(a, rin, i);
  __syncthreads();
  fft_smem_forward(s, N, a.logL, a.tw, tid, tid < T);
  for (int i = tid; i < N; i += blockDim.x) store_out(a, rout, i, s[i]);
}

__global__ void fft_bluestein_kernel(FftArgs a) {
  extern __shared__ double2 s[];
  const int N = a.N, L = a.L, T = L >> 3, tid = threadIdx.x;
  const double* rin = a.in + (size_t)blockIdx.x * a.in_stride * (a.in_real ? 1 : 2);
  double* rout = a.out + (size_t)blockIdx.x * a.out_stride * (a.out_real ? 1 : 2);
  for (int i = tid; i < L; i += blockDim.x)
    s[i] = i < N ? cmulf(load_in(a, rin, i), __ldg(a.chirp + i)) : make_double2(0.0, 0.0);
  __syncthreads();
  fft_smem_forward(s, L, a.logL, a.tw, tid, tid < T);
  // pointwise product with FFT(b)/L, conjugated so that the second forward FFT
  // acts as the inverse: ifft(z) = conj(fft(conj(z)))
  for (int i = tid; i < L; i += blockDim.x) s[i] = cconj(cmulf(s[i], __ldg(a.bfft + i)));
  __syncthreads();
  fft_smem_forward(s, L, a.logL, a.tw, tid, tid < T);
  for (int i = tid; i < N; i += blockDim.x)
    store_out(a, rout, i, cmulf(cconj(s[i]), __ldg(a.chirp + i)));
}

}  // namespace chb

using namespace chb;

extern "C" {

int chb_fft_max_pow2(void) { return 8192; }

int chb_fft_x(const double* in, double* out, uint32_t rows, uint32_t Nx, size_t in_stride,
              size_t out_stride, int inverse, int in_real, int out_real,
              const double* phase, int phase_on_input, const double* twiddles,
              uint32_t L, const double* chirp, const double* bfft, void* stream) {
  if (rows == 0 || Nx == 0) return CHB_OK;
  if (L < 8 || (L & (L - 1)) || L > 8192) return CHB_ERR_ARG;
  const bool pow2 = (L == Nx);
  if (!pow2 && (L < 2 * Nx - 1 || !chirp || !bfft)) return CHB_ERR_ARG;
  FftArgs a;
  a.in = in; a.out = out;
  a.in_stride = in_stride; a.out_stride = out_stride;
  a.tw = (const double2*)twiddles;
  a.phase = (const double2*)phase;
  a.chirp = (const double2*)chirp;
  a.bfft = (const double2*)bfft;
  a.N = (int)Nx; a.L = (int)L;
  a.logL = 0;
  while ((1u << a.logL) < L) ++a.logL;
  a.inverse = inverse;
  a.in_real = in_real; a.out_real = out_real;
  a.phase_in = (phase && phase_on_input) ? 1 : 0;
  a.phase_out = (phase && !phase_on_input) ? 1 : 0;
  a.scale = inverse ? 1.0 / (double)Nx : 1.0;
  int threads = (int)(L >> 3);
  if (threads < 32) threads = 32;
  size_t smem = (size_t)L * sizeof(double2);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if (pow2) {
    e = cudaFuncSetAttribute(fft_pow2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    fft_pow2_kernel<<<rows, threads, smem, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(fft_bluestein_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    fft_bluestein_kernel<<<rows, threads, smem, st>>>(a);
  }
  CHB_RETURN_LAST_ERROR();
}

}  // extern "C"
