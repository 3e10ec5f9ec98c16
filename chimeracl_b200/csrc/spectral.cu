// chimera-b200 element-wise spectral / field kernels: the complex AXPY family,
// the m = -1 mirror, x phase shift, edge damping profile and the PSATD advance.
//
// Replaces (behaviour, not code) the reference's
//   kernels/generic.cl:4-112, kernels/transformer_generic.cl:3-80,
//   kernels/solver_ms_pic.cl:5-143.
//
// All of these are pure HBM streaming (0.1-0.5 flop/B): 16-byte vector accesses,
// grid-stride loops over a grid sized from the SM count, no shared memory.
// Arithmetic keeps the reference's operation order; contraction into FMA is left
// to the compiler here (results agree to ~1 ulp, tolerance stated in the tests).
#include "common.cuh"
#include "../../include/chimera_b200.h"

namespace chb {

constexpr int kEB = 256;

static inline int ew_grid(size_t n) {
  size_t need = (n + kEB - 1) / kEB;
  size_t cap = (size_t)kSMs * 8;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

#define CHB_GRID_STRIDE(i, n) \
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (n); i += (size_t)gridDim.x * blockDim.x)

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// generic.cl:100-112
__global__ void cast_c2d_kernel(const double2* __restrict__ in, double* __restrict__ out, size_t n) {
  CHB_GRID_STRIDE(i, n) out[i] = in[i].x;
}
// d -> c (astype(complex128) in transformer_methods_cl.py:301)
__global__ void cast_d2c_kernel(const double* __restrict__ in, double2* __restrict__ out, size_t n) {
  CHB_GRID_STRIDE(i, n) out[i] = make_double2(in[i], 0.0);
}
// generic.cl:18-30
__global__ void append_c2c_kernel(double2* __restrict__ base, const double2* __restrict__ add, size_t n) {
  CHB_GRID_STRIDE(i, n) {
    double2 b = base[i], a = add[i];
    base[i] = make_double2(b.x + a.x, b.y + a.y);
  }
}
// generic.cl:33-45: z = z + a*x
__global__ void zpaxz_kernel(double2 a, const double2* __restrict__ x, double2* __restrict__ z, size_t n) {
  CHB_GRID_STRIDE(i, n) {
    double2 xv = x[i], zv = z[i];
    z[i] = make_double2(zv.x + a.x * xv.x - a.y * xv.y, zv.y + a.x * xv.y + a.y * xv.x);
  }
}
// generic.cl:47-58: z = x*z, x real
__global__ void mult_d2c_kernel(const double* __restrict__ x, double2* __restrict__ z, size_t n) {
  CHB_GRID_STRIDE(i, n) {
    double xv = x[i];
    double2 zv = z[i];
    z[i] = make_double2(xv * zv.x, xv * zv.y);
  }
}
// generic.cl:60-77: z = a*x + b*y
__global__ void axpbyz_kernel(double2 a, const double2* __restrict__ x, double2 b,
                              const double2* __restrict__ y, double2* __restrict__ z, size_t n) {
  CHB_GRID_STRIDE(i, n) {
    double2 xv = x[i], yv = y[i];
    z[i] = make_double2(a.x * xv.x - a.y * xv.y + b.x * yv.x - b.y * yv.y,
                        a.x * xv.y + a.y * xv.x + b.x * yv.y + b.y * yv.x);
  }
}
// generic.cl:80-98: z[ir,ix] = b[ix]*(a*x[ir,ix])
__global__ void ab_dot_x_kernel(double2 a, const double* __restrict__ b, const double2* __restrict__ x,
                                double2* __restrict__ z, size_t n, uint32_t Nx) {
  CHB_GRID_STRIDE(i, n) {
    uint32_t ix = (uint32_t)(i % Nx);
    double2 xv = x[i];
    double bv = b[ix];
    z[i] = make_double2(bv * (a.x * xv.x - a.y * xv.y), bv * (a.x * xv.y + a.y * xv.x));
  }
}
// transformer_generic.cl:3-24: F_{-1}(ix) = -conj(F_{+1}((Nx - ix) mod Nx))
__global__ void get_m1_kernel(double2* __restrict__ dst, const double2* __restrict__ src,
                              size_t n, uint32_t Nx) {
  CHB_GRID_STRIDE(i, n) {
    size_t ir = i / Nx;
    uint32_t ix = (uint32_t)(i - ir * Nx);
    uint32_t ixo = ix == 0 ? 0u : Nx - ix;
    double2 v = src[ir * Nx + ixo];
    dst[i] = make_double2(-v.x, v.y);
  }
}
// out (op)= alpha*b(ix) + beta*conj(b((Nx-ix) mod Nx)), row by row.  With the m=-1
// spectrum F_{-1}(kx) = -conj(F_1(-kx)) (get_m1 above) and a REAL operator matrix D,
// D.F_{-1} = -conj(mirror(D.F_1)): the m=0 "minus" terms of field_grad / field_rot
// follow from the "plus" product without a second contraction.
__global__ void mirror_axpy_kernel(double2* __restrict__ out, const double2* __restrict__ b,
                                   double2 alpha, double2 beta, int accumulate, size_t n,
                                   uint32_t Nx) {
  CHB_GRID_STRIDE(i, n) {
    size_t ir = i / Nx;
    uint32_t ix = (uint32_t)(i - ir * Nx);
    uint32_t ixo = ix == 0 ? 0u : Nx - ix;
    const double2 v = b[i];
    double2 w = b[ir * Nx + ixo];
    w.y = -w.y;
    double2 r = make_double2(alpha.x * v.x - alpha.y * v.y + beta.x * w.x - beta.y * w.y,
                             alpha.x * v.y + alpha.y * v.x + beta.x * w.y + beta.y * w.x);
    if (accumulate) { const double2 o = out[i]; r.x += o.x; r.y += o.y; }
    out[i] = r;
  }
}
// transformer_generic.cl:28-55: exp(sign * i * x0 * kx)
__global__ void phase_kernel(double2* __restrict__ phs, const double* __restrict__ kx, double x0,
                             double sign, uint32_t Nx) {
  CHB_GRID_STRIDE(i, Nx) {
    double s, c;
    sincos(x0 * kx[i], &s, &c);
    phs[i] = make_double2(c, sign * s);
  }
}
// transformer_generic.cl:58-80
__global__ void mult_phase_kernel(double2* __restrict__ arr, const double2* __restrict__ phs,
                                  size_t n, uint32_t Nx) {
  CHB_GRID_STRIDE(i, n) arr[i] = cmul(arr[i], phs[i % Nx]);
}

// solver_ms_pic.cl:5-55, batched over arrays: only the 2*Nf edge columns are touched.
struct EdgeList {
  double* ptr[CHB_MAX_FIELDS];
  int is_complex[CHB_MAX_FIELDS];
};
__global__ void profile_edges_kernel(EdgeList f, const double* __restrict__ prof, uint32_t Nr,
                                     uint32_t Nx, uint32_t Nf) {
  const int k = blockIdx.y;
  double* arr = f.ptr[k];
  const bool cplx = f.is_complex[k] != 0;
  // columns ix < Nf and ix > Nx-Nf  (the two sets may overlap when 2*Nf > Nx)
  const uint32_t n_lo = Nf < Nx ? Nf : Nx;
  const uint32_t first_hi = Nx - Nf + 1 > n_lo ? Nx - Nf + 1 : n_lo;  // handled once below
  const uint32_t n_cols_hi = Nx > first_hi ? Nx - first_hi : 0;
  const size_t per_row = (size_t)n_lo + n_cols_hi;
  CHB_GRID_STRIDE(i, per_row * Nr) {
    size_t ir = i / per_row;
    uint32_t c = (uint32_t)(i - ir * per_row);
    uint32_t ix = c < n_lo ? c : first_hi + (c - n_lo);
    size_t e = ir * Nx + ix;
    if (cplx) {
      double2* a = (double2*)arr;
      double2 v = a[e];
      if (ix < Nf) { v.x *= prof[ix]; v.y *= prof[ix]; }
      if (ix > Nx - Nf) { v.x *= prof[Nx - ix]; v.y *= prof[Nx - ix]; }
      a[e] = v;
    } else {
      double v = arr[e];
      if (ix < Nf) v *= prof[ix];
      if (ix > Nx - Nf) v *= prof[Nx - ix];
      arr[e] = v;
    }
  }
}

// solver_ms_pic.cl:57-143: PSATD update of (E, G) from (J, dN0, dN1), one mode.
struct PsatdArgs {
  const double* __restrict__ c1;
  const double* __restrict__ c2;
  const double* __restrict__ c3;
  double2* e[3];
  double2* g[3];
  const double2* j[3];
  const double2* n0[3];
  const double2* n1[3];
  const double* __restrict__ dt_inv;
  size_t n;
};
__global__ void __launch_bounds__(kEB)
psatd_kernel(PsatdArgs a) {
  const double dt_inv = __ldg(a.dt_inv);
  const double pi2 = 2 * 3.14159265358979323846;
  CHB_GRID_STRIDE(i, a.n) {
    const double c1 = a.c1[i], c2 = a.c2[i], c3 = a.c3[i];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double2 e0 = a.e[k][i], g0 = a.g[k][i];
      double2 j0 = a.j[k][i], n0 = a.n0[k][i], n1 = a.n1[k][i];
      double e0v[2] = {e0.x, e0.y}, g0v[2] = {g0.x, g0.y};
      double j0v[2] = {j0.x * pi2, j0.y * pi2};
      double n0v[2] = {n0.x * pi2, n0.y * pi2};
      double n1v[2] = {n1.x * pi2, n1.y * pi2};
      double e1[2], g1[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        e1[c] = c1 * e0v[c] + c2 * c3 * (g0v[c] - j0v[c]) +
                c3 * (c1 * n0v[c] - n1v[c] - (n0v[c] - n1v[c]) * dt_inv * c2 * c3);
        g1[c] = -c2 * e0v[c] + c1 * (g0v[c] - j0v[c]) + j0v[c] +
                c3 * (dt_inv * (1. - c1) * (n0v[c] - n1v[c]) - c2 * n0v[c]);
      }
      a.e[k][i] = make_double2(e1[0], e1[1]);
      a.g[k][i] = make_double2(g1[0], g1[1]);
    }
  }
}

}  // namespace chb

using namespace chb;

#define CHB_ST ((cudaStream_t)stream)

extern "C" {

int chb_cast_c2d(const double* in_c, double* out_d, size_t n, void* stream) {
  if (!n) return CHB_OK;
  cast_c2d_kernel<<<ew_grid(n), kEB, 0, CHB_ST>>>((const double2*)in_c, out_d, n);
  CHB_RETURN_LAST_ERROR();
}
int chb_cast_d2c(const double* in_d, double* out_c, size_t n, void* stream) {
  if (!n) return CHB_OK;
  cast_d2c_kernel<<<ew_grid(n), kEB, 0, CHB_ST>>>(in_d, (double2*)out_c, n);
  CHB_RETURN_LAST_ERROR();
}
int chb_append_c2c(double* base, const double* add, size_t n, void* stream) {
  if (!n) return CHB_OK;
  append_c2c_kernel<<<ew_grid(n), kEB, 0, CHB_ST>>>((double2*)base, (const double2*)add, n);
  CHB_RETURN_LAST_ERROR();
}
int chb_zpaxz_c2c(double a_re, double a_im, const double* x, double* z, size_t n, void* stream) {
  if (!n) return CHB_OK;
  zpaxz_kernel<<<ew_grid(n), kEB, 0, CHB_ST>>>(make_double2(a_re, a_im), (const double2*)x,
                                               (double2*)z, n);
  CHB_RETURN_LAST_ERROR();
}
int chb_mult_elementwise_d2c(const double* x, double* z, size_t n, void* stream) {
  if (!n) return CHB_OK;
  mult_d2c_kernel<<<ew_grid(n), kEB, 0, CHB_ST>>>(x, (double2*)z, n);
  CHB_RETURN_LAST_ERROR();
}
int chb_axpbyz_c2c(double a_re, double a_im, const double* x, double b_re, double b_im,
                   const double* y, double* z, size_t n, void* stream) {
  if (!n) return CHB_OK;
  axpbyz_kernel<<<ew_grid(n), kEB, 0, CHB_ST>>>(make_double2(a_re, a_im), (const double2*)x,
                                                make_double2(b_re, b_im), (const double2*)y,
                                                (double2*)z, n);
  CHB_RETURN_LAST_ERROR();
}
int chb_ab_dot_x(double a_re, double a_im, const double* b, const double* x, double* z,
                 size_t n, uint32_t Nx, void* stream) {
  if (!n) return CHB_OK;
  ab_dot_x_kernel<<<ew_grid(n), kEB, 0, CHB_ST>>>(make_double2(a_re, a_im), b, (const double2*)x,
                                                  (double2*)z, n, Nx);
  CHB_RETURN_LAST_ERROR();
}
int chb_get_m1(double* dst, const double* src, size_t n, uint32_t Nx, void* stream) {
  if (!n) return CHB_OK;
  get_m1_kernel<<<ew_grid(n), kEB, 0, CHB_ST>>>((double2*)dst, (const double2*)src, n, Nx);
  CHB_RETURN_LAST_ERROR();
}
int chb_mirror_axpy(double* out, const double* b, double a_re, double a_im, double b_re,
                    double b_im, int accumulate, size_t n, uint32_t Nx, void* stream) {
  if (!n) return CHB_OK;
  mirror_axpy_kernel<<<ew_grid(n), kEB, 0, CHB_ST>>>((double2*)out, (const double2*)b,
                                                     make_double2(a_re, a_im),
                                                     make_double2(b_re, b_im), accumulate, n, Nx);
  CHB_RETURN_LAST_ERROR();
}
int chb_get_phase(double* phs, const double* kx, double x0, int dir, uint32_t Nx, void* stream) {
  if (!Nx) return CHB_OK;
  phase_kernel<<<ew_grid(Nx), kEB, 0, CHB_ST>>>((double2*)phs, kx, x0, dir == 1 ? 1.0 : -1.0, Nx);
  CHB_RETURN_LAST_ERROR();
}
int chb_multiply_by_phase(double* arr, const double* phs, size_t n, uint32_t Nx, void* stream) {
  if (!n) return CHB_OK;
  mult_phase_kernel<<<ew_grid(n), kEB, 0, CHB_ST>>>((double2*)arr, (const double2*)phs, n, Nx);
  CHB_RETURN_LAST_ERROR();
}
int chb_profile_edges(double* const* fld_host, const int* is_complex_host, int nfld,
                      const double* prof, uint32_t Nr, uint32_t Nx, uint32_t Nf, void* stream) {
  if (nfld == 0 || Nf == 0) return CHB_OK;
  if (nfld < 0 || nfld > CHB_MAX_FIELDS || Nf > Nx) return CHB_ERR_ARG;
  EdgeList f;
  for (int k = 0; k < CHB_MAX_FIELDS; ++k) {
    f.ptr[k] = k < nfld ? fld_host[k] : nullptr;
    f.is_complex[k] = k < nfld ? is_complex_host[k] : 0;
  }
  size_t work = (size_t)2 * Nf * Nr;
  profile_edges_kernel<<<dim3(ew_grid(work), nfld), kEB, 0, CHB_ST>>>(f, prof, Nr, Nx, Nf);
  CHB_RETURN_LAST_ERROR();
}
int chb_psatd_advance(size_t n, const double* dt_inv_dev, const double* c1, const double* c2,
                      const double* c3, double* const* e_host, double* const* g_host,
                      const double* const* j_host, const double* const* n0_host,
                      const double* const* n1_host, void* stream) {
  if (!n) return CHB_OK;
  PsatdArgs a;
  a.c1 = c1; a.c2 = c2; a.c3 = c3; a.dt_inv = dt_inv_dev; a.n = n;
  for (int k = 0; k < 3; ++k) {
    a.e[k] = (double2*)e_host[k];
    a.g[k] = (double2*)g_host[k];
    a.j[k] = (const double2*)j_host[k];
    a.n0[k] = (const double2*)n0_host[k];
    a.n1[k] = (const double2*)n1_host[k];
  }
  psatd_kernel<<<ew_grid(n), kEB, 0, CHB_ST>>>(a);
  CHB_RETURN_LAST_ERROR();
}

}  // extern "C"
