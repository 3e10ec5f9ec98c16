// chimera-b200 field gather + relativistic Boris push.
//
// Replaces (behaviour, not code) gather_and_push of the reference,
//   kernels/grid_deposit_m0.cl:280-427 and kernels/grid_deposit_m1.cl:330-511,
// launched from methods/grid_methods_cl.py:168-192.
//
// This file is compiled with -fmad=false and keeps the reference's operation
// order, so px, py, pz, g_inv are bit-identical to the reference kernels built
// without FMA contraction (oracle/_ref) when the field arrays are identical.
//
// HBM-bound: 52 B read + 32 B written per particle; the 2x2 node stencil of the
// 6 field components x (M+1) modes is read through the read-only path and is
// L1/L2 resident because particles are visited in cell-sorted order.
#include "common.cuh"
#include "../../include/chimera_b200.h"

namespace chb {

template <int M>
struct GatherArgs {
  const double* __restrict__ x;
  const double* __restrict__ y;
  const double* __restrict__ z;
  double* __restrict__ px;
  double* __restrict__ py;
  double* __restrict__ pz;
  double* __restrict__ g_inv;
  const uint32_t* __restrict__ sort_indx;
  const double* __restrict__ factor_push;
  const uint32_t* __restrict__ np_stay;
  const double* eb[6 * (M + 1)];  // [m][E,B][x,y,z]
  GridGeom geom;
  uint32_t np;
};

template <int M>
__global__ void __launch_bounds__(256)
gather_push_kernel(GatherArgs<M> a) {
  const GridVals g = load_geom(a.geom);
  const uint32_t np_stay = __ldg(a.np_stay);
  const double dt_2 = 0.5 * __ldg(a.factor_push);
  const int Nx_cell = g.Nx - 1, Nr_cell = g.Nr - 1;

  for (uint32_t ip = blockIdx.x * blockDim.x + threadIdx.x; ip < a.np;
       ip += gridDim.x * blockDim.x) {
    const uint32_t s = __ldg(a.sort_indx + ip);
    if (s >= np_stay) continue;  // gate on the STORAGE index (grid_deposit_m1.cl:367-368)
    const double xp = __ldg(a.x + s), yp = __ldg(a.y + s), zp = __ldg(a.z + s);
    double rp;
    int ix, ir;
    cell_coords(xp, yp, zp, g, rp, ix, ir);
    if (!(ix > 0 && ix < Nx_cell - 1 && ir < Nr_cell - 1)) continue;
    // the reference has no ir >= 0 test here; r >= 0 > Rmin makes it moot, and a
    // NaN radius is rejected by the test above on the host but would index out of
    // bounds there with a negative ir -- we skip it.
    if (ir < 0) continue;

    double u_p[3] = {a.px[s], a.py[s], a.pz[s]};
    const double sX1 = (xp - g.xmin) * g.dx_inv - ix;
    const double sX0 = 1.0 - sX1;
    const double sR1 = (rp - g.rmin) * g.dr_inv - ir;
    const double sR0 = 1.0 - sR1;
    const double C[4] = {sR0 * sX0, sR0 * sX1, sR1 * sX0, sR1 * sX1};

    double er[M > 0 ? M : 1], ei[M > 0 ? M : 1];
    if (M > 0) {
      const double rp_inv = 1. / rp;
      er[0] = yp * rp_inv;       // exp_m1[0]
      ei[0] = -zp * rp_inv;      // exp_m1[1]
#pragma unroll
      for (int m = 1; m < (M > 0 ? M : 1); ++m) {  // e^{-i(m+1)theta}
        er[m] = er[m - 1] * er[0] - ei[m - 1] * ei[0];
        ei[m] = er[m - 1] * ei[0] + ei[m - 1] * er[0];
      }
    }

    const size_t i_grid = (size_t)ix + (size_t)ir * (size_t)g.Nx;
    double e_p[3] = {0, 0, 0}, b_p[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const size_t i_loc = i_grid + (n & 1) + (size_t)g.Nx * (n >> 1);
        e_p[k] += C[n] * __ldg(a.eb[k] + i_loc);
        b_p[k] += C[n] * __ldg(a.eb[3 + k] + i_loc);
        if (M > 0) {
#pragma unroll
          for (int m = 0; m < (M > 0 ? M : 1); ++m) {
            const double2 ev = __ldg((const double2*)a.eb[6 * (m + 1) + k] + i_loc);
            const double2 bv = __ldg((const double2*)a.eb[6 * (m + 1) + 3 + k] + i_loc);
            // factor 2: Hermitian symmetry of the m >= 1 modes (grid_deposit_m1.cl:435)
            e_p[k] += C[n] * (2 * ev.x) * er[m];
            e_p[k] -= C[n] * (2 * ev.y) * ei[m];
            b_p[k] += C[n] * (2 * bv.x) * er[m];
            b_p[k] -= C[n] * (2 * bv.y) * ei[m];
          }
        }
      }
    }

    // Boris rotation, grid_deposit_m1.cl:472-507
    double um[3], up[3], u0[3], t[3], sv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) um[k] = u_p[k] + dt_2 * e_p[k];
    double g_p_inv = 1. / sqrt(1. + um[0] * um[0] + um[1] * um[1] + um[2] * um[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) t[k] = dt_2 * b_p[k] * g_p_inv;
    const double t2p1_m1_05 = 2. / (1. + t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) sv[k] = t[k] * t2p1_m1_05;

    u0[0] = um[0] + um[1] * t[2] - um[2] * t[1];
    u0[1] = um[1] - um[0] * t[2] + um[2] * t[0];
    u0[2] = um[2] + um[0] * t[1] - um[1] * t[0];

    up[0] = um[0] + u0[1] * sv[2] - u0[2] * sv[1];
    up[1] = um[1] - u0[0] * sv[2] + u0[2] * sv[0];
    up[2] = um[2] + u0[0] * sv[1] - u0[1] * sv[0];

#pragma unroll
    for (int k = 0; k < 3; ++k) u_p[k] = up[k] + dt_2 * e_p[k];
    g_p_inv = 1. / sqrt(1. + u_p[0] * u_p[0] + u_p[1] * u_p[1] + u_p[2] * u_p[2]);

    a.px[s] = u_p[0];
    a.py[s] = u_p[1];
    a.pz[s] = u_p[2];
    a.g_inv[s] = g_p_inv;
  }
}

template <int M>
static int launch_gather(const double* x, const double* y, const double* z, double* px,
                         double* py, double* pz, double* g_inv, const uint32_t* sort_indx,
                         const double* factor_push, uint32_t np, const uint32_t* np_stay,
                         GridGeom g, const double* const* eb, cudaStream_t st) {
  GatherArgs<M> a{x, y, z, px, py, pz, g_inv, sort_indx, factor_push, np_stay, {}, g, np};
  for (int k = 0; k < 6 * (M + 1); ++k) a.eb[k] = eb[k];
  uint64_t need = ((uint64_t)np + 255) / 256;
  uint64_t cap = (uint64_t)kSMs * 16;
  int grid = (int)(need < cap ? need : cap);
  gather_push_kernel<M><<<grid, 256, 0, st>>>(a);
  CHB_RETURN_LAST_ERROR();
}

}  // namespace chb

using namespace chb;

extern "C" int chb_gather_push(int M, const double* x, const double* y, const double* z,
                               double* px, double* py, double* pz, double* g_inv,
                               const uint32_t* sort_indx, const double* factor_push_dev,
                               uint32_t np, const uint32_t* np_stay_dev, uint32_t Nx,
                               uint32_t Nr, const double* xmin, const double* dx_inv,
                               const double* rmin, const double* dr_inv,
                               const double* const* eb_host, void* stream) {
  if (M < 0 || M >= CHB_MAX_MODES) return CHB_ERR_ARG;
  if (np == 0) return CHB_OK;
  GridGeom g{xmin, dx_inv, rmin, dr_inv, Nx, Nr};
  cudaStream_t st = (cudaStream_t)stream;
  switch (M) {
    case 0: return launch_gather<0>(x, y, z, px, py, pz, g_inv, sort_indx, factor_push_dev, np, np_stay_dev, g, eb_host, st);
    case 1: return launch_gather<1>(x, y, z, px, py, pz, g_inv, sort_indx, factor_push_dev, np, np_stay_dev, g, eb_host, st);
    default: return launch_gather<2>(x, y, z, px, py, pz, g_inv, sort_indx, factor_push_dev, np, np_stay_dev, g, eb_host, st);
  }
}
