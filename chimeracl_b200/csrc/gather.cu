// chimera-b200 field gather + relativistic Boris push.
//
// Replaces (behaviour, not code) gather_and_push of the reference,
//   kernels/grid_deposit_m0.cl:280-427 and kernels/grid_deposit_m1.cl:330-511,
// launched from methods/grid_methods_cl.py:168-192.
//
// HBM-bound FP64 streaming (52 B read + 32 B written per particle).  One CTA owns
// a tile of kGatCells consecutive cells of one grid row, i.e. one contiguous range
// of the cell-sorted particle list:
//   * the 2 x (cells+1) node stencil of all 6 field components x (M+1) modes is
//     staged once per CTA in shared memory (coalesced row segments), so the 24
//     node reads per particle are LDS broadcasts instead of L1/L2 requests;
//   * particles are visited in sorted order through sort_indx (coalesced index
//     reads, near-sequential attribute reads);
//   * the mode sum uses per-node factors C, 2C*Re(e^{-im theta}), 2C*Im(...) and
//     fused multiply-adds: 3 DFMA per component and node instead of 6 DMUL+4 DADD.
// Results agree with the reference arithmetic to a few ulp (tolerance 1e-13 in
// the tests); the gating quirks of the reference are kept (storage-index gate
// `sort_indx[ip] < Np_stay`, no ir>=0 test needed, factor 2 on m>=1 modes,
// dt_2 = 0.5*FactorPush).
#include "common.cuh"
#include "../../include/chimera_b200.h"

namespace chb {

constexpr int kGatCells = 64;
constexpr int kGatThreads = 256;
constexpr int kGatCols = kGatCells + 1;

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
  const uint32_t* __restrict__ cell_offset;
  const double* __restrict__ factor_push;
  const uint32_t* __restrict__ np_stay;
  const double* eb[6 * (M + 1)];  // [m][E,B][x,y,z]
  GridGeom geom;
  uint32_t tiles_per_row;
};

constexpr int kGatSlots = 4;   // particles in flight per thread (async landing zone)

template <int M>
struct GatSmem {
  static constexpr int kMM = M > 0 ? M : 1;
  static constexpr int kStencilM = M > 0 ? M * 6 * 2 * kGatCols * 2 : 0;   // doubles
  static constexpr int kStencil0 = 6 * 2 * kGatCols;                       // doubles
  static constexpr int kLanding = 6 * kGatSlots * kGatThreads;             // doubles
  static constexpr int kBytes = (kStencilM + kStencil0 + kLanding) * (int)sizeof(double);
};

template <int M>
__global__ void __launch_bounds__(kGatThreads, 3)
gather_push_kernel(GatherArgs<M> a) {
  // node planes: [field 0..5][row 0..1][col], complex for m>=1 (first, 16-byte
  // aligned), real for m=0; then the landing zone [attr 0..5][slot][thread]
  extern __shared__ double2 gat_smem[];
  typedef double2 (*SmT)[6][2][kGatCols];
  typedef double (*S0T)[2][kGatCols];
  SmT sm = reinterpret_cast<SmT>(gat_smem);
  S0T s0 = reinterpret_cast<S0T>(reinterpret_cast<double*>(gat_smem) + GatSmem<M>::kStencilM);
  double* land = reinterpret_cast<double*>(gat_smem) + GatSmem<M>::kStencilM + GatSmem<M>::kStencil0;

  const GridVals g = load_geom(a.geom);
  const int Nx_cell = g.Nx - 1, Nr_cell = g.Nr - 1;
  const int ir_t = blockIdx.x / a.tiles_per_row;
  const int ix0 = (blockIdx.x - ir_t * a.tiles_per_row) * kGatCells;
  const int ncell = min(kGatCells, Nx_cell - ix0);
  const uint32_t c0 = (uint32_t)ir_t * (uint32_t)Nx_cell + (uint32_t)ix0;
  const uint32_t P0 = a.cell_offset[c0], P1 = a.cell_offset[c0 + ncell];
  if (P0 == P1) return;

  const uint32_t np_stay = __ldg(a.np_stay);
  const double dt_2 = 0.5 * __ldg(a.factor_push);

  // Particles of the tile in rounds of kGatSlots*kGatThreads: every thread first
  // issues the asynchronous copies of ALL its particles of the round (6 attributes
  // x kGatSlots in flight per thread, no registers held), then consumes them.
  // A thread only ever reads the slots it filled itself, so no barrier is needed.
  auto issue_round = [&](uint32_t base) {
#pragma unroll
    for (int k = 0; k < kGatSlots; ++k) {
      const uint32_t ip = base + threadIdx.x + k * kGatThreads;
      const uint32_t s = ip < P1 ? __ldg(a.sort_indx + ip) : 0xffffffffu;
      if (s >= np_stay) continue;   // also skips the padding
      double* l = land + k * kGatThreads + threadIdx.x;
      cp_async8(l + 0 * kGatSlots * kGatThreads, a.x + s);
      cp_async8(l + 1 * kGatSlots * kGatThreads, a.y + s);
      cp_async8(l + 2 * kGatSlots * kGatThreads, a.z + s);
      cp_async8(l + 3 * kGatSlots * kGatThreads, a.px + s);
      cp_async8(l + 4 * kGatSlots * kGatThreads, a.py + s);
      cp_async8(l + 5 * kGatSlots * kGatThreads, a.pz + s);
    }
    cp_async_commit();
  };
  issue_round(P0);   // particle data starts flowing while the stencil is staged

  // ---- stage the node stencil (rows ir_t, ir_t+1; columns ix0 .. ix0+ncell)
  const int ncol = min(ncell + 1, g.Nx - ix0);
  for (int i = threadIdx.x; i < 6 * 2 * kGatCols; i += kGatThreads) {
    const int f = i / (2 * kGatCols);
    const int rem = i - f * 2 * kGatCols;
    const int r = rem / kGatCols, c = rem - r * kGatCols;
    if (c < ncol) {
      const size_t node = (size_t)(ir_t + r) * g.Nx + ix0 + c;
      s0[f][r][c] = __ldg(a.eb[f] + node);
#pragma unroll
      for (int m = 0; m < (M > 0 ? M : 1); ++m)
        if (M > 0) sm[m][f][r][c] = __ldg((const double2*)a.eb[6 * (m + 1) + f] + node);
    }
  }
  __syncthreads();


  for (uint32_t base = P0; base < P1; base += kGatSlots * kGatThreads) {
    if (base != P0) issue_round(base);
    cp_async_wait_all();
#pragma unroll 1
    for (int k = 0; k < kGatSlots; ++k) {
    const uint32_t ip = base + threadIdx.x + k * kGatThreads;
    if (ip >= P1) break;
    const uint32_t s = __ldg(a.sort_indx + ip);   // L1 hit (read by issue_round)
    if (s >= np_stay) continue;  // gate on the STORAGE index (grid_deposit_m1.cl:367-368)
    const double* l = land + k * kGatThreads + threadIdx.x;
    const double xp = l[0], yp = l[1 * kGatSlots * kGatThreads], zp = l[2 * kGatSlots * kGatThreads];
    double u_p[3] = {l[3 * kGatSlots * kGatThreads], l[4 * kGatSlots * kGatThreads],
                     l[5 * kGatSlots * kGatThreads]};
    double rp;
    int ix, ir;
    cell_coords(xp, yp, zp, g, rp, ix, ir);
    if (!(ix > 0 && ix < Nx_cell - 1 && ir < Nr_cell - 1) || ir < 0) continue;

    const double sX1 = __dsub_rn(__dmul_rn(__dsub_rn(xp, g.xmin), g.dx_inv), (double)ix);
    const double sX0 = 1.0 - sX1;
    const double sR1 = __dsub_rn(__dmul_rn(__dsub_rn(rp, g.rmin), g.dr_inv), (double)ir);
    const double sR0 = 1.0 - sR1;
    const double C[4] = {sR0 * sX0, sR0 * sX1, sR1 * sX0, sR1 * sX1};

    // 2*C*e^{-i m theta} per node (factor 2: Hermitian symmetry, grid_deposit_m1.cl:435)
    double cr[M > 0 ? M : 1][4], ci[M > 0 ? M : 1][4];
    if (M > 0) {
      const double rp_inv = 1. / rp;          // unguarded, as the reference (:396)
      double er = yp * rp_inv, ei = -zp * rp_inv;
      const double e1r = er, e1i = ei;
#pragma unroll
      for (int m = 0; m < (M > 0 ? M : 1); ++m) {
        if (m > 0) {
          const double t = er * e1r - ei * e1i;
          ei = er * e1i + ei * e1r;
          er = t;
        }
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          cr[m][n] = 2.0 * C[n] * er;
          ci[m][n] = 2.0 * C[n] * ei;
        }
      }
    }

    double f_p[6] = {0, 0, 0, 0, 0, 0};   // E x,y,z then B x,y,z at the particle
    const bool in_tile = (ir == ir_t) && (ix >= ix0) && (ix < ix0 + ncell);
    if (in_tile) {
      const int cl = ix - ix0;
#pragma unroll
      for (int f = 0; f < 6; ++f) {
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const int r = n >> 1, c = cl + (n & 1);
          f_p[f] = fma(C[n], s0[f][r][c], f_p[f]);
#pragma unroll
          for (int m = 0; m < (M > 0 ? M : 1); ++m) {
            if (M > 0) {
              const double2 v = sm[m][f][r][c];
              f_p[f] = fma(cr[m][n], v.x, f_p[f]);
              f_p[f] = fma(-ci[m][n], v.y, f_p[f]);
            }
          }
        }
      }
    } else {
      // particle whose coordinates no longer match its sorted cell (the caller
      // pushed it after sorting): same sum straight from global memory
      const size_t i_grid = (size_t)ix + (size_t)ir * (size_t)g.Nx;
#pragma unroll
      for (int f = 0; f < 6; ++f) {
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const size_t node = i_grid + (n & 1) + (size_t)g.Nx * (n >> 1);
          f_p[f] = fma(C[n], __ldg(a.eb[f] + node), f_p[f]);
#pragma unroll
          for (int m = 0; m < (M > 0 ? M : 1); ++m) {
            if (M > 0) {
              const double2 v = __ldg((const double2*)a.eb[6 * (m + 1) + f] + node);
              f_p[f] = fma(cr[m][n], v.x, f_p[f]);
              f_p[f] = fma(-ci[m][n], v.y, f_p[f]);
            }
          }
        }
      }
    }
    const double* e_p = f_p;
    const double* b_p = f_p + 3;

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
}

template <int M>
static int launch_gather(const double* x, const double* y, const double* z, double* px,
                         double* py, double* pz, double* g_inv, const uint32_t* sort_indx,
                         const uint32_t* cell_offset, const double* factor_push,
                         const uint32_t* np_stay, GridGeom g, const double* const* eb,
                         cudaStream_t st) {
  GatherArgs<M> a{x, y, z, px, py, pz, g_inv, sort_indx, cell_offset, factor_push, np_stay,
                  {}, g, 0};
  for (int k = 0; k < 6 * (M + 1); ++k) a.eb[k] = eb[k];
  a.tiles_per_row = (g.Nx - 1 + kGatCells - 1) / kGatCells;
  uint32_t grid = a.tiles_per_row * (g.Nr - 1);
  constexpr int smem = GatSmem<M>::kBytes;
  cudaError_t e = cudaFuncSetAttribute(gather_push_kernel<M>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  gather_push_kernel<M><<<grid, kGatThreads, smem, st>>>(a);
  CHB_RETURN_LAST_ERROR();
}

}  // namespace chb

using namespace chb;

extern "C" int chb_gather_push(int M, const double* x, const double* y, const double* z,
                               double* px, double* py, double* pz, double* g_inv,
                               const uint32_t* sort_indx, const uint32_t* cell_offset,
                               const double* factor_push_dev, uint32_t np,
                               const uint32_t* np_stay_dev, uint32_t Nx, uint32_t Nr,
                               const double* xmin, const double* dx_inv, const double* rmin,
                               const double* dr_inv, const double* const* eb_host,
                               void* stream) {
  if (M < 0 || M >= CHB_MAX_MODES || Nx < 3 || Nr < 3) return CHB_ERR_ARG;
  if (np == 0) return CHB_OK;
  GridGeom g{xmin, dx_inv, rmin, dr_inv, Nx, Nr};
  cudaStream_t st = (cudaStream_t)stream;
  switch (M) {
    case 0: return launch_gather<0>(x, y, z, px, py, pz, g_inv, sort_indx, cell_offset, factor_push_dev, np_stay_dev, g, eb_host, st);
    case 1: return launch_gather<1>(x, y, z, px, py, pz, g_inv, sort_indx, cell_offset, factor_push_dev, np_stay_dev, g, eb_host, st);
    default: return launch_gather<2>(x, y, z, px, py, pz, g_inv, sort_indx, cell_offset, factor_push_dev, np_stay_dev, g, eb_host, st);
  }
}
