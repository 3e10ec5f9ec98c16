// chimera-b200 field gather + relativistic Boris push.
//
// Replaces (behaviour, not code) gather_and_push of the reference,
//   kernels/grid_deposit_m0.cl:280-427 and kernels/grid_deposit_m1.cl:330-511,
// launched from methods/grid_methods_cl.py:168-192.
//
// HBM-bound FP64 streaming (52 B read + 32 B written per particle).  Persistent
// CTAs (one per SM) walk over tiles of <= 64 consecutive cells of one grid row, i.e.
// contiguous ranges of the cell-sorted particle list, with a three-stage software
// pipeline so that no global-load latency is exposed:
//   stage A  tile k+3: cell_offset -> particle range            (registers)
//   stage B  tile k+2: sort_indx of its particles               (registers)
//   stage C  tile k+1: cp.async of the particle attributes and of the 2 x (cells+1)
//                      E/B node stencil (6 comps x (M+1) modes) into the other
//                      shared-memory buffer
//   stage D  tile k  : gather from shared memory + Boris push, results stored
// One __syncthreads per tile.  The mode sum uses per-node factors C,
// 2C*Re(e^{-im theta}), 2C*Im(...) and fused multiply-adds (3 DFMA per component and
// node).  Results agree with the reference arithmetic to a few ulp (tolerance 1e-13
// in the tests); the gating quirks of the reference are kept (storage-index gate
// `sort_indx[ip] < Np_stay`, factor 2 on m>=1 modes, dt_2 = 0.5*FactorPush).
#include "common.cuh"
#include "../../include/chimera_b200.h"

namespace chb {

// mean filling of a tile's particle slots the tile width aims at (measured on cfg3:
// 0.80 -> 0.979 ms, 0.90 -> 0.936, 0.98 -> 0.905, 1.0 -> 0.935; tiles that end up with more
// particles than slots finish the remainder through the synchronous path)
#ifndef CHB_GAT_FILL
#define CHB_GAT_FILL 0.97
#endif
#ifndef CHB_GAT_CELLS
#define CHB_GAT_CELLS 64
#endif
#ifndef CHB_GAT_SLOTS
#define CHB_GAT_SLOTS 2
#endif
constexpr int kGatCells = CHB_GAT_CELLS;      // max cells per tile
constexpr int kGatCols = kGatCells + 1;
#ifndef CHB_GAT_THREADS
#define CHB_GAT_THREADS 512
#endif
constexpr int kGatThreads = CHB_GAT_THREADS;
constexpr int kGatCtasPerSm = 1;              // persistent CTAs per SM (2 x 256 threads measured slower)
constexpr int kGatSlots = CHB_GAT_SLOTS;      // pipelined particles per thread and tile
constexpr int kGatRound = kGatSlots * kGatThreads;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem_src) : "memory");
}

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
  uint32_t tiles_per_row, ntiles, cells_per_tile;
  // walk over empty tiles without a pipeline slot (costs a dependent load per tile: only
  // worth it when most tiles are expected to be empty)
  uint32_t skip_empty;
};

// shared-memory layout of ONE pipeline buffer
template <int M>
struct GatBuf {
  double2 sm[M > 0 ? M : 1][6][2][kGatCols];   // m >= 1 node planes
  double s0[6][2][kGatCols];                   // m = 0 node planes
  double land[6][kGatRound];                   // x y z px py pz of the staged particles
  uint32_t sidx[kGatRound];                    // their storage indices (0xffffffff: none)
};

struct TileDesc {
  uint32_t P0, P1;     // particle range in sorted order
  int ir, ix0, ncell;  // row, first cell column, cells
  bool valid;          // false: the CTA's tile sequence is exhausted
};

struct GatherCtx {
  GridVals g;
  uint32_t np_stay;
  double dt_2;
  int Nx_cell, Nr_cell;
};

// Gather + Boris for one particle whose tile stencil is in `B`.
template <int M>
__device__ __forceinline__ void gather_one(const GatherArgs<M>& a, const GatherCtx& cx,
                                           const GatBuf<M>& B, const TileDesc& d, uint32_t s,
                                           double xp, double yp, double zp, double ux, double uy,
                                           double uz) {
  constexpr int MM = M > 0 ? M : 1;
  const GridVals& g = cx.g;
  double rp;
  int ix, ir;
  cell_coords(xp, yp, zp, g, rp, ix, ir);
  if (!(ix > 0 && ix < cx.Nx_cell - 1 && ir < cx.Nr_cell - 1) || ir < 0) return;

  const double sX1 = __dsub_rn(__dmul_rn(__dsub_rn(xp, g.xmin), g.dx_inv), (double)ix);
  const double sX0 = 1.0 - sX1;
  const double sR1 = __dsub_rn(__dmul_rn(__dsub_rn(rp, g.rmin), g.dr_inv), (double)ir);
  const double sR0 = 1.0 - sR1;
  const double C[4] = {sR0 * sX0, sR0 * sX1, sR1 * sX0, sR1 * sX1};

  // 2*C*e^{-i m theta} per node (factor 2: Hermitian symmetry, grid_deposit_m1.cl:435)
  double cr[MM][4], ci[MM][4];
  if (M > 0) {
    const double rp_inv = 1. / rp;          // unguarded, as the reference (:396)
    double er = yp * rp_inv, ei = -zp * rp_inv;
    const double e1r = er, e1i = ei;
#pragma unroll
    for (int m = 0; m < MM; ++m) {
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
  const bool in_tile = (ir == d.ir) && (ix >= d.ix0) && (ix < d.ix0 + d.ncell);
  if (in_tile) {
    const int cl = ix - d.ix0;
#pragma unroll
    for (int f = 0; f < 6; ++f) {
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const int r = n >> 1, c = cl + (n & 1);
        f_p[f] = fma(C[n], B.s0[f][r][c], f_p[f]);
        if (M > 0) {
#pragma unroll
          for (int m = 0; m < MM; ++m) {
            const double2 v = B.sm[m][f][r][c];
            f_p[f] = fma(cr[m][n], v.x, f_p[f]);
            f_p[f] = fma(-ci[m][n], v.y, f_p[f]);
          }
        }
      }
    }
  } else {
    // particle whose coordinates no longer match its sorted cell (the caller moved
    // it after sorting): same sum straight from global memory
    const size_t i_grid = (size_t)ix + (size_t)ir * (size_t)g.Nx;
#pragma unroll
    for (int f = 0; f < 6; ++f) {
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const size_t node = i_grid + (n & 1) + (size_t)g.Nx * (n >> 1);
        f_p[f] = fma(C[n], __ldg(a.eb[f] + node), f_p[f]);
        if (M > 0) {
#pragma unroll
          for (int m = 0; m < MM; ++m) {
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
  const double dt_2 = cx.dt_2;

  // Boris rotation, grid_deposit_m1.cl:472-507
  double u_p[3] = {ux, uy, uz};
  double um[3], up[3], u0[3], t[3], sv[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) um[k] = u_p[k] + dt_2 * e_p[k];
  double g_p_inv = rsqrt(1. + um[0] * um[0] + um[1] * um[1] + um[2] * um[2]);
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
  g_p_inv = rsqrt(1. + u_p[0] * u_p[0] + u_p[1] * u_p[1] + u_p[2] * u_p[2]);

  a.px[s] = u_p[0];
  a.py[s] = u_p[1];
  a.pz[s] = u_p[2];
  a.g_inv[s] = g_p_inv;
}

// SKIP: tiles without particles are stepped over while fetching descriptors (a dependent
// load per tile, outside the pipeline) -- for particle sets that occupy a fraction of the
// grid; !SKIP: every tile takes a pipeline slot (uniform fillings: nothing to skip, and the
// descriptor fetch stays free of data-dependent control flow)
template <int M, bool SKIP>
__global__ void __launch_bounds__(kGatThreads, kGatCtasPerSm)
gather_push_kernel(GatherArgs<M> a) {
  constexpr int MM = M > 0 ? M : 1;
  extern __shared__ __align__(16) unsigned char gat_smem_raw[];
  GatBuf<M>* bufs = reinterpret_cast<GatBuf<M>*>(gat_smem_raw);

  GatherCtx cx;
  cx.g = load_geom(a.geom);
  cx.np_stay = __ldg(a.np_stay);
  cx.dt_2 = 0.5 * __ldg(a.factor_push);
  cx.Nx_cell = cx.g.Nx - 1;
  cx.Nr_cell = cx.g.Nr - 1;
  const int tid = threadIdx.x;

  // descriptor of the next NON-EMPTY tile of this CTA's stride sequence at or after
  // `tnext` (a rank of a multi-GPU run holds particles in a fraction of the rows: walking
  // the empty tiles through the pipeline cost half of the kernel time at 8 ranks);
  // the loads are CTA-uniform
  uint32_t tnext = blockIdx.x;
  const uint32_t dt = gridDim.x;
  auto load_desc = [&]() {
    TileDesc d;
    d.P0 = d.P1 = 0;
    d.ir = d.ix0 = d.ncell = 0;
    d.valid = false;
    if (SKIP) {
      while (tnext < a.ntiles) {
        const uint32_t t = tnext;
        tnext += dt;
        d.ir = (int)(t / a.tiles_per_row);
        d.ix0 = (int)(t - (uint32_t)d.ir * a.tiles_per_row) * (int)a.cells_per_tile;
        d.ncell = min((int)a.cells_per_tile, cx.Nx_cell - d.ix0);
        const uint32_t c0 = (uint32_t)d.ir * (uint32_t)cx.Nx_cell + (uint32_t)d.ix0;
        d.P0 = __ldg(a.cell_offset + c0);
        d.P1 = __ldg(a.cell_offset + c0 + d.ncell);
        if (d.P1 > d.P0) { d.valid = true; break; }
      }
    } else {
      // every tile takes a pipeline slot: nothing here depends on the loaded offsets
      const uint32_t t = tnext;
      tnext += dt;
      if (t < a.ntiles) {
        d.valid = true;
        d.ir = (int)(t / a.tiles_per_row);
        d.ix0 = (int)(t - (uint32_t)d.ir * a.tiles_per_row) * (int)a.cells_per_tile;
        d.ncell = min((int)a.cells_per_tile, cx.Nx_cell - d.ix0);
        const uint32_t c0 = (uint32_t)d.ir * (uint32_t)cx.Nx_cell + (uint32_t)d.ix0;
        d.P0 = __ldg(a.cell_offset + c0);
        d.P1 = __ldg(a.cell_offset + c0 + d.ncell);
      }
    }
    if (!d.valid) d.P0 = d.P1 = 0;
    return d;
  };
  auto load_sidx = [&](const TileDesc& d, uint32_t* sv) {
#pragma unroll
    for (int j = 0; j < kGatSlots; ++j) {
      const uint32_t ip = d.P0 + tid + j * kGatThreads;
      sv[j] = (ip < d.P1) ? __ldg(a.sort_indx + ip) : 0xffffffffu;
    }
  };
  // stage C: asynchronous copies of one tile into buffer B
  auto issue = [&](const TileDesc& d, const uint32_t* sv, GatBuf<M>& B) {
    if (d.P1 > d.P0) {
#pragma unroll
      for (int j = 0; j < kGatSlots; ++j) {
        const int slot = tid + j * kGatThreads;
        uint32_t s = sv[j];
        if (s >= cx.np_stay) s = 0xffffffffu;   // gate on the STORAGE index (:367-368)
        B.sidx[slot] = s;
        if (s != 0xffffffffu) {
          cp_async8(&B.land[0][slot], a.x + s);
          cp_async8(&B.land[1][slot], a.y + s);
          cp_async8(&B.land[2][slot], a.z + s);
          cp_async8(&B.land[3][slot], a.px + s);
          cp_async8(&B.land[4][slot], a.py + s);
          cp_async8(&B.land[5][slot], a.pz + s);
        }
      }
      const int ncol = min(d.ncell + 1, cx.g.Nx - d.ix0);
      for (int i = tid; i < 6 * 2 * kGatCols; i += kGatThreads) {
        const int f = i / (2 * kGatCols);
        const int rem = i - f * 2 * kGatCols;
        const int r = rem / kGatCols, c = rem - r * kGatCols;
        if (c < ncol) {
          const size_t node = (size_t)(d.ir + r) * cx.g.Nx + d.ix0 + c;
          cp_async8(&B.s0[f][r][c], a.eb[f] + node);
          if (M > 0) {
#pragma unroll
            for (int m = 0; m < MM; ++m)
              cp_async16(&B.sm[m][f][r][c], (const double2*)a.eb[6 * (m + 1) + f] + node);
          }
        }
      }
    }
    cp_async_commit();
  };

  // ---- pipeline prologue
  TileDesc d0 = load_desc();
  uint32_t sv0[kGatSlots], sv1[kGatSlots];
  load_sidx(d0, sv0);
  issue(d0, sv0, bufs[0]);
  TileDesc d1 = load_desc();
  load_sidx(d1, sv1);
  TileDesc d2 = load_desc();
  int cur = 0;

  while (d0.valid) {
    cp_async_wait_all();     // this thread's copies for tile d0 have landed
    __syncthreads();         // everybody's have; the previous tile is fully consumed
    issue(d1, sv1, bufs[cur ^ 1]);                 // stage C for the next tile
    uint32_t sv2[kGatSlots];
    load_sidx(d2, sv2);                            // stage B for the one after
    const TileDesc d3 = load_desc();               // stage A

    // ---- stage D: tile t
    const GatBuf<M>& B = bufs[cur];
    if (d0.P1 > d0.P0) {
#pragma unroll 1
      for (int j = 0; j < kGatSlots; ++j) {
        const int slot = tid + j * kGatThreads;
        const uint32_t s = B.sidx[slot];
        if (s == 0xffffffffu) continue;
        gather_one<M>(a, cx, B, d0, s, B.land[0][slot], B.land[1][slot], B.land[2][slot],
                      B.land[3][slot], B.land[4][slot], B.land[5][slot]);
      }
      // tiles holding more than kGatRound particles: remaining ones synchronously
      for (uint32_t ip = d0.P0 + kGatRound + tid; ip < d0.P1; ip += kGatThreads) {
        const uint32_t s = __ldg(a.sort_indx + ip);
        if (s >= cx.np_stay) continue;
        gather_one<M>(a, cx, B, d0, s, __ldg(a.x + s), __ldg(a.y + s), __ldg(a.z + s),
                      a.px[s], a.py[s], a.pz[s]);
      }
    }
    d0 = d1; d1 = d2; d2 = d3;
#pragma unroll
    for (int j = 0; j < kGatSlots; ++j) { sv1[j] = sv2[j]; }
    cur ^= 1;
  }
  cp_async_wait_all();
}

template <int M>
static int launch_gather(const double* x, const double* y, const double* z, double* px,
                         double* py, double* pz, double* g_inv, const uint32_t* sort_indx,
                         const uint32_t* cell_offset, const double* factor_push,
                         const uint32_t* np_stay, GridGeom g, uint32_t np,
                         const double* const* eb, cudaStream_t st) {
  GatherArgs<M> a{x, y, z, px, py, pz, g_inv, sort_indx, cell_offset, factor_push, np_stay,
                  {}, g, 0, 0, 0, 0};
  for (int k = 0; k < 6 * (M + 1); ++k) a.eb[k] = eb[k];
  // tile width from the mean filling, so that a tile's particles fit one pipelined round
  const double ncells = (double)(g.Nx - 1) * (double)(g.Nr - 1);
  const double ppc = ncells > 0 ? (double)np / ncells : 1.0;
  uint32_t cpt = (uint32_t)(CHB_GAT_FILL * kGatRound / (ppc > 1e-9 ? ppc : 1e-9));
  // mean filling far below what one tile holds although the tile is as wide as it gets:
  // the particles sit in a fraction of the grid (a rank's band of a multi-GPU run, an LWFA
  // slab) -- most tiles are empty
  a.skip_empty = ppc < 0.5 * kGatRound / kGatCells ? 1u : 0u;
  if (cpt > (uint32_t)kGatCells) cpt = kGatCells;
  if (cpt < 4) cpt = 4;
  a.cells_per_tile = cpt;
  a.tiles_per_row = (g.Nx - 1 + cpt - 1) / cpt;
  a.ntiles = a.tiles_per_row * (g.Nr - 1);
  const int smem = 2 * (int)sizeof(GatBuf<M>);
  cudaError_t e = cudaFuncSetAttribute(gather_push_kernel<M, false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(gather_push_kernel<M, true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  uint32_t grid = (uint32_t)(kSMs * kGatCtasPerSm);
  if (a.ntiles < grid) grid = a.ntiles;
  if (a.skip_empty) gather_push_kernel<M, true><<<grid, kGatThreads, smem, st>>>(a);
  else gather_push_kernel<M, false><<<grid, kGatThreads, smem, st>>>(a);
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
    case 0: return launch_gather<0>(x, y, z, px, py, pz, g_inv, sort_indx, cell_offset, factor_push_dev, np_stay_dev, g, np, eb_host, st);
    case 1: return launch_gather<1>(x, y, z, px, py, pz, g_inv, sort_indx, cell_offset, factor_push_dev, np_stay_dev, g, np, eb_host, st);
    default: return launch_gather<2>(x, y, z, px, py, pz, g_inv, sort_indx, cell_offset, factor_push_dev, np_stay_dev, g, np, eb_host, st);
  }
}
