// chimera-b200 charge / current deposition and the small grid fix-up kernels.
//
// Replaces (behaviour, not code) the reference's 4-colour, thread-per-cell
//   depose_scalar  kernels/grid_deposit_m0.cl:20-129, grid_deposit_m1.cl:20-152
//   depose_vector  kernels/grid_deposit_m0.cl:148-277, grid_deposit_m1.cl:171-327
//   treat_axis_*, divide_by_dv_*, warp_axis_*  kernels/grid_generic.cl:4-86
//
// Design (HBM-bound FP64 scatter-add, not GEMM-shaped):
//   * one CTA owns CELLS consecutive cells = one contiguous range of the cell-sorted
//     particle list (cell_offset), processed in batches (one batch in the 32-cell
//     geometry used for the current, see DepGeom);
//   * stage (thread per particle, through sort_indx): the attributes are copied
//     global->shared ASYNCHRONOUSLY (cp.async); on arrival the owning thread converts
//     them in place (sqrt, 1/r, scaled coordinates), once per particle -- and, in the
//     fused variants, pushes the coordinates and computes the next cell index;
//   * accumulate (thread per cell and component): the cell's 2x2 node stencil for all
//     modes in registers from shared memory (padded layout, no bank conflicts for
//     ~uniform fillings) -- no atomics at all inside a cell;
//   * one FP64 RED per node value and CELL (not per particle) to the L2.
// The four colour passes, their launches and the memsets between them are gone;
// sums agree with the reference up to FP64 summation order.
#include <cstdlib>
#include "common.cuh"
#include "../../include/chimera_b200.h"

namespace chb {

// Two CTA geometries (template parameter CELLS):
//   128 cells, batches of 512 (current) / 256 (charge) particles, double buffered (the staging of batch b+1
//        overlaps the accumulation of batch b; per batch only the warps owning its ~32
//        cells accumulate, the others wait at the barrier);
//    32 cells, ONE batch of up to 544 particles, single buffer: no CTA-wide pipeline at
//        all -- every warp of the CTA accumulates at the same time, and the overlap of
//        loads and arithmetic comes from the many small CTAs resident on an SM.
// charge deposit (128-cell geometry): 256-particle batches keep the double buffer at
// 22 KB, so that 8 CTAs (32 warps, 63 registers) fit an SM -- measured 0.35 ms per call
// against 0.42 (320 / 512 per batch) and 0.43 / 0.50 (192 / 128)
#ifndef CHB_SCALAR_BATCH
#define CHB_SCALAR_BATCH 256
#endif
template <int CELLS, bool VEC = true>
struct DepGeom {
  static constexpr int kCells = CELLS;
  static constexpr bool kDouble = CELLS >= 128;
  static constexpr int kBatch = kDouble ? (VEC ? 512 : CHB_SCALAR_BATCH) : CELLS * 17;
  static constexpr int kPad = kBatch + kBatch / 16;
};

__device__ __forceinline__ int pidx(int j) { return j + (j >> 4); }

template <int M, bool VEC>
struct DepArgs {
  const uint32_t* __restrict__ sort_indx;
  const double* __restrict__ x;
  const double* __restrict__ y;
  const double* __restrict__ z;
  const double* __restrict__ px;
  const double* __restrict__ py;
  const double* __restrict__ pz;
  const double* __restrict__ g_inv;
  const double* __restrict__ w;
  const uint32_t* __restrict__ cell_offset;
  double* out[(VEC ? 3 : 1) * (M + 1)];  // [m][comp]
  GridGeom geom;
  int charge;
  uint32_t ncells;
  // PUSH variant only: coordinates are advanced by dt*g_inv*p before the deposit
  double* xw;
  double* yw;
  double* zw;
  const double* __restrict__ dt_dev;
  uint32_t np_total;
  // PUSH: particles that are no longer in the cell the traversal order assumes are
  // queued as records of their 8 derived values (ax ar wp px py pz e0 e1) and
  // deposited one by one by depose_push_tail_kernel.  The recommended queue holds one
  // record per particle and cannot overflow; a caller may pass a smaller one and must then
  // check the counter (*exc_count > exc_cap: records were dropped, J is incomplete).  A
  // slow path inside the main kernel instead would cost registers on its hot loop
  // (measured: 4x the local-memory traffic, 2x the time).
  uint32_t* exc_count;
  double* exc_rec;
  uint32_t exc_cap;
  // PUSH == 2: the second half push, the cell index and the histogram of the final
  // coordinates are produced in the same pass (what chb_push_index would do next)
  uint32_t* indx_in_cell;
  uint32_t* sum_in_cell;
};

__device__ __forceinline__ bool cell_valid(int ix, int ir, const GridVals& g) {
  return ix > 0 && ix < g.Nx - 2 && ir < g.Nr - 2 && ir >= 0;
}

// Deposit of ONE particle straight to global memory (REDs), from its derived values
// ax = (x-xmin)*dx_inv, ar = (r-rmin)*dr_inv, wp = w*g_inv*q, e = (y,z)/r: used for the
// few particles whose cell is not the one the traversal order assumes.  Same operation
// order as the reference depose_vector (grid_deposit_m1.cl:268-310).
// (`a` is the kernel parameter block, declared __grid_constant__ so that passing it by
// reference does not force a per-thread local-memory copy of the whole block.)
template <int M, class Args>
__device__ __noinline__ void deposit_derived(const Args& a, const GridVals& g,
                                                double ax, double ar, double wp, double jx,
                                                double jy, double jz, double e0, double e1) {
  constexpr int MM = M > 0 ? M : 1;
  const int ix = floor_to_int(ax), ir = floor_to_int(ar);
  if (!cell_valid(ix, ir, g)) return;
  double er[MM], ei[MM];
  er[0] = e0;
  ei[0] = e1;
#pragma unroll
  for (int m = 1; m < MM; ++m) {
    er[m] = er[m - 1] * er[0] - ei[m - 1] * ei[0];
    ei[m] = er[m - 1] * ei[0] + ei[m - 1] * er[0];
  }
  double sX1 = __dsub_rn(ax, (double)ix);
  double sX0 = __dsub_rn(1.0, sX1);
  const double sR1 = __dsub_rn(ar, (double)ir);
  const double sR0 = __dsub_rn(1.0, sR1);
  sX0 = __dmul_rn(sX0, wp);
  sX1 = __dmul_rn(sX1, wp);
  const double C[4] = {__dmul_rn(sR0, sX0), __dmul_rn(sR0, sX1), __dmul_rn(sR1, sX0),
                       __dmul_rn(sR1, sX1)};
  const double jk[3] = {jx, jy, jz};
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const size_t node = (size_t)(ix + (n & 1)) + (size_t)(ir + (n >> 1)) * (size_t)g.Nx;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double pj = __dmul_rn(C[n], jk[k]);
      red_add_f64(a.out[k] + node, pj);
      if (M > 0) {
#pragma unroll
        for (int m = 0; m < MM; ++m) {
          double* o = a.out[(m + 1) * 3 + k] + 2 * node;
          red_add_f64(o, pj * er[m]);
          red_add_f64(o + 1, pj * ei[m]);
        }
      }
    }
  }
}

// Threads per CTA: one thread per (cell, component) -- the three current components
// of a cell are accumulated by three threads (each in a different warp, so the
// component index is warp-uniform), which keeps the accumulators at 4*(1+2M) doubles
// per thread and the occupancy high.
// Threads per cell: the three current components (VEC), or -- small geometry, scalar --
// the 2M+1 real mode values (m = 0, Re m = 1, Im m = 1, ...) of the charge density; each
// thread then keeps only the four node accumulators of its value.
template <int M, bool VEC, int CELLS>
struct DepShape {
  static constexpr bool kSplit = !VEC && CELLS < 128 && M > 0;
  static constexpr int kPerCell = VEC ? 3 : (kSplit ? 2 * M + 1 : 1);
  static constexpr int kThreads = CELLS * kPerCell;
  // resident CTAs the launch bounds ask for (registers per thread follow from it)
  static constexpr int kCtas = CELLS >= 128 ? (VEC ? 3 : 8) : (VEC ? 6 : (M > 1 ? 5 : 8));
};

template <int M, bool VEC, int CELLS>
struct DepSmem {
  // staged doubles per particle: raw attributes land here asynchronously and are
  // converted in place to (ax, ar, wp, [px, py, pz], [e0, e1])
  static constexpr int kSlots = VEC ? 8 : 5;
  static constexpr int kBytes = (DepGeom<CELLS, VEC>::kDouble ? 2 : 1) * kSlots * DepGeom<CELLS, VEC>::kPad *
                                (int)sizeof(double);
};

// PUSH >= 1 (J only): the traversal order (sort_indx / cell_offset) is the one of the
// PREVIOUS sort; every particle is first advanced by the half step, then deposited --
// through the cell-ordered fast path if it is still in the cell the order assumes, else
// (a few per cent) through deposit_derived.  This folds push_coords('half') +
// sort_parts + depose_currents of pic_loop.py:70-81 into one pass and removes one full
// sort per step.
// PUSH == 2: the momenta do not change between the two half pushes of a step
// (pic_loop.py:70,75), so the second one is applied while the particle is still in
// registers -- x2 = (x0 + d) + d with the same rounded increment d the two push_xyz
// launches compute -- and x2's cell index and the cell histogram are produced here as
// well: the pass also replaces the second push_coords and index_and_sum_in_cell.
template <int M, bool VEC, int PUSH = 0, int CELLS = 128>
__global__ void __launch_bounds__(DepShape<M, VEC, CELLS>::kThreads, DepShape<M, VEC, CELLS>::kCtas)
depose_kernel(const __grid_constant__ DepArgs<M, VEC> a) {
  static_assert(!PUSH || VEC, "the fused push exists for the current deposit only");
  constexpr int NC = VEC ? 3 : 1;
  constexpr int NT = DepShape<M, VEC, CELLS>::kThreads;
  constexpr bool SPLIT = DepShape<M, VEC, CELLS>::kSplit;
  constexpr int MM = M > 0 ? M : 1;
  constexpr int NS = DepSmem<M, VEC, CELLS>::kSlots;
  constexpr int kDepCells = CELLS;
  constexpr int kDepBatch = DepGeom<CELLS, VEC>::kBatch;
  constexpr int kDepPad = DepGeom<CELLS, VEC>::kPad;
  constexpr bool DB = DepGeom<CELLS, VEC>::kDouble;
  constexpr int KP = (kDepBatch + NT - 1) / NT;     // particles loaded per thread and batch
  // slot layout (VEC):   raw x y z px py pz w g_inv -> ax ar wp px py pz e0 e1
  // slot layout (!VEC):  raw x y z w  -          -> ax ar wp e0 e1
  constexpr int SL_E0 = VEC ? 6 : 3, SL_E1 = VEC ? 7 : 4;
  extern __shared__ double dep_smem[];
  auto slot = [&](int buf, int sl, int p) -> double& {
    return dep_smem[(buf * NS + sl) * kDepPad + p];
  };

  const GridVals g = load_geom(a.geom);
  const int Nx_cell = g.Nx - 1;
  const double q = (double)a.charge;
  const double dt = PUSH ? __ldg(a.dt_dev) : 0.0;

  const uint32_t c0 = blockIdx.x * kDepCells;
  const uint32_t c1 = min(c0 + (uint32_t)kDepCells, a.ncells);
  const uint32_t P0 = a.cell_offset[c0], P1 = a.cell_offset[c1];
  if (P0 == P1) return;

  const int comp = threadIdx.x / kDepCells;           // warp-uniform
  const uint32_t c = c0 + (threadIdx.x - comp * kDepCells);
  uint32_t S = 0, E = 0;
  int ix = 0, ir = 0;
  if (c < c1) {
    S = a.cell_offset[c];
    E = a.cell_offset[c + 1];
    ir = (int)(c / (uint32_t)Nx_cell);
    ix = (int)(c - (uint32_t)ir * (uint32_t)Nx_cell);
  }
  const double dix = (double)ix, dir_ = (double)ir;

  double acc0[4] = {0.0, 0.0, 0.0, 0.0};
  double accm[MM][4][2];
#pragma unroll
  for (int m = 0; m < MM; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n) accm[m][n][0] = accm[m][n][1] = 0.0;

  // sorted indices of the particles this thread stages in the NEXT issued batch
  uint32_t sidx[KP];
  uint32_t sprev[PUSH ? KP : 1];   // ... and of the batch in flight (PUSH: write-back)
  const int lane = threadIdx.x & 31;
  auto load_sidx = [&](uint32_t b0) {
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const uint32_t j = b0 + threadIdx.x + k * NT;
      sidx[k] = (j < P1 && threadIdx.x + k * NT < kDepBatch) ? __ldg(a.sort_indx + j) : 0xffffffffu;
    }
  };
  // asynchronous stage of batch [b0, b0+kDepBatch) into buffer `buf`
  auto issue = [&](uint32_t b0, int buf) {
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const uint32_t s = sidx[k];
      if (PUSH) sprev[k] = s;
      if (s == 0xffffffffu) continue;
      const int p = pidx((int)(threadIdx.x + k * NT));
      cp_async8(&slot(buf, 0, p), a.x + s);
      cp_async8(&slot(buf, 1, p), a.y + s);
      cp_async8(&slot(buf, 2, p), a.z + s);
      if (VEC) {
        cp_async8(&slot(buf, 3, p), a.px + s);
        cp_async8(&slot(buf, 4, p), a.py + s);
        cp_async8(&slot(buf, 5, p), a.pz + s);
        cp_async8(&slot(buf, 6, p), a.w + s);
        cp_async8(&slot(buf, 7, p), a.g_inv + s);
      } else {
        cp_async8(&slot(buf, 3, p), a.w + s);
      }
    }
    cp_async_commit();
  };
  // in-place raw -> derived conversion of the particles this thread staged
  auto convert = [&](uint32_t b0, int buf) {
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      // warp-uniform (kDepBatch and NT are multiples of 32)
      if (!(threadIdx.x - lane + k * NT < kDepBatch)) continue;
      const uint32_t j = b0 + threadIdx.x + k * NT;
      const bool valid = j < P1;
      uint32_t cell2 = 0xffffffffu;
      if (valid) {
      const int p = pidx((int)(threadIdx.x + k * NT));
      double xp = slot(buf, 0, p), yp = slot(buf, 1, p), zp = slot(buf, 2, p);
      double wp;
      if (VEC) wp = __dmul_rn(__dmul_rn(slot(buf, 6, p), slot(buf, 7, p)), q);
      else wp = __dmul_rn(slot(buf, 3, p), q);
      if (PUSH) {
        // half push, same arithmetic as push_xyz (particles_generic.cl:142-151)
        const uint32_t s = sprev[k];
        const double dt_g = __dmul_rn(dt, slot(buf, 7, p));
        const double ddx = __dmul_rn(slot(buf, 3, p), dt_g);
        const double ddy = __dmul_rn(slot(buf, 4, p), dt_g);
        const double ddz = __dmul_rn(slot(buf, 5, p), dt_g);
        xp = __dadd_rn(xp, ddx);
        yp = __dadd_rn(yp, ddy);
        zp = __dadd_rn(zp, ddz);
        if (PUSH == 2) {
          const double x2 = __dadd_rn(xp, ddx), y2 = __dadd_rn(yp, ddy), z2 = __dadd_rn(zp, ddz);
          a.xw[s] = x2; a.yw[s] = y2; a.zw[s] = z2;
          cell2 = cell_index(x2, y2, z2, g);
          a.indx_in_cell[s] = cell2;
        } else {
          a.xw[s] = xp; a.yw[s] = yp; a.zw[s] = zp;
        }
      }
      const double rp = __dsqrt_rn(__dadd_rn(__dmul_rn(yp, yp), __dmul_rn(zp, zp)));
      slot(buf, 0, p) = __dmul_rn(__dsub_rn(xp, g.xmin), g.dx_inv);
      slot(buf, 1, p) = __dmul_rn(__dsub_rn(rp, g.rmin), g.dr_inv);
      slot(buf, 2, p) = wp;
      if (M > 0) {
        // depose_scalar uses the unguarded 1/r (grid_deposit_m1.cl:115),
        // depose_vector guards it (grid_deposit_m1.cl:280-281)
        const double rinv = (VEC && !(rp > 0.0)) ? 0.0 : __drcp_rn(rp);
        slot(buf, SL_E0, p) = __dmul_rn(yp, rinv);
        slot(buf, SL_E1, p) = __dmul_rn(zp, rinv);
      }
      }
      if (PUSH == 2) histogram_add(cell2, valid, a.sum_in_cell);
    }
  };

  load_sidx(P0);
  issue(P0, 0);
  if (DB && P0 + kDepBatch < P1) load_sidx(P0 + kDepBatch);

  int buf = 0;
  for (uint32_t b0 = P0; b0 < P1; b0 += kDepBatch, buf ^= (DB ? 1 : 0)) {
    const uint32_t b1 = min(b0 + (uint32_t)kDepBatch, P1);
    if (!DB && b0 != P0) {        // crowded cells: further batches, one after the other
      __syncthreads();            // the single buffer has been consumed
      load_sidx(b0);
      issue(b0, 0);
    }
    cp_async_wait_all();          // this thread's copies of batch b0 have landed
    convert(b0, buf);
    __syncthreads();              // batch b0 ready for everyone; batch b0-1 fully consumed
    if (DB && b1 < P1) {          // overlap: stage the next batch while accumulating this one
      issue(b1, buf ^ 1);
      if (b1 + kDepBatch < P1) load_sidx(b1 + kDepBatch);
    }
    // ---------------- accumulate: thread per (cell, component)
    const uint32_t js = max(S, b0), je = min(E, b1);
    for (uint32_t j = js; j < je; ++j) {
      const int p = pidx((int)(j - b0));
      const double wp = slot(buf, 2, p);
      double sX1 = __dsub_rn(slot(buf, 0, p), dix);
      double sX0 = __dsub_rn(1.0, sX1);
      const double sR1 = __dsub_rn(slot(buf, 1, p), dir_);
      const double sR0 = __dsub_rn(1.0, sR1);
      if (PUSH) {
        // floor(ax) == ix  <=>  0 <= ax - ix < 1 (the subtraction is exact): a particle
        // that left this cell during the push goes to the exception list instead
        if (!(sX1 >= 0.0 && sX1 < 1.0 && sR1 >= 0.0 && sR1 < 1.0)) {
          if (comp == 0) {
            const uint32_t e = atomicAdd(a.exc_count, 1u);
            if (e < a.exc_cap) {
              double* rec = a.exc_rec + (size_t)e * 8;
#pragma unroll
              for (int sl = 0; sl < 8; ++sl) rec[sl] = slot(buf, sl, p);
            }
          }
          continue;
        }
      }
      sX0 = __dmul_rn(sX0, wp);
      sX1 = __dmul_rn(sX1, wp);
      double pj[4] = {__dmul_rn(sR0, sX0), __dmul_rn(sR0, sX1),
                      __dmul_rn(sR1, sX0), __dmul_rn(sR1, sX1)};
      if (VEC) {
        const double jk = slot(buf, 3 + comp, p);
#pragma unroll
        for (int n = 0; n < 4; ++n) pj[n] = __dmul_rn(pj[n], jk);
      }
      double er[MM], ei[MM];
      if (M > 0 && !(SPLIT && comp == 0)) {
        er[0] = slot(buf, SL_E0, p);
        ei[0] = slot(buf, SL_E1, p);
#pragma unroll
        for (int m = 1; m < MM; ++m) {  // e^{i(m+1)theta}
          er[m] = er[m - 1] * er[0] - ei[m - 1] * ei[0];
          ei[m] = er[m - 1] * ei[0] + ei[m - 1] * er[0];
        }
      }
      if (SPLIT) {
        // this thread's value: comp = 0 -> m = 0; 2m-1 -> Re(mode m); 2m -> Im(mode m)
        if (comp == 0) {
#pragma unroll
          for (int n = 0; n < 4; ++n) acc0[n] = __dadd_rn(acc0[n], pj[n]);
        } else {
          double f = 0.0;
#pragma unroll
          for (int m = 0; m < MM; ++m) {
            if (comp == 2 * m + 1) f = er[m];
            if (comp == 2 * m + 2) f = ei[m];
          }
#pragma unroll
          for (int n = 0; n < 4; ++n) acc0[n] = fma(pj[n], f, acc0[n]);
        }
        continue;
      }
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        acc0[n] = __dadd_rn(acc0[n], pj[n]);
        if (M > 0) {
#pragma unroll
          for (int m = 0; m < MM; ++m) {
            accm[m][n][0] = fma(pj[n], er[m], accm[m][n][0]);
            accm[m][n][1] = fma(pj[n], ei[m], accm[m][n][1]);
          }
        }
      }
    }
  }

  // ---------------- flush: one RED per node value per cell
  if (SPLIT) {
    if (E > S) {
      const int m = (comp + 1) >> 1;                 // mode of this thread's value
      double* base = a.out[0];
#pragma unroll
      for (int k = 1; k <= MM; ++k)
        if (m == k) base = a.out[k];
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const size_t node = (size_t)(ix + (n & 1)) + (size_t)(ir + (n >> 1)) * (size_t)g.Nx;
        red_add_f64(comp == 0 ? base + node : base + 2 * node + ((comp & 1) ? 0 : 1), acc0[n]);
      }
    }
    return;
  }
  if (E > S) {
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const size_t node = (size_t)(ix + (n & 1)) + (size_t)(ir + (n >> 1)) * (size_t)g.Nx;
      red_add_f64(a.out[comp] + node, acc0[n]);
      if (M > 0) {
#pragma unroll
        for (int m = 0; m < MM; ++m) {
          double* o = a.out[(m + 1) * NC + comp] + 2 * node;
          red_add_f64(o, accm[m][n][0]);
          red_add_f64(o + 1, accm[m][n][1]);
        }
      }
    }
  }
}

// PUSH: particles beyond the last real cell of the previous sort (its trash bin) are
// pushed too and, should they have re-entered the box, deposited directly; then the
// queued cell changers are deposited from their records.
template <int M, int PUSH>
__global__ void __launch_bounds__(256)
depose_push_tail_kernel(const __grid_constant__ DepArgs<M, true> a) {
  const GridVals g = load_geom(a.geom);
  const double dt = __ldg(a.dt_dev);
  const double q = (double)a.charge;
  const uint32_t first = __ldg(a.cell_offset + a.ncells);
  for (uint32_t j = first + blockIdx.x * blockDim.x + threadIdx.x; j < a.np_total;
       j += gridDim.x * blockDim.x) {
    const uint32_t s = __ldg(a.sort_indx + j);
    const double ux = a.px[s], uy = a.py[s], uz = a.pz[s], gi = a.g_inv[s];
    const double dt_g = __dmul_rn(dt, gi);
    const double ddx = __dmul_rn(ux, dt_g), ddy = __dmul_rn(uy, dt_g), ddz = __dmul_rn(uz, dt_g);
    const double xp = __dadd_rn(a.xw[s], ddx);
    const double yp = __dadd_rn(a.yw[s], ddy);
    const double zp = __dadd_rn(a.zw[s], ddz);
    if (PUSH == 2) {
      const double x2 = __dadd_rn(xp, ddx), y2 = __dadd_rn(yp, ddy), z2 = __dadd_rn(zp, ddz);
      a.xw[s] = x2; a.yw[s] = y2; a.zw[s] = z2;
      const uint32_t cell2 = cell_index(x2, y2, z2, g);
      a.indx_in_cell[s] = cell2;
      atomicAdd(&a.sum_in_cell[cell2], 1u);
    } else {
      a.xw[s] = xp; a.yw[s] = yp; a.zw[s] = zp;
    }
    const double rp = __dsqrt_rn(__dadd_rn(__dmul_rn(yp, yp), __dmul_rn(zp, zp)));
    const double rinv = !(rp > 0.0) ? 0.0 : __drcp_rn(rp);
    deposit_derived<M>(a, g, __dmul_rn(__dsub_rn(xp, g.xmin), g.dx_inv),
                       __dmul_rn(__dsub_rn(rp, g.rmin), g.dr_inv),
                       __dmul_rn(__dmul_rn(a.w[s], gi), q), ux, uy, uz, __dmul_rn(yp, rinv),
                       __dmul_rn(zp, rinv));
  }
  const uint32_t nexc = min(*a.exc_count, a.exc_cap);
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < nexc; e += gridDim.x * blockDim.x) {
    const double* rec = a.exc_rec + (size_t)e * 8;
    deposit_derived<M>(a, g, rec[0], rec[1], rec[2], rec[3], rec[4], rec[5], rec[6], rec[7]);
  }
}

template <int M, bool VEC, int PUSH, int CELLS>
static int launch_depose_geom(DepArgs<M, VEC>& a, cudaStream_t st) {
  uint32_t grid = (a.ncells + CELLS - 1) / CELLS;
  constexpr int smem = DepSmem<M, VEC, CELLS>::kBytes;
  cudaError_t e = cudaFuncSetAttribute(depose_kernel<M, VEC, PUSH, CELLS>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  depose_kernel<M, VEC, PUSH, CELLS><<<grid, DepShape<M, VEC, CELLS>::kThreads, smem, st>>>(a);
  CHB_RETURN_LAST_ERROR();
}

template <int M, bool VEC, int PUSH = 0>
static int launch_depose(DepArgs<M, VEC>& a, cudaStream_t st) {
  // measured on B200 (cfg3): the small geometry wins for the current (1.53 -> 1.11 ms,
  // 41 % of the stall samples of the 128-cell kernel were warps waiting at the batch
  // barrier), the large one for the charge (0.40 vs 0.43 ms).  CHB_DEP_CELLS=32|128
  // forces one of them (profiling).
  static const int forced = getenv("CHB_DEP_CELLS") ? atoi(getenv("CHB_DEP_CELLS")) : 0;
  const int geom = forced ? forced : (VEC ? 32 : 128);
  if (geom == 32) return launch_depose_geom<M, VEC, PUSH, 32>(a, st);
  return launch_depose_geom<M, VEC, PUSH, 128>(a, st);
}

// ------------------------------------------------------------------ grid fix-ups
struct FieldList {
  double* ptr[CHB_MAX_FIELDS];
  int is_complex[CHB_MAX_FIELDS];
  int n;
};

// (row1 - row0) * dV_inv[1] and row * dV_inv[ir]: the thread that owns column
// ix of row 0 also handles row 1, so no ordering hazard between the two rows.
__global__ void __launch_bounds__(256)
postproc_kernel(FieldList f, uint32_t Nx, uint32_t Nr, const double* __restrict__ dV_inv) {
  const int k = blockIdx.y;
  double* arr = f.ptr[k];
  const int w = f.is_complex[k] ? 2 : 1;          // doubles per grid point
  const size_t row = (size_t)Nx * w;
  const size_t total = row * Nr;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t ir = (uint32_t)(i / row);
    if (ir == 1) continue;
    if (ir == 0) {
      const double v0 = arr[i], v1 = arr[i + row];
      arr[i + row] = __dmul_rn(__dsub_rn(v1, v0), dV_inv[1]);
      arr[i] = __dmul_rn(v0, dV_inv[0]);
    } else {
      arr[i] = __dmul_rn(arr[i], dV_inv[ir]);
    }
  }
}

__global__ void __launch_bounds__(256)
warp_axis_kernel(FieldList f, uint32_t Nx) {
  const int k = blockIdx.y;
  double* arr = f.ptr[k];
  const bool cplx = f.is_complex[k] != 0;
  const uint32_t row = Nx * (cplx ? 2u : 1u);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < row; i += gridDim.x * blockDim.x)
    arr[i] = cplx ? -arr[i + row] : arr[i + row];
}

static int fill_list(FieldList& f, double* const* ptrs, const int* is_complex, int n) {
  if (n < 0 || n > CHB_MAX_FIELDS) return CHB_ERR_ARG;
  f.n = n;
  for (int k = 0; k < CHB_MAX_FIELDS; ++k) {
    f.ptr[k] = k < n ? ptrs[k] : nullptr;
    f.is_complex[k] = k < n ? is_complex[k] : 0;
  }
  return CHB_OK;
}

}  // namespace chb

using namespace chb;

extern "C" {

int chb_depose_scalar(int M, const uint32_t* sort_indx, const double* x, const double* y,
                      const double* z, const double* w, const uint32_t* cell_offset,
                      int charge, uint32_t Nx, uint32_t Nr, const double* xmin,
                      const double* dx_inv, const double* rmin, const double* dr_inv,
                      double* const* rho_host, void* stream) {
  if (M < 0 || M >= CHB_MAX_MODES || Nx < 3 || Nr < 3) return CHB_ERR_ARG;
  GridGeom g{xmin, dx_inv, rmin, dr_inv, Nx, Nr};
  cudaStream_t st = (cudaStream_t)stream;
#define CHB_GO(MM)                                                              \
  {                                                                             \
    DepArgs<MM, false> a{sort_indx, x, y, z, nullptr, nullptr, nullptr, nullptr, w, \
                         cell_offset, {}, g, charge, (Nx - 1) * (Nr - 1),       \
                         nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0,   \
                         nullptr, nullptr};                                     \
    for (int k = 0; k < MM + 1; ++k) a.out[k] = rho_host[k];                    \
    return launch_depose<MM, false>(a, st);                                     \
  }
  switch (M) {
    case 0: CHB_GO(0)
    case 1: CHB_GO(1)
    default: CHB_GO(2)
  }
#undef CHB_GO
}

int chb_depose_vector(int M, const uint32_t* sort_indx, const double* x, const double* y,
                      const double* z, const double* px, const double* py,
                      const double* pz, const double* g_inv, const double* w,
                      const uint32_t* cell_offset, int charge, uint32_t Nx, uint32_t Nr,
                      const double* xmin, const double* dx_inv, const double* rmin,
                      const double* dr_inv, double* const* j_host, void* stream) {
  if (M < 0 || M >= CHB_MAX_MODES || Nx < 3 || Nr < 3) return CHB_ERR_ARG;
  GridGeom g{xmin, dx_inv, rmin, dr_inv, Nx, Nr};
  cudaStream_t st = (cudaStream_t)stream;
#define CHB_GO(MM)                                                              \
  {                                                                             \
    DepArgs<MM, true> a{sort_indx, x, y, z, px, py, pz, g_inv, w, cell_offset,  \
                        {}, g, charge, (Nx - 1) * (Nr - 1),                     \
                        nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0,   \
                        nullptr, nullptr};                                      \
    for (int k = 0; k < 3 * (MM + 1); ++k) a.out[k] = j_host[k];                \
    return launch_depose<MM, true>(a, st);                                      \
  }
  switch (M) {
    case 0: CHB_GO(0)
    case 1: CHB_GO(1)
    default: CHB_GO(2)
  }
#undef CHB_GO
}

size_t chb_push_depose_workspace_bytes(uint32_t np) {
  // counter (16 bytes) + one record of 8 doubles per particle (worst case: every
  // particle changed cell)
  return 16 + (size_t)np * 8 * sizeof(double);
}

static int push_depose(int M, int push, const uint32_t* sort_indx, double* x, double* y,
                       double* z, const double* px, const double* py, const double* pz,
                       const double* g_inv, const double* w, const uint32_t* cell_offset,
                       const double* dt_dev, uint32_t np, int charge, uint32_t Nx, uint32_t Nr,
                       const double* xmin, const double* dx_inv, const double* rmin,
                       const double* dr_inv, double* const* j_host, uint32_t* indx_in_cell,
                       uint32_t* sum_in_cell, void* workspace, size_t workspace_bytes,
                       void* stream) {
  if (M < 0 || M >= CHB_MAX_MODES || Nx < 3 || Nr < 3) return CHB_ERR_ARG;
  if (np == 0) return CHB_OK;
  if (workspace_bytes < 16 + 8 * sizeof(double) ||
      (reinterpret_cast<uintptr_t>(workspace) & 7u) != 0)
    return CHB_ERR_WORKSPACE;
  GridGeom g{xmin, dx_inv, rmin, dr_inv, Nx, Nr};
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* exc_count = (uint32_t*)workspace;
  double* exc_rec = (double*)((char*)workspace + 16);
  size_t cap = (workspace_bytes - 16) / (8 * sizeof(double));
  if (cap > 0xffffffffull) cap = 0xffffffffull;
  {
    cudaError_t e = cudaMemsetAsync(exc_count, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) return (int)e;
  }
#define CHB_GO2(MM, PP)                                                         \
  {                                                                             \
    DepArgs<MM, true> a{sort_indx, x, y, z, px, py, pz, g_inv, w, cell_offset,  \
                        {}, g, charge, (Nx - 1) * (Nr - 1), x, y, z, dt_dev, np,  \
                        exc_count, exc_rec, (uint32_t)cap, indx_in_cell, sum_in_cell}; \
    for (int k = 0; k < 3 * (MM + 1); ++k) a.out[k] = j_host[k];                \
    int rc = launch_depose<MM, true, PP>(a, st);                                \
    if (rc) return rc;                                                          \
    depose_push_tail_kernel<MM, PP><<<kSMs, 256, 0, st>>>(a);                   \
    CHB_RETURN_LAST_ERROR();                                                    \
  }
#define CHB_GO(MM) { if (push == 2) CHB_GO2(MM, 2) else CHB_GO2(MM, 1) }
  switch (M) {
    case 0: CHB_GO(0)
    case 1: CHB_GO(1)
    default: CHB_GO(2)
  }
#undef CHB_GO
#undef CHB_GO2
}

int chb_push_depose_vector(int M, const uint32_t* sort_indx, double* x, double* y, double* z,
                           const double* px, const double* py, const double* pz,
                           const double* g_inv, const double* w,
                           const uint32_t* cell_offset, const double* dt_dev, uint32_t np,
                           int charge, uint32_t Nx, uint32_t Nr, const double* xmin,
                           const double* dx_inv, const double* rmin, const double* dr_inv,
                           double* const* j_host, void* workspace, size_t workspace_bytes,
                           void* stream) {
  return push_depose(M, 1, sort_indx, x, y, z, px, py, pz, g_inv, w, cell_offset, dt_dev, np,
                     charge, Nx, Nr, xmin, dx_inv, rmin, dr_inv, j_host, nullptr, nullptr,
                     workspace, workspace_bytes, stream);
}

int chb_push_depose_push_index(int M, const uint32_t* sort_indx, double* x, double* y,
                               double* z, const double* px, const double* py,
                               const double* pz, const double* g_inv, const double* w,
                               const uint32_t* cell_offset, const double* dt_dev, uint32_t np,
                               int charge, uint32_t Nx, uint32_t Nr, const double* xmin,
                               const double* dx_inv, const double* rmin, const double* dr_inv,
                               double* const* j_host, uint32_t* indx_in_cell,
                               uint32_t* sum_in_cell, void* workspace, size_t workspace_bytes,
                               void* stream) {
  if (!indx_in_cell || !sum_in_cell) return CHB_ERR_ARG;
  return push_depose(M, 2, sort_indx, x, y, z, px, py, pz, g_inv, w, cell_offset, dt_dev, np,
                     charge, Nx, Nr, xmin, dx_inv, rmin, dr_inv, j_host, indx_in_cell,
                     sum_in_cell, workspace, workspace_bytes, stream);
}

int chb_postproc_depose(double* const* fld_host, const int* is_complex_host, int nfld,
                        uint32_t Nx, uint32_t Nr, const double* dV_inv, void* stream) {
  if (nfld == 0) return CHB_OK;
  if (Nr < 2) return CHB_ERR_ARG;
  FieldList f;
  int rc = fill_list(f, fld_host, is_complex_host, nfld);
  if (rc) return rc;
  size_t total = (size_t)Nx * Nr * 2;
  int gx = (int)((total + 255) / 256);
  if (gx > kSMs * 8) gx = kSMs * 8;
  postproc_kernel<<<dim3(gx, nfld), 256, 0, (cudaStream_t)stream>>>(f, Nx, Nr, dV_inv);
  CHB_RETURN_LAST_ERROR();
}

int chb_warp_axis(double* const* fld_host, const int* is_complex_host, int nfld,
                  uint32_t Nx, void* stream) {
  if (nfld == 0) return CHB_OK;
  FieldList f;
  int rc = fill_list(f, fld_host, is_complex_host, nfld);
  if (rc) return rc;
  int gx = (int)((2 * Nx + 255) / 256);
  warp_axis_kernel<<<dim3(gx, nfld), 256, 0, (cudaStream_t)stream>>>(f, Nx);
  CHB_RETURN_LAST_ERROR();
}

}  // extern "C"
