/* TEST INFRASTRUCTURE ONLY -- not part of the product path.
 *
 * Serial / OpenMP "NDRange" drivers around the reference's own OpenCL C
 * kernels, which are #included UNMODIFIED from /root/reference (path given
 * by -DCHIMERA_REF_KERNELS=...).  One shared object per azimuthal-mode
 * count M, because grid_deposit_m0.cl and grid_deposit_m1.cl define the
 * same kernel names (reference: methods/grid_methods_cl.py:21-23 picks the
 * file by M at run time).
 *
 * nd_<kernel>(n_global, parallel, <kernel args>) runs the kernel body for
 * get_global_id(0) = 0..n_global-1, serially (parallel=0, deterministic,
 * gives the stable order for `sort`) or with an OpenMP parallel-for
 * (parallel=1, used only for CPU-baseline timing).
 */
#include "cl_shim.h"

#define CHIMERA_STR2(x) #x
#define CHIMERA_STR(x) CHIMERA_STR2(x)
#define CHIMERA_KFILE(name) CHIMERA_STR(CHIMERA_REF_KERNELS/name)

#include CHIMERA_KFILE(generic.cl)
#include CHIMERA_KFILE(particles_generic.cl)
#include CHIMERA_KFILE(grid_generic.cl)
#include CHIMERA_KFILE(transformer_generic.cl)
#include CHIMERA_KFILE(solver_ms_pic.cl)
#if CHIMERA_REF_M == 0
#include CHIMERA_KFILE(grid_deposit_m0.cl)
#else
#include CHIMERA_KFILE(grid_deposit_m1.cl)
#endif

template <typename... P, typename... A>
static inline void nd_run(size_t n, int par, void (*k)(P...), A... a) {
  if (par) {
#pragma omp parallel for schedule(static)
    for (size_t g = 0; g < n; ++g) {
      chimera_ref_gid0 = g;
      k(a...);
    }
  } else {
    for (size_t g = 0; g < n; ++g) {
      chimera_ref_gid0 = g;
      k(a...);
    }
  }
}

extern "C" {

int nd_ref_mode_count(void) { return CHIMERA_REF_M; }

/* ---- particles_generic.cl ---- */
void nd_push_xyz(size_t n, int par, double *x, double *y, double *z, double *px,
                 double *py, double *pz, double *g_inv, const double *dt,
                 const uint *num_p) {
  nd_run(n, par, push_xyz, x, y, z, px, py, pz, g_inv, dt, num_p);
}
void nd_index_and_sum_in_cell(size_t n, int par, double *x, double *y, double *z,
                              uint *sum_in_cell, const uint *num_p,
                              uint *indx_in_cell, const uint *Nx,
                              const double *xmin, const double *dx_inv,
                              const uint *Nr, const double *rmin,
                              const double *dr_inv) {
  nd_run(n, par, index_and_sum_in_cell, x, y, z, sum_in_cell, num_p,
         indx_in_cell, Nx, xmin, dx_inv, Nr, rmin, dr_inv);
}
void nd_sort(size_t n, int par, uint *cell_offset, uint *indx_in_cell,
             uint *new_sum_in_cell, uint *sorted_indx, uint num_p) {
  nd_run(n, par, sort, cell_offset, indx_in_cell, new_sum_in_cell, sorted_indx,
         num_p);
}
void nd_data_align_dbl(size_t n, int par, double *x, double *x_new,
                       uint *sorted_indx, uint num_p) {
  nd_run(n, par, data_align_dbl, x, x_new, sorted_indx, num_p);
}
void nd_fill_grid(size_t n, int par, double *x, double *y, double *z, double *w,
                  double *theta_var, double *xgrid, double *rgrid, uint Nx,
                  uint ncells, uint Nppc_x, uint Nppc_r, uint Nppc_th) {
  nd_run(n, par, fill_grid, x, y, z, w, theta_var, xgrid, rgrid, Nx, ncells,
         Nppc_x, Nppc_r, Nppc_th);
}
void nd_profile_by_interpolant(size_t n, int par, double *x, double *w, uint Np,
                               double *xx_loc, double *ff_loc, double *dxm1_loc,
                               uint Nx_loc) {
  nd_run(n, par, profile_by_interpolant, x, w, Np, xx_loc, ff_loc, dxm1_loc,
         Nx_loc);
}

/* ---- grid_generic.cl ---- */
void nd_divide_by_dv_d(size_t n, int par, double *arr, const uint *NxNr,
                       const uint *Nx, double *dv_inv) {
  nd_run(n, par, divide_by_dv_d, arr, NxNr, Nx, dv_inv);
}
void nd_divide_by_dv_c(size_t n, int par, double2 *arr, const uint *NxNr,
                       const uint *Nx, double *dv_inv) {
  nd_run(n, par, divide_by_dv_c, arr, NxNr, Nx, dv_inv);
}
void nd_treat_axis_d(size_t n, int par, double *arr, uint Nx) {
  nd_run(n, par, treat_axis_d, arr, Nx);
}
void nd_treat_axis_c(size_t n, int par, double2 *arr, uint Nx) {
  nd_run(n, par, treat_axis_c, arr, Nx);
}
void nd_warp_axis_m0_d(size_t n, int par, double *arr, uint Nx) {
  nd_run(n, par, warp_axis_m0_d, arr, Nx);
}
void nd_warp_axis_m1plus_c(size_t n, int par, double2 *arr, uint Nx) {
  nd_run(n, par, warp_axis_m1plus_c, arr, Nx);
}

/* ---- grid_deposit_m{0,1}.cl ---- */
#if CHIMERA_REF_M == 0
void nd_depose_scalar(size_t n, int par, uint cell_offset, uint *sorting_indx,
                      double *x, double *y, double *z, double *w,
                      uint *indx_offset, char charge, const uint *Nx,
                      const double *xmin, const double *dx_inv, const uint *Nr,
                      const double *rmin, const double *dr_inv,
                      const uint *NxNr_4, double *scl_m0) {
  nd_run(n, par, depose_scalar, cell_offset, sorting_indx, x, y, z, w,
         indx_offset, charge, Nx, xmin, dx_inv, Nr, rmin, dr_inv, NxNr_4,
         scl_m0);
}
void nd_depose_vector(size_t n, int par, uint cell_offset, uint *sorting_indx,
                      double *x, double *y, double *z, double *ux, double *uy,
                      double *uz, double *g_inv, double *w, uint *indx_offset,
                      char charge, const uint *Nx, const double *xmin,
                      const double *dx_inv, const uint *Nr, const double *rmin,
                      const double *dr_inv, const uint *NxNr_4, double *vx0,
                      double *vy0, double *vz0) {
  nd_run(n, par, depose_vector, cell_offset, sorting_indx, x, y, z, ux, uy, uz,
         g_inv, w, indx_offset, charge, Nx, xmin, dx_inv, Nr, rmin, dr_inv,
         NxNr_4, vx0, vy0, vz0);
}
void nd_gather_and_push(size_t n, int par, double *x, double *y, double *z,
                        double *px, double *py, double *pz, double *g_inv,
                        uint *sorting_indx, uint *indx_offset, const double *dt,
                        uint Np, uint Np_stay, const uint *Nx,
                        const double *xmin, const double *dx_inv,
                        const uint *Nr, const double *rmin,
                        const double *dr_inv, const uint *Nxm1Nrm1, double *ex0,
                        double *ey0, double *ez0, double *bx0, double *by0,
                        double *bz0) {
  nd_run(n, par, gather_and_push, x, y, z, px, py, pz, g_inv, sorting_indx,
         indx_offset, dt, Np, Np_stay, Nx, xmin, dx_inv, Nr, rmin, dr_inv,
         Nxm1Nrm1, ex0, ey0, ez0, bx0, by0, bz0);
}
#else
void nd_depose_scalar(size_t n, int par, uint cell_offset, uint *sorting_indx,
                      double *x, double *y, double *z, double *w,
                      uint *indx_offset, char charge, const uint *Nx,
                      const double *xmin, const double *dx_inv, const uint *Nr,
                      const double *rmin, const double *dr_inv,
                      const uint *NxNr_4, double *scl_m0, double2 *scl_m1) {
  nd_run(n, par, depose_scalar, cell_offset, sorting_indx, x, y, z, w,
         indx_offset, charge, Nx, xmin, dx_inv, Nr, rmin, dr_inv, NxNr_4,
         scl_m0, scl_m1);
}
void nd_depose_vector(size_t n, int par, uint cell_offset, uint *sorting_indx,
                      double *x, double *y, double *z, double *ux, double *uy,
                      double *uz, double *g_inv, double *w, uint *indx_offset,
                      char charge, const uint *Nx, const double *xmin,
                      const double *dx_inv, const uint *Nr, const double *rmin,
                      const double *dr_inv, const uint *NxNr_4, double *vx0,
                      double *vy0, double *vz0, double2 *vx1, double2 *vy1,
                      double2 *vz1) {
  nd_run(n, par, depose_vector, cell_offset, sorting_indx, x, y, z, ux, uy, uz,
         g_inv, w, indx_offset, charge, Nx, xmin, dx_inv, Nr, rmin, dr_inv,
         NxNr_4, vx0, vy0, vz0, vx1, vy1, vz1);
}
void nd_gather_and_push(size_t n, int par, double *x, double *y, double *z,
                        double *px, double *py, double *pz, double *g_inv,
                        uint *sorting_indx, uint *indx_offset, const double *dt,
                        uint Np, uint Np_stay, const uint *Nx,
                        const double *xmin, const double *dx_inv,
                        const uint *Nr, const double *rmin,
                        const double *dr_inv, const uint *Nxm1Nrm1, double *ex0,
                        double *ey0, double *ez0, double *bx0, double *by0,
                        double *bz0, double2 *ex1, double2 *ey1, double2 *ez1,
                        double2 *bx1, double2 *by1, double2 *bz1) {
  nd_run(n, par, gather_and_push, x, y, z, px, py, pz, g_inv, sorting_indx,
         indx_offset, dt, Np, Np_stay, Nx, xmin, dx_inv, Nr, rmin, dr_inv,
         Nxm1Nrm1, ex0, ey0, ez0, bx0, by0, bz0, ex1, ey1, ez1, bx1, by1, bz1);
}
#endif

/* ---- generic.cl ---- */
void nd_set_cdouble_to(size_t n, int par, double2 *x, double re, double im,
                       uint arr_size) {
  nd_run(n, par, set_cdouble_to, x, double2{re, im}, arr_size);
}
void nd_append_c2c(size_t n, int par, double2 *arr_base, double2 *arr_add,
                   uint arr_size) {
  nd_run(n, par, append_c2c, arr_base, arr_add, arr_size);
}
void nd_zpaxz_c2c(size_t n, int par, double a_re, double a_im, double2 *x,
                  double2 *z, uint arr_size) {
  nd_run(n, par, zpaxz_c2c, double2{a_re, a_im}, x, z, arr_size);
}
void nd_mult_elementwise_d2c(size_t n, int par, double *x, double2 *z,
                             uint arr_size) {
  nd_run(n, par, mult_elementwise_d2c, x, z, arr_size);
}
void nd_axpbyz_c2c(size_t n, int par, double a_re, double a_im, double2 *x,
                   double b_re, double b_im, double2 *y, double2 *z,
                   uint arr_size) {
  nd_run(n, par, axpbyz_c2c, double2{a_re, a_im}, x, double2{b_re, b_im}, y, z,
         arr_size);
}
void nd_ab_dot_x(size_t n, int par, double a_re, double a_im, double *b,
                 double2 *x, double2 *z, uint NxNr, uint Nx) {
  nd_run(n, par, ab_dot_x, double2{a_re, a_im}, b, x, z, NxNr, Nx);
}
void nd_cast_array_d2c(size_t n, int par, double2 *arr_in, double *arr_out,
                       uint arr_size) {
  nd_run(n, par, cast_array_d2c, arr_in, arr_out, arr_size);
}

/* ---- transformer_generic.cl ---- */
void nd_get_m1(size_t n, int par, double2 *fld_m_m1, double2 *fld_m_p1,
               const uint *Nx, const uint *NxNrm1) {
  nd_run(n, par, get_m1, fld_m_m1, fld_m_p1, Nx, NxNrm1);
}
void nd_get_phase_plus(size_t n, int par, double2 *phs_shft, double *kx,
                       double x0, uint Nx) {
  nd_run(n, par, get_phase_plus, phs_shft, kx, x0, Nx);
}
void nd_get_phase_minus(size_t n, int par, double2 *phs_shft, double *kx,
                        double x0, uint Nx) {
  nd_run(n, par, get_phase_minus, phs_shft, kx, x0, Nx);
}
void nd_multiply_by_phase(size_t n, int par, double2 *arr, const uint *NxNr,
                          const uint *Nx, double2 *phs_shft) {
  nd_run(n, par, multiply_by_phase, arr, NxNr, Nx, phs_shft);
}

/* ---- solver_ms_pic.cl ---- */
void nd_profile_edges_c(size_t n, int par, double2 *x, double *f, uint NxNr,
                        uint Nx, uint Nf) {
  nd_run(n, par, profile_edges_c, x, f, NxNr, Nx, Nf);
}
void nd_profile_edges_d(size_t n, int par, double *x, double *f, uint NxNr,
                        uint Nx, uint Nf) {
  nd_run(n, par, profile_edges_d, x, f, NxNr, Nx, Nf);
}
void nd_advance_e_g_m(size_t n, int par, const uint *NxNr, const double *dt_inv,
                      double *c1, double *c2, double *c3, double2 *ex,
                      double2 *ey, double2 *ez, double2 *gx, double2 *gy,
                      double2 *gz, double2 *jx, double2 *jy, double2 *jz,
                      double2 *n0x, double2 *n0y, double2 *n0z, double2 *n1x,
                      double2 *n1y, double2 *n1z) {
  nd_run(n, par, advance_e_g_m, NxNr, dt_inv, c1, c2, c3, ex, ey, ez, gx, gy,
         gz, jx, jy, jz, n0x, n0y, n0z, n1x, n1y, n1z);
}

} /* extern "C" */
