/* chimera_b200.h -- C ABI of libchimera_b200.so (sm_100a).
 *
 * Drop-in boundary for chimeraCL's per-step PIC hot path.  The reference has no
 * FFI: its methods/ mixins launch OpenCL kernels through PyOpenCL.  Each entry
 * point below replaces one such launch site (cited as file:line relative to the
 * reference repository) and is what a chimeraCL maintainer would bind from the
 * corresponding mixin method (see INTEGRATION.md for the ctypes stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - "device scalars" (xmin, dx_inv, rmin, dr_inv, dt) are 1-element device
 *     arrays, exactly the `__constant double *` arguments of the reference
 *     kernels, i.e. DataDev['Xmin'] etc. (generic_methods_cl.py:56-76);
 *   - arrays are C-ordered (Nr, Nx); complex = interleaved (re, im) doubles;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - functions never allocate, never synchronise and never throw: they return
 *     0 on success, a positive cudaError_t, or a negative CHB_ERR_* code;
 *   - workspaces are sized by the matching *_workspace_bytes() call.
 */
#ifndef CHIMERA_B200_H
#define CHIMERA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHB_MAX_ATTRS 8      /* x y z px py pz g_inv w */
#define CHB_MAX_MODES 3      /* azimuthal modes m = 0..2 */
#define CHB_MAX_FIELDS 16    /* arrays per batched grid launch */
#define CHB_GIANT_CAP 4096   /* cells with > 8192 particles handled per sort */

int chb_version(void);
const char* chb_error_string(int code);

/* ---------------------------------------------------------------- particles */

/* x += px*g_inv*dt (and y, z).  Replaces push_xyz,
 * kernels/particles_generic.cl:129-153, launched from
 * methods/particles_methods_cl.py:206-223. */
int chb_push_xyz(double* x, double* y, double* z, const double* px, const double* py,
                 const double* pz, const double* g_inv, const double* dt_dev,
                 uint32_t np, void* stream);

/* Cell index with trash bin + histogram.  sum_in_cell (nbins = (Nx-1)(Nr-1)+1)
 * must be zeroed by the caller.  Replaces index_and_sum_in_cell,
 * kernels/particles_generic.cl:88-126 (particles_methods_cl.py:225-243). */
int chb_index_and_sum(const double* x, const double* y, const double* z,
                      uint32_t* indx_in_cell, uint32_t* sum_in_cell, uint32_t np,
                      uint32_t Nx, uint32_t Nr, const double* xmin, const double* dx_inv,
                      const double* rmin, const double* dr_inv, void* stream);

/* Fusion of the two launches above (push_coords immediately followed by
 * sort_parts in pic_loop.py:70-76): one pass over the particles. */
int chb_push_index(double* x, double* y, double* z, const double* px, const double* py,
                   const double* pz, const double* g_inv, const double* dt_dev,
                   uint32_t* indx_in_cell, uint32_t* sum_in_cell, uint32_t np,
                   uint32_t Nx, uint32_t Nr, const double* xmin, const double* dx_inv,
                   const double* rmin, const double* dr_inv, void* stream);

/* cell_offset = [0, inclusive_scan(sum_in_cell)] (nbins+1 entries); cursor (nbins,
 * may be NULL) receives the exclusive offsets; *np_stay_dev = cell_offset[nbins-1].
 * Replaces pyopencl.array.cumsum + the 0-prepend of
 * methods/particles_methods_cl.py:303-311 and the read of cell_offset[-2] (:250). */
size_t chb_cell_offsets_workspace_bytes(uint32_t nbins);
int chb_cell_offsets(const uint32_t* sum_in_cell, uint32_t nbins, uint32_t* cell_offset,
                     uint32_t* cursor, uint32_t* np_stay_dev, void* workspace,
                     size_t workspace_bytes, void* stream);

/* Counting-sort scatter in the STABLE order (what serial execution of
 * kernels/particles_generic.cl:186-201 yields): sort_indx[cell_offset[c]+k] =
 * k-th particle, in ascending storage index, of cell c.  `cursor` must hold the
 * exclusive offsets on entry (chb_cell_offsets) and is consumed. */
size_t chb_sort_workspace_bytes(uint32_t np, uint32_t nbins);
int chb_sort_scatter_stable(const uint32_t* indx_in_cell, const uint32_t* cell_offset,
                            uint32_t* cursor, uint32_t* sort_indx, uint32_t np,
                            uint32_t nbins, void* workspace, size_t workspace_bytes,
                            void* stream);

/* dst[a][i] = src[a][sort_indx[i]], i < np_stay, for nattr <= CHB_MAX_ATTRS
 * attributes in one launch; optionally writes sort_indx_out = arange(np_stay).
 * src/dst are HOST arrays of device pointers.  Replaces data_align_dbl,
 * kernels/particles_generic.cl:156-169 (particles_methods_cl.py:263-286). */
int chb_align(const double* const* src_host, double* const* dst_host, int nattr,
              const uint32_t* sort_indx, uint32_t np_stay, uint32_t* sort_indx_out,
              void* stream);

/* ---------------------------------------------------------------- grid */

/* Linear (bilinear r-x) deposition of w*charge onto rho_m0 (real) and rho_m{1..M}
 * (complex), accumulating into the arrays (caller zero-fills).  Same sums as the
 * four colour passes of depose_scalar, kernels/grid_deposit_m0.cl:20-129 and
 * kernels/grid_deposit_m1.cl:20-152 (grid_methods_cl.py:44-69), up to summation
 * order.  rho_host: HOST array of M+1 device pointers. */
int chb_depose_scalar(int M, const uint32_t* sort_indx, const double* x, const double* y,
                      const double* z, const double* w, const uint32_t* cell_offset,
                      int charge, uint32_t Nx, uint32_t Nr, const double* xmin,
                      const double* dx_inv, const double* rmin, const double* dr_inv,
                      double* const* rho_host, void* stream);

/* Same for the current: (w*g_inv*charge)*p_k onto J{x,y,z}_m*.  j_host: HOST array
 * of 3*(M+1) device pointers ordered [m][component] as in
 * grid_methods_cl.py:79-82.  Replaces depose_vector,
 * kernels/grid_deposit_m0.cl:148-277, kernels/grid_deposit_m1.cl:171-327. */
int chb_depose_vector(int M, const uint32_t* sort_indx, const double* x, const double* y,
                      const double* z, const double* px, const double* py,
                      const double* pz, const double* g_inv, const double* w,
                      const uint32_t* cell_offset, int charge, uint32_t Nx, uint32_t Nr,
                      const double* xmin, const double* dx_inv, const double* rmin,
                      const double* dr_inv, double* const* j_host, void* stream);

/* Fusion of push_coords(dt) + sort_parts + depose_currents (pic_loop.py:70-81): every
 * particle is advanced by dt*g_inv*p (x, y, z updated in place, same arithmetic as
 * chb_push_xyz) and its current deposited at the NEW position.  sort_indx /
 * cell_offset are those of the PREVIOUS sort and only serve as traversal order:
 * particles still in the cell that order assumes take the cell-ordered fast path, the
 * others (and the previous trash bin) are deposited one by one.  Same sums as
 * chb_push_xyz -> sort -> chb_depose_vector up to summation order; np = all particles;
 * workspace (8-byte aligned): a u32 counter (16 bytes reserved) followed by the queue of
 * the cell changers, 64 bytes each; chb_push_depose_workspace_bytes(np) is the size that
 * can never overflow (one record per particle).  With a smaller workspace the caller must
 * read the counter back: a value above (workspace_bytes - 16) / 64 means records were
 * dropped and the deposited current is incomplete. */
size_t chb_push_depose_workspace_bytes(uint32_t np);
int chb_push_depose_vector(int M, const uint32_t* sort_indx, double* x, double* y, double* z,
                           const double* px, const double* py, const double* pz,
                           const double* g_inv, const double* w,
                           const uint32_t* cell_offset, const double* dt_dev, uint32_t np,
                           int charge, uint32_t Nx, uint32_t Nr, const double* xmin,
                           const double* dx_inv, const double* rmin, const double* dr_inv,
                           double* const* j_host, void* workspace, size_t workspace_bytes,
                           void* stream);

/* The whole particle side of pic_loop.py:70-76 in ONE pass: push_coords('half') +
 * sort_parts + depose_currents as chb_push_depose_vector, then -- the momenta being
 * unchanged between the two half pushes of a step -- the second push_coords('half')
 * (x2 = (x0 + d) + d with the same rounded d = p*(dt*g_inv) both chb_push_xyz calls
 * would use, so x, y, z are bit-identical) and the index_and_sum_in_cell of the second
 * sort_parts: indx_in_cell[np] and sum_in_cell (zeroed by the caller, (Nx-1)*(Nr-1)+1
 * bins) are those chb_push_index would produce.  The caller continues with
 * chb_cell_offsets and chb_sort_scatter_stable. */
int chb_push_depose_push_index(int M, const uint32_t* sort_indx, double* x, double* y,
                               double* z, const double* px, const double* py,
                               const double* pz, const double* g_inv, const double* w,
                               const uint32_t* cell_offset, const double* dt_dev, uint32_t np,
                               int charge, uint32_t Nx, uint32_t Nr, const double* xmin,
                               const double* dx_inv, const double* rmin, const double* dr_inv,
                               double* const* j_host, uint32_t* indx_in_cell,
                               uint32_t* sum_in_cell, void* workspace, size_t workspace_bytes,
                               void* stream);

/* row1 -= row0, then arr[ir,:] *= dV_inv[ir], for nfld arrays in one launch
 * (is_complex_host[k] != 0 for complex arrays).  Replaces treat_axis_{d,c} and
 * divide_by_dv_{d,c}, kernels/grid_generic.cl:4-59 (grid_methods_cl.py:98-151). */
int chb_postproc_depose(double* const* fld_host, const int* is_complex_host, int nfld,
                        uint32_t Nx, uint32_t Nr, const double* dV_inv, void* stream);

/* Ghost row: row0 = +row1 (real, m=0) / -row1 (complex, m>=1).  Replaces
 * warp_axis_m0_d / warp_axis_m1plus_c, kernels/grid_generic.cl:63-86
 * (grid_methods_cl.py:153-166). */
int chb_warp_axis(double* const* fld_host, const int* is_complex_host, int nfld,
                  uint32_t Nx, void* stream);

/* Field gather (bilinear, modes 0..M, m>=1 weighted 2*Re(F e^{-i m theta})) and
 * relativistic Boris push of px, py, pz, g_inv.  eb_host: HOST array of 6*(M+1)
 * device pointers ordered [m][E,B][x,y,z] (grid_methods_cl.py:176-180).
 * factor_push_dev: DataDev['FactorPush']; np_stay_dev: device scalar written by
 * chb_cell_offsets (the reference passes Args['Np_stay'] by value).  Particles are
 * visited cell tile by cell tile through cell_offset / sort_indx.  Replaces
 * gather_and_push, kernels/grid_deposit_m0.cl:280-427,
 * kernels/grid_deposit_m1.cl:330-511 (grid_methods_cl.py:168-192). */
int chb_gather_push(int M, const double* x, const double* y, const double* z, double* px,
                    double* py, double* pz, double* g_inv, const uint32_t* sort_indx,
                    const uint32_t* cell_offset, const double* factor_push_dev, uint32_t np,
                    const uint32_t* np_stay_dev, uint32_t Nx, uint32_t Nr,
                    const double* xmin, const double* dx_inv, const double* rmin,
                    const double* dr_inv, const double* const* eb_host, void* stream);

/* ---------------------------------------------------------------- spectral: element-wise
 * n = number of complex (or real) elements; complex arrays are (re, im) doubles. */

/* out_d[i] = Re(in_c[i]).  Replaces cast_array_d2c, kernels/generic.cl:100-112
 * (generic_methods_cl.py:89-94). */
int chb_cast_c2d(const double* in_c, double* out_d, size_t n, void* stream);
/* out_c[i] = in_d[i] + 0i.  Replaces `.astype(np.complex128)`,
 * methods/transformer_methods_cl.py:301-302. */
int chb_cast_d2c(const double* in_d, double* out_c, size_t n, void* stream);
/* base += add.  kernels/generic.cl:18-30 (generic_methods_cl.py:107-112). */
int chb_append_c2c(double* base, const double* add, size_t n, void* stream);
/* z += a*x.  kernels/generic.cl:33-45 (generic_methods_cl.py:128-134). */
int chb_zpaxz_c2c(double a_re, double a_im, const double* x, double* z, size_t n, void* stream);
/* z *= x, x real.  kernels/generic.cl:47-58 (generic_methods_cl.py:114-118). */
int chb_mult_elementwise_d2c(const double* x, double* z, size_t n, void* stream);
/* z = a*x + b*y.  kernels/generic.cl:60-77 (generic_methods_cl.py:120-126). */
int chb_axpbyz_c2c(double a_re, double a_im, const double* x, double b_re, double b_im,
                   const double* y, double* z, size_t n, void* stream);
/* z[ir,ix] = b[ix]*(a*x[ir,ix]).  kernels/generic.cl:80-98
 * (generic_methods_cl.py:136-142). */
int chb_ab_dot_x(double a_re, double a_im, const double* b, const double* x, double* z,
                 size_t n, uint32_t Nx, void* stream);
/* dst(ix) = -conj(src((Nx-ix) mod Nx)) per row: the m=-1 spectrum.
 * kernels/transformer_generic.cl:3-24 (transformer_methods_cl.py:265-288). */
int chb_get_m1(double* dst, const double* src, size_t n, uint32_t Nx, void* stream);
/* out (op)= alpha*b(ix) + beta*conj(b((Nx-ix) mod Nx)) per row (op: '=' or '+=').
 * Expresses the m=0 "m-1" terms of field_grad / field_rot
 * (transformer_methods_cl.py:104-121, :208-235) through the "m+1" product, using
 * F_{-1} = -conj(mirror F_1) (get_m1) and dDHT_minus_m0 == dDHT_plus_m0. */
int chb_mirror_axpy(double* out, const double* b, double a_re, double a_im, double b_re,
                    double b_im, int accumulate, size_t n, uint32_t Nx, void* stream);
/* phs[ix] = exp(+i*x0*kx[ix]) (dir=1) or exp(-i*x0*kx[ix]) (dir=0).
 * kernels/transformer_generic.cl:28-55 (transformer_methods_cl.py:38-44). */
int chb_get_phase(double* phs, const double* kx, double x0, int dir, uint32_t Nx, void* stream);
/* arr[ir,ix] *= phs[ix].  kernels/transformer_generic.cl:58-80. */
int chb_multiply_by_phase(double* arr, const double* phs, size_t n, uint32_t Nx, void* stream);
/* Edge damping of nfld arrays (Nr x Nx) in one launch: ix<Nf: *= prof[ix];
 * ix>Nx-Nf: *= prof[Nx-ix].  kernels/solver_ms_pic.cl:5-55
 * (solver_methods_cl.py:64-83). */
int chb_profile_edges(double* const* fld_host, const int* is_complex_host, int nfld,
                      const double* prof, uint32_t Nr, uint32_t Nx, uint32_t Nf, void* stream);
/* PSATD advance of (E, G) for one azimuthal mode; e/g/j/n0/n1_host: HOST arrays of 3
 * device pointers (x, y, z).  kernels/solver_ms_pic.cl:57-143
 * (solver_methods_cl.py:43-62). */
int chb_psatd_advance(size_t n, const double* dt_inv_dev, const double* c1, const double* c2,
                      const double* c3, double* const* e_host, double* const* g_host,
                      const double* const* j_host, const double* const* n0_host,
                      const double* const* n1_host, void* stream);

/* ---------------------------------------------------------------- spectral: DHT and FFT */

/* Host-only query (no GPU work): width in real columns (64..128, multiple of 16) of the
 * 128-row output tiles the wide contraction kernel uses for an M x N_doubles result
 * (N_doubles = 2*Nx for complex data) and `nbatch` right-hand sides -- the width whose tile
 * count fills whole waves of the 148 SMs best.  No reference counterpart. */
int chb_dht_tile_columns(uint32_t M, uint32_t N_doubles, int nbatch);

/* Measurement aid (bench.py): launches a register-only FP64 DMMA loop on every SM;
 * flops_out = floating-point operations it executes (time it with events on `stream`).
 * scratch: >= 148*512 doubles of device memory.  No reference counterpart. */
int chb_dmma_peak(double* scratch, size_t scratch_doubles, int iters, double* flops_out,
                  void* stream);

/* C (op)= alpha * A . B with A real (M x K, leading dimension lda), B and C real
 * (is_complex=0) or complex (is_complex=1) with N columns and leading dimensions
 * ldb/ldc given in ELEMENTS of their type; op is '=' or '+=' (accumulate).  A complex
 * alpha requires is_complex.  FP64 DMMA tensor-pipe kernel.  Replaces Reikna
 * MatrixMul `_ddot`/`_cdot`, methods/transformer_methods_cl.py:458-480, and, through
 * alpha/accumulate, the zpaxz/append_c2c passes that follow it in field_grad /
 * field_rot (:111-133, :221-263). */
int chb_dht(const double* A, uint32_t lda, const double* B, uint32_t ldb, double* C,
            uint32_t ldc, uint32_t M, uint32_t K, uint32_t N, int is_complex,
            double alpha_re, double alpha_im, int accumulate, void* stream);

/* C_k = A . B_k for nbatch <= CHB_MAX_FIELDS right-hand sides sharing A, dimensions and
 * leading dimensions, in one launch (B_host / C_host: HOST arrays of device pointers):
 * all components of one fb_transform call (transformer.py:17-26). */
int chb_dht_batched(const double* A, uint32_t lda, const double* const* B_host,
                    double* const* C_host, int nbatch, uint32_t ldb, uint32_t ldc, uint32_t M,
                    uint32_t K, uint32_t N, int is_complex, void* stream);

/* Same product written to two outputs, C1 (op1)= a1*A.B and C2 (op2)= a2*A.B (same
 * leading dimension): the "b = dDHT.x; y += a1*b; z += a2*b" pattern of field_grad /
 * field_rot (transformer_methods_cl.py:111-133, :228-263) in one pass. */
int chb_dht2(const double* A, uint32_t lda, const double* B, uint32_t ldb, double* C1,
             double a1_re, double a1_im, int accumulate1, double* C2, double a2_re,
             double a2_im, int accumulate2, uint32_t ldc, uint32_t M, uint32_t K, uint32_t N,
             int is_complex, void* stream);

/* chb_dht2 for a complex B that is the x-spectrum of a REAL field (an m = 0 spectral
 * array: B(:, Nx-k) = conj(B(:, k))): only the columns k <= Nx/2 are contracted, every
 * result is also written, conjugated (before alpha), to column Nx-k.  Same results as
 * chb_dht2 up to the rounding-level asymmetry of B; half the arithmetic.  The caller
 * vouches for the symmetry (PIC_loop does: its m = 0 spectra come from real grid fields). */
int chb_dht2_hermitian(const double* A, uint32_t lda, const double* B, uint32_t ldb,
                       double* C1, double a1_re, double a1_im, int accumulate1, double* C2,
                       double a2_re, double a2_im, int accumulate2, uint32_t ldc, uint32_t M,
                       uint32_t K, uint32_t N, void* stream);

/* Batched FFT along x of `rows` rows of length Nx (numpy conventions; inverse is
 * normalised).  Strides are in elements of the row's type.  in_real: input rows are
 * real; out_real: keep only the real part.  phase (Nx complex, may be NULL) multiplies
 * the input (phase_on_input=1, backward path) or the output (forward path).
 * twiddles: L complex roots exp(-2 pi i j/L).  L == Nx for power-of-two lengths;
 * otherwise Bluestein with L = pow2 >= 2Nx-1, chirp[n] = exp(-i pi n^2/Nx) (Nx
 * entries) and bfft = FFT_L(b)/L (L entries) supplied by the caller (host-side plan,
 * see chimeracl_b200/methods/transformer_methods_cl.py).  8 <= L <= 8192, plus
 * L == Nx == 16384, for which twiddles holds 8192 roots exp(-2 pi i j/8192) followed
 * by 8192 values exp(-2 pi i j/16384) (one radix-2 stage is split over two CTAs).
 * Replaces Reikna FFT `_fft`, methods/transformer_methods_cl.py:482-509, plus the
 * cast / phase / slice-copy passes around it (:295-311, :338-358). */
int chb_fft_max_pow2(void);
/* damp_fields (solver.py:32-35) on the spectral arrays, in place and on chip: half
 * backward transform (x phase `phase_bwd`, inverse FFT, real part where real_x_host[k]),
 * profile_edges (kernels/solver_ms_pic.cl:5-55; columns ix < Nf and ix > Nx-Nf times
 * prof[ix] / prof[Nx-ix]), half forward transform (FFT, x phase `phase_fwd`).  Replaces
 * the sequence transform_field(dir=1,'half') -> profile_edges -> transform_field(dir=0,
 * 'half') (transformer_methods_cl.py:385-455, solver_methods_cl.py:64-83) with the same
 * per-element arithmetic; the x-space intermediate is not written to the grid arrays.
 * Nx: power of two in [256, 8192]; stride in complex elements. */
int chb_fft_damp_x_batched(double* const* spec_host, const int* real_x_host, int nbatch,
                           uint32_t rows, uint32_t Nx, size_t stride, const double* phase_bwd,
                           const double* phase_fwd, const double* prof, uint32_t Nf,
                           const double* twiddles, void* stream);

/* The same transform applied to nbatch <= CHB_MAX_FIELDS arrays in one launch;
 * out_filter (rows x Nx real, may be NULL) multiplies the output: the spectral
 * smoothing of fields_smooth, transformer_methods_cl.py:79-85, folded into the
 * forward transform that precedes it in pic_loop.py:99-103. */
int chb_fft_x_batched(const double* const* in_host, double* const* out_host, int nbatch,
                      uint32_t rows, uint32_t Nx, size_t in_stride, size_t out_stride,
                      int inverse, int in_real, int out_real, const double* phase,
                      int phase_on_input, const double* twiddles, uint32_t L,
                      const double* chirp, const double* bfft, const double* out_filter,
                      void* stream);
int chb_fft_x(const double* in, double* out, uint32_t rows, uint32_t Nx, size_t in_stride,
              size_t out_stride, int inverse, int in_real, int out_real,
              const double* phase, int phase_on_input, const double* twiddles,
              uint32_t L, const double* chirp, const double* bfft, void* stream);

/* ---------------------------------------------------------------- multi-GPU: peer memory */

/* In-place sum over ranks of one flat FP64 buffer of n doubles that every rank holds as
 * SYMMETRIC memory: peer_ptrs_host[r] = device address of rank r's buffer as mapped into
 * this process (world <= 16 entries, 16-byte aligned, n even), multicast_ptr = the
 * NVSwitch multicast address of the same buffers or 0.  Rank `rank` sums its 1/world
 * block over all ranks and stores the total into every rank's buffer (P2P loads / stores,
 * or multimem.ld_reduce / multimem.st when multicast_ptr != 0).  The caller must place a
 * cross-rank barrier on `stream` before (all partial sums written) and after (all totals
 * stored) the call.  Used for the E / B partial backward transforms of the kr-row sharded
 * field solve instead of an NCCL all-reduce; no reference counterpart. */
int chb_peer_allreduce_f64(const uint64_t* peer_ptrs_host, int world, int rank,
                           uint64_t multicast_ptr, size_t n, void* stream);

/* All-gather of blocks of one symmetric FP64 buffer: this rank's doubles [begin, begin+count)
 * (begin, count even) are copied to the same place in every other rank's buffer (16-byte P2P
 * stores, or one multimem.st per element when multicast_ptr != 0).  The caller places a
 * cross-rank barrier on `stream` after the call of all ranks, before the gathered data is
 * read.  Used for the rho / G spectra of the kr-row sharded field solve (the owned kr rows
 * are one contiguous block) instead of an NCCL all-gather; no reference counterpart. */
int chb_peer_allgather_f64(const uint64_t* peer_ptrs_host, int world, int rank,
                           uint64_t multicast_ptr, size_t begin, size_t count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CHIMERA_B200_H */
