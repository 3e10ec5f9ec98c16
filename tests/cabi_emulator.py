"""TEST INFRASTRUCTURE ONLY -- a NumPy stand-in for the entry points of
libchimera_b200.so (spectral side: NumPy; particle side: the oracle's restatement of the
reference kernels), so that the HOST orchestration (the mixins, PIC_loop's phase sequence
including the one-pass particle side, the kr-row sharding and the collectives) can be
exercised without a GPU: single process against the golden vectors, and several ranks
over gloo.  Every function takes
exactly the arguments of its C-ABI namesake in include/chimera_b200.h (raw addresses,
sizes, scalars, stream) and works on host memory.  Nothing under chimeracl_b200/ imports
this file; the product has no CPU path (Communicator raises without CUDA)."""
import ctypes

import numpy as np
import torch

from chimeracl_b200 import _lib as real_lib

C16, F8 = np.complex128, np.float64


def _vec(ptr, n, dtype):
    n = int(n)
    if n == 0:
        return np.empty(0, dtype)
    nbytes = n * np.dtype(dtype).itemsize
    buf = (ctypes.c_char * nbytes).from_address(int(ptr))
    return np.frombuffer(buf, dtype=dtype)


def _mat(ptr, rows, cols, ld, dtype):
    """rows x cols view with leading dimension ld (elements of dtype)."""
    rows, cols, ld = int(rows), int(cols), int(ld)
    flat = _vec(ptr, (rows - 1) * ld + cols if rows else 0, dtype)
    item = np.dtype(dtype).itemsize
    return np.lib.stride_tricks.as_strided(flat, (rows, cols), (ld * item, item))


def _store(c, val, alpha, acc):
    val = alpha * val
    if acc:
        c += val
    else:
        c[...] = val


class EmulatedLib:
    """Attribute-compatible with chimeracl_b200._lib's proxy for the calls the field solve
    makes; host-only queries fall through to the real library."""

    def __init__(self):
        self._real = real_lib.load()

    def __getattr__(self, name):
        if name in ("chb_fft_max_pow2", "chb_version", "chb_error_string",
                    "chb_dht_tile_columns", "chb_cell_offsets_workspace_bytes",
                    "chb_sort_workspace_bytes", "chb_push_depose_workspace_bytes"):
            return getattr(self._real, name)
        raise AttributeError("cabi_emulator: %s is not emulated" % name)

    # ---- particles and grid (oracle restatement of the reference kernels)
    @staticmethod
    def _geom(Nx, Nr, xmin, dx_inv, rmin, dr_inv):
        return {"Nx": int(Nx), "Nr": int(Nr), "Xmin": float(_vec(xmin, 1, F8)[0]),
                "dx_inv": float(_vec(dx_inv, 1, F8)[0]), "Rmin": float(_vec(rmin, 1, F8)[0]),
                "dr_inv": float(_vec(dr_inv, 1, F8)[0]),
                "Nxm1Nrm1": (int(Nx) - 1) * (int(Nr) - 1)}

    @staticmethod
    def _kern(M):
        from oracle.np_kernels import NumpyKernels
        return NumpyKernels(int(M))

    def chb_push_xyz(self, x, y, z, px, py, pz, g_inv, dt, n, stream):
        v = [_vec(p, n, F8) for p in (x, y, z, px, py, pz, g_inv)]
        self._kern(0).push_xyz(*v, float(_vec(dt, 1, F8)[0]))
        return 0

    def chb_index_and_sum(self, x, y, z, indx, summ, n, Nx, Nr, xmin, dx_inv, rmin, dr_inv,
                          stream):
        g = self._geom(Nx, Nr, xmin, dx_inv, rmin, dr_inv)
        i, s = self._kern(0).index_and_sum(_vec(x, n, F8), _vec(y, n, F8), _vec(z, n, F8), g)
        _vec(indx, n, np.uint32)[...] = i
        _vec(summ, g["Nxm1Nrm1"] + 1, np.uint32)[...] += s
        return 0

    def chb_push_index(self, x, y, z, px, py, pz, g_inv, dt, indx, summ, n, Nx, Nr, *geom_st):
        self.chb_push_xyz(x, y, z, px, py, pz, g_inv, dt, n, None)
        return self.chb_index_and_sum(x, y, z, indx, summ, n, Nx, Nr, *geom_st)

    def chb_cell_offsets(self, summ, nbins, cell_offset, cursor, np_stay, ws, ws_bytes, stream):
        s = _vec(summ, nbins, np.uint32)
        off = _vec(cell_offset, nbins + 1, np.uint32)
        off[0] = 0
        off[1:] = np.cumsum(s, dtype=np.uint32)
        if cursor:
            _vec(cursor, nbins, np.uint32)[...] = off[:-1]
        if np_stay:
            _vec(np_stay, 1, np.uint32)[0] = off[nbins - 1]
        return 0

    def chb_sort_scatter_stable(self, indx, cell_offset, cursor, sort_indx, n, nbins, ws, wsb,
                                stream):
        out, _ = self._kern(0).sort_scatter(_vec(cell_offset, nbins + 1, np.uint32),
                                            _vec(indx, n, np.uint32))
        _vec(sort_indx, n, np.uint32)[...] = out
        return 0

    def chb_align(self, src, dst, nattr, sort_indx, np_stay, sort_out, stream):
        s = _vec(sort_indx, np_stay, np.uint32).astype(np.int64)
        hi = int(s.max()) + 1 if np_stay else 0
        for k in range(nattr):
            _vec(dst[k], np_stay, F8)[...] = _vec(src[k], hi, F8)[s]
        if sort_out:
            _vec(sort_out, np_stay, np.uint32)[...] = np.arange(np_stay, dtype=np.uint32)
        return 0

    def _mode_arrays(self, ptrs, count, M, per_mode, Nx, Nr):
        out = []
        for k in range(count):
            m = k // per_mode
            out.append(_vec(ptrs[k], Nx * Nr, C16 if m else F8).reshape(Nr, Nx))
        return out

    def chb_depose_scalar(self, M, sort_indx, x, y, z, w, cell_offset, charge, Nx, Nr, xmin,
                          dx_inv, rmin, dr_inv, rho, stream, n=None):
        g = self._geom(Nx, Nr, xmin, dx_inv, rmin, dr_inv)
        off = _vec(cell_offset, g["Nxm1Nrm1"] + 2, np.uint32)
        n = int(off[-1])
        flds = self._mode_arrays(rho, M + 1, M, 1, Nx, Nr)
        self._kern(M).depose_scalar(_vec(sort_indx, n, np.uint32), _vec(x, n, F8), _vec(y, n, F8),
                                    _vec(z, n, F8), _vec(w, n, F8), off, charge, g, flds)
        return 0

    def chb_depose_vector(self, M, sort_indx, x, y, z, px, py, pz, g_inv, w, cell_offset, charge,
                          Nx, Nr, xmin, dx_inv, rmin, dr_inv, j, stream):
        g = self._geom(Nx, Nr, xmin, dx_inv, rmin, dr_inv)
        off = _vec(cell_offset, g["Nxm1Nrm1"] + 2, np.uint32)
        n = int(off[-1])
        flds = self._mode_arrays(j, 3 * (M + 1), M, 3, Nx, Nr)
        v = [_vec(p, n, F8) for p in (x, y, z, px, py, pz, g_inv, w)]
        self._kern(M).depose_vector(_vec(sort_indx, n, np.uint32), *v, off, charge, g, flds)
        return 0

    def _push_depose(self, M, x, y, z, px, py, pz, g_inv, w, dt, n, charge, Nx, Nr, geom, j):
        """push by dt, then the deposit a fresh sort at the new positions would give."""
        K = self._kern(M)
        g = self._geom(Nx, Nr, *geom)
        v = [_vec(p, n, F8) for p in (x, y, z, px, py, pz, g_inv, w)]
        K.push_xyz(*v[:7], float(_vec(dt, 1, F8)[0]))
        indx, summ = K.index_and_sum(v[0], v[1], v[2], g)
        off = np.concatenate(([0], np.cumsum(summ, dtype=np.uint32))).astype(np.uint32)
        srt, _ = K.sort_scatter(off, indx)
        K.depose_vector(srt, *v, off, charge, g, self._mode_arrays(j, 3 * (M + 1), M, 3, Nx, Nr))
        return g, v

    def chb_push_depose_vector(self, M, sort_indx, x, y, z, px, py, pz, g_inv, w, cell_offset,
                               dt, n, charge, Nx, Nr, xmin, dx_inv, rmin, dr_inv, j, ws, wsb,
                               stream):
        self._push_depose(M, x, y, z, px, py, pz, g_inv, w, dt, n, charge, Nx, Nr,
                          (xmin, dx_inv, rmin, dr_inv), j)
        return 0

    def chb_push_depose_push_index(self, M, sort_indx, x, y, z, px, py, pz, g_inv, w,
                                   cell_offset, dt, n, charge, Nx, Nr, xmin, dx_inv, rmin,
                                   dr_inv, j, indx, summ, ws, wsb, stream):
        g, v = self._push_depose(M, x, y, z, px, py, pz, g_inv, w, dt, n, charge, Nx, Nr,
                                 (xmin, dx_inv, rmin, dr_inv), j)
        K = self._kern(M)
        K.push_xyz(*v[:7], float(_vec(dt, 1, F8)[0]))
        i, s = K.index_and_sum(v[0], v[1], v[2], g)
        _vec(indx, n, np.uint32)[...] = i
        _vec(summ, g["Nxm1Nrm1"] + 1, np.uint32)[...] += s
        return 0

    def chb_postproc_depose(self, flds, is_complex, nfld, Nx, Nr, dV_inv, stream):
        dv = _vec(dV_inv, Nr, F8)
        for k in range(nfld):
            a = _vec(flds[k], Nx * Nr, C16 if is_complex[k] else F8).reshape(Nr, Nx)
            a[1] -= a[0]
            a *= dv[:, None]
        return 0

    def chb_warp_axis(self, flds, is_complex, nfld, Nx, stream):
        for k in range(nfld):
            a = _vec(flds[k], 2 * Nx, C16 if is_complex[k] else F8).reshape(2, Nx)
            a[0] = -a[1] if is_complex[k] else a[1]
        return 0

    def chb_gather_push(self, M, x, y, z, px, py, pz, g_inv, sort_indx, cell_offset, factor, n,
                        np_stay, Nx, Nr, xmin, dx_inv, rmin, dr_inv, eb, stream):
        g = self._geom(Nx, Nr, xmin, dx_inv, rmin, dr_inv)
        flds = self._mode_arrays(eb, 6 * (M + 1), M, 6, Nx, Nr)
        v = [_vec(p, n, F8) for p in (x, y, z, px, py, pz, g_inv)]
        self._kern(M).gather_and_push(*v, _vec(sort_indx, n, np.uint32),
                                      _vec(cell_offset, g["Nxm1Nrm1"] + 2, np.uint32),
                                      float(_vec(factor, 1, F8)[0]), n,
                                      int(_vec(np_stay, 1, np.uint32)[0]), g, flds)
        return 0

    # ---- element-wise
    def chb_mult_elementwise_d2c(self, x, z, n, stream):
        _vec(z, n, C16)[...] *= _vec(x, n, F8)
        return 0

    def chb_axpbyz_c2c(self, are, aim, x, bre, bim, y, z, n, stream):
        _vec(z, n, C16)[...] = complex(are, aim) * _vec(x, n, C16) + \
            complex(bre, bim) * _vec(y, n, C16)
        return 0

    def chb_ab_dot_x(self, are, aim, b, x, z, n, Nx, stream):
        rows = int(n) // int(Nx)
        _vec(z, n, C16).reshape(rows, Nx)[...] = _vec(b, Nx, F8)[None, :] * (
            complex(are, aim) * _vec(x, n, C16).reshape(rows, Nx))
        return 0

    def chb_get_m1(self, dst, src, n, Nx, stream):
        rows = int(n) // int(Nx)
        idx = (Nx - np.arange(Nx)) % Nx
        _vec(dst, n, C16).reshape(rows, Nx)[...] = -np.conj(_vec(src, n, C16).reshape(rows, Nx)[:, idx])
        return 0

    def chb_mirror_axpy(self, out, b, are, aim, bre, bim, acc, n, Nx, stream):
        rows = int(n) // int(Nx)
        idx = (Nx - np.arange(Nx)) % Nx
        B = _vec(b, n, C16).reshape(rows, Nx)
        val = complex(are, aim) * B + complex(bre, bim) * np.conj(B[:, idx])
        _store(_vec(out, n, C16).reshape(rows, Nx), val, 1.0, acc)
        return 0

    def chb_get_phase(self, phs, kx, x0, direction, Nx, stream):
        k = _vec(kx, Nx, F8)
        sgn = 1.0 if direction == 1 else -1.0
        _vec(phs, Nx, C16)[...] = np.cos(x0 * k) + 1j * (sgn * np.sin(x0 * k))
        return 0

    def chb_profile_edges(self, flds, is_complex, nfld, prof, Nr, Nx, Nf, stream):
        f = _vec(prof, Nf, F8)
        ix = np.arange(Nx)
        fac = np.ones(Nx)
        lo = ix < Nf
        fac[lo] *= f[ix[lo]]
        hi = ix > Nx - Nf
        fac[hi] *= f[Nx - ix[hi]]
        for k in range(nfld):
            dt = C16 if is_complex[k] else F8
            _vec(flds[k], Nr * Nx, dt).reshape(Nr, Nx)[...] *= fac[None, :]
        return 0

    def chb_psatd_advance(self, n, dt_inv, c1, c2, c3, e, g, j, n0, n1, stream):
        from oracle.np_kernels import NumpyKernels
        f = [_vec(grp[k], n, C16) for grp in (e, g, j, n0, n1) for k in range(3)]
        NumpyKernels(0).advance_e_g(n, float(_vec(dt_inv, 1, F8)[0]), _vec(c1, n, F8),
                                    _vec(c2, n, F8), _vec(c3, n, F8), f)
        return 0

    # ---- contractions
    def _gemm(self, A, lda, B, ldb, M, K, N, cplx):
        a = _mat(A, M, K, lda, F8)
        b = _mat(B, K, N, ldb, C16 if cplx else F8)
        return a @ b

    def chb_dht(self, A, lda, B, ldb, C, ldc, M, K, N, cplx, are, aim, acc, stream):
        if M == 0 or N == 0:
            return 0
        dt = C16 if cplx else F8
        alpha = complex(are, aim) if cplx else are
        _store(_mat(C, M, N, ldc, dt), self._gemm(A, lda, B, ldb, M, K, N, cplx), alpha, acc)
        return 0

    def chb_dht_batched(self, A, lda, Bs, Cs, nbatch, ldb, ldc, M, K, N, cplx, stream):
        for k in range(nbatch):
            self.chb_dht(A, lda, Bs[k], ldb, Cs[k], ldc, M, K, N, cplx, 1.0, 0.0, 0, stream)
        return 0

    def chb_dht2(self, A, lda, B, ldb, C1, a1re, a1im, acc1, C2, a2re, a2im, acc2, ldc, M, K, N,
                 cplx, stream):
        if M == 0 or N == 0:
            return 0
        dt = C16 if cplx else F8
        p = self._gemm(A, lda, B, ldb, M, K, N, cplx)
        _store(_mat(C1, M, N, ldc, dt), p, complex(a1re, a1im) if cplx else a1re, acc1)
        _store(_mat(C2, M, N, ldc, dt), p, complex(a2re, a2im) if cplx else a2re, acc2)
        return 0

    def chb_dht2_hermitian(self, A, lda, B, ldb, C1, a1re, a1im, acc1, C2, a2re, a2im, acc2,
                           ldc, M, K, N, stream):
        if M == 0 or N == 0:
            return 0
        half = N // 2 + 1                       # columns k = 0 .. Nx/2 are contracted
        ph = self._gemm(A, lda, B, ldb, M, K, half, 1)
        p = np.empty((M, N), C16)
        p[:, :half] = ph
        k = np.arange(1, N - half + 1)
        p[:, N - k] = np.conj(ph[:, k])
        _store(_mat(C1, M, N, ldc, C16), p, complex(a1re, a1im), acc1)
        _store(_mat(C2, M, N, ldc, C16), p, complex(a2re, a2im), acc2)
        return 0

    # ---- FFT along x
    def chb_fft_x_batched(self, ins, outs, nbatch, rows, Nx, in_stride, out_stride, inverse,
                          in_real, out_real, phase, phase_on_input, tw, L, chirp, bfft,
                          out_filter, stream):
        phs = _vec(phase, Nx, C16) if phase else None
        flt = _vec(out_filter, rows * Nx, F8).reshape(rows, Nx) if out_filter else None
        for k in range(nbatch):
            src = _mat(ins[k], rows, Nx, in_stride, F8 if in_real else C16).astype(C16)
            if phs is not None and phase_on_input:
                src = src * phs[None, :]
            res = np.fft.ifft(src, axis=1) if inverse else np.fft.fft(src, axis=1)
            if phs is not None and not phase_on_input:
                res = res * phs[None, :]
            if flt is not None:
                res = res * flt
            if out_real:
                _mat(outs[k], rows, Nx, out_stride, F8)[...] = res.real
            else:
                _mat(outs[k], rows, Nx, out_stride, C16)[...] = res
        return 0

    def chb_fft_damp_x_batched(self, specs, real_x, nbatch, rows, Nx, stride, phase_bwd,
                               phase_fwd, prof, Nf, tw, stream):
        pb, pf = _vec(phase_bwd, Nx, C16), _vec(phase_fwd, Nx, C16)
        f = _vec(prof, Nf, F8)
        ix = np.arange(Nx)
        fac = np.ones(Nx)
        lo = ix < Nf
        fac[lo] *= f[ix[lo]]
        hi = ix > Nx - Nf
        fac[hi] *= f[Nx - ix[hi]]
        for k in range(nbatch):
            s = _mat(specs[k], rows, Nx, stride, C16)
            x = np.fft.ifft(s * pb[None, :], axis=1)
            if real_x[k]:
                x = x.real.astype(C16)
            s[...] = np.fft.fft(x * fac[None, :], axis=1) * pf[None, :]
        return 0


class EmulatedComm:
    """Stands in for Communicator in the CPU tests (host memory, no stream)."""

    def __init__(self, process_group=None):
        self.lib = EmulatedLib()
        self.device = torch.device("cpu")
        self.ctx = self.queue = self.thr = self
        self.dev_type, self.plat_name = "CPU-emulated", "tests"
        self.process_group = process_group
        self.stream = None
        self.generator = torch.Generator(device="cpu")
        self.generator.manual_seed(99)

    def synchronize(self):
        pass

    finish = synchronize


def make_solver(cfg, process_group=None, comm=None):
    """A chimeracl_b200 Solver whose C-ABI calls land in EmulatedLib."""
    from chimeracl_b200.solver import Solver
    comm = comm or EmulatedComm(process_group)
    saved = real_lib._lib
    real_lib._lib = comm.lib        # init_generic_methods() re-reads _lib.load()
    try:
        return Solver(dict(cfg), comm)
    finally:
        real_lib._lib = saved


def make_particles(cfg, comm):
    from chimeracl_b200.particles import Particles
    saved = real_lib._lib
    real_lib._lib = comm.lib
    try:
        return Particles(dict(cfg), comm)
    finally:
        real_lib._lib = saved


class _HostEvent:
    """torch.cuda.Event stand-in (the host code records one for the Np_stay read-back)."""

    def __init__(self, *a, **k):
        pass

    def record(self, *a):
        pass

    def synchronize(self):
        pass


def patch_cuda_host_calls(monkeypatch):
    """The particle mixin touches two CUDA-only host facilities (an event and pinned
    memory for the asynchronous Np_stay read-back); give them host stand-ins."""
    monkeypatch.setattr(torch.cuda, "Event", _HostEvent)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
