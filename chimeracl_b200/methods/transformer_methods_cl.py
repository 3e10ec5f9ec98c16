"""Fourier-Bessel transformer methods mixin.

Interface of the reference's chimeraCL/methods/transformer_methods_cl.py (same
method names, DataDev keys and argument meaning).  Differences are only in how the
work is scheduled on the device:

  transform_field (:38-64) / _transform_forward (:290-332) / _transform_backward
  (:334-383) / _half_transform_* (:385-455):
      DHT  = chb_dht  (FP64 DMMA contraction reading / writing rows [1:] in place)
      FFT  = chb_fft_x (in-smem FFT with the cast, phase shift and real-part
             extraction fused into its load / store) -- 2 launches per
             component-mode instead of 5-6 full-array passes.
  field_grad (:87-133), field_rot (:185-263): same operator algebra, with the
      "b = dDHT.x; y += a*b; z += c*b" triples folded into one two-output
      contraction (chb_dht2).

kr-row sharding (multi-GPU field solve, no reference counterpart; see
Solver.enable_spectral_sharding): while `self._shard` is set, every method here works on
the kr rows [lo, hi) this rank owns -- element-wise work, FFTs (row-local) and the OUTPUT
rows of the forward / grad / rot / div contractions (operator rows [lo, hi), full
right-hand sides, which the caller all-gathers first) -- and the backward transform
contracts over the owned rows only (operator columns [lo, hi)), leaving a partial sum in
the grid arrays that the caller adds up over the ranks.  All of it is pointer arithmetic
on the same C-ABI calls: a row block of a C-ordered array is contiguous.
"""
import numpy as np

from .. import _lib
from ..devarray import DevArray
from .generic_methods_cl import GenericMethodsCL


def _next_pow2(n):
    p = 1
    while p < n:
        p <<= 1
    return p


def fft_plan_tables(Nx):
    """Host-side plan for chb_fft_x: (L, twiddles[L], chirp[Nx] | None, bfft[L] | None).
    Power-of-two Nx -> direct; otherwise Bluestein with L = pow2 >= 2Nx-1."""
    pow2 = Nx >= 8 and (Nx & (Nx - 1)) == 0
    L = Nx if pow2 else max(_next_pow2(2 * Nx - 1), 8)
    j = np.arange(L)
    tw = np.exp(-2j * np.pi * j / L)
    if pow2 and L == 16384:
        # split rows: roots of the half-length transform, then W_L^i for i < L/2
        h = np.arange(L // 2)
        tw = np.concatenate((np.exp(-2j * np.pi * h / (L // 2)), np.exp(-2j * np.pi * h / L)))
    if pow2:
        return L, tw, None, None
    n = np.arange(Nx, dtype=np.int64)
    # exp(-i pi n^2 / Nx) with the exponent reduced mod 2Nx for accuracy
    chirp = np.exp(-1j * np.pi * ((n * n) % (2 * Nx)) / Nx)
    b = np.zeros(L, dtype=np.complex128)
    b[:Nx] = np.conj(chirp)
    b[L - Nx + 1:] = np.conj(chirp[1:][::-1])
    bfft = np.fft.fft(b) / L
    return L, tw, chirp, bfft


class TransformerMethodsCL(GenericMethodsCL):
    def init_transformer_methods(self):
        self.init_generic_methods()
        self._prepare_fft()
        self._prepare_dot()

    # ------------------------------------------------------------------ plans
    def _prepare_fft(self):
        Nx = int(self.Args['Nx'])
        L, tw, chirp, bfft = fft_plan_tables(Nx)
        if L > self._lib.chb_fft_max_pow2():
            raise ValueError("chimera_b200: Nx=%d needs an FFT of length %d > %d "
                             "(multi-pass FFT not implemented yet)" %
                             (Nx, L, self._lib.chb_fft_max_pow2()))
        dev = self.comm.device
        self._fft_L = L
        self._fft_tw = DevArray.from_numpy(tw, dev)
        self._fft_chirp = DevArray.from_numpy(chirp, dev) if chirp is not None else None
        self._fft_bfft = DevArray.from_numpy(bfft, dev) if bfft is not None else None

    def _fft_rows(self, src, dst, direction, phase=None, in_real=False, out_real=False,
                  out_filter=None):
        """dst = FFT_x(src) row by row; src/dst are 2-D DevArrays (views allowed) or
        equally long lists of them (one batched launch)."""
        srcs = src if isinstance(src, (list, tuple)) else [src]
        dsts = dst if isinstance(dst, (list, tuple)) else [dst]
        rows, Nx = srcs[0].shape
        for i in range(0, len(srcs), 16):
            sl, dl = srcs[i:i + 16], dsts[i:i + 16]
            self._call('chb_fft_x_batched', _lib.ptr_array([a.ptr for a in sl]),
                       _lib.ptr_array([a.ptr for a in dl]), len(sl), rows, Nx,
                       sl[0].t.stride(0), dl[0].t.stride(0),
                       1 if direction == 1 else 0, int(in_real), int(out_real),
                       phase.ptr if phase is not None else None,
                       1 if direction == 1 else 0, self._fft_tw.ptr, self._fft_L,
                       self._fft_chirp.ptr if self._fft_chirp is not None else None,
                       self._fft_bfft.ptr if self._fft_bfft is not None else None,
                       out_filter.ptr if out_filter is not None else None)

    def _fft(self, arr_out, arr, dir):
        """Plain FFT along axis 1 (numpy conventions, normalised inverse): the
        reference's Reikna plan call `_fft(out, in, inverse)` (:507-509)."""
        self._fft_rows(arr, arr_out, dir)
        return arr_out

    def _prepare_dot(self):
        pass

    def _dot(self, c, a, b, alpha=1.0, accumulate=False):
        cplx = b.dtype == np.complex128
        M, K = a.shape
        N = b.shape[1]
        alpha = complex(alpha)
        self._call('chb_dht', a.ptr, a.t.stride(0), b.ptr, b.t.stride(0), c.ptr, c.t.stride(0),
                   M, K, N, int(cplx), alpha.real, alpha.imag, int(accumulate))

    def _ddot(self, c, a, b):
        self._dot(c, a, b)

    def _cdot(self, c, a, b):
        self._dot(c, a, b)

    def _cdot2(self, a, b, c1, alpha1, acc1, c2, alpha2, acc2, hermitian=False):
        M, K = a.shape
        N = b.shape[1]
        a1, a2 = complex(alpha1), complex(alpha2)
        if hermitian:
            # b is the x-spectrum of a real field: half of the columns are contracted
            self._call('chb_dht2_hermitian', a.ptr, a.t.stride(0), b.ptr, b.t.stride(0), c1.ptr,
                       a1.real, a1.imag, int(acc1), c2.ptr, a2.real, a2.imag, int(acc2),
                       c1.t.stride(0), M, K, N)
            return
        self._call('chb_dht2', a.ptr, a.t.stride(0), b.ptr, b.t.stride(0), c1.ptr, a1.real,
                   a1.imag, int(acc1), c2.ptr, a2.real, a2.imag, int(acc2), c1.t.stride(0),
                   M, K, N, 1)

    def _dot_batched(self, cs, a, bs, accumulate=False):
        """cs[k] = a . bs[k] for equally shaped right-hand sides, one launch
        (accumulate: cs[k] += ..., one launch per right-hand side)."""
        if accumulate:
            for c, b in zip(cs, bs):
                self._dot(c, a, b, accumulate=True)
            return
        cplx = bs[0].dtype == np.complex128
        M, K = a.shape
        N = bs[0].shape[1]
        for i in range(0, len(bs), 16):
            bl, cl = bs[i:i + 16], cs[i:i + 16]
            self._call('chb_dht_batched', a.ptr, a.t.stride(0),
                       _lib.ptr_array([x.ptr for x in bl]), _lib.ptr_array([x.ptr for x in cl]),
                       len(bl), bl[0].t.stride(0), cl[0].t.stride(0), M, K, N, int(cplx))

    # ------------------------------------------------------------------ kr-row shard views
    def _own(self, a):
        """Rows of a (Nr-1, ...) spectral array / filter / operator matrix owned by the
        current shard (the array itself when the solve is not sharded)."""
        sh = self.__dict__.get('_shard')
        if sh is None:
            return a
        # the views are cached (keyed by storage address and extent, so that a fresh
        # `arr[1:]` view of the same storage hits too): ~170 slices per step otherwise
        cache = self.__dict__.setdefault('_own_cache', {})
        t = a.t
        key = (t.data_ptr(), t.dtype, t.shape, t.stride(0), sh.lo, sh.hi)
        view = cache.get(key)
        if view is None:
            view = cache[key] = a[sh.lo:sh.hi]
        return view

    def _kcols(self, mat):
        """Operator columns matching the owned rows of a right-hand side."""
        sh = self.__dict__.get('_shard')
        if sh is None:
            return mat
        cache = self.__dict__.setdefault('_own_cache', {})
        key = (mat.t.data_ptr(), 'cols', sh.lo, sh.hi)
        view = cache.get(key)
        if view is None:
            view = cache[key] = mat[:, sh.lo:sh.hi]
        return view

    def _shard_is_empty(self):
        sh = self.__dict__.get('_shard')
        return sh is not None and sh.hi <= sh.lo

    # ------------------------------------------------------------------ transforms
    def _phase(self, dir):
        """exp(-+ i kx Xmin) from the HOST Xmin (moving-window aware, reference
        :40-44); recomputed only when Xmin changed."""
        xmin = float(self.Args['Xmin'])
        cache = self.__dict__.setdefault('_phase_cache', {})
        hit = cache.get(dir)
        if hit is None or hit[0] != xmin:
            arr = hit[1] if hit is not None else DevArray.empty(self.Args['Nx'], np.complex128,
                                                                self.comm.device)
            self._call('chb_get_phase', arr.ptr, self.DataDev['kx'].ptr, xmin, int(dir),
                       int(self.Args['Nx']))
            cache[dir] = (xmin, arr)
            hit = cache[dir]
        self.DataDev['phs_shft'] = hit[1]
        return hit[1]

    def _tmp(self, kind, n):
        """Pool of (Nr-1, Nx) scratch arrays for batched transforms."""
        pool = self.__dict__.setdefault('_tmp_pool', {'d': [], 'c': []})[kind]
        shape = (self.Args['Nr'] - 1, self.Args['Nx'])
        while len(pool) < n:
            pool.append(DevArray.empty(shape, np.double if kind == 'd' else np.complex128,
                                       self.comm.device))
        return pool[:n]

    def transform_field(self, arg_cmp, dir, mode):
        self.transform_fields([arg_cmp], dir, mode)

    def transform_fields(self, comps, dir, mode, smooth=False):
        """Forward (dir=0) / backward (dir=1) Fourier-Bessel transform of several
        components at once: per azimuthal mode one batched DHT and one batched FFT
        launch (mode='half': FFT only).  Same arithmetic per component as
        reference transformer_methods_cl.py:290-455.  smooth=True (forward only) also
        applies SmoothingFilter_m to the result, i.e. the fields_smooth() call that
        follows the forward transform in pic_loop.py:99-103, without an extra pass."""
        if not comps:
            return
        D = self.DataDev
        phs = self._phase(dir)
        full = (mode == 'full')
        n = len(comps)
        own = self._own
        sh = self.__dict__.get('_shard')
        # sharded backward transform: the grid arrays receive this shard's partial sum
        # (virtual shards of one process add to what the earlier ones left)
        acc = sh is not None and not sh.first
        for m in range(self.Args['M'] + 1):
            ms = str(m)
            real = (m == 0)
            grid = [D[c + '_m' + ms][1:] for c in comps]
            spec = [own(D[c + '_fb_m' + ms]) for c in comps]
            if self._shard_is_empty():
                if dir == 1 and full and not acc:
                    for g in grid:
                        g.fill(0.)
                continue
            if dir == 0:
                flt = own(D['SmoothingFilter_m' + ms]) if smooth else None
                if not full:
                    self._fft_rows([own(g) for g in grid], spec, 0, phs, in_real=real,
                                   out_filter=flt)
                elif real:
                    tmp = [own(t) for t in self._tmp('d', n)]
                    self._dot_batched(tmp, own(D['DHT_m0']), grid)
                    self._fft_rows(tmp, spec, 0, phs, in_real=True, out_filter=flt)
                else:
                    self._dot_batched(spec, own(D['DHT_m' + ms]), grid)
                    self._fft_rows(spec, spec, 0, phs, out_filter=flt)   # in place
            else:
                if not full:
                    self._fft_rows(spec, [own(g) for g in grid], 1, phs, out_real=real)
                else:
                    tmp = [own(t) for t in self._tmp('d' if real else 'c', n)]
                    self._fft_rows(spec, tmp, 1, phs, out_real=real)
                    self._dot_batched(grid, self._kcols(D['DHT_inv_m' + ms]), tmp,
                                      accumulate=acc)

    # the reference's per-component helpers (transformer_methods_cl.py:290-455); kept
    # as entry points with the reference signature, served by the batched path above
    def _split_names(self, dht_arg, arg_in, arg_out, forward):
        grid, spec = (arg_in, arg_out) if forward else (arg_out, arg_in)
        comp = grid[:-2]
        want = 'DHT_m' if forward else 'DHT_inv_m'
        if not (grid.endswith('_m') and spec == comp + '_fb_m' and dht_arg == want):
            raise ValueError("chimera_b200: unsupported array naming in transform helper: "
                             "%s, %s, %s" % (dht_arg, arg_in, arg_out))
        return comp

    def _transform_forward(self, dht_arg, arg_in, arg_out, phs_shft=None):
        self.transform_fields([self._split_names(dht_arg, arg_in, arg_out, True)], 0, 'full')

    def _transform_backward(self, dht_arg, arg_in, arg_out, phs_shft=None):
        self.transform_fields([self._split_names(dht_arg, arg_in, arg_out, False)], 1, 'full')

    def _half_transform_forward(self, dht_arg, arg_in, arg_out, phs_shft=None):
        self.transform_fields([self._split_names(dht_arg, arg_in, arg_out, True)], 0, 'half')

    def _half_transform_backward(self, dht_arg, arg_in, arg_out, phs_shft=None):
        self.transform_fields([self._split_names(dht_arg, arg_in, arg_out, False)], 1, 'half')

    # ------------------------------------------------------------------ spectral operators
    def field_poiss_vec(self, fld):
        if self._shard_is_empty():
            return
        for m in range(self.Args['M'] + 1):
            for comp in self.Args['vec_comps']:
                self.mult_elementwise(self._own(self.DataDev['Poiss_m' + str(m)]),
                                      self._own(self.DataDev[fld + comp + '_fb_m' + str(m)]))

    def field_poiss_scl(self, fld):
        if self._shard_is_empty():
            return
        for m in range(self.Args['M'] + 1):
            self.mult_elementwise(self._own(self.DataDev['Poiss_m' + str(m)]),
                                  self._own(self.DataDev[fld + '_fb_m' + str(m)]))

    def fields_smooth(self, flds):
        if self._shard_is_empty():
            return
        for m in range(self.Args['M'] + 1):
            for fld in flds:
                self.mult_elementwise(self._own(self.DataDev['SmoothingFilter_m' + str(m)]),
                                      self._own(self.DataDev[fld + '_fb_m' + str(m)]))

    def _mirror_axpy(self, out, b, alpha, beta, accumulate):
        alpha, beta = complex(alpha), complex(beta)
        self._call('chb_mirror_axpy', out.ptr, b.ptr, alpha.real, alpha.imag, beta.real,
                   beta.imag, int(accumulate), b.size, int(self.Args['Nx']))

    def _m0_real(self):
        """True while the caller vouches that the m = 0 spectral arrays are spectra of REAL
        grid fields (F(kr, -kx) = conj F(kr, kx)), which lets the contractions of m = 0
        sources run on half of the kx columns.  PIC_loop.step() sets it for its own calls
        (its m = 0 spectra are FFTs of real arrays, and damp_fields projects E and G onto
        real x-space fields every step); the reference API itself accepts arbitrary
        complex m = 0 spectra, so the default is False."""
        return bool(self.__dict__.get('m0_spectra_of_real_fields', False))

    def _m0_pm_identical(self):
        """dDHT_minus_m0 == dDHT_plus_m0 (jn_zeros(-1, n) == jn_zeros(1, n)): then
        D.F_{-1} = -conj(mirror(D.F_{+1})) and the m=0 'minus' contractions are free."""
        flag = self.__dict__.get('_m0_pm_flag')
        if flag is None:
            D = self.DataDev
            flag = bool(self.Args['M'] > 0 and
                        np.array_equal(D['dDHT_minus_m0'].get(), D['dDHT_plus_m0'].get()))
            self._m0_pm_flag = flag
        return flag

    def field_grad(self, scl_in, vec_out):
        """Sharded solve: outputs and the element-wise source are the owned rows; the
        contraction sources scl_in_fb_m* must be valid on ALL rows (all-gathered)."""
        D, M = self.DataDev, self.Args['M']
        own = self._own
        fast0 = self._m0_pm_identical()
        if self._shard_is_empty():
            return
        if not fast0:
            self._get_mm1_scl(scl_in)
        for m in range(M + 1):
            ox, oy, oz = (own(D[vec_out + c + '_fb_m' + str(m)]) for c in 'xyz')
            self.ab_dot_x(1.j, D['kx'], own(D[scl_in + '_fb_m' + str(m)]), ox)
            if m == 0 and fast0:
                # b+ = dDHT+ . scl_1 ; b- = dDHT- . scl_{-1} = -conj(mirror(b+))
                bp = own(D['fld_buff0_c'])
                self._dot(bp, own(D['dDHT_plus_m0']), D[scl_in + '_fb_m1'])
                self._mirror_axpy(oy, bp, 1., 1., False)       # oy = -b- + b+
                self._mirror_axpy(oz, bp, -1.j, 1.j, False)    # oz = -i b- - i b+
                continue
            if m > 0:
                src = D[scl_in + '_fb_m' + str(m - 1)]
            elif M > 0:
                src = D['buff_fb_m-1_x']
            else:
                self.set_to(oy, 0.)
                self.set_to(oz, 0.)
                continue
            # oy = -b, oz = -i b with b = dDHT_minus . src
            self._cdot2(own(D['dDHT_minus_m' + str(m)]), src, oy, -1., False, oz, -1.j, False,
                        hermitian=(m == 1 and self._m0_real()))
            if m < M:
                # oy += b, oz -= i b with b = dDHT_plus . scl_{m+1}
                self._cdot2(own(D['dDHT_plus_m' + str(m)]), D[scl_in + '_fb_m' + str(m + 1)],
                            oy, 1., True, oz, -1.j, True)

    def field_div(self, vec_in, scl_out):
        """Sharded solve: vec_in{y,z}_fb_m* must be valid on all rows (see field_grad)."""
        D, M = self.DataDev, self.Args['M']
        own = self._own
        if self._shard_is_empty():
            return
        for comp in ['y', 'z']:
            self._get_mm1_scl(vec_in + comp, comp)
        for m in range(M + 1):
            out = own(D[scl_out + '_fb_m' + str(m)])
            self.ab_dot_x(1.j, D['kx'], own(D[vec_in + 'x' + '_fb_m' + str(m)]), out)
            if m > 0:
                fy, fz = (D[vec_in + c + '_fb_m' + str(m - 1)] for c in 'yz')
            elif M > 0:
                fy, fz = D['buff_fb_m-1_y'], D['buff_fb_m-1_z']
            else:
                continue
            self.axpbyz(-1.j, fz, -1.0, fy, D['fld_buff0_c'])
            self._dot(out, own(D['dDHT_minus_m' + str(m)]), D['fld_buff0_c'], accumulate=True)
            if m < M:
                fy, fz = (D[vec_in + c + '_fb_m' + str(m + 1)] for c in 'yz')
                self.axpbyz(-1.j, fz, 1.0, fy, D['fld_buff0_c'])
                self._dot(out, own(D['dDHT_plus_m' + str(m)]), D['fld_buff0_c'],
                          accumulate=True)

    def field_rot(self, fld_in, fld_out):
        """Sharded solve: outputs and the element-wise sources are the owned rows; the
        contraction sources fld_in{x,y,z}_fb_m* must be valid on ALL rows (all-gathered);
        their linear combinations (fld_buff0_c) are formed on all rows by every shard."""
        D, M = self.DataDev, self.Args['M']
        own = self._own
        fast0 = self._m0_pm_identical()
        if self._shard_is_empty():
            return
        if not fast0:
            self._get_mm1_vec(fld_in)
        for m in range(M + 1):
            ox, oy, oz = (own(D[fld_out + c + '_fb_m' + str(m)]) for c in 'xyz')
            self.ab_dot_x(-1.j, D['kx'], own(D[fld_in + 'z' + '_fb_m' + str(m)]), oy)
            self.ab_dot_x(1.j, D['kx'], own(D[fld_in + 'y' + '_fb_m' + str(m)]), oz)
            if m == 0 and fast0:
                fx, fy, fz = (D[fld_in + c + '_fb_m1'] for c in 'xyz')
                dp = own(D['dDHT_plus_m0'])
                b0, b1 = own(D['fld_buff0_c']), own(D['fld_buff1_c'])
                # X+ = dDHT+.(fz + i fy); the m-1 term is conj(mirror(X+))
                self.axpbyz(1, fz, 1.j, fy, D['fld_buff0_c'])
                self._dot(b1, dp, D['fld_buff0_c'])
                self._mirror_axpy(ox, b1, 1., 1., False)
                # b' = dDHT+.fx ; b = dDHT-.fx_{-1} = -conj(mirror(b'))
                self._dot(b0, dp, fx)
                self._mirror_axpy(oy, b0, -1.j, 1.j, True)   # oy -= i b + i b'
                self._mirror_axpy(oz, b0, -1., -1., True)    # oz += b - b'
                continue
            if m > 0:
                fx, fy, fz = (D[fld_in + c + '_fb_m' + str(m - 1)] for c in 'xyz')
            elif M > 0:
                fx, fy, fz = (D['buff_fb_m-1_' + c] for c in 'xyz')
            else:
                self.set_to(ox, 0.0)
                continue
            dm = own(D['dDHT_minus_m' + str(m)])
            self.axpbyz(-1, fz, 1.j, fy, D['fld_buff0_c'])
            self._dot(ox, dm, D['fld_buff0_c'])                       # ox  = dDHT-.(-fz + i fy)
            self._cdot2(dm, fx, oy, -1.j, True, oz, 1., True,         # oy -= i b, oz += b
                        hermitian=(m == 1 and self._m0_real()))
            if m < M:
                fx, fy, fz = (D[fld_in + c + '_fb_m' + str(m + 1)] for c in 'xyz')
                dp = own(D['dDHT_plus_m' + str(m)])
                self.axpbyz(1, fz, 1.j, fy, D['fld_buff0_c'])
                self._dot(ox, dp, D['fld_buff0_c'], accumulate=True)  # ox += dDHT+.(fz + i fy)
                self._cdot2(dp, fx, oy, -1.j, True, oz, -1., True)    # oy -= i b, oz -= b

    def _get_mm1_vec(self, fld):
        if self.Args['M'] == 0:
            return
        for comp in self.Args['vec_comps']:
            self._get_mm1_scl(fld + comp, comp)

    def _get_mm1_scl(self, fld, comp='x'):
        # for a scalar the X component of the buffer is used (reference :277-288)
        if self.Args['M'] == 0:
            return
        self._call('chb_get_m1', self.DataDev['buff_fb_m-1_' + comp].ptr,
                   self.DataDev[fld + '_fb_m1'].ptr, int(self.Args['NxNrm1']),
                   int(self.Args['Nx']))
