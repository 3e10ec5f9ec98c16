"""Maxwell (PSATD) solver methods mixin -- interface of the reference's
chimeraCL/methods/solver_methods_cl.py.

  advance_fields (:43-62) -> chb_psatd_advance, one launch per azimuthal mode
  profile_edges  (:64-83) -> chb_profile_edges, all components/modes in one launch
                             touching only the 2*DampCells edge columns per side
"""
import numpy as np

from .. import _lib
from .generic_methods_cl import GenericMethodsCL


def damping_profile(N):
    """sin^2 absorbing-edge profile over 2N cells (reference :32-41)."""
    N = int(N)
    z = np.arange(2 * N)
    z_shft = 3. * (z - N + 1) / (N + 1)
    return ((z < 4. * N / 3) * (z >= N) * np.sin(0.5 * np.pi * z_shft) ** 2
            + (z >= 4. * N / 3)).astype(np.double)


class SolverMethodsCL(GenericMethodsCL):
    def init_solver_methods(self):
        self.init_generic_methods()
        if 'DampCells' in self.Args:
            self._init_field_damping()

    def _init_field_damping(self):
        self.Args['DampProfile'] = damping_profile(self.Args['DampCells'])
        self.Args['dont_keep'].append('DampProfile')
        self.Args['dont_send'].append('DampCells')

    def advance_fields(self, vecs):
        """Element-wise in (kr, kx): with a kr-row sharded solve (transformer_methods_cl.py)
        only the owned rows are advanced."""
        D = self.DataDev
        comps = self.Args['vec_comps']
        own = self._own
        if self._shard_is_empty():
            return
        for m in range(self.Args['M'] + 1):
            ms = '_m' + str(m)
            groups = [_lib.ptr_array([own(D[v + c + '_fb' + ms]).ptr for c in comps])
                      for v in vecs]
            c1 = own(D['MxSlv_cos(wdt)' + ms])
            self._call('chb_psatd_advance', int(c1.size), D['dt_inv'].ptr,
                       c1.ptr, own(D['MxSlv_sin(wdt)*w' + ms]).ptr,
                       own(D['MxSlv_1/w**2' + ms]).ptr, *groups)

    def damp_fields_fused(self, flds):
        """The three calls of damp_fields (reference solver.py:32-35) as one in-place,
        on-chip pass per spectral row (chb_fft_damp_x_batched): same arithmetic per
        element, but the x-space intermediate never goes to the E/G grid arrays (the
        reference only uses them as scratch at this point).  Returns False when the
        shape is not covered (Nx not a power of two in [256, 8192], no damping) and
        the caller should run the three separate calls."""
        Nx = int(self.Args['Nx'])
        if 'DampProfile' not in self.DataDev or Nx < 256 or Nx > 8192 or Nx & (Nx - 1):
            return False
        if getattr(self, '_fft_L', None) != Nx:
            return False
        D = self.DataDev
        if self._shard_is_empty():
            return True
        phs_b, phs_f = self._phase(1), self._phase(0)
        arrays, real_x = [], []
        for fld in flds:
            for comp in self.Args['vec_comps']:
                for m in range(self.Args['M'] + 1):
                    arrays.append(self._own(D[fld + comp + '_fb_m' + str(m)]))
                    real_x.append(1 if m == 0 else 0)
        rows = arrays[0].shape[0]            # Nr-1, or the owned kr rows
        for i in range(0, len(arrays), 16):
            chunk = arrays[i:i + 16]
            self._call('chb_fft_damp_x_batched', _lib.ptr_array([a.ptr for a in chunk]),
                       _lib.int_array(real_x[i:i + 16]), len(chunk), rows, Nx,
                       chunk[0].t.stride(0), phs_b.ptr, phs_f.ptr, D['DampProfile'].ptr,
                       2 * int(self.Args['DampCells']), self._fft_tw.ptr)
        return True

    def profile_edges(self, flds):
        sh = self.__dict__.get('_shard')
        if self._shard_is_empty():
            return
        arrays = []
        for fld in flds:
            for comp in self.Args['vec_comps']:
                for m in range(self.Args['M'] + 1):
                    a = self.DataDev[fld + comp + '_m' + str(m)]
                    # sharded half transforms only hold the owned rows of grid[1:]
                    arrays.append(a if sh is None else a[1 + sh.lo:1 + sh.hi])
        for i in range(0, len(arrays), 16):
            chunk = arrays[i:i + 16]
            self._call('chb_profile_edges', _lib.ptr_array([a.ptr for a in chunk]),
                       _lib.int_array([1 if a.dtype == np.complex128 else 0 for a in chunk]),
                       len(chunk), self.DataDev['DampProfile'].ptr, int(chunk[0].shape[0]),
                       int(self.Args['Nx']), 2 * int(self.Args['DampCells']))
