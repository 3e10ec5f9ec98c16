"""Grid methods mixin: deposition, axis fix-ups, gather + Boris push.

Interface of the reference's chimeraCL/methods/grid_methods_cl.py; compute goes to
libchimera_b200.so:

  depose_scalar / depose_vector       :44-95   -> chb_depose_scalar / chb_depose_vector
        (one launch instead of four colour passes)
  postproc_depose_scalar / _vector    :98-151  -> chb_postproc_depose (all arrays, one launch)
  preproc_project_vec                 :153-166 -> chb_warp_axis
  _gather_and_push                    :168-192 -> chb_gather_push

Multi-GPU: when the communicator carries a process group, the raw deposits are
summed over ranks (NCCL all-reduce) before the axis / volume post-processing; both
are linear, so the result equals a single-device deposit of all particles.
"""
import numpy as np

from .. import _lib
from .generic_methods_cl import GenericMethodsCL


class GridMethodsCL(GenericMethodsCL):
    def init_grid_methods(self):
        self.init_generic_methods()
        if self.Args['M'] > 2:
            raise ValueError("chimera_b200 supports azimuthal modes M <= 2")
        if 'vec_comps' not in self.Args:
            self.Args['vec_comps'] = ['x', 'y', 'z']

    def _geom(self):
        D = self.DataDev
        return (int(self.Args['Nx']), int(self.Args['Nr']), D['Xmin'].ptr, D['dx_inv'].ptr,
                D['Rmin'].ptr, D['dr_inv'].ptr)

    def _mode_fields(self, name):
        return [self.DataDev[name + '_m' + str(m)] for m in range(self.Args['M'] + 1)]

    # ------------------------------------------------------------------ deposition
    def depose_scalar(self, parts, src_scalar, dest_fld, charge):
        if parts.Args['Np'] <= 0:
            return
        P = parts.DataDev
        flds = _lib.ptr_array([f.ptr for f in self._mode_fields(dest_fld)])
        self._call('chb_depose_scalar', int(self.Args['M']), P['sort_indx'].ptr, P['x'].ptr,
                   P['y'].ptr, P['z'].ptr, P[src_scalar].ptr, P['cell_offset'].ptr,
                   int(np.int8(charge)), *self._geom(), flds)

    def depose_vector(self, parts, vec, factors, vec_fld, charge, push_dt=None,
                      second_push_index=False):
        if parts.Args['Np'] <= 0:
            return
        P = parts.DataDev
        flds = []
        for m in range(self.Args['M'] + 1):
            for comp in self.Args['vec_comps']:
                flds.append(self.DataDev[vec_fld + comp + '_m' + str(m)].ptr)
        if push_dt is not None and second_push_index:
            # one pass: push, deposit, second push, cell index + histogram
            # (chb_push_depose_push_index); the caller finishes the sort
            indx, hist = parts.prepare_index(self)
            self._call('chb_push_depose_push_index', int(self.Args['M']), P['sort_indx'].ptr,
                       P['x'].ptr, P['y'].ptr, P['z'].ptr, P[vec[0]].ptr, P[vec[1]].ptr,
                       P[vec[2]].ptr, P[factors[0]].ptr, P[factors[1]].ptr,
                       P['cell_offset'].ptr, P[push_dt].ptr, int(parts.Args['Np']),
                       int(np.int8(charge)), *self._geom(), _lib.ptr_array(flds),
                       indx.ptr, hist.ptr, *parts.exception_workspace())
            parts.exception_count_readback()
            parts.flag_sorted = False
            parts._index_prefilled = True
            return
        if push_dt is not None:
            # coordinates advance by push_dt ('dt_2' | 'dt') inside the deposit; the
            # previous sort only provides the traversal order (chb_push_depose_vector)
            self._call('chb_push_depose_vector', int(self.Args['M']), P['sort_indx'].ptr,
                       P['x'].ptr, P['y'].ptr, P['z'].ptr, P[vec[0]].ptr, P[vec[1]].ptr,
                       P[vec[2]].ptr, P[factors[0]].ptr, P[factors[1]].ptr,
                       P['cell_offset'].ptr, P[push_dt].ptr, int(parts.Args['Np']),
                       int(np.int8(charge)), *self._geom(), _lib.ptr_array(flds),
                       *parts.exception_workspace())
            parts.exception_count_readback()
            parts.flag_sorted = False
            return
        # factors = ['g_inv', 'w'] in the reference call (grid.py:49-51)
        self._call('chb_depose_vector', int(self.Args['M']), P['sort_indx'].ptr, P['x'].ptr,
                   P['y'].ptr, P['z'].ptr, P[vec[0]].ptr, P[vec[1]].ptr, P[vec[2]].ptr,
                   P[factors[0]].ptr, P[factors[1]].ptr, P['cell_offset'].ptr,
                   int(np.int8(charge)), *self._geom(), _lib.ptr_array(flds))

    def _allreduce(self, group, arrays):
        pg = getattr(self.comm, 'process_group', None)
        if pg is None:
            return
        from ..parallel import allreduce_sum
        flat = getattr(self, '_flat', {}).get(group)
        if flat is not None:
            allreduce_sum([flat], pg)          # in place, one collective
        else:
            allreduce_sum([a.t for a in arrays], pg)

    def _postproc(self, names, reduce=True):
        arrays = []
        for name in names:
            arrays += self._mode_fields(name)
        if reduce:
            self._allreduce('rho' if names == ['rho'] else ('J' if names[0][0] == 'J' else None),
                            arrays)
        ptrs = _lib.ptr_array([a.ptr for a in arrays])
        flags = _lib.int_array([1 if a.dtype == np.complex128 else 0 for a in arrays])
        self._call('chb_postproc_depose', ptrs, flags, len(arrays), int(self.Args['Nx']),
                   int(self.Args['Nr']), self.DataDev['dV_inv'].ptr)

    def postproc_depose_scalar(self, fld):
        self._postproc([fld])

    def postproc_depose_vector(self, vec_fld, reduce=True):
        self._postproc([vec_fld + comp for comp in self.Args['vec_comps']], reduce=reduce)

    def start_reduce_currents(self):
        """Multi-GPU: launch the sum of the raw J deposits over ranks asynchronously."""
        pg = getattr(self.comm, 'process_group', None)
        flat = getattr(self, '_flat', {}).get('J')
        if pg is None or flat is None:
            return None
        peer = getattr(self, 'peer_reduce_flat', None)
        work = peer('J') if peer is not None else None
        if work is not None:
            return work
        from ..parallel import allreduce_sum_async
        return allreduce_sum_async(flat, pg)

    # ------------------------------------------------------------------ gather
    def preproc_project_vec(self, vec_fld):
        arrays = []
        for comp in self.Args['vec_comps']:
            arrays += self._mode_fields(vec_fld + comp)
        ptrs = _lib.ptr_array([a.ptr for a in arrays])
        flags = _lib.int_array([1 if a.dtype == np.complex128 else 0 for a in arrays])
        self._call('chb_warp_axis', ptrs, flags, len(arrays), int(self.Args['Nx']))

    def _gather_and_push(self, parts, flds):
        P = parts.DataDev
        Np = int(parts.Args['Np'])
        if Np == 0:
            return
        ptrs = []
        for m in range(self.Args['M'] + 1):
            for fld in flds:
                for comp in self.Args['vec_comps']:
                    ptrs.append(self.DataDev[fld + comp + '_m' + str(m)].ptr)
        if 'Np_stay_dev' not in P:
            raise RuntimeError("gather_and_push needs sorted particles (call sort_parts first)")
        self._call('chb_gather_push', int(self.Args['M']), P['x'].ptr, P['y'].ptr, P['z'].ptr,
                   P['px'].ptr, P['py'].ptr, P['pz'].ptr, P['g_inv'].ptr, P['sort_indx'].ptr,
                   P['cell_offset'].ptr, P['FactorPush'].ptr, Np, P['Np_stay_dev'].ptr, *self._geom(),
                   _lib.ptr_array(ptrs))
