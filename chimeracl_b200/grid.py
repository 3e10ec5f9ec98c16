"""Grid wrapper class (API of the reference's chimeraCL/grid.py)."""
import numpy as np

from .methods.generic_methods_cl import ArgsDict
from .methods.grid_methods_cl import GridMethodsCL


def grid_geometry(A):
    """r-x grid geometry and index products, reference grid.py:68-103.  Pure host
    math on a dict (no device needed)."""
    if 'M' not in A:
        A['M'] = 0
    Nx, Nr = A['Nx'], A['Nr']
    A['dx'] = (A['Xmax'] - A['Xmin']) / (Nx - 1)
    A['dx_inv'] = 1. / A['dx']
    A['dr'] = A['Rmax'] / (Nr - 1.5)
    A['dr_inv'] = 1. / A['dr']
    if 'dt' not in A:
        A['dt'] = A['dx']
    A['dt_inv'] = 1.0 / A['dt']
    A['Xgrid'] = A['Xmin'] + A['dx'] * np.arange(Nx)
    A['Rmin'] = -0.5 * A['dr']               # ghost row at r = -dr/2
    A['Rgrid'] = A['Rmin'] + A['dr'] * np.arange(Nr)
    A['Rmax'] = A['Rgrid'].max()
    with np.errstate(divide='ignore'):
        A['dV_inv'] = (A['Rgrid'] > 0) / (2 * np.pi * A['dx'] * A['dr'] * A['Rgrid'])
    A['NxNr'] = Nr * Nx
    A['Nxm1Nrm1'] = (Nr - 1) * (Nx - 1)
    A['NxNrm1'] = (Nr - 1) * Nx
    A['NxNr_4'] = Nr // 2 * Nx // 2
    A['dont_send'] = []
    A['dont_keep'] = []
    return A


class Grid(GridMethodsCL):
    def __init__(self, configs_in, comm):
        self.import_comm(comm)
        self._process_configs(configs_in)
        if 'vec_comps' not in self.Args:
            self.Args['vec_comps'] = ['x', 'y', 'z']
        self.init_grid_methods()
        self.DataDev = {}
        self._init_grid_data_on_dev()
        self.send_args_to_dev()

    def finish_charge(self):
        """Complete a depose_charge(..., defer=True) (see finish_currents)."""
        pending = self.__dict__.pop('_pending_rho', None)
        if pending is None:
            return
        if pending is not True:
            pending.wait()
        self._postproc(['rho'], reduce=False)

    def depose_charge(self, species=[], defer=False):
        self._flat['rho'].zero_()
        for parts in species:
            self.depose_scalar(parts, 'w', 'rho', charge=parts.Args['charge'])
        if defer:
            pg = getattr(self.comm, 'process_group', None)
            work = None
            if pg is not None:
                peer = getattr(self, 'peer_reduce_flat', None)
                work = peer('rho') if peer is not None else None
                if work is None:
                    from .parallel import allreduce_sum_async
                    work = allreduce_sum_async(self._flat['rho'], pg)
            self._pending_rho = work if work is not None else True
            return
        self.postproc_depose_scalar('rho')

    def finish_currents(self):
        """Complete a depose_currents(..., defer=True): wait for the all-reduce of the
        raw deposits (multi-GPU) and apply the axis / volume post-processing."""
        pending = self.__dict__.pop('_pending_J', None)
        if pending is None:
            return
        # a bounded cell-changer queue (more particles than a quarter of the device memory
        # holds records for) must not have overflowed: checked here, before J is consumed
        for parts in self.__dict__.pop('_fused_species', ()):
            parts._check_exception_overflow()
        if pending is not True:
            pending.wait()
        self.postproc_depose_vector('J', reduce=False)

    def depose_currents(self, species=[], defer=False, push_mode=None):
        comps = self.Args['vec_comps']
        self._flat['J'].zero_()
        for parts in species:
            if 'Immobile' in parts.Args.keys():
                continue
            push_dt = None
            if push_mode is not None:
                # fused push_coords + sort_parts + deposit for this species; with
                # 'half+half' also the second half push and the cell index of the
                # following sort_parts (pic_loop.py:70-76 in one pass)
                push_dt = 'dt' if push_mode == 'full' else 'dt_2'
            self.depose_vector(parts, ['p' + comp for comp in comps], ['g_inv', 'w'], 'J',
                               charge=parts.Args['charge'], push_dt=push_dt,
                               second_push_index=(push_mode == 'half+half'))
            if push_dt is not None:
                self.__dict__.setdefault('_fused_species', []).append(parts)
        if not defer:
            for parts in self.__dict__.pop('_fused_species', ()):
                parts._check_exception_overflow()
        if defer:
            # the sum over ranks runs while the caller goes on (second push + sort);
            # finish_currents() must be called before J is used
            work = self.start_reduce_currents()
            self._pending_J = work if work is not None else True
            return
        self.postproc_depose_vector('J')

    def gather_and_push(self, species=[]):
        for fld in ['E', 'B']:
            self.preproc_project_vec(fld)
        for parts in species:
            if 'Immobile' in parts.Args.keys():
                continue
            self._gather_and_push(parts, ['E', 'B'])

    def _process_configs(self, configs_in):
        self.Args = grid_geometry(ArgsDict(configs_in))

    def _init_grid_data_on_dev(self):
        comps = self.Args['vec_comps']
        shape = (self.Args['Nr'], self.Args['Nx'])
        for name in [f + comp for f in ('E', 'B', 'G') for comp in comps]:
            self.DataDev[name + '_m0'] = self.dev_arr(val=0, dtype=np.double, shape=shape)
            for m in range(1, self.Args['M'] + 1):
                self.DataDev[name + '_m' + str(m)] = self.dev_arr(val=0, dtype=np.complex128,
                                                                 shape=shape)
        # The deposited fields of one group (J: 3 comps x modes, rho: modes) are views
        # of ONE flat buffer each, so that the multi-GPU sum over ranks is a single
        # in-place all-reduce and the zero-fill a single memset.
        self._flat = {}
        for group, names in (('J', ['J' + comp for comp in comps]), ('rho', ['rho'])):
            self._flat[group] = self._alloc_group(names, shape)

    def _alloc_group(self, names, shape, alloc=None):
        """alloc(n) -> zeroed flat float64 device tensor of n elements (default: torch.zeros;
        the peer-memory exchange passes a symmetric-memory allocator)."""
        import torch
        from .devarray import DevArray
        n = shape[0] * shape[1]
        n_pad = (n + 1) // 2 * 2                       # keep complex views 16-byte aligned
        sizes = []
        for name in names:
            for m in range(self.Args['M'] + 1):
                sizes.append((name + '_m' + str(m), n_pad if m == 0 else 2 * n_pad, m > 0))
        total = sum(sz for _, sz, _ in sizes)
        flat = alloc(total) if alloc is not None else \
            torch.zeros(total, dtype=torch.float64, device=self.comm.device)
        off = 0
        for key, sz, cplx in sizes:
            if cplx:
                view = torch.view_as_complex(flat[off:off + 2 * n].view(n, 2)).view(shape)
            else:
                view = flat[off:off + n].view(shape)
            self.DataDev[key] = DevArray(view)
            off += sz
        return flat
