"""Particle methods mixin: coordinate push, cell sort, align, particle creation.

Interface of the reference's chimeraCL/methods/particles_methods_cl.py (same method
names / DataDev keys); compute goes to libchimera_b200.so:

  push_coords     :206-223 -> chb_push_xyz
  index_sort      :225-261 -> chb_index_and_sum (or fused chb_push_index),
                              chb_cell_offsets, chb_sort_scatter_stable
  align_and_damp  :263-286 -> chb_align (all attributes in one launch)

Particle creation (make_new_domain / make_new_beam / dens_profile, :66-204) is the
init path: same lattice / profile formulas evaluated with torch ops on the device.
"""
import numpy as np
import torch

from .. import _lib
from ..devarray import DevArray, sqrt  # noqa: F401  (sqrt re-exported like the reference)
from .generic_methods_cl import GenericMethodsCL


class ParticleMethodsCL(GenericMethodsCL):

    def init_particle_methods(self):
        self.init_generic_methods()
        self._ws = {}           # persistent sort workspaces (replace the MemoryPools)
        self._pending_push = None

    # ------------------------------------------------------------------ buffers
    def _buf(self, name, n, dtype):
        """Persistent, grow-only device buffer; returns a DevArray view of n items."""
        cur = self._ws.get(name)
        if cur is None or cur.size < n or cur.dtype != np.dtype(dtype):
            cap = max(int(n * 1.1) + 16, 16)
            cur = DevArray.empty(cap, dtype, self.comm.device)
            self._ws[name] = cur
        return cur[:n]

    def _attr_names(self):
        if 'Immobile' in self.Args.keys():
            return ['x', 'y', 'z', 'w']
        return ['x', 'y', 'z', 'px', 'py', 'pz', 'w', 'g_inv']

    # ------------------------------------------------------------------ creation
    def traversal_order_valid(self, grid):
        """True if sort_indx / cell_offset still describe a permutation of the current
        storage computed on `grid` (no particles added since the last sort): they can
        then serve as the traversal order of the fused push + deposit."""
        return (self.Args['Np'] > 0 and getattr(self, '_order_np', -1) == self.Args['Np']
                and getattr(self, '_order_grid', None) is grid
                and 'sort_indx' in self.DataDev
                and self.DataDev['sort_indx'].size == self.Args['Np'])

    def graph_safe(self, grid):
        """True if this species' part of a PIC step may be captured in a CUDA graph: no
        host read-back that a later step has to wait for (the bounded cell-changer queue
        of very large species), sorted once if immobile."""
        Np = int(self.Args['Np'])
        if Np == 0:
            return True
        if 'Immobile' in self.Args.keys():
            return bool(self.flag_sorted)
        full = int(self._lib.chb_push_depose_workspace_bytes(Np))
        limit = getattr(self, '_exc_full_limit', None)
        if limit is None:
            limit = getattr(self.comm, 'device_memory_bytes', 16 << 30) // 4
        return full <= limit and getattr(self, '_exc_check', None) is None

    def after_graph_replay(self):
        """Host-side effect of a replayed step: Args['Np_stay'] resolves to the value the
        replayed scan copied to pinned memory."""
        if self.Args['Np'] == 0 or 'Immobile' in self.Args.keys() or \
                getattr(self, '_np_stay_host', None) is None:
            return
        host = self._np_stay_host
        ev = torch.cuda.Event()
        ev.record()

        def _resolve():
            ev.synchronize()
            return int(host.item())
        self.Args.set_lazy('Np_stay', _resolve)

    def exception_workspace(self):
        """(pointer, bytes) of the scratch chb_push_depose_vector / _push_index need: the
        queue of the particles that changed cell.  Sized for the worst case (one 64-byte
        record per particle) up to a quarter of the device memory; beyond that a quarter of
        the particles, with the device-side counter read back asynchronously and checked
        by Grid.finish_currents() BEFORE the deposited current is used in the same step
        (an overflow means that current is incomplete: raise)."""
        Np = int(self.Args['Np'])
        full = int(self._lib.chb_push_depose_workspace_bytes(Np))
        limit = getattr(self, '_exc_full_limit', None)
        if limit is None:
            # worst case (every particle changes cell) as long as it takes at most a quarter
            # of the device memory: 64 B per particle, 34 GB for 5.4e8 particles on 180 GB
            limit = getattr(self.comm, 'device_memory_bytes', 16 << 30) // 4
        nbytes = full if full <= limit else 16 + 64 * (Np // 4 + 4096)
        self._check_exception_overflow()
        ws = self._buf('exc_ws', (nbytes + 7) // 8, np.double)
        if nbytes < full:
            cap = (nbytes - 16) // 64
            if getattr(self, '_exc_host', None) is None:
                self._exc_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._exc_pending = (cap, ws)      # read back after the launch, see below
        return ws.ptr, nbytes

    def exception_count_readback(self):
        """Enqueue the asynchronous read-back of the cell-changer counter (only needed
        when the queue is smaller than the worst case)."""
        pend = self.__dict__.pop('_exc_pending', None)
        if pend is None:
            return
        cap, ws = pend
        self._exc_host.copy_(ws.t[:1].view(torch.int32)[:1], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._exc_check = (cap, ev)

    def _check_exception_overflow(self):
        chk = self.__dict__.pop('_exc_check', None)
        if chk is None:
            return
        cap, ev = chk
        ev.synchronize()
        n = int(self._exc_host.item()) & 0xffffffff
        if n > cap:
            raise RuntimeError("chimera_b200: %d particles changed cell in one step but the "
                               "queue holds %d; the deposited current was incomplete" % (n, cap))

    def add_new_particles(self, source=None):
        self._order_np = -1
        self._index_prefilled = False
        DataSrc = self.DataDev if source is None else source.DataDev
        for arg in self._attr_names():
            self.DataDev[arg] = DevArray(torch.cat((self.DataDev[arg].t,
                                                    DataSrc[arg + '_new'].t)))
        self.reset_num_parts()
        self.flag_sorted = False

    def make_new_domain(self, parts_in, density_profiles=None):
        """parts_in['r_shard'] = (rank, world) (optional, multi-GPU): only this rank's
        contiguous band of the domain's radial cell rows is created -- every rank gets the
        same number of particles per row, so the bands are balanced."""
        dev = self.comm.device
        xmin, xmax, rmin, rmax = [parts_in[k] for k in ('Xmin', 'Xmax', 'Rmin', 'Rmax')]
        dx, dr = self.Args['dx'], self.Args['dr']
        Nx_loc = int(np.ceil((xmax - xmin) / dx) + 1)
        Nr_loc = int(np.round((rmax - rmin) / dr) + 1)
        Xgrid = xmin + dx * np.arange(Nx_loc)
        Rgrid = rmin + dr * np.arange(Nr_loc)
        self.Args['right_lim'] = Xgrid[-1]
        if parts_in.get('r_shard') is not None:
            from ..parallel import shard_range
            j0, j1 = shard_range(Nr_loc - 1, *parts_in['r_shard'])
            Rgrid = Rgrid[j0:j1 + 1]
            Nr_loc = Rgrid.size

        npx, npr, npt = (int(v) for v in self.Args['Nppc'])
        ncx, ncr = Nx_loc - 1, Nr_loc - 1
        ncells = ncx * ncr
        theta_var = self._theta_offsets(ncells)
        self._fill_grid(theta_var, Xgrid, Rgrid, (npx, npr, npt))
        self.DataDev['w_new'] *= self.Args['w0']

        if density_profiles is not None:
            for profile in density_profiles:
                if profile['coord'] != 'x':
                    print('Only longitudinal profiling is implemented')
                    continue
                self.dens_profile(profile['points'], profile['values'],
                                  parts_in['Xmin'], parts_in['Xmax'],
                                  coord='x_new', weight='w_new')

        if 'Immobile' in self.Args.keys():
            return
        Np = ncells * npx * npr * npt
        for arg in ('px', 'py', 'pz'):
            centre = parts_in.setdefault(arg + '_c', 0)
            spread = parts_in.setdefault('d' + arg, 0)
            buf = DevArray.empty(Np, np.double, dev)
            if spread != 0:
                self._fill_arr_randn(buf, mu=centre, sigma=spread)
            else:
                buf.fill(centre)
            self.DataDev[arg + '_new'] = buf
        self._set_g_inv_new()

    def _theta_offsets(self, ncells):
        """Per-cell azimuthal offsets of the new lattice, uniform in [0, 2 pi)
        (reference :80-82, `_fill_arr_rand` on the Threefry stream)."""
        return torch.rand(ncells, dtype=torch.float64, device=self.comm.device,
                          generator=self.comm.generator) * (2 * np.pi)

    def _fill_grid(self, theta_var, Xgrid, Rgrid, nppc):
        """Regular (x, r, theta) lattice in every cell with a per-cell theta offset:
        the layout of fill_grid, reference kernels/particles_generic.cl:33-84
        (particle order inside a cell: theta slowest, then r, then x)."""
        dev = self.comm.device
        npx, npr, npt = nppc
        ncx = Xgrid.size - 1
        xg = torch.from_numpy(Xgrid).to(dev)
        rg = torch.from_numpy(Rgrid).to(dev)
        ncells = ncx * (Rgrid.size - 1)
        ic = torch.arange(ncells, device=dev)
        ir = torch.div(ic, ncx, rounding_mode='floor')
        ix = ic - ncx * ir
        x_lo, r_lo = xg[ix], rg[ir]
        Lx, Lr = xg[ix + 1] - x_lo, rg[ir + 1] - r_lo
        fx = (0.5 + torch.arange(npx, device=dev, dtype=torch.float64)) * (1. / npx)
        fr = (0.5 + torch.arange(npr, device=dev, dtype=torch.float64)) * (1. / npr)
        th = theta_var[:, None] + torch.arange(npt, device=dev, dtype=torch.float64)[None, :] \
            * (2 * np.pi / npt)
        shape = (ncells, npt, npr, npx)
        rp = (r_lo[:, None] + fr[None, :] * Lr[:, None])[:, None, :, None].expand(shape)
        xp = (x_lo[:, None] + fx[None, :] * Lx[:, None])[:, None, None, :].expand(shape)
        sin_t = torch.sin(th)[:, :, None, None].expand(shape)
        cos_t = torch.cos(th)[:, :, None, None].expand(shape)
        self.DataDev['x_new'] = DevArray(xp.reshape(-1).contiguous())
        self.DataDev['y_new'] = DevArray((rp * sin_t).reshape(-1).contiguous())
        self.DataDev['z_new'] = DevArray((rp * cos_t).reshape(-1).contiguous())
        self.DataDev['w_new'] = DevArray(rp.reshape(-1).contiguous())

    def _set_g_inv_new(self):
        D = self.DataDev
        D['g_inv_new'] = DevArray(torch.rsqrt(1 + D['px_new'].t ** 2 + D['py_new'].t ** 2
                                              + D['pz_new'].t ** 2))

    def make_new_beam(self, parts_in):
        dev = self.comm.device
        Np = int(parts_in['Np'])
        for arg in ('x', 'y', 'z'):
            buf = DevArray.empty(Np, np.double, dev)
            self._fill_arr_randn(buf, mu=parts_in[arg + '_c'], sigma=parts_in['L' + arg])
            self.DataDev[arg + '_new'] = buf
        for arg in ('px', 'py', 'pz'):
            buf = DevArray.empty(Np, np.double, dev)
            self._fill_arr_randn(buf, mu=parts_in.setdefault(arg + '_c', 0),
                                 sigma=parts_in.setdefault('d' + arg, 0))
            self.DataDev[arg + '_new'] = buf
        w = DevArray.empty(Np, np.double, dev)
        w.fill(parts_in['FullCharge'] / parts_in['Np'])
        self.DataDev['w_new'] = w
        self._set_g_inv_new()

    def dens_profile(self, x_prf, f_prf, xmin, xmax, coord='x', weight='w'):
        """Piecewise-linear density profile (reference :179-204 and
        kernels/particles_generic.cl:6-30)."""
        x_prf = np.array(x_prf, dtype=np.double)
        f_prf = np.array(f_prf, dtype=np.double)
        i_start = (x_prf < xmin).sum() - 1
        i_stop = (x_prf < xmax).sum() + 1
        dev = self.comm.device
        x_loc = torch.from_numpy(x_prf[i_start:i_stop]).to(dev)
        f_loc = torch.from_numpy(f_prf[i_start:i_stop]).to(dev)
        dxm1 = 1. / (x_loc[1:] - x_loc[:-1])
        x = self.DataDev[coord].t
        # interval ix with x_loc[ix] < x <= x_loc[ix+1]
        ix = torch.searchsorted(x_loc, x, right=False) - 1
        ix = ix.clamp_(0, x_loc.numel() - 2)
        f_minus = f_loc[ix] * dxm1[ix]
        f_plus = f_loc[ix + 1] * dxm1[ix]
        self.DataDev[weight].t.mul_(f_minus * (x_loc[ix + 1] - x) + f_plus * (x - x_loc[ix]))

    # ------------------------------------------------------------------ hot path
    def push_coords(self, mode='half'):
        if self.Args['Np'] == 0:
            return
        if 'Immobile' in self.Args.keys():
            return
        which_dt = 'dt_2' if mode == 'half' else 'dt'
        D = self.DataDev
        self._call('chb_push_xyz', D['x'].ptr, D['y'].ptr, D['z'].ptr, D['px'].ptr,
                   D['py'].ptr, D['pz'].ptr, D['g_inv'].ptr, D[which_dt].ptr,
                   int(self.Args['Np']))
        self.flag_sorted = False

    def push_and_sort(self, grid, mode='half'):
        """push_coords(mode) immediately followed by sort_parts(grid) -- the pair
        pic_loop.py:70-76 issues twice per step -- as ONE pass over the particles
        (chb_push_index computes the cell index and histogram while the pushed
        coordinates are still in registers).  Same results as the two calls."""
        if self.Args['Np'] == 0 or 'Immobile' in self.Args.keys():
            self.push_coords(mode)
            self.sort_parts(grid)
            return
        if getattr(self, '_index_prefilled', False):
            # the push happened inside chb_push_depose_push_index
            self.flag_sorted = False
            self.sort_parts(grid)
            return
        self._pending_push = 'dt_2' if mode == 'half' else 'dt'
        self.flag_sorted = False
        try:
            self.sort_parts(grid)
        finally:
            self._pending_push = None

    def prepare_index(self, grid):
        """Allocate (persistent workspaces) and zero the products of the cell index pass;
        returns (indx_in_cell, sum_in_cell)."""
        D = self.DataDev
        Np = int(self.Args['Np'])
        nbins = int(grid.Args['Nxm1Nrm1']) + 1
        D['indx_in_cell'] = self._buf('indx_in_cell', Np, np.uint32)
        D['sum_in_cell'] = self._buf('sum_in_cell', nbins, np.uint32)
        D['sum_in_cell'].t.zero_()
        return D['indx_in_cell'], D['sum_in_cell']

    def index_sort(self, grid):
        lib, st = self._lib, self._stream
        D, G = self.DataDev, grid.DataDev
        Np = int(self.Args['Np'])
        nbins = int(grid.Args['Nxm1Nrm1']) + 1
        Nx, Nr = int(grid.Args['Nx']), int(grid.Args['Nr'])

        # chb_push_depose_push_index already pushed the coordinates and filled
        # indx_in_cell / sum_in_cell (PIC_loop's one-pass particle side)
        prefilled, self._index_prefilled = getattr(self, '_index_prefilled', False), False
        self._order_np, self._order_grid = Np, grid
        if not prefilled:
            self.prepare_index(grid)
        D['cell_offset'] = self._buf('cell_offset', nbins + 1, np.uint32)
        D['sort_indx'] = self._buf('sort_indx', Np, np.uint32)
        cursor = self._buf('cursor', nbins, np.uint32)
        if 'Np_stay_dev' not in D:
            D['Np_stay_dev'] = DevArray.zeros(1, np.uint32, self.comm.device)
            self._np_stay_host = torch.zeros(1, dtype=torch.int32).pin_memory()

        geom = (G['Xmin'].ptr, G['dx_inv'].ptr, G['Rmin'].ptr, G['dr_inv'].ptr)
        which_dt, self._pending_push = self._pending_push, None
        if prefilled:
            pass
        elif which_dt is not None:
            _lib.check(lib.chb_push_index(
                D['x'].ptr, D['y'].ptr, D['z'].ptr, D['px'].ptr, D['py'].ptr, D['pz'].ptr,
                D['g_inv'].ptr, D[which_dt].ptr, D['indx_in_cell'].ptr, D['sum_in_cell'].ptr,
                Np, Nx, Nr, *geom, st), 'chb_push_index')
        else:
            _lib.check(lib.chb_index_and_sum(
                D['x'].ptr, D['y'].ptr, D['z'].ptr, D['indx_in_cell'].ptr,
                D['sum_in_cell'].ptr, Np, Nx, Nr, *geom, st), 'chb_index_and_sum')

        ws_bytes = lib.chb_cell_offsets_workspace_bytes(nbins)
        ws = self._buf('scan_ws', (ws_bytes + 3) // 4, np.uint32)
        _lib.check(lib.chb_cell_offsets(D['sum_in_cell'].ptr, nbins, D['cell_offset'].ptr,
                                        cursor.ptr, D['Np_stay_dev'].ptr, ws.ptr, ws_bytes, st),
                   'chb_cell_offsets')

        # Np_stay = cell_offset[-2]: asynchronous read-back, waited for on demand
        host = self._np_stay_host
        host.copy_(D['Np_stay_dev'].t, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()

        def _resolve():
            ev.synchronize()
            return int(host.item())
        self.Args.set_lazy('Np_stay', _resolve)

        sb = lib.chb_sort_workspace_bytes(Np, nbins)
        sws = self._buf('sort_ws', (sb + 3) // 4, np.uint32)
        _lib.check(lib.chb_sort_scatter_stable(D['indx_in_cell'].ptr, D['cell_offset'].ptr,
                                               cursor.ptr, D['sort_indx'].ptr, Np, nbins,
                                               sws.ptr, sb, st), 'chb_sort_scatter_stable')

    def align_and_damp(self, comps_align):
        Np_stay = int(self.Args['Np_stay'])
        dev = self.comm.device
        if Np_stay == 0:
            for comp in comps_align + ['sort_indx', ]:
                self.DataDev[comp] = DevArray.empty(0, self.DataDev[comp].dtype, dev)
            self.reset_num_parts()
            return
        new = [DevArray.empty(Np_stay, np.double, dev) for _ in comps_align]
        src = _lib.ptr_array([self.DataDev[c].ptr for c in comps_align])
        dst = _lib.ptr_array([a.ptr for a in new])
        sort_new = DevArray.empty(Np_stay, np.uint32, dev)
        self._call('chb_align', src, dst, len(comps_align), self.DataDev['sort_indx'].ptr,
                   Np_stay, sort_new.ptr)
        for comp, arr in zip(comps_align, new):
            self.DataDev[comp] = arr
        self.DataDev['sort_indx'] = sort_new
        self.reset_num_parts()
        # storage now IS the sorted order; cell_offset stays valid for the real cells,
        # but the trash tail is gone: cell_offset[-1] must end at the new Np
        self.DataDev['cell_offset'][-1:] = self.DataDev['cell_offset'][-2:-1]
        self._order_np = Np_stay

    def reset_num_parts(self, Np=None):
        if Np is None:
            Np = self.DataDev['x'].size
        self.DataDev['Np'].fill(Np)
        self.Args['Np'] = Np
        self.Args['Np_stay'] = Np
        if 'Np_stay_dev' in self.DataDev:
            self.DataDev['Np_stay_dev'].fill(Np)

    # ------------------------------------------------------------------ misc
    def _fill_arr_randn(self, arr, mu=0, sigma=1):
        arr.t.normal_(mean=float(mu), std=float(sigma), generator=self.comm.generator) \
            if sigma != 0 else arr.t.fill_(float(mu))

    def _fill_arr_rand(self, arr, xmin=0, xmax=1):
        arr.t.uniform_(float(xmin), float(xmax), generator=self.comm.generator)

    def _cumsum(self, arr_in, allocator=None, output_dtype=np.uint32):
        """[0, inclusive_scan(arr_in)] (reference :303-311) via chb_cell_offsets."""
        lib, n = self._lib, arr_in.size
        out = DevArray.empty(n + 1, np.uint32, self.comm.device)
        ws_bytes = lib.chb_cell_offsets_workspace_bytes(n)
        ws = self._buf('scan_ws', (ws_bytes + 3) // 4, np.uint32)
        _lib.check(lib.chb_cell_offsets(arr_in.ptr, n, out.ptr, None, None, ws.ptr,
                                        ws_bytes, self._stream), 'chb_cell_offsets')
        return out

    def free_mp(self):
        # sort temporaries live in persistent workspaces; nothing to release
        pass

    def free_added(self):
        for arg in self._attr_names():
            self.DataDev[arg + '_new'] = None
