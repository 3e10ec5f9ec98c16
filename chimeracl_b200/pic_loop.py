"""PIC time-step driver (API of the reference's chimeraCL/pic_loop.py).

The step sequence is the reference's (pic_loop.py:57-142); every phase only enqueues
work on the rank's CUDA stream -- there is no host synchronisation inside a step.
With timit=True the phases are bracketed by CUDA events and accumulated (in seconds)
under the reference's Timer keys.  A solver with a kr-row sharded field solve
(Solver.enable_spectral_sharding, multi-GPU) takes _deposit_and_solve_sharded: the same
phases on the owned rows plus the exchanges the sharding needs."""
import numpy as np
import torch

loop_steps = ['frame', 'push-x', 'sort', 'depose',
              'transform', 'smooth', 'data_copy',
              'grad', 'push-eb', 'damp-eb', 'restore_B',
              'gather + push-p']


class PIC_loop:
    def __init__(self, solvers=[], species=[], frames=[], diags=[], timit=False,
                 fuse_push_sort=True, real_m0_symmetry=True, align_every=None,
                 use_cuda_graph=False):
        self.solvers = solvers
        self.mainsolver = self.solvers[0]
        self.species = species
        self.frames = frames
        self.diags = diags
        self.timit = timit
        self.it = 0
        self.fuse_push_sort = bool(fuse_push_sort)
        self.real_m0_symmetry = bool(real_m0_symmetry)
        # align_every=N: every N steps the loop calls species.align_parts() itself (the
        # reference API call, particles.py:42-51, which the reference only issues on plasma
        # injection, frame.py:59).  Between aligns the storage order drifts away from the
        # cell order and every kernel that visits particles through sort_indx loses
        # coalescing (cfg3: one-pass particle kernel 1.3 -> 1.9 ms 40 steps after an align).
        # Opt-in, because align_parts() reorders the DataDev arrays and drops the
        # particles of the trash bin, as it does in the reference.
        self.align_every = int(align_every) if align_every else 0
        self.use_cuda_graph = bool(use_cuda_graph)
        import os
        self.early_rho_gather = os.environ.get('CHB_EARLY_RHO_GATHER', '1') != '0'
        self._graph = None
        self._graph_pool = None
        self._graph_stream = None
        self._graph_keep = None
        self.graph_captures = 0
        self.graph_replays = 0
        if self.timit is True:
            self.Timer = {key: 0 for key in loop_steps}
            self._events = []

    # ---- phase timer (CUDA events; resolved lazily by timer_collect)
    def timer_start(self):
        if self.timit is True:
            self._t0 = torch.cuda.Event(enable_timing=True)
            self._t0.record()

    def timer_record(self, method_str):
        if self.timit is True:
            t1 = torch.cuda.Event(enable_timing=True)
            t1.record()
            self._events.append((method_str, self._t0, t1))

    def timer_collect(self):
        if self.timit is not True:
            return {}
        torch.cuda.synchronize()
        for key, t0, t1 in self._events:
            self.Timer[key] += t0.elapsed_time(t1) * 1e-3
        self._events = []
        return self.Timer

    def step(self):
        for diag in self.diags:
            diag.make_record(self.it)

        self.timer_start()
        for frame in self.frames:
            if np.mod(self.it, frame.Args['Steps']) == 0:
                frame.shift_grids(grids=self.solvers)
                frame.inject_plasma(species=self.species, grid=self.mainsolver)
        if self.align_every and self.it > 0 and self.it % self.align_every == 0:
            for parts in self.species:
                if 'Immobile' in parts.Args.keys():
                    continue
                parts.sort_parts(grid=self.mainsolver)
                parts.align_parts()
        self.timer_record('frame')

        if self.use_cuda_graph and self._graph_eligible():
            self._step_graphed()
        else:
            self._graph = None
            self._step_body()

        for parts in self.species:
            parts.free_mp()

        self.it += 1
        return self.it

    def _step_body(self):
        """One step without the frame / diagnostics / align prologue: only enqueues device
        work (what a CUDA graph of the step captures)."""
        # first half push + sort + current deposit.  When every mobile species still
        # has the previous step's sort as a valid traversal order, the three are ONE
        # pass and the first sort of the step is not needed; the same pass applies the
        # second half push (the momenta do not change in between) and computes the cell
        # indices of the second sort (chb_push_depose_push_index), which then only has
        # its scan + scatter left.
        fuse_first = self.fuse_push_sort and len(self.solvers) == 1 and \
            hasattr(self.mainsolver, 'finish_currents') and self._can_fuse_first_half()
        if not fuse_first:
            self._push_and_sort()

        self.timer_start()
        for solver in self.solvers:
            if fuse_first:
                solver.depose_currents(species=self.species, defer=True,
                                       push_mode='half+half')
            elif hasattr(solver, 'finish_currents'):
                solver.depose_currents(species=self.species, defer=True)
            else:
                solver.depose_currents(species=self.species)
        self.timer_record('depose')

        self._push_and_sort()
        # x, y, z now hold their end-of-step values (the rest of the step only changes
        # the momenta): a caller streaming results to the host can start copying them
        hook = getattr(self, 'on_coordinates_final', None)
        if hook is not None:
            hook(self)

        for solver in self.solvers:
            if getattr(solver, 'spectral_sharding_enabled', lambda: False)():
                self._deposit_and_solve_sharded(solver)
                self.timer_start()
                solver.gather_and_push(species=self.species)
                self.timer_record('gather + push-p')
                continue
            self.timer_start()
            multi = getattr(solver.comm, 'process_group', None) is not None and \
                hasattr(solver, 'finish_charge')
            if multi:
                # the sum of J over the ranks (started right after the one-pass particle
                # side) keeps running under the scan + scatter AND the charge deposits
                solver.depose_charge(species=self.species, defer=True)
                solver.finish_currents()       # J: wait for the all-reduce, axis / dV
            else:
                if hasattr(solver, 'finish_currents'):
                    solver.finish_currents()
                solver.depose_charge(species=self.species)
            self.timer_record('depose')

            # forward transform with the spectral smoothing (reference: a separate
            # fields_smooth(['rho','Jx','Jy','Jz']) pass) folded into its last stage.
            # Multi-GPU: J is transformed while the all-reduce of rho is in flight.
            self.timer_start()
            if multi:
                solver.fb_transform(vects=['J', ], dir=0, smooth=True)
                solver.finish_charge()
                solver.fb_transform(scals=['rho', ], dir=0, smooth=True)
            else:
                solver.fb_transform(scals=['rho', ], vects=['J', ], dir=0, smooth=True)
            self.timer_record('transform')

            self.timer_start()
            for m in range(0, solver.Args['M'] + 1):
                for comp in solver.Args['vec_comps']:
                    key = comp + '_fb_m' + str(m)
                    # dN0 <- dN1: swap the buffers instead of copying them
                    # (field_grad overwrites every element of dN1 right after)
                    solver.DataDev['dN0' + key], solver.DataDev['dN1' + key] = \
                        solver.DataDev['dN1' + key], solver.DataDev['dN0' + key]
            self.timer_record('data_copy')

            # the m = 0 spectra of this loop are spectra of real grid fields: their
            # contractions run on half of the kx columns (see Transformer._m0_real)
            solver.m0_spectra_of_real_fields = self.real_m0_symmetry
            self.timer_start()
            solver.field_grad('rho', 'dN1')
            self.timer_record('grad')

            self.timer_start()
            solver.push_fields()
            self.timer_record('push-eb')

            self.timer_start()
            solver.damp_fields()
            self.timer_record('damp-eb')

            self.timer_start()
            solver.restore_B_fb()
            self.timer_record('restore_B')

            solver.m0_spectra_of_real_fields = False
            self.timer_start()
            solver.fb_transform(vects=['E', 'B'], dir=1)
            self.timer_record('transform')

            self.timer_start()
            solver.gather_and_push(species=self.species)
            self.timer_record('gather + push-p')


    # ---- CUDA-graph replay of the step (small configurations are launch-bound: a step
    # is ~60 kernel launches issued from Python; cfg1 / cfg4 spend more time enqueueing
    # than computing).  PIC_loop(use_cuda_graph=True) captures the device work of two
    # consecutive steps (the dN0 <-> dN1 buffer swap has period 2) and replays them
    # alternately as long as nothing the captured launches depend on changes: particle
    # counts, array addresses, Xmin (moving window).  An injection / align / window shift
    # changes those: that step runs eagerly, the next one re-captures.
    def _graph_eligible(self):
        if len(self.solvers) != 1 or self.timit is True or not self.fuse_push_sort:
            return False
        if getattr(self, 'on_coordinates_final', None) is not None:
            return False
        solver = self.mainsolver
        if getattr(solver.comm, 'process_group', None) is not None:
            return False
        if not hasattr(solver, 'finish_currents'):
            return False
        for parts in self.species:
            if not hasattr(parts, 'graph_safe') or not parts.graph_safe(solver):
                return False
        return self._can_fuse_first_half()

    def _graph_signature(self):
        solver = self.mainsolver
        sig = [float(solver.Args['Xmin']), int(solver.Args['Nx']), int(solver.Args['Nr']),
               len(solver.DataDev)]
        # a sample of the field arrays (replacing DataDev entries between steps is legal in
        # the reference API; code that swaps others should call invalidate_graph())
        for key in ('Ex_m0', 'Bx_m0', 'Jx_m0', 'rho_m0', 'Ex_fb_m0', 'Gx_fb_m0'):
            if key in solver.DataDev:
                sig.append(solver.DataDev[key].ptr)
        for parts in self.species:
            D = parts.DataDev
            sig += [int(parts.Args['Np'])] + [D[k].ptr for k in ('x', 'w', 'sort_indx', 'cell_offset')
                                              if k in D and D[k] is not None]
        return tuple(sig)

    def invalidate_graph(self):
        """Drop the captured step (the next eligible step runs eagerly, the one after it
        re-captures): for callers that replace device arrays the signature does not watch."""
        self._graph = None

    def _dn_swap(self):
        solver = self.mainsolver
        for m in range(0, solver.Args['M'] + 1):
            for comp in solver.Args['vec_comps']:
                key = comp + '_fb_m' + str(m)
                solver.DataDev['dN0' + key], solver.DataDev['dN1' + key] = \
                    solver.DataDev['dN1' + key], solver.DataDev['dN0' + key]

    def _dn_ptr(self):
        solver = self.mainsolver
        return solver.DataDev['dN0' + solver.Args['vec_comps'][0] + '_fb_m0'].ptr

    def _step_graphed(self):
        sig = self._graph_signature()
        G = self._graph
        if G is not None and G['sig'] == sig and self._dn_ptr() in G['dn_ptr']:
            phase = G['dn_ptr'].index(self._dn_ptr())
            G['graphs'][phase].replay()
            self._dn_swap()                       # host-side effects of the captured step
            for parts in self.species:
                if hasattr(parts, 'after_graph_replay'):
                    parts.after_graph_replay()
            self.graph_replays += 1
            return
        if G is None or G.get('pending') != sig:
            # something changed (or first use): this step runs eagerly -- it also brings
            # every persistent workspace to its final size -- and the next one captures
            self._step_body()
            self._graph = {'sig': None, 'pending': self._graph_signature()}
            return
        graphs, ptrs = [], []
        # low-level capture (torch.cuda.graph() would run gc.collect() and empty_cache()
        # on every re-capture, i.e. after every plasma injection): a side stream ordered
        # after the work already enqueued, one private memory pool for the loop's lifetime
        # (the previous pair of graphs is kept alive until the new one exists, so that the
        # pool -- released with its last graph -- can be shared)
        old = self._graph_keep
        if self._graph_stream is None:
            self._graph_stream = torch.cuda.Stream()
        self._graph_pool = old[0].pool() if old else torch.cuda.graph_pool_handle()
        cur = torch.cuda.current_stream()
        side = self._graph_stream
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):
                g = torch.cuda.CUDAGraph()
                ptrs.append(self._dn_ptr())
                g.capture_begin(pool=self._graph_pool, capture_error_mode="thread_local")
                try:
                    self._step_body()
                finally:
                    g.capture_end()
                graphs.append(g)
        cur.wait_stream(side)
        self._graph_keep = graphs
        self._graph = {'sig': sig, 'graphs': graphs, 'dn_ptr': ptrs}
        self.graph_captures += 1
        self._step_graphed()

    def _deposit_and_solve_sharded(self, solver):
        """Charge deposit + field solve of one step with the kr-row sharded spectral solve
        (Solver.enable_spectral_sharding): the same phases as in step(), each on the kr
        rows this rank owns, plus the three exchanges the sharding needs.  What overlaps:
        the J all-reduce with the charge deposits, the rho all-reduce with the forward
        transform of J, the all-gather of G with the backward transform of E, and the
        sum of the E partials with the B side (field_rot, backward transform of B)."""
        self.timer_start()
        solver.depose_charge(species=self.species, defer=True)
        solver.finish_currents()
        self.timer_record('depose')
        self._solve_fields_sharded(solver, solver.finish_charge)

    def _solve_fields_sharded(self, solver, charge_ready=None):
        """From the deposited J (and, once charge_ready() has returned, rho) grids to the
        E and B grids."""
        S = solver.shards
        self.timer_start()
        if getattr(self, 'early_rho_gather', True):
            # one component of J first (the sum of rho over the ranks is still in flight),
            # then rho, whose all-gather then runs under the other two components of J
            comps = ['J' + c for c in solver.Args['vec_comps']]
            for _ in S():
                solver.fb_transform(scals=comps[:1], dir=0, smooth=True)
            if charge_ready is not None:
                charge_ready()
            for _ in S():
                solver.fb_transform(scals=['rho', ], dir=0, smooth=True)
            rho_ready = solver.gather_spectral(['rho'])
            for _ in S():
                solver.fb_transform(scals=comps[1:], dir=0, smooth=True)
        else:
            for _ in S():
                solver.fb_transform(vects=['J', ], dir=0, smooth=True)
            if charge_ready is not None:
                charge_ready()
            for _ in S():
                solver.fb_transform(scals=['rho', ], dir=0, smooth=True)
            rho_ready = solver.gather_spectral(['rho'])
        self.timer_record('transform')

        self.timer_start()
        for m in range(0, solver.Args['M'] + 1):
            for comp in solver.Args['vec_comps']:
                key = comp + '_fb_m' + str(m)
                solver.DataDev['dN0' + key], solver.DataDev['dN1' + key] = \
                    solver.DataDev['dN1' + key], solver.DataDev['dN0' + key]
        self.timer_record('data_copy')

        solver.m0_spectra_of_real_fields = self.real_m0_symmetry
        self.timer_start()
        rho_ready.wait()
        for _ in S():
            solver.field_grad('rho', 'dN1')
        self.timer_record('grad')

        self.timer_start()
        for _ in S():
            solver.push_fields()
        self.timer_record('push-eb')

        # G first: its all-gather then also runs under the damping of E
        self.timer_start()
        for _ in S():
            solver.damp_fields(['G'])
        g_ready = solver.gather_spectral(['G' + c for c in solver.Args['vec_comps']])
        for _ in S():
            solver.damp_fields(['E'])
        self.timer_record('damp-eb')

        self.timer_start()
        for _ in S():
            solver.fb_transform(vects=['E'], dir=1, partial=True)
        e_ready = solver.reduce_grid_fields(['E'])
        self.timer_record('transform')

        self.timer_start()
        g_ready.wait()
        for _ in S():
            solver.restore_B_fb(gathered=True)
        self.timer_record('restore_B')
        solver.m0_spectra_of_real_fields = False

        self.timer_start()
        for _ in S():
            solver.fb_transform(vects=['B'], dir=1, partial=True)
        b_ready = solver.reduce_grid_fields(['B'])
        e_ready.wait()
        b_ready.wait()
        self.timer_record('transform')

    def _can_fuse_first_half(self):
        for parts in self.species:
            if 'Immobile' in parts.Args.keys() or parts.Args['Np'] == 0:
                continue
            if not (hasattr(parts, 'traversal_order_valid')
                    and parts.traversal_order_valid(self.mainsolver)):
                return False
        return True

    def _push_and_sort(self):
        for parts in self.species:
            if self.fuse_push_sort and hasattr(parts, 'push_and_sort'):
                self.timer_start()
                parts.push_and_sort(self.mainsolver, mode='half')   # one fused pass
                self.timer_record('sort')
                continue
            self.timer_start()
            parts.push_coords(mode='half')
            self.timer_record('push-x')

            self.timer_start()
            parts.sort_parts(grid=self.mainsolver)
            self.timer_record('sort')
