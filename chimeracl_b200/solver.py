"""Solver wrapper class (API of the reference's chimeraCL/solver.py)."""
import numpy as np

from .grid import Grid
from .transformer import Transformer
from .methods.solver_methods_cl import SolverMethodsCL


def psatd_coefficients(A):
    """cos(w dt), w sin(w dt), 1/w^2 per azimuthal mode (reference solver.py:41-52)."""
    dt = A['dt']
    for m in range(A['M'] + 1):
        w = A['w_m' + str(m)]
        ms = '_m' + str(m)
        A['MxSlv_cos(wdt)' + ms] = np.cos(w * dt)
        A['MxSlv_sin(wdt)*w' + ms] = np.sin(w * dt) * w
        A['MxSlv_1/w**2' + ms] = 1. / w ** 2
        A['dont_keep'] += ['MxSlv_cos(wdt)' + ms, 'MxSlv_sin(wdt)*w' + ms, 'MxSlv_1/w**2' + ms]
    return A


class _Shard:
    """kr rows [lo, hi) one (real or virtual) rank owns; `first`: its partial backward
    contractions overwrite the grid arrays (later virtual shards of a process add)."""
    __slots__ = ("lo", "hi", "first")

    def __init__(self, lo, hi, first=True):
        self.lo, self.hi, self.first = int(lo), int(hi), bool(first)


class _Done:
    def wait(self):
        pass


class _EventWait:
    """Work enqueued on the exchange stream; wait() makes the current stream wait for it."""

    def __init__(self, event):
        self.event = event

    def wait(self):
        import torch
        torch.cuda.current_stream().wait_event(self.event)


class SpectralSharding:
    """kr-row sharded field solve over the ranks of the Communicator's process group (no
    reference counterpart: the reference is single-device).  Every rank keeps full-size
    spectral arrays but only computes, and only keeps valid, the rows spectral_rows()
    assigns to it: 1/world of the forward and backward contractions, FFTs, PSATD, damping
    and grad / rot outputs.  Three exchanges per step make up for it, all on contiguous
    row blocks (no packing): gather_spectral(['rho']) before field_grad,
    gather_spectral(['Gx','Gy','Gz']) before field_rot (equal chunks of row-padded
    storage, in-place all-gather), and reduce_grid_fields(['E']), (['B']) after the
    backward transform, whose contraction over the owned kr rows leaves partial sums
    (one all-reduce per vector field, on a flat buffer holding its components and modes).
    PIC_loop.step() drives the sequence.

    emulate=True runs all `world` shards one after the other in ONE process
    (`for _ in solver.shards(): ...`): the single-GPU test of the row arithmetic."""

    def enable_spectral_sharding(self, world=None, rank=None, emulate=False):
        import torch
        from .devarray import DevArray
        from .parallel import spectral_rows
        pg = getattr(self.comm, 'process_group', None)
        if emulate:
            world = int(world)
        elif pg is not None:
            import torch.distributed as dist
            world, rank = dist.get_world_size(pg), dist.get_rank(pg)
        else:
            world, rank = 1, 0
        K = int(self.Args['Nr']) - 1
        R = spectral_rows(K, 0, world)[2]
        self._sharding = {'world': world, 'rank': rank, 'emulate': bool(emulate), 'R': R,
                          'stores': {}}
        # CHB_PEER_EXCHANGE=1 | multimem (opt-in; run on 2 and 8 B200 in round 2, parity
        # 1.5e-14 / 8.4e-14 against the replicated solve, profiles/README.md): the exchanged
        # buffers are symmetric memory and the exchanges are own kernels over NVLink peer
        # memory (csrc/peer.cu: P2P loads / stores, or NVSwitch multimem) between cross-rank
        # barriers of the symmetric-memory handles, instead of NCCL collectives
        import ctypes
        import os
        # CHB_PEER_EXCHANGE: 0 = NCCL collectives, 1 = own kernels with P2P loads / stores,
        # multimem = own kernels through the NVSwitch multicast address, auto (default)
        peer = os.environ.get('CHB_PEER_EXCHANGE', 'auto')
        if peer == 'auto':
            # measured (profiles/README.md, round 2): through the NVSwitch multicast address
            # the six exchanges of a step take 1.5 ms instead of NCCL's 2.3 ms (f64 sums run
            # NCCL's RING_LL protocol, not NVLS) on 8 GPUs, 2.87 against 3.16 ms per step;
            # on 2 GPUs NCCL is as fast.  Only used where a probe allocation shows that
            # symmetric memory with a multicast address is available; anything else: NCCL.
            peer = '0'
            if world >= 8 and not emulate and pg is not None:
                try:
                    import torch.distributed._symmetric_memory as _symm
                    probe = _symm.empty(1024, dtype=torch.float64, device=self.comm.device)
                    if int(getattr(_symm.rendezvous(probe, group=pg), 'multicast_ptr', 0) or 0):
                        peer = 'multimem'
                except Exception:
                    peer = '0'
        symm = None
        if peer != '0' and not emulate and world > 1:
            import torch.distributed._symmetric_memory as symm
            self._sharding['peer'] = {}
            # the exchanges run on their own (high-priority) stream, like NCCL's, so that
            # they overlap the compute the step enqueues before it waits for them
            self._sharding['peer_stream'] = torch.cuda.Stream(priority=-1)

        def register(key, flat):
            hdl = symm.rendezvous(flat, group=pg)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            mc = int(getattr(hdl, 'multicast_ptr', 0) or 0) if peer == 'multimem' else 0
            self._sharding['peer'][key] = (hdl, (ctypes.c_uint64 * world)(*ptrs), mc)

        # all-gathered arrays live in storage of world*R >= K rows, so that the chunks
        # are equal; DataDev keeps showing the (K, Nx) prefix under the same key
        Nx = int(self.Args['Nx'])
        for name in ['rho'] + ['G' + c for c in self.Args['vec_comps']]:
            for m in range(self.Args['M'] + 1):
                key = name + '_fb_m' + str(m)
                if symm is not None:
                    flat = symm.empty(world * R * Nx * 2, dtype=torch.float64,
                                      device=self.comm.device).zero_()
                    store = torch.view_as_complex(flat.view(world * R * Nx, 2)).view(world * R, Nx)
                    register(key, flat)
                else:
                    store = torch.zeros((world * R, Nx), dtype=torch.complex128,
                                        device=self.comm.device)
                store[:K] = self.DataDev[key].t
                self.DataDev[key] = DevArray(store[:K])
                self._sharding['stores'][key] = store
        # E and B grids: one flat buffer per vector field (like J and rho), so that the sum
        # of the partial backward transforms is ONE collective per field.  Row 0 of every
        # array rides along; the backward transform never writes it and warp_axis
        # overwrites it before the gather reads it.
        shape = (int(self.Args['Nr']), Nx)
        alloc = None
        if symm is not None:
            alloc = lambda n: symm.empty(n, dtype=torch.float64, device=self.comm.device).zero_()  # noqa: E731
        for v in ('E', 'B'):
            names = [v + c for c in self.Args['vec_comps']]
            old = {n + '_m' + str(m): self.DataDev[n + '_m' + str(m)].t
                   for n in names for m in range(self.Args['M'] + 1)}
            self._flat[v] = self._alloc_group(names, shape, alloc)
            for key, t in old.items():
                self.DataDev[key].t.copy_(t)
            if alloc is not None:
                register(v, self._flat[v])
        if alloc is not None:
            # ... and the raw J / rho deposits, whose sum over the ranks then also goes
            # through the peer-memory kernel instead of NCCL
            comps = self.Args['vec_comps']
            for group, names in (('J', ['J' + c for c in comps]), ('rho', ['rho'])):
                old = {n + '_m' + str(m): self.DataDev[n + '_m' + str(m)].t
                       for n in names for m in range(self.Args['M'] + 1)}
                self._flat[group] = self._alloc_group(names, shape, alloc)
                for key, t in old.items():
                    self.DataDev[key].t.copy_(t)
                register(group, self._flat[group])
        self._shard = None if emulate else _Shard(*spectral_rows(K, rank, world)[:2])
        return self

    def _on_exchange_stream(self, fn):
        """Run fn() on the exchange stream, ordered after everything enqueued so far on the
        current stream; returns a handle whose wait() orders the current stream after it."""
        import torch
        st = self._sharding
        side = st['peer_stream']
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(side):
            side.wait_event(ready)
            fn()
            done = torch.cuda.Event()
            done.record(side)
        return _EventWait(done)

    def peer_reduce_flat(self, key):
        """Sum over the ranks of the flat symmetric buffer `key` ('J', 'rho', 'E', 'B') with
        chb_peer_allreduce_f64, asynchronously on the exchange stream; None when the
        peer-memory exchange is off (the caller then uses NCCL)."""
        st = self.__dict__.get('_sharding')
        if st is None or 'peer' not in st or key not in st['peer']:
            return None
        hdl, ptrs, mc = st['peer'][key]
        n = self._flat[key].numel()

        def run():
            hdl.barrier(channel=0)             # every rank's partial sums are written
            self._call('chb_peer_allreduce_f64', ptrs, st['world'], st['rank'], mc, n)
            hdl.barrier(channel=1)             # every block's totals are stored
        return self._on_exchange_stream(run)

    def spectral_sharding_enabled(self):
        return self.__dict__.get('_sharding') is not None

    def shards(self):
        """Iterate over the shards this process computes: its own one (real ranks, or no
        sharding at all), or all of them in turn (emulation)."""
        from .parallel import spectral_rows
        st = self.__dict__.get('_sharding')
        if st is None or not st['emulate']:
            yield st['rank'] if st is not None else 0
            return
        K = int(self.Args['Nr']) - 1
        try:
            for r in range(st['world']):
                self._shard = _Shard(*spectral_rows(K, r, st['world'])[:2], first=(r == 0))
                yield r
        finally:
            self._shard = None

    def gather_spectral(self, names):
        """Start the all-gather of the owned kr rows of names[i]_fb_m*; .wait() on the
        result before the arrays are used as contraction sources."""
        st = self.__dict__.get('_sharding')
        if st is None or st['emulate'] or st['world'] == 1:
            return _Done()
        keys = [n + '_fb_m' + str(m) for n in names for m in range(self.Args['M'] + 1)]
        if 'peer' in st:
            # push the owned rows into every rank's array (P2P stores or one multimem.st),
            # on the exchange stream, between two cross-rank barriers: the first orders the
            # push after every rank's last use of the previous contents (write-after-read;
            # also for direct callers such as Diagnostics), the second makes all ranks'
            # rows visible
            chunk = st['R'] * int(self.Args['Nx']) * 2
            own = (self._shard.hi - self._shard.lo) * int(self.Args['Nx']) * 2
            hdl0 = st['peer'][keys[0]][0]

            def run():
                hdl0.barrier(channel=2)
                for key in keys:
                    _, ptrs, mc = st['peer'][key]
                    self._call('chb_peer_allgather_f64', ptrs, st['world'], st['rank'], mc,
                               st['rank'] * chunk, own)
                hdl0.barrier(channel=0)
            return self._on_exchange_stream(run)
        from .parallel import allgather_rows_async
        stores = [st['stores'][k] for k in keys]
        return allgather_rows_async(stores, st['rank'], self.comm.process_group) or _Done()

    def reduce_grid_fields(self, vects):
        """Start the sum over ranks of the partial backward transforms of the vector
        fields `vects` (every component / mode, one collective per field); .wait()
        before use."""
        st = self.__dict__.get('_sharding')
        if st is None or st['emulate'] or st['world'] == 1:
            return _Done()
        if 'peer' in st:
            works = [self.peer_reduce_flat(v) for v in vects]
            from .parallel import _Works
            return _Works(works)
        from .parallel import allreduce_each_async
        return allreduce_each_async([self._flat[v] for v in vects],
                                    self.comm.process_group) or _Done()


class Solver(Grid, Transformer, SolverMethodsCL, SpectralSharding):
    def __init__(self, configs_in, comm):
        self.import_comm(comm)
        self._process_configs(configs_in)
        self.Args['vec_comps'] = ['x', 'y', 'z']
        self.init_solver_methods()
        self.init_grid_methods()
        self.DataDev = {}
        self._init_grid_data_on_dev()
        self.init_transformer()
        self._make_ms_coefficients()
        self.send_args_to_dev()
        self.pad_operator_matrices()

    def push_fields(self):
        self.advance_fields(vecs=['E', 'G', 'J', 'dN0', 'dN1'])

    def damp_fields(self, vects=('E', 'G')):
        """Reference solver.py:32-35 (which always damps E and G); `vects` lets the sharded
        step damp G ahead of E."""
        vects = list(vects)
        if self.damp_fields_fused(vects):
            return
        self.fb_transform(vects=vects, dir=1, mode='half')
        self.profile_edges(vects)
        self.fb_transform(vects=vects, dir=0, mode='half')

    def restore_B_fb(self, gathered=False):
        """Reference solver.py:37-39.  On a kr-row sharded solver field_rot needs every kr
        row of G: unless the caller has gathered them already (PIC_loop does, under the
        backward transform of E), they are gathered here."""
        if not gathered:
            self.gather_spectral(['G' + c for c in self.Args['vec_comps']]).wait()
        self.field_rot('G', 'B')
        self.field_poiss_vec('B')

    def _make_ms_coefficients(self):
        psatd_coefficients(self.Args)
