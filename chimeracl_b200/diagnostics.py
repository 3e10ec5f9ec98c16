"""Minimal field / species output (API of the reference's chimeraCL/diagnostics.py).

Same constructor, same `make_record(it)` hook called by PIC_loop.step(), same record
layout as reference diagnostics.py:57-141:

    /data/fields/<name>          (1 + 2M, Nr, Nx): m = 0, Re m = 1, Im m = 1, ...
    /data/species/species_<i>/<component>
    /data/info/{iteration, Xgrid, Rgrid, dx, dr, dt, Nx, Nr, M, FrameVelocity}

The reference writes HDF5 through h5py.  h5py is used when it is importable; otherwise
the identical key -> array mapping goes to `<iteration>.npz` (numpy.load gives it back
with the same path strings as keys).  This is host-side I/O around the hot path: it
only needs `.get()` on the device arrays and the solver's own transform methods.
"""
import os

import numpy as np

try:                                    # pragma: no cover - not present in the B200 image
    import h5py
except ImportError:                     # the record goes to .npz instead
    h5py = None


class Diagnostics:
    def __init__(self, configs_in, solver, species=[], frame=None,
                 path='diags', dtype_flds=np.float32, dtype_parts=np.float64):
        self.Args = configs_in
        self.solver = solver
        self.species = species
        self.frame = frame
        self.dtype_flds = dtype_flds
        self.dtype_parts = dtype_parts

        self.base_str = '/data/'
        self.flds_str = 'fields/'
        self.parts_str = 'species/'
        self.info_str = 'info/'
        self.generic_keys = ['Xgrid', 'Rgrid', 'dx', 'dr', 'dt', 'Nx', 'Nr', 'M']

        # multi-GPU (one process per rank): the fields are complete on every rank after
        # the solver's collectives, so rank 0 alone writes them (and cleans the directory);
        # the species arrays are each rank's particle shard and go to per-rank files
        comm = getattr(solver, 'comm', None)
        self.rank = int(getattr(comm, 'rank', 0))
        self.world = int(getattr(comm, 'world_size', 1))
        self._group = getattr(comm, 'process_group', None)

        self.path = os.path.join(os.getcwd(), path) + '/'
        if self.rank == 0:
            if not os.path.exists(self.path):
                os.makedirs(self.path)
            else:
                for fl in os.listdir(self.path):
                    os.remove(self.path + fl)
        self._barrier()

        self.Args.setdefault('ScalarFields', [])
        self.Args.setdefault('VectorFields', [])
        self.Args.setdefault('Species', {'Components': [], })

    def _barrier(self):
        if self._group is not None and self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=self._group)

    # ------------------------------------------------------------------ record
    def make_record(self, it):
        if np.mod(it, self.Args['Interval']) != 0:
            return
        self.record = {}
        self.record[self.base_str + self.info_str + 'iteration'] = it
        self.add_generic_info()
        # every rank takes part in the deposits / transforms (they are collective) ...
        for fld in self.Args['ScalarFields']:
            self.add_field(fld)
        for fld in self.Args['VectorFields']:
            for comp in ['x', 'y', 'z']:
                self.add_field(fld + comp)
        self.add_species()
        stem = str(it).rjust(9, '0')
        if self.rank > 0:
            # ... but only rank 0 writes the fields; the others write their species shard
            flds = self.base_str + self.flds_str
            self.record = {k: v for k, v in self.record.items() if not k.startswith(flds)}
            stem += '_rank%d' % self.rank
        self._write(stem)
        self.record = None

    def _write(self, stem):
        if h5py is not None:            # pragma: no cover
            with h5py.File(self.path + stem + '.h5', 'w') as f:
                for key, val in self.record.items():
                    f[key] = val
            return self.path + stem + '.h5'
        np.savez(self.path + stem + '.npz', **{k: np.asarray(v) for k, v in self.record.items()})
        return self.path + stem + '.npz'

    def _selection(self, part):
        """Indices of the particles passing every [component, vmin, vmax] window of
        Args['Species']['Selections'] (strict inequalities, None = open end); None when
        no selection is configured."""
        windows = self.Args['Species'].get('Selections')
        if windows is None:
            return None
        keep = np.ones(part.Args['Np'], dtype=bool)
        for name, lo, hi in windows:
            vals = part.DataDev[name].get()
            if lo is not None:
                keep &= vals > lo
            if hi is not None:
                keep &= vals < hi
        return np.flatnonzero(keep)

    def add_species(self):
        comps = self.Args['Species']['Components']
        for i, part in enumerate(self.species):
            group = '%s%sspecies_%d/' % (self.base_str, self.parts_str, i)
            picked = self._selection(part)
            for name in comps:
                if part.Args['Np'] == 0:
                    vals = np.zeros(0, dtype=self.dtype_parts)
                else:
                    host = part.DataDev[name].get() if picked is None \
                        else part.DataDev[name].map_to_host()[picked]
                    vals = host.astype(self.dtype_parts)
                if name == 'w':         # weights in pC per normalisation length
                    vals = vals * self.dtype_parts(self.Args['w2pC'])
                self.record[group + name] = vals

    def add_generic_info(self):
        h5_path = self.base_str + self.info_str
        for key in self.generic_keys:
            self.record[h5_path + key] = self.solver.Args[key]
        self.record[h5_path + 'FrameVelocity'] = \
            0. if self.frame is None else self.frame.Args['Velocity']

    def add_field(self, fld):
        """Field on the r-x grid, modes stacked as [m0, Re m1, Im m1, ...]
        (reference diagnostics.py:120-141, including its treatment of rho / J: the charge
        is re-deposited, transformed forward and smoothed before the backward transform)."""
        h5_path = self.base_str + self.flds_str + fld
        if fld == 'rho' or fld[0] == 'J':
            self.solver.depose_charge(self.species)
            self.solver.fb_transform(scals=[fld, ], dir=0)
            self.solver.fields_smooth(flds=[fld, ])
        self.solver.fb_transform(scals=[fld, ], dir=1)

        D = self.solver.DataDev
        planes = [D[fld + '_m0'].get()]
        for m in range(1, self.solver.Args['M'] + 1):
            mode = D[fld + '_m' + str(m)].get()
            planes += [mode.real, mode.imag]
        self.record[h5_path] = np.stack([p.astype(self.dtype_flds) for p in planes], axis=0)
