"""Moving window + plasma injector (API of the reference's chimeraCL/frame.py)."""
import numpy as np


class Frame():
    def __init__(self, configs_in, comm=None):
        self._process_configs(configs_in)

    def _process_configs(self, configs_in):
        self.Args = configs_in
        for key, default in (('Steps', 1.), ('Velocity', 0.), ('dt', 1),
                             ('DensityProfiles', None)):
            if key not in self.Args:
                self.Args[key] = default

    def _shift(self, steps):
        if steps is None:
            steps = self.Args['Steps']
        return steps * self.Args['dt'] * self.Args['Velocity']

    def shift_grids(self, grids=[], steps=None):
        """Move Xmin/Xmax/Xgrid on host AND device (reference frame.py:22-30)."""
        x_shift = self._shift(steps)
        for grid in grids:
            for store in (grid.Args, grid.DataDev):
                for arg in ('Xmax', 'Xmin', 'Xgrid'):
                    store[arg] += x_shift

    def inject_plasma(self, species, grid, steps=None):
        """New plasma slab at the right edge, appended, sorted and aligned
        (reference frame.py:32-64)."""
        x_shift = self._shift(steps)
        for specie in species:
            if specie.Args['Np'] == 0:
                specie.Args['right_lim'] = grid.Args['Xmax'] - x_shift
            left = specie.Args['right_lim']
            domain = {'Xmin': left, 'Xmax': left + x_shift,
                      'Rmin': grid.Args['Rmin'] * (grid.Args['Rmin'] > 0),
                      'Rmax': grid.Args['Rmax']}
            # multi-GPU: the new slab is dealt to the ranks by bands of radial cell rows
            # (equal particle counts; the deposits are summed over ranks anyway)
            pg = getattr(getattr(specie, 'comm', None), 'process_group', None)
            if pg is not None:
                import torch.distributed as dist
                if dist.get_world_size(pg) > 1:
                    domain['r_shard'] = (dist.get_rank(pg), dist.get_world_size(pg))
            specie.make_new_domain(domain, density_profiles=self.Args['DensityProfiles'])
            if 'InjectorSource' in specie.Args.keys():
                specie.add_new_particles(specie.Args['InjectorSource'])
            else:
                specie.add_new_particles()

        for specie in species:
            specie.free_added()
            specie.sort_parts(grid=grid)
            specie.align_parts()
            self._update_right_lim(specie)

    @staticmethod
    def _update_right_lim(specie):
        """right_lim = largest x of the last Nppc+1 (aligned) particles + ddx/2 (reference
        frame.py:61-64), made safe and consistent for multi-GPU runs."""
        Num_ppc = np.int32(np.prod(specie.Args['Nppc']) + 1)
        tail = specie.DataDev['x'][-Num_ppc:].get() if specie.Args['Np'] > 0 else np.empty(0)
        x_max = float(tail.max()) if tail.size else -np.inf
        # multi-GPU: a rank whose radial band is empty (more ranks than cell rows, or
        # everything of it left the box) has no particles to look at, and the next slab
        # must start at the same x on every rank: largest x_max over the ranks
        pg = getattr(getattr(specie, 'comm', None), 'process_group', None)
        if pg is not None:
            import torch
            import torch.distributed as dist
            if dist.get_world_size(pg) > 1:
                dev = specie.comm.device if dist.get_backend(pg) == 'nccl' else 'cpu'
                t = torch.tensor([x_max], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=pg)
                x_max = float(t.item())
        if np.isfinite(x_max):
            specie.Args['right_lim'] = x_max + 0.5 * specie.Args['ddx']
        # else: nothing anywhere -- keep the lattice edge make_new_domain recorded
