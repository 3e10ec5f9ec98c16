"""Fourier-Bessel round-trip check -- the reference's examples/test_transformer.py
with only the import lines changed."""
import sys
import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from chimeracl_b200.methods.generic_methods_cl import Communicator
from chimeracl_b200.particles import Particles
from chimeracl_b200.solver import Solver


def run_test(answers=[0, 0], verb=False):
    comm = Communicator(answers=answers)
    grid_in = {'Xmin': -1., 'Xmax': 1., 'Nx': 1024, 'Rmin': 0, 'Rmax': 1., 'Nr': 200, 'M': 1}
    parts = Particles(grid_in, comm)
    grid = Solver(grid_in, comm)
    beam_in = {'Np': int(7 * 10 ** 6), 'FullCharge': 1, 'x_c': 0., 'Lx': 0.3, 'y_c': 0.2,
               'Ly': 0.3, 'z_c': 0.2, 'Lz': 0.3}
    parts.add_particles(beam_in=beam_in)
    parts.sort_parts(grid=grid)
    parts.align_parts()
    grid.depose_charge([parts, ])
    tmp0 = grid.DataDev['rho_m0'].get().copy()
    tmp1 = grid.DataDev['rho_m1'].get().copy()
    grid.fb_transform(scals=['rho', ], dir=0)
    grid.set_to(grid.DataDev['rho_m0'], 0)
    grid.set_to(grid.DataDev['rho_m1'], 0)
    grid.fb_transform(scals=['rho', ], dir=1)
    err_xr = (np.abs(grid.DataDev['rho_m0'].get() - tmp0)[1:] / np.abs(tmp0[1:]).max() +
              np.abs(grid.DataDev['rho_m1'].get() - tmp1)[1:] / np.abs(tmp1[1:]).max()).max()
    comm.thr.synchronize()
    if verb:
        print("Error in transform is {:g}".format(err_xr))
    return err_xr


if __name__ == "__main__":
    run_test(verb=True)
