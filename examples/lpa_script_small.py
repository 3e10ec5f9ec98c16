"""Laser-wakefield run with moving window -- the reference's examples/lpa_script_small.py
with only the import lines changed (chimeraCL -> chimeracl_b200) and diagnostics
dropped (HDF5 output is out of scope here).  Usage: python examples/lpa_script_small.py [Nsteps]"""
import sys
from time import time
from copy import deepcopy
import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from chimeracl_b200.methods.generic_methods_cl import Communicator
from chimeracl_b200.particles import Particles
from chimeracl_b200.solver import Solver
from chimeracl_b200.frame import Frame
from chimeracl_b200.laser import add_gausian_pulse
from chimeracl_b200.pic_loop import PIC_loop


def build(Nx=900, Nr=90, M=1, comm=None, Lx=10., x0=0., profile_start=43.1):
    xmin, xmax = -43., 43.
    rmin, rmax = 0., 36.
    a0 = 3
    w0 = 12.
    x_foc = 100.
    dens = 7e18 / (1.1e21 / 0.8 ** 2)
    Npx, Npr, Npth = 2, 2, 4
    frame_velocity = 1.
    frameSteps = 20
    dens_profiles = [{'coord': 'x', 'points': [-100, profile_start, 90, 5e5], 'values': [0, 0, 1, 1]}, ]

    comm = comm or Communicator(answers=[0, 0])
    grid_in = {'Xmin': xmin, 'Xmax': xmax, 'Nx': Nx, 'Rmin': rmin, 'Rmax': rmax, 'Nr': Nr,
               'M': M, 'DampCells': 50}
    laser_in = {'k0': 1., 'a0': a0, 'x0': x0, 'Lx': Lx, 'R': w0, 'x_foc': x_foc}
    grid_in['dt'] = (grid_in['Xmax'] - grid_in['Xmin']) / grid_in['Nx']

    solver = Solver(grid_in, comm)
    add_gausian_pulse(solver, laser=laser_in)

    eons_in = {'Nppc': (Npx, Npr, Npth), 'dx': solver.Args['dx'], 'dr': solver.Args['dr'],
               'dt': solver.Args['dt'], 'dens': dens, 'charge': -1}
    ions_in = deepcopy(eons_in)
    ions_in['charge'] = 1
    ions_in['Immobile'] = True
    eons = Particles(eons_in, comm)
    ions = Particles(ions_in, comm)
    ions.Args['InjectorSource'] = eons

    frame = Frame({'Velocity': frame_velocity, 'dt': solver.Args['dt'], 'Steps': frameSteps,
                   'DensityProfiles': dens_profiles})
    loop = PIC_loop(solvers=[solver, ], species=[eons, ions], frames=[frame, ], diags=[])
    return comm, solver, eons, ions, frame, loop


if __name__ == "__main__":
    Nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    comm, solver, eons, ions, frame, loop = build()
    t0 = time()
    while loop.it < Nsteps + 1:
        loop.step()
        if np.mod(loop.it, 10) == 0:
            sys.stdout.write("\rstep {:d} of {:d}".format(loop.it, Nsteps))
            sys.stdout.flush()
    comm.queue.finish()
    t0 = time() - t0
    print("\nTotal time is {:g} mins \nMean step time is {:g} ms, {:d} electrons".format(
        t0 / 60., t0 / Nsteps * 1e3, eons.Args['Np']))
