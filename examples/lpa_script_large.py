"""BASELINE configs[3]/[4]: the reference's examples/lpa_script_large.py (Nx=4096, Nr=252,
M=1, moving window, 6-point density profile) with only the imports changed and the HDF5
diagnostics dropped; `--cfg5` switches to the scaled shape of configs[4] (Nx=16384,
Nr=1024, M=2, 32 ppc, plasma pre-filled over the box, particles sharded over the ranks
by x-slabs).  Usage:  [torchrun --nproc-per-node N] python examples/lpa_script_large.py
[--cfg5] [--steps K]"""
import argparse
import os
import sys
from copy import deepcopy
from time import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chimeracl_b200.methods.generic_methods_cl import Communicator
from chimeracl_b200.particles import Particles
from chimeracl_b200.solver import Solver
from chimeracl_b200.frame import Frame
from chimeracl_b200.laser import add_gausian_pulse
from chimeracl_b200.pic_loop import PIC_loop
from chimeracl_b200.parallel import init_distributed


def build(comm, cfg5=False, replicated_solve=False):
    rank, world = comm.rank, comm.world_size
    xmin, xmax, rmax = -100., 40., 50.
    if cfg5:
        Nx, Nr, M, nppc = 16384, 1024, 2, (2, 4, 4)
    else:
        Nx, Nr, M, nppc = 4096, 252, 1, (2, 2, 4)
    a0, Lx, w0, x0, x_foc = 5, 10., 16., 0., 100.
    dens = 0.5e18 / 1.1e21
    dens_profiles = [{'coord': 'x', 'points': [-200, 40, 140, 340, 440, 1000],
                      'values': [0, 0, 2, 2, 1, 1]}, ]
    grid_in = {'Xmin': xmin, 'Xmax': xmax, 'Nx': Nx, 'Rmin': 0., 'Rmax': rmax, 'Nr': Nr,
               'M': M, 'DampCells': 50}
    grid_in['dt'] = (xmax - xmin) / Nx
    solver = Solver(grid_in, comm)
    add_gausian_pulse(solver, {'k0': 1., 'a0': a0, 'x0': x0, 'Lx': Lx, 'R': w0, 'x_foc': x_foc})

    eons_in = {'Nppc': nppc, 'dx': solver.Args['dx'], 'dr': solver.Args['dr'],
               'dt': solver.Args['dt'], 'dens': dens, 'charge': -1}
    ions_in = deepcopy(eons_in)
    ions_in['charge'] = 1
    ions_in['Immobile'] = True
    eons, ions = Particles(eons_in, comm), Particles(ions_in, comm)
    ions.Args['InjectorSource'] = eons

    frames = []
    if cfg5:
        # plasma pre-filled over the box; every rank fills its own x-slab
        A = solver.Args
        ncx = A['Nx'] - 3
        lo, hi = ncx * rank // world, ncx * (rank + 1) // world
        dom = {'Xmin': A['Xmin'] + A['dx'] * (1 + lo),
               'Xmax': A['Xmin'] + A['dx'] * (1 + hi - 0.5),
               'Rmin': 0.0, 'Rmax': (A['Nr'] - 2) * A['dr']}
        eons.make_new_domain(dom)
        eons.add_new_particles()
        ions.add_new_particles(source=eons)
        eons.free_added()
        for p in (eons, ions):
            p.sort_parts(solver)
            p.align_parts()
    else:
        frames = [Frame({'Velocity': 1., 'dt': solver.Args['dt'], 'Steps': 20,
                         'DensityProfiles': dens_profiles})]
    if world > 1 and not replicated_solve:
        # after the laser initialiser: the spectra it wrote are complete on every rank
        solver.enable_spectral_sharding()
    loop = PIC_loop(solvers=[solver, ], species=[eons, ions], frames=frames, diags=[])
    return solver, eons, ions, loop


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg5", action="store_true")
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--replicated-solve", action="store_true",
                    help="multi-GPU: every rank runs the whole field solve (default: kr-row sharded)")
    ap.add_argument("--sharded-solve", action="store_true", help="the default, spelled out")
    a = ap.parse_args()
    comm = Communicator(answers=[0, 0])
    init_distributed(comm)
    rank, world = comm.rank, comm.world_size
    solver, eons, ions, loop = build(comm, a.cfg5, a.replicated_solve)
    Nx, Nr, M = solver.Args['Nx'], solver.Args['Nr'], solver.Args['M']

    for _ in range(3):
        loop.step()
    comm.synchronize()
    t0 = time()
    for _ in range(a.steps):
        loop.step()
    comm.synchronize()
    dt_ms = (time() - t0) / a.steps * 1e3
    n_loc = int(eons.Args['Np'])
    ok = all(bool(torch.isfinite(solver.DataDev[k].t.abs().sum()).item())
             for k in ('Ez_m0', 'rho_m0', 'Ex_m%d' % M))
    print("rank %d/%d  Nx=%d Nr=%d M=%d  electrons(rank)=%d  %.2f ms/step  %.3g particle-steps/s/GPU  "
          "finite=%s  free=%.1f GB" % (rank, world, Nx, Nr, M, n_loc, dt_ms, n_loc / dt_ms * 1e3, ok,
                                       torch.cuda.mem_get_info()[0] / 1e9))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
