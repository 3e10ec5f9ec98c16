"""Multi-GPU parity of the kr-row sharded field solve against the replicated one.

    torchrun --nproc-per-node N tools/sharded_parity.py [--full] [--steps 3]

Every rank builds the same plasma shard twice (same seed), runs `steps` PIC steps with
the replicated solve and with Solver.enable_spectral_sharding(), and rank 0 prints the
largest relative difference of the E / B grids and of the electron momenta over all
ranks (expected: rounding level, < 1e-10)."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import plasma_domain, species_cfgs, workload          # noqa: E402
from chimeracl_b200.methods.generic_methods_cl import Communicator  # noqa: E402
from chimeracl_b200.parallel import init_distributed               # noqa: E402
from chimeracl_b200.particles import Particles                     # noqa: E402
from chimeracl_b200.pic_loop import PIC_loop                       # noqa: E402
from chimeracl_b200.solver import Solver                           # noqa: E402


def run(comm, small, steps, sharded, seed):
    comm.generator.manual_seed(seed)
    solver = Solver(dict(workload(small)), comm)
    ecfg, icfg = species_cfgs(solver.Args)
    eons, ions = Particles(ecfg, comm), Particles(icfg, comm)
    ions.Args["InjectorSource"] = eons
    eons.make_new_domain(plasma_domain(solver.Args))
    eons.add_new_particles()
    ions.add_new_particles(source=eons)
    eons.free_added()
    for p in (eons, ions):
        p.sort_parts(solver)
        p.align_parts()
    if sharded:
        solver.enable_spectral_sharding()
    loop = PIC_loop(solvers=[solver], species=[eons, ions])
    for _ in range(steps):
        loop.step()
    comm.synchronize()
    out = {k: solver.DataDev[k].t.clone() for k in solver.DataDev
           if k[0] in "EB" and k[1] in "xyz" and "_fb_" not in k}
    for k in ("px", "py", "pz"):
        out[k] = eons.DataDev[k].t.clone()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="cfg3 grid instead of the 512 x 128 one")
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    comm = Communicator(answers=[0, 0])
    init_distributed(comm)
    seed = 4321 + comm.rank
    ref = run(comm, not a.full, a.steps, False, seed)
    got = run(comm, not a.full, a.steps, True, seed)
    worst = torch.zeros(1, dtype=torch.float64, device=comm.device)
    for k in ref:
        r, g = ref[k], got[k]
        rows = slice(1, None) if k[0] in "EB" else slice(None)
        scale = r[rows].abs().max().clamp_min(1e-300)
        worst = torch.maximum(worst, ((g[rows] - r[rows]).abs().max() / scale).reshape(1))
    if dist.is_initialized():
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    if comm.rank == 0:
        print("sharded vs replicated field solve, world %d, %d steps: max rel diff %.3e  %s"
              % (comm.world_size, a.steps, worst.item(), "OK" if worst.item() < 1e-10 else "FAIL"))
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if worst.item() < 1e-10 else 1)


if __name__ == "__main__":
    main()
