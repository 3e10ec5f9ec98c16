"""Sets up the bench workload (or --small) and runs a few PIC steps; meant to be run
under ncu on the GPU box:  ncu ... python tools/profile_step.py --steps 2"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--presteps", type=int, default=30,
                    help="unprofiled steps first: the lattice start (16 per cell exactly, nobody "
                         "changes cell for ~25 steps) is not the steady state")
    ap.add_argument("--align-every", type=int, default=0)
    ap.add_argument("--time", action="store_true",
                    help="no profiler: CUDA-event time of every C-ABI call and of the step")
    a = ap.parse_args()
    import torch
    from chimeracl_b200.methods.generic_methods_cl import Communicator
    from chimeracl_b200.particles import Particles
    from chimeracl_b200.solver import Solver
    from chimeracl_b200.pic_loop import PIC_loop
    comm = Communicator(answers=[0, 0], seed=1234)
    solver = Solver(dict(bench.workload(a.small)), comm)
    ecfg, icfg = bench.species_cfgs(solver.Args)
    eons, ions = Particles(ecfg, comm), Particles(icfg, comm)
    eons.make_new_domain(bench.plasma_domain(solver.Args))
    eons.add_new_particles()
    ions.add_new_particles(source=eons)
    eons.free_added()
    for p in (eons, ions):
        p.sort_parts(solver)
        p.align_parts()
    loop = PIC_loop(solvers=[solver], species=[eons, ions], align_every=a.align_every)
    for _ in range(max(1, a.presteps)):   # (the first step cannot use the one-pass particle side)
        loop.step()
    torch.cuda.synchronize()
    if a.time:
        from chimeracl_b200 import _lib
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            loop.step()
        e1.record()
        torch.cuda.synchronize()
        print("step %.3f ms" % (e0.elapsed_time(e1) / a.steps))
        lib = _lib.load()
        lib.enable_profiling()
        for _ in range(a.steps):
            loop.step()
        rep = lib.profile_report()
        for k, (n, t) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
            print("  %-32s %5.1f calls/step %8.3f ms/step" % (k, n / a.steps, t / a.steps))
        return
    torch.cuda.profiler.start()
    for _ in range(a.steps):
        loop.step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
