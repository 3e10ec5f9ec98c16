"""Experiment: how much do the sort_indx-indirected kernels depend on storage order?
Times depose_currents / depose_charge / gather_and_push right after align_parts()
(sort_indx == identity) and after k PIC steps without re-aligning."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from chimeracl_b200.methods.generic_methods_cl import Communicator
from chimeracl_b200.particles import Particles
from chimeracl_b200.solver import Solver
from chimeracl_b200.pic_loop import PIC_loop

comm = Communicator(answers=[0, 0], seed=1234)
solver = Solver(dict(bench.workload(False)), comm)
ecfg, icfg = bench.species_cfgs(solver.Args)
eons, ions = Particles(ecfg, comm), Particles(icfg, comm)
eons.make_new_domain(bench.plasma_domain(solver.Args)); eons.add_new_particles()
ions.add_new_particles(source=eons); eons.free_added()
for p in (eons, ions):
    p.sort_parts(solver); p.align_parts()
loop = PIC_loop(solvers=[solver], species=[eons, ions])

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def report(tag):
    for p in (eons, ions):
        p.flag_sorted = False; p.sort_parts(solver)
    srt = eons.DataDev['sort_indx'].t.long()
    disp = (srt - torch.arange(srt.numel(), device=srt.device)).abs().double()
    print("%-28s J %.3f ms  rho(2 species) %.3f ms  gather %.3f ms | displaced %.1f%% mean|d| %.1f" % (
        tag, timeit(lambda: solver.depose_currents([eons, ions])),
        timeit(lambda: solver.depose_charge([eons, ions])),
        timeit(lambda: solver.gather_and_push([eons, ions])),
        100 * (disp > 0).double().mean().item(), disp.mean().item()))

report("aligned (identity)")
for k in (1, 4, 15):
    for _ in range(k): loop.step()
    report("after +%d steps" % k)
eons.align_parts(); ions.align_parts()
report("re-aligned")
