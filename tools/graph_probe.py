"""Per-step wall time (with a device synchronize) of the reduced / full lpa_script_small run,
eager against CUDA-graph replay.  GPU box: python tools/graph_probe.py [Nx Nr]"""
import importlib.util
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from chimeracl_b200.methods.generic_methods_cl import Communicator  # noqa: E402

spec = importlib.util.spec_from_file_location("lpa_small", os.path.join(ROOT, "examples", "lpa_script_small.py"))
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)
Nx = int(sys.argv[1]) if len(sys.argv) > 1 else 900
Nr = int(sys.argv[2]) if len(sys.argv) > 2 else 90
import numpy as np  # noqa: E402
final = []
for graph in (False, True):
    c = Communicator(answers=[0, 0], seed=11)
    _, solver, eons, ions, frame, loop = mod.build(Nx=Nx, Nr=Nr, M=1, comm=c)
    loop.use_cuda_graph = graph
    for _ in range(200):
        loop.step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(45):
        t0 = time.perf_counter()
        loop.step()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        ts.append(((t1 - t0) * 1e3, (time.perf_counter() - t0) * 1e3))
    print("graph" if graph else "eager", "captures", loop.graph_captures, "replays", loop.graph_replays,
          "Np", eons.Args["Np"])
    print("  host ms :", " ".join("%.2f" % a for a, _ in ts))
    print("  total ms:", " ".join("%.2f" % b for _, b in ts))
    final.append({k: solver.DataDev[k].get() for k in ("Ez_m0", "Ex_m1", "Bz_m1", "rho_m0")})
    final[-1].update({k: eons.DataDev[k].get() for k in ("x", "px", "g_inv")})
for k in final[0]:
    a, b = final[0][k], final[1][k]
    print("%-8s max rel diff eager/graph %.3e" % (k, np.abs(a - b).max() / np.abs(a).max()))
