"""Eager and CUDA-graph loops advanced in lockstep on the reduced lpa_script_small run;
prints the first steps where they differ.  GPU box: python tools/graph_lockstep.py [Nx Nr nsteps]"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from chimeracl_b200.methods.generic_methods_cl import Communicator  # noqa: E402

spec = importlib.util.spec_from_file_location("lpa_small", os.path.join(ROOT, "examples", "lpa_script_small.py"))
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)
Nx = int(sys.argv[1]) if len(sys.argv) > 1 else 300
Nr = int(sys.argv[2]) if len(sys.argv) > 2 else 48
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 45
runs = []
modes = (False, False) if os.environ.get('BOTH_EAGER') else (False, True)
for graph in modes:
    c = Communicator(answers=[0, 0], seed=11)
    _, solver, eons, ions, frame, loop = mod.build(Nx=Nx, Nr=Nr, M=1, comm=c)
    loop.use_cuda_graph = graph
    runs.append((solver, eons, ions, loop))
for it in range(nsteps):
    for r in runs:
        r[3].step()
    if it % int(os.environ.get("EVERY", "1")):
        continue
    torch.cuda.synchronize()
    out = []
    for k in ("rho_m0", "Jx_m1", "Ez_m0", "Bz_m1"):
        a, b = runs[0][0].DataDev[k].get(), runs[1][0].DataDev[k].get()
        out.append("%s %.1e" % (k, np.abs(a - b).max() / max(np.abs(a).max(), 1e-300)))
    for k in ("x", "px"):
        a, b = runs[0][1].DataDev[k].get(), runs[1][1].DataDev[k].get()
        out.append("%s %s" % (k, "%.1e" % (np.abs(a - b).max() / np.abs(a).max()) if a.shape == b.shape else "SHAPE"))
    if it in (42, 50):
        P0, P1 = runs[0][1].DataDev, runs[1][1].DataDev
        d = np.abs(P0["px"].get() - P1["px"].get())
        idx = np.argsort(d)[-5:]
        y, z = P0["y"].get()[idx], P0["z"].get()[idx]
        print("   worst px diffs", d[idx], "px", P0["px"].get()[idx], "r", np.sqrt(y * y + z * z),
              "x", P0["x"].get()[idx], "w", P0["w"].get()[idx], "n differing", int((d > 0).sum()),
              "Xmin", runs[0][0].Args["Xmin"], "dr", runs[0][0].Args["dr"])
    print("step %3d Np %d/%d stay %d/%d replays %d | %s" % (
        it, runs[0][1].Args["Np"], runs[1][1].Args["Np"], int(runs[0][1].Args["Np_stay"]),
        int(runs[1][1].Args["Np_stay"]), runs[1][3].graph_replays, "  ".join(out)))
