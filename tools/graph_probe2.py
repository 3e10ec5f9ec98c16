"""cfg1 (lpa_script_small): 100 steps timed as bench.py does (CUDA events, no per-step sync):
eager, graph from the start, graph switched on late.  GPU box: python tools/graph_probe2.py"""
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
for mode in ("eager", "graph", "late"):
    c = Communicator(answers=[0, 0], seed=11)
    _, solver, eons, ions, frame, loop = mod.build(comm=c)
    loop.use_cuda_graph = mode == "graph"
    for _ in range(200):
        loop.step()
    if mode == "late":
        loop.use_cuda_graph = True
        for _ in range(4):
            loop.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(100):
        loop.step()
    e1.record()
    th = time.perf_counter() - t0
    torch.cuda.synchronize()
    tw = time.perf_counter() - t0
    print("%-6s events %.3f ms/step  host %.3f  wall %.3f  replays %d captures %d Np %d" % (
        mode, e0.elapsed_time(e1) / 100, th * 10, tw * 10, loop.graph_replays, loop.graph_captures,
        eons.Args["Np"]))
