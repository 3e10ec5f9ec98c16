"""Times the kr-row sharded field solve of cfg3 (Nx=4096, Nr=512, M=1) with all 8 shards
run one after the other on ONE GPU (Solver.enable_spectral_sharding(emulate=True)), i.e.
8x the per-rank solve work of an 8-GPU run without its exchanges, and prints the time per
C-ABI entry point.  Run twice, with CHB_DHT_TILE64=0 / 1, to compare the 128-row and
64-row contraction tiles on the 64-row outputs the sharding produces."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from chimeracl_b200 import _lib                                      # noqa: E402
from chimeracl_b200.methods.generic_methods_cl import Communicator   # noqa: E402
from chimeracl_b200.solver import Solver                             # noqa: E402
from bench import workload                                           # noqa: E402
from test_sharded_solve_cpu import _Loop, _random_state              # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
comm = Communicator(answers=[0, 0])
S = Solver(dict(workload(False)), comm)
_random_state(S, 5)
if world > 0:
    S.enable_spectral_sharding(world=world, emulate=True)
loop = _Loop()
run = (lambda: loop.solve_sharded(S)) if world > 0 else (lambda: loop.solve_plain(S))
for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    run()
e1.record()
torch.cuda.synchronize()
total = e0.elapsed_time(e1) / n
lib = _lib.load()
lib.enable_profiling()
for _ in range(n):
    run()
rep = lib.profile_report()
lib.disable_profiling()
print("tile64=%s world=%d: %.3f ms per solve (all shards)" % (os.environ.get("CHB_DHT_TILE64", "auto"), world, total))
for k, (c, t) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    print("  %-28s %5.1f calls  %.3f ms" % (k, c / n, t / n))
