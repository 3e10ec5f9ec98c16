"""Timeline of one PIC step on rank 0 of a multi-GPU run (no nsys in the image: torch.profiler /
CUPTI).  torchrun --nproc-per-node N tools/trace_step.py [--weak] -> gpurun_out/trace_N.txt:
every kernel / memcpy of one steady-state step with stream, start and duration, and per
stream the busy time, so that exposed collective time can be read off."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--weak", action="store_true")
    ap.add_argument("--replicated-solve", action="store_true")
    a = ap.parse_args()
    from chimeracl_b200.methods.generic_methods_cl import Communicator
    from chimeracl_b200.parallel import init_distributed
    from chimeracl_b200.pic_loop import PIC_loop
    comm = Communicator(answers=[0, 0], seed=1234 + int(os.environ.get("RANK", "0")))
    init_distributed(comm)
    world = comm.world_size
    solver, eons, ions = bench.build_case(comm, bench.workload(False), "weak" if a.weak else "strong",
                                          1234 + comm.rank, world > 1 and not a.replicated_solve)
    loop = PIC_loop(solvers=[solver], species=[eons, ions])
    for _ in range(8):
        loop.step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            loop.step()
        torch.cuda.synchronize()
    if comm.rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda e: e.time_range.start)
        t0 = evs[0].time_range.start
        # second step only: from the first kernel after the middle
        half = [e for e in evs if e.name.startswith("void chb::depose_kernel<1, 1, 2")
                or "depose_kernel<1, true, 2" in e.name or "depose_kernel<1, 1, 2" in e.name]
        start = half[-1].time_range.start if half else t0
        out = ["# world %d, one step on rank 0 (us from the one-pass particle kernel)" % world]
        busy = {}
        last_end = start
        for e in evs:
            if e.time_range.start < start:
                continue
            s, d = e.time_range.start - start, e.time_range.end - e.time_range.start
            stream = getattr(e, "stream", None)
            if stream is None:
                stream = -1
            busy[stream] = busy.get(stream, 0.0) + d
            out.append("%9.1f %8.1f  s%-3s %s" % (s, d, stream, e.name[:90]))
            last_end = max(last_end, e.time_range.end)
        out.append("# step span %.1f us; busy per stream: %s" % (
            last_end - start, {k: round(v, 1) for k, v in busy.items()}))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "trace_%d%s.txt" % (world, "_weak" if a.weak else "")), "w") as f:
            f.write("\n".join(out) + "\n")
        print(out[-1])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
