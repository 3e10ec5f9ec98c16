"""What the exchanges of the sharded step cost on this box, alone and under a concurrent
FP64 matmul (torchrun, one rank per GPU):  all-reduce (J 144 MB, E/B 150 MB, rho 48 MB),
all-gather (G 201 MB), reduce-scatter and all-to-all of the same volumes (the transposed
layout), in f64.  Prints algorithm bandwidth = bytes of the full array / time."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    dev = torch.device("cuda")
    a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    side = torch.cuda.Stream()

    def timed(fn, n=10, busy=False):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if busy:
            with torch.cuda.stream(side):
                for _ in range(4):
                    torch.matmul(a, a)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    rows = []
    for mb in (48, 144, 201, 300):
        n = mb * 1000 * 1000 // 8 // world * world
        x = torch.randn(n, dtype=torch.float64, device=dev)
        out = torch.empty_like(x)
        chunk = x[: n // world].clone()
        for busy in (False, True):
            rows.append(("all_reduce", mb, busy, timed(lambda: dist.all_reduce(x), busy=busy)))
            rows.append(("all_gather", mb, busy,
                         timed(lambda: dist.all_gather_into_tensor(out, chunk), busy=busy)))
            rows.append(("reduce_scatter", mb, busy,
                         timed(lambda: dist.reduce_scatter_tensor(chunk, x), busy=busy)))
            rows.append(("all_to_all", mb, busy,
                         timed(lambda: dist.all_to_all_single(out, x), busy=busy)))
        x.normal_()
    if rank == 0:
        print("world %d" % world)
        for name, mb, busy, ms in rows:
            print("%-15s %4d MB  %s  %.3f ms  %.0f GB/s" % (name, mb, "under dgemm" if busy else "alone      ",
                                                          ms, mb / ms))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
