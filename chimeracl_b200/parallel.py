"""Multi-GPU plumbing: one process per GPU (torchrun), particles sharded over ranks,
every rank holds the full grid; the exchange step of the hot path is the sum of the raw
rho / J deposits over ranks (NCCL all-reduce over NVLink), issued before the axis /
volume post-processing.  With the kr-row sharded field solve (Solver.
enable_spectral_sharding; what bench.py uses for N > 1) three more exchanges appear per step: the all-gather
of the rho spectrum (for field_grad), the all-gather of the G spectra (for field_rot)
and the sum of the partial backward contractions of E and B.  The same code runs on
gloo/CPU tensors for the world_size-2 tests."""
import os

import torch
import torch.distributed as dist


def bind_to_gpu_cpus(device):
    """Pin the calling thread (and the threads it starts later) to the CPUs NVML reports as
    local to `device`, so that host-side launches and, above all, the pinned staging buffers
    of the host-buffer path (first touch) sit on the GPU's own NUMA node / PCIe root when
    several ranks share a two-socket host.  Best effort: returns False (and changes nothing)
    if NVML, the device or the CPU set is not available (e.g. a cpuset-restricted container);
    CHB_NO_AFFINITY=1 disables it."""
    if os.environ.get("CHB_NO_AFFINITY", "0") == "1":
        return False
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device)
        try:
            handle = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(props.uuid))
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(torch.device(device).index or 0)
        before = os.sched_getaffinity(0)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        if not os.sched_getaffinity(0):
            os.sched_setaffinity(0, before)
            return False
        return True
    except Exception:
        return False


def init_distributed(comm=None, backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for a
    single process) and attach the process group to the Communicator."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None
    if comm is not None and torch.cuda.is_available():
        comm.cpu_affinity_bound = bind_to_gpu_cpus(comm.device)
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl" and comm is not None:
            kw["device_id"] = comm.device
        dist.init_process_group(backend=backend, **kw)
    pg = dist.group.WORLD
    if comm is not None:
        comm.process_group = pg
    return pg


def shard_range(n, rank, world):
    """Contiguous-by-index split of n particles: [lo, hi) of this rank."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_async(flat, group=None):
    """Start an in-place sum over ranks of one flat real tensor on the collective's own
    stream and return the work handle (None for a single rank); `handle.wait()` makes
    the caller's stream wait for it.  Lets the J all-reduce overlap the second
    push + sort of the step."""
    if group is None and not dist.is_initialized():
        return None
    if dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)


def allreduce_sum(tensors, group=None):
    """In-place sum over ranks of a list of (real or complex) tensors.  The arrays
    are packed into one flat FP64 buffer so that the exchange is a single
    collective per deposit (payload: SURVEY.md section 8e)."""
    if group is None and not dist.is_initialized():
        return
    if dist.get_world_size(group) == 1:
        return
    if len(tensors) == 1 and tensors[0].is_contiguous() and not tensors[0].is_complex():
        dist.all_reduce(tensors[0], op=dist.ReduceOp.SUM, group=group)   # in place
        return
    views = [torch.view_as_real(t).reshape(-1) if t.is_complex() else t.reshape(-1)
             for t in tensors]
    flat = torch.cat(views)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for v in views:
        v.copy_(flat[off:off + v.numel()])
        off += v.numel()


def spectral_rows(K, rank, world):
    """kr rows [lo, hi) of the K = Nr-1 spectral rows owned by `rank`, and the chunk
    height R all ranks share (a multiple of 8 -- operator-matrix slices stay 16-byte
    aligned for the 128-bit operand loads of the contraction kernel; the last ranks may
    own fewer than R rows, or none)."""
    K, rank, world = int(K), int(rank), int(world)
    R = (-(-K // world) + 7) // 8 * 8
    lo = min(K, rank * R)
    return lo, min(K, lo + R), R


class _Works:
    """Several async collectives waited for as one."""

    def __init__(self, works):
        self.works = [w for w in works if w is not None]

    def wait(self):
        for w in self.works:
            w.wait()


def allgather_rows_async(stores, rank, group=None):
    """In-place all-gather of equally sized row chunks: every tensor of `stores` is a
    contiguous (world*R, ...) array whose chunk [rank*R, (rank+1)*R) holds this rank's
    rows; afterwards all chunks are valid on all ranks.  One collective per array, all in
    flight together; returns one handle (None for a single rank)."""
    if group is None and not dist.is_initialized():
        return None
    world = dist.get_world_size(group)
    if world == 1:
        return None
    works = []
    for st in stores:
        flat = (torch.view_as_real(st) if st.is_complex() else st).reshape(-1)
        n = flat.numel() // world
        works.append(dist.all_gather_into_tensor(flat, flat[rank * n:(rank + 1) * n],
                                                 group=group, async_op=True))
    return _Works(works)


def allreduce_each_async(tensors, group=None):
    """Sum over ranks of every (contiguous) tensor in place, one async collective each;
    returns one handle (None for a single rank)."""
    if group is None and not dist.is_initialized():
        return None
    if dist.get_world_size(group) == 1:
        return None
    works = []
    for t in tensors:
        flat = (torch.view_as_real(t) if t.is_complex() else t).reshape(-1)
        works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True))
    return _Works(works)
