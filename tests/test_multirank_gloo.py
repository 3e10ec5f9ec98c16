"""world_size-2 test of the N>1 path on CPU (gloo): particles are sharded
contiguous-by-index over the ranks, every rank deposits its shard on the full grid
(here with the oracle as the per-rank depositor -- the GPU kernel is covered by the
-m gpu tests), the raw rho/J arrays are summed with chimeracl_b200.parallel.
allreduce_sum and only then post-processed.  The result must equal the
single-process deposit of all particles."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import orchestration as O
from oracle.np_kernels import NumpyKernels

from helpers import load_golden, oracle_case_from_golden, rel_err


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    from chimeracl_b200.parallel import allreduce_sum, init_distributed, shard_range
    pg = init_distributed(backend="gloo")
    M = 1
    K = NumpyKernels(M)
    G = load_golden(M)
    S, P, I = oracle_case_from_golden(G, K)
    lo, hi = shard_range(P.Args["Np"], rank, world)
    for p in (P, I):                       # this rank's shard of every species
        p.set_particles(**{a: p.D[a][lo:hi] for a in p.attrs})
        p.push_coords("half")              # electrons move off the ions: rho != 0
        p.sort_parts(S)
    names = ["rho"] + ["J" + c for c in "xyz"]
    arrays = [S.D["%s_m%d" % (n, m)] for n in names for m in range(M + 1)]
    for a in arrays:
        a[...] = 0
    flds_j = [S.D["J%s_m%d" % (c, m)] for m in range(M + 1) for c in "xyz"]
    flds_r = [S.D["rho_m%d" % m] for m in range(M + 1)]
    D = P.D
    K.depose_vector(D["sort_indx"], D["x"], D["y"], D["z"], D["px"], D["py"], D["pz"],
                    D["g_inv"], D["w"], D["cell_offset"], -1, S.Args, flds_j)
    for p, q in ((P, -1), (I, 1)):
        D = p.D
        K.depose_scalar(D["sort_indx"], D["x"], D["y"], D["z"], D["w"], D["cell_offset"], q,
                        S.Args, flds_r)
    tensors = [torch.from_numpy(a) for a in arrays]   # share memory with the arrays
    allreduce_sum(tensors, pg)
    # the asynchronous flavour PIC_loop uses for the packed J / rho buffers: two sums in
    # flight at once (J under the charge deposits, rho under the J transform), waited for
    # in the order they were started
    from chimeracl_b200.parallel import allreduce_sum_async
    flat_j = torch.full((1000,), float(rank + 1), dtype=torch.float64)
    flat_r = torch.arange(50, dtype=torch.float64) * (rank + 1)
    wj = allreduce_sum_async(flat_j, pg)
    wr = allreduce_sum_async(flat_r, pg)
    wj.wait()
    wr.wait()
    assert torch.all(flat_j == world * (world + 1) / 2)
    assert torch.equal(flat_r, torch.arange(50, dtype=torch.float64) * (world * (world + 1) / 2))
    for n in names:
        S.postproc_depose(n)
    if rank == 0:
        np.savez(os.path.join(out_dir, "dist.npz"),
                 **{"%s_m%d" % (n, m): S.D["%s_m%d" % (n, m)] for n in names for m in range(M + 1)})
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_deposit_allreduce_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "dist.npz")
    M = 1
    K = NumpyKernels(M)
    S, P, I = oracle_case_from_golden(load_golden(M), K)
    for p in (P, I):
        p.push_coords("half")
        p.sort_parts(S)
    S.depose_currents([P, I])
    S.depose_charge([P, I])
    for k in got.files:
        assert rel_err(got[k], S.D[k]) < 1e-13, k


# ------------------------------------------------------------------ Frame.right_lim over ranks
class _StubArr:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, sl):
        return _StubArr(self.a[sl])

    def get(self):
        return self.a


class _StubComm:
    def __init__(self, pg):
        self.process_group, self.device = pg, "cpu"


class _StubSpecies:
    def __init__(self, x, pg):
        self.Args = {"Nppc": (2, 2, 4), "Np": x.size, "ddx": 0.5, "right_lim": 7.0}
        self.DataDev = {"x": _StubArr(x)}
        self.comm = _StubComm(pg)


def _right_lim_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    from chimeracl_b200.parallel import init_distributed
    from chimeracl_b200.frame import Frame
    pg = init_distributed(backend="gloo")
    # rank 1 holds no particles (its radial band is empty)
    x = np.linspace(0.0, 10.0 + rank, 40) if rank != 1 else np.empty(0)
    sp = _StubSpecies(x, pg)
    Frame._update_right_lim(sp)
    # nobody has particles: the previous limit stays
    empty = _StubSpecies(np.empty(0), pg)
    Frame._update_right_lim(empty)
    np.savez(os.path.join(out_dir, "lim%d.npz" % rank), lim=sp.Args["right_lim"],
             lim_empty=empty.Args["right_lim"])
    dist.barrier()
    dist.destroy_process_group()


def test_right_lim_is_the_same_on_every_rank_even_with_an_empty_band(tmp_path):
    """a rank whose band holds no particles must not fail on max() of an
    empty array, and all ranks must continue the injection from the same x."""
    world = 3
    mp.spawn(_right_lim_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [np.load(tmp_path / ("lim%d.npz" % r)) for r in range(world)]
    for g in got:
        assert abs(float(g["lim"]) - (12.0 + 0.25)) < 1e-12      # max over ranks (rank 2: 12.0) + ddx/2
        assert float(g["lim_empty"]) == 7.0
