"""CPU tests of the HOST orchestration of the full PIC step (PIC_loop.step(): one-pass
particle side, sorts, deposits, field solve, gather) with every C-ABI call served by
tests/cabi_emulator.py (oracle kernels / NumPy on host memory; test infrastructure only):
the real wrapper classes and mixins run unchanged.

  * two steps against the golden vectors generated from oracle/_ref (the second step
    takes the fused half push + deposit + half push + index path),
  * two ranks over gloo (particles split by index, rho / J all-reduced, kr-row sharded
    field solve -- the N > 1 configuration of bench.py) against the single-process run.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ATTR, golden_cfgs, load_golden, rel_err
import cabi_emulator as emu


def _set_particles(P, arrays):
    from chimeracl_b200.devarray import DevArray
    for a in P._attr_names():
        P.DataDev[a] = DevArray.from_numpy(np.ascontiguousarray(arrays[a], dtype=np.float64),
                                           P.comm.device)
    P.reset_num_parts()
    P.flag_sorted = False


def _interior(G, margin=4):
    """Indices of the golden particles at least `margin` cells away from every edge of the
    valid region, so that none of them reaches the trash bin within a few steps.  (The
    gather gates on the STORAGE index, sort_indx[ip] < Np_stay -- a reference quirk the
    kernels reproduce -- so with particles in the trash bin the result depends on how the
    particles are partitioned; real runs align them away.)"""
    cfg, _ = golden_cfgs(G)
    Nx, Nr = cfg['Nx'], cfg['Nr']
    dx = (cfg['Xmax'] - cfg['Xmin']) / (Nx - 1)
    dr = cfg['Rmax'] / (Nr - 1.5)
    ix = np.floor((G['in/P/x'] - cfg['Xmin']) / dx)
    ir = np.floor((np.hypot(G['in/P/y'], G['in/P/z']) + 0.5 * dr) / dr)
    return np.flatnonzero((ix >= margin) & (ix < Nx - 2 - margin) & (ir < Nr - 2 - margin))


def _case(G, comm, lo=None, hi=None, keep=None):
    cfg, pcfg = golden_cfgs(G)
    S = emu.make_solver(cfg, comm=comm)
    for k in G.files:
        if k.startswith("in/S/"):
            S.DataDev[k[5:]][:] = G[k]
    P = emu.make_particles(pcfg, comm)
    I = emu.make_particles(dict(pcfg, charge=1, Immobile=True), comm)
    sl = slice(lo, hi) if keep is None else keep[lo:hi]
    _set_particles(P, {a: G["in/P/" + a][sl] for a in ATTR})
    _set_particles(I, {a: G["in/P/" + a][sl] for a in ("x", "y", "z", "w")})
    return S, P, I


@pytest.mark.parametrize("M", [0, 1])
def test_host_pic_steps_match_golden(monkeypatch, M):
    from chimeracl_b200.pic_loop import PIC_loop
    emu.patch_cuda_host_calls(monkeypatch)
    G = load_golden(M)
    S, P, I = _case(G, emu.EmulatedComm())
    loop = PIC_loop(solvers=[S], species=[P, I], frames=[], diags=[])
    loop.step()
    n = 0
    for k in G.files:
        if k.startswith("step1/S/"):
            assert rel_err(S.DataDev[k[8:]].get(), G[k]) < 1e-10, k
            n += 1
        elif k.startswith("step1/P/"):
            assert rel_err(P.DataDev[k[8:]].get(), G[k]) < 1e-10, k
            n += 1
    assert n > 20
    assert P.traversal_order_valid(S)            # the next step takes the one-pass path
    loop.step()
    P.align_parts()
    for k in G.files:
        if k.startswith("step2_aligned/S/"):
            assert rel_err(S.DataDev[k[16:]].get(), G[k]) < 1e-10, k
        elif k.startswith("step2_aligned/P/"):
            name = k[16:]
            got = P.DataDev[name].get()
            if name == "sort_indx":
                assert np.array_equal(got, G[k])
            else:
                assert rel_err(got, G[k]) < 1e-10, k


def test_host_align_every_matches_oracle(monkeypatch):
    """Host logic of PIC_loop(align_every=2) over the emulated C ABI: the loop's own
    sort_parts + align_parts calls at steps 2 and 4 against the oracle given the same calls."""
    from chimeracl_b200.pic_loop import PIC_loop
    from oracle import orchestration as O
    from oracle.np_kernels import NumpyKernels
    from helpers import oracle_case_from_golden
    emu.patch_cuda_host_calls(monkeypatch)
    G = load_golden(1)
    S, P, I = _case(G, emu.EmulatedComm())
    loop = PIC_loop(solvers=[S], species=[P, I], align_every=2)
    So, Po, Io = oracle_case_from_golden(G, NumpyKernels(1))
    for it in range(5):
        loop.step()
        if it > 0 and it % 2 == 0:
            Po.sort_parts(So)
            Po.align_parts()
        O.pic_step(So, [Po, Io])
    assert int(P.Args["Np"]) == Po.Args["Np"] < G["in/P/x"].size     # the trash bin was dropped
    assert np.array_equal(P.DataDev["sort_indx"].get(), Po.D["sort_indx"])
    for k in ("Ex_m0", "Bz_m1", "rho_m0", "Jx_m1"):
        assert rel_err(S.DataDev[k].get(), So.D[k]) < 1e-10, k
    for k in ("x", "px", "g_inv", "w"):
        assert rel_err(P.DataDev[k].get(), Po.D[k]) < 1e-10, k


# ------------------------------------------------------------------ two ranks over gloo
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, sharded):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    torch.cuda.Event = emu._HostEvent
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    from chimeracl_b200.parallel import init_distributed, shard_range
    from chimeracl_b200.pic_loop import PIC_loop
    pg = init_distributed(backend="gloo")
    G = load_golden(1)
    keep = _interior(G)
    lo, hi = shard_range(keep.size, rank, world)
    S, P, I = _case(G, emu.EmulatedComm(pg), lo, hi, keep)
    if sharded:
        S.enable_spectral_sharding()
    loop = PIC_loop(solvers=[S], species=[P, I], frames=[], diags=[])
    for _ in range(3):
        loop.step()
    assert int(P.Args["Np_stay"]) == int(P.Args["Np"])
    out = {k: S.DataDev[k].get() for k in S.DataDev
           if k[0] in "EBJr" and "_fb_" not in k and k.split("_m")[0] in
           ("Ex", "Ey", "Ez", "Bx", "By", "Bz", "Jx", "Jy", "Jz", "rho")}
    for a in ("x", "px", "py", "pz", "g_inv"):
        out["P/" + a] = P.DataDev[a].get()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lo=lo, hi=hi, **out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("sharded", [False, True])
def test_two_rank_pic_steps_equal_single_process(tmp_path, monkeypatch, sharded):
    from chimeracl_b200.pic_loop import PIC_loop
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), sharded), nprocs=world,
             join=True)
    emu.patch_cuda_host_calls(monkeypatch)
    G = load_golden(1)
    S, P, I = _case(G, emu.EmulatedComm(), keep=_interior(G))
    loop = PIC_loop(solvers=[S], species=[P, I], frames=[], diags=[])
    for _ in range(3):
        loop.step()
    assert int(P.Args["Np_stay"]) == int(P.Args["Np"]) > 500
    for rank in range(world):
        got = np.load(tmp_path / ("rank%d.npz" % rank))
        lo, hi = int(got["lo"]), int(got["hi"])
        for k in got.files:
            if k in ("lo", "hi"):
                continue
            if k.startswith("P/"):
                # no particle leaves the box in these steps: storage order is kept
                assert rel_err(got[k], P.DataDev[k[2:]].get()[lo:hi]) < 1e-10, (rank, k)
            else:
                assert rel_err(got[k][1:], S.DataDev[k].get()[1:]) < 1e-10, (rank, k)


# ------------------------------------------------------------------ moving window over ranks
def _lwfa(comm, steps):
    import importlib.util
    from chimeracl_b200 import _lib as real_lib
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "examples", "lpa_script_small.py")
    spec = importlib.util.spec_from_file_location("lpa_small", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    real_lib._lib = comm.lib                 # the wrapper classes bind _lib.load()
    _, solver, eons, ions, frame, loop = mod.build(Nx=64, Nr=20, M=1, comm=comm)
    for _ in range(steps):
        loop.step()
    # the plasma is neutral (ions sit on the electrons): look at the electron charge alone
    solver.depose_charge(species=[eons])
    return solver, eons, ions


def _lwfa_worker(rank, world, port, out_dir, steps):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    torch.cuda.Event = emu._HostEvent
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    from chimeracl_b200.parallel import init_distributed
    pg = init_distributed(backend="gloo")
    solver, eons, ions = _lwfa(emu.EmulatedComm(pg), steps)
    np.savez(os.path.join(out_dir, "lwfa%d.npz" % rank), rho=solver.DataDev["rho_m0"].get(),
             ne=int(eons.Args["Np"]), ni=int(ions.Args["Np"]),
             lim=float(eons.Args["right_lim"]), xmin=float(solver.Args["Xmin"]),
             r_max=float(np.hypot(eons.DataDev["y"].get(), eons.DataDev["z"].get()).max()))
    dist.barrier()
    dist.destroy_process_group()


def test_moving_window_injection_over_two_ranks(tmp_path, monkeypatch):
    """examples/lpa_script_small.py (reduced grid) for 21 steps = two injections, on two
    gloo ranks: every new slab is dealt to the ranks by radial bands (Frame.inject_plasma ->
    make_new_domain(r_shard=...)).  Particle counts add up to the single-process run, both
    ranks keep the same right_lim / window position, rank 0 holds the inner band, and the
    summed m = 0 electron charge equals the single-process one (it does not depend on the random
    per-cell theta offsets, and the new plasma has barely moved)."""
    from chimeracl_b200 import _lib as real_lib
    world, steps = 2, 21
    mp.spawn(_lwfa_worker, args=(world, _free_port(), str(tmp_path), steps), nprocs=world,
             join=True)
    emu.patch_cuda_host_calls(monkeypatch)
    monkeypatch.setattr(real_lib, "_lib", real_lib._lib)      # restored after the test
    solver, eons, ions = _lwfa(emu.EmulatedComm(), steps)
    got = [np.load(tmp_path / ("lwfa%d.npz" % r)) for r in range(world)]
    assert sum(int(g["ne"]) for g in got) == int(eons.Args["Np"]) > 0
    assert sum(int(g["ni"]) for g in got) == int(ions.Args["Np"])
    assert abs(int(got[0]["ne"]) - int(got[1]["ne"])) <= 2 * 21 * 16      # one cell row
    for g in got:
        assert abs(float(g["lim"]) - float(eons.Args["right_lim"])) < 1e-9
        assert abs(float(g["xmin"]) - float(solver.Args["Xmin"])) < 1e-12
    assert float(got[0]["r_max"]) < float(got[1]["r_max"])
    rho = solver.DataDev["rho_m0"].get()
    assert np.abs(rho).max() > 0
    for g in got:                                   # rho is all-reduced: complete on both
        assert rel_err(g["rho"][1:], rho[1:]) < 1e-6


# ------------------------------------------------------------------ laser initialiser
@pytest.mark.parametrize("M", [0, 1])
def test_laser_initialiser_matches_oracle(M):
    """chimeracl_b200.laser.add_gausian_pulse (reference laser.py:3-37: host NumPy math
    around fb_transform / restore_B_fb) through the real Solver mixins on the emulated
    C ABI, against the oracle's restatement."""
    from oracle import orchestration as O
    from oracle.np_kernels import NumpyKernels
    from chimeracl_b200.laser import add_gausian_pulse
    cfg = {'Xmin': -20.0, 'Xmax': 20.0, 'Nx': 128, 'Rmin': 0.0, 'Rmax': 12.0, 'Nr': 24, 'M': M,
           'DampCells': 10}
    cfg['dt'] = (cfg['Xmax'] - cfg['Xmin']) / cfg['Nx']
    laser = {'k0': 1., 'a0': 2.0, 'x0': 2.0, 'Lx': 4.0, 'R': 3.0, 'x_foc': 15.0}
    S = emu.make_solver(cfg)
    add_gausian_pulse(S, dict(laser))
    So = O.OracleSolver(dict(cfg), NumpyKernels(M))
    O.add_gaussian_pulse(So, dict(laser))
    assert np.abs(So.D['Ez_m0']).max() > 0.5
    for k in ('Ez_fb_m0', 'Gz_fb_m0', 'Bx_fb_m0', 'By_fb_m0'):
        assert rel_err(S.DataDev[k].get(), So.D[k]) < 1e-11, k
    for f in 'EB':
        for c in 'xyz':
            k = '%s%s_m0' % (f, c)
            ref = So.D[k][1:]
            scale = max(np.abs(So.D['Ez_m0']).max(), 1e-300)
            assert np.abs(S.DataDev[k].get()[1:] - ref).max() / scale < 1e-11, k


# ------------------------------------------------------------------ diagnostics hook
def test_diagnostics_records_host_side(tmp_path, monkeypatch):
    """The checks of tests/test_gpu_parity.py::test_diagnostics_records (record layout of
    reference diagnostics.py:57-141, selections, weight scaling, the re-deposit / transform
    sequence of add_field) with the C ABI emulated: the Diagnostics class and the calls
    it makes into the solver are host logic."""
    from chimeracl_b200 import _lib as real_lib
    from test_gpu_parity import test_diagnostics_records as body
    emu.patch_cuda_host_calls(monkeypatch)
    comm = emu.EmulatedComm()
    monkeypatch.setattr(real_lib, "_lib", comm.lib)
    body(comm, tmp_path, monkeypatch)
