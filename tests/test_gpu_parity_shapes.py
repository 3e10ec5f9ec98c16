"""GPU parity at REAL shapes: the CUDA path (through the C ABI) against the reference's
own kernels (oracle/_ref, OpenMP over work-items; falls back to the NumPy restatement
where the prebuilt library is absent) on grids whose rows are far longer than any
kernel tile -- Nx = 1024 .. 4096 cells per row -- so that the full gather tiles, several
tiles per row, the tile-overflow slots, the 32- and 128-cell deposit CTAs inside one
row, uneven fillings, cells with thousands of particles and the trash bin are all
compared number by number, not just run.

Tolerances: integer products and coordinates bit-exact; deposits <= 1e-12 max|field|
(summation order); gather + Boris PER PARTICLE <= 1e-13 (|dp| / |p_ref| of that
particle, and |d g_inv| / g_inv) given identical fields; full steps <= 1e-10 on fields
and per particle; N-step runs on moments (sum w, sum w p, sum w (gamma-1), field
energy) <= 1e-8.
"""
import os

import numpy as np
import pytest

from oracle import orchestration as O
from oracle.np_kernels import NumpyKernels
from oracle.ref_kernels import RefKernels, ref_available

from helpers import ATTR, rel_err, per_particle_rel_err, moments, field_energy
from test_gpu_parity import INT_KEYS, check_sort_products, set_particles

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def comm():
    from chimeracl_b200.methods.generic_methods_cl import Communicator
    return Communicator(answers=[0, 0], seed=7)


def oracle_kernels(M):
    """The reference's own kernels where oracle/_ref was built (it ships prebuilt to
    the GPU box), else the NumPy restatement pinned against them."""
    if M in (0, 1) and ref_available(M):
        return RefKernels(M, parallel=True)
    return NumpyKernels(M)


def uneven_particles(cfg, n, seed, p_scale=1.0):
    """~n particles on the (Xmin..Xmax, 0..Rmax) box with very uneven fillings: a
    uniform background, a dense blob (thousands per cell: single-CTA radix / several
    deposit batches per cell), a moderately dense ring (hundreds per cell: bitonic
    path, > 544 per 32-cell deposit CTA), an empty band, on-axis particles and
    ~3 % outside the box (trash bin)."""
    rng = np.random.default_rng(seed)
    X0, X1, R1 = cfg["Xmin"], cfg["Xmax"], cfg["Rmax"]
    Lx = X1 - X0
    dx = Lx / (cfg["Nx"] - 1)
    dr = R1 / (cfg["Nr"] - 1.5)
    n_bg, n_blob, n_ring = int(0.55 * n), int(0.25 * n), int(0.17 * n)
    n_out = n - n_bg - n_blob - n_ring
    # background: uniform in x and in r^2, with an empty band in x
    x = rng.uniform(X0 - 0.01 * Lx, X1 + 0.01 * Lx, n_bg)
    band = (x > X0 + 0.30 * Lx) & (x < X0 + 0.34 * Lx)
    x[band] += 0.2 * Lx
    r = R1 * 1.02 * np.sqrt(rng.uniform(0, 1, n_bg))
    # dense blob: sigma of 3 cells in x and r around r = 20 dr
    xb = rng.normal(X0 + 0.61 * Lx, 3 * dx, n_blob)
    rb = np.abs(rng.normal(20 * dr, 3 * dr, n_blob))
    # ring sheet: 200 cells long, 6 cells thick
    xr = rng.uniform(X0 + 0.1 * Lx, X0 + 0.1 * Lx + 200 * dx, n_ring)
    rr = rng.uniform(60 * dr, 66 * dr, n_ring)
    # outside
    xo = rng.uniform(X1 + 0.02 * Lx, X1 + 0.3 * Lx, n_out)
    ro = rng.uniform(0, 2 * R1, n_out)
    x = np.concatenate((x, xb, xr, xo))
    r = np.concatenate((r, rb, rr, ro))
    th = rng.uniform(0, 2 * np.pi, x.size)
    perm = rng.permutation(x.size)               # random storage order
    x, r, th = x[perm], r[perm], th[perm]
    a = {"x": x, "y": r * np.sin(th), "z": r * np.cos(th)}
    a["y"][:5] = 1e-9 * dr                       # (almost) on the axis; exactly r = 0 is
    a["z"][:5] = 0.0                             # NaN in the reference's unguarded 1/r
    for k in ("px", "py", "pz"):
        a[k] = rng.normal(0, p_scale, x.size)
    a["w"] = rng.uniform(0.5, 1.5, x.size)
    a["g_inv"] = 1 / np.sqrt(1 + a["px"] ** 2 + a["py"] ** 2 + a["pz"] ** 2)
    return a


def _pair(cfg, pcfg, comm, arrays, K):
    from chimeracl_b200.particles import Particles
    from chimeracl_b200.solver import Solver
    S = Solver(dict(cfg), comm)
    So = O.OracleSolver(dict(cfg), K)
    P = Particles(dict(pcfg), comm)
    set_particles(P, arrays)
    Po = O.OracleParticles(dict(pcfg), K)
    Po.set_particles(**arrays)
    return S, So, P, Po


def _random_fields(S, So, seed):
    rng = np.random.default_rng(seed)
    for k in sorted(So.D):
        if k[0] in "EB" and "_fb_" not in k:
            a = rng.normal(size=So.D[k].shape)
            if So.D[k].dtype == np.complex128:
                a = a + 1j * rng.normal(size=a.shape)
            So.D[k][...] = a
            S.DataDev[k][:] = a


def _compare_deposits(S, So, names, tol=1e-12):
    for k in So.D:
        if k.startswith(names) and "_fb_" not in k:
            e = rel_err(S.DataDev[k].get(), So.D[k])
            assert e < tol, (k, e)


def _compare_momenta(P, Po, tol):
    got = {k: P.DataDev[k].get() for k in ("px", "py", "pz", "g_inv")}
    ep, eg = per_particle_rel_err(got, Po.D)
    assert ep < tol, ("momentum, per particle", ep)
    assert eg < tol, ("g_inv, per particle", eg)


@pytest.mark.parametrize("M", [0, 1])
def test_particle_kernels_long_rows_uneven_filling(comm, M, Nx=1024, Nr=128, n=2_000_000):
    """Nx = 1024, Nr = 128, ~2 M particles in random storage order, uneven fillings:
    sort, both deposits, gather + Boris; then the aligned storage through the one-pass
    particle side of PIC_loop (half push + J deposit + half push + index, chb_push_
    depose_push_index) against the reference sequence pic_loop.py:70-76.
    (Nx, Nr, n are only lowered by the CPU re-run of this body on the C-ABI emulator.)"""
    K = oracle_kernels(M)
    cfg = {"Xmin": -25.6, "Xmax": 25.6, "Nx": Nx, "Rmin": 0.0, "Rmax": 12.8, "Nr": Nr,
           "M": M, "dt": 0.05}
    pcfg = {"charge": -1, "dt": 0.05}
    arrays = uneven_particles(cfg, n, seed=100 + M)
    S, So, P, Po = _pair(cfg, pcfg, comm, arrays, K)
    P.sort_parts(S)
    Po.sort_parts(So)
    check_sort_products(P, Po)
    counts = Po.D["sum_in_cell"]
    assert counts[-1] > 0                                           # trash bin
    if n >= 2_000_000:
        assert counts[:-1].max() > 8192                             # giant cells
        assert (counts[:-1] == 0).sum() > 1000                      # and empty ones
        nrow = cfg["Nx"] - 1
        per32 = np.add.reduceat(counts[:nrow * 126], np.arange(0, nrow * 126, 32))
        assert per32.max() > 544                                    # multi-batch deposit CTAs

    S.depose_currents([P])
    S.depose_charge([P])
    So.depose_currents([Po])
    So.depose_charge([Po])
    _compare_deposits(S, So, ("rho_m", "Jx_m", "Jy_m", "Jz_m"))

    _random_fields(S, So, 11 + M)
    S.gather_and_push([P])
    So.gather_and_push([Po])
    _compare_momenta(P, Po, 1e-13)

    # aligned storage (what a production step sees), then the one-pass particle side
    P.align_parts()
    Po.align_parts()
    for k in ("x", "y", "z", "w"):
        assert np.array_equal(P.DataDev[k].get(), Po.D[k]), k
    for k in ("px", "py", "pz", "g_inv"):
        # equal to rounding after the gather; made identical so that everything that
        # follows compares kernels on the same inputs again
        P.DataDev[k][:] = Po.D[k]
    P.flag_sorted = False
    Po.flag_sorted = False
    P.sort_parts(S)
    Po.sort_parts(So)
    assert P.traversal_order_valid(S)
    S.depose_currents([P], push_mode="half+half")
    P.push_and_sort(S, mode="half")
    Po.push_coords("half")
    Po.sort_parts(So)
    So.depose_currents([Po])
    Po.push_coords("half")
    Po.sort_parts(So)
    for k in ("x", "y", "z"):
        assert np.array_equal(P.DataDev[k].get(), Po.D[k]), k
    check_sort_products(P, Po)
    _compare_deposits(S, So, ("Jx_m", "Jy_m", "Jz_m"))
    moved = (Po.D["indx_in_cell"] != np.repeat(
        np.arange(counts.size, dtype=np.uint32), counts)[:Po.Args["Np"]]).mean()
    assert moved > 0.05                                         # the cell-changer queue is used
    # gather along the new order on the aligned storage (full 64-cell tiles per row)
    S.gather_and_push([P])
    So.gather_and_push([Po])
    _compare_momenta(P, Po, 1e-13)
    S.depose_charge([P])
    So.depose_charge([Po])
    _compare_deposits(S, So, ("rho_m",))


def lattice_plasma(cfg, nppc, seed, dp):
    """BASELINE configs[2] generator at a chosen ppc: fill_grid lattice over all valid
    cells with a per-cell theta offset, w = r * w0, Gaussian thermal momenta."""
    K = NumpyKernels(1)
    A = O.grid_args(dict(cfg))
    rng = np.random.default_rng(seed)
    xg = A["Xgrid"][1:cfg["Nx"] - 1]                 # cells ix = 1 .. Nx-3
    rg = A["dr"] * np.arange(cfg["Nr"] - 1)          # cells ir = 0 .. Nr-3
    th = rng.uniform(0, 2 * np.pi, (xg.size - 1) * (rg.size - 1))
    x, y, z, w = K.fill_grid(th, xg, rg, nppc)
    w0 = 2 * np.pi * A["dx"] * A["dr"] * 0.01 / np.prod(nppc)
    a = {"x": x, "y": y, "z": z, "w": w * w0}
    for k in ("px", "py", "pz"):
        a[k] = rng.normal(0, dp, x.size)
    a["g_inv"] = 1 / np.sqrt(1 + a["px"] ** 2 + a["py"] ** 2 + a["pz"] ** 2)
    return a


def test_cfg3_shape_two_steps_against_reference_kernels(comm, Nx=4096, Nr=512):
    """BASELINE configs[2] grid (Nx = 4096, Nr = 512, M = 1, DampCells 50) with the
    uniform thermal plasma subsampled to 2 particles per cell (4.2 M electrons + as many
    immobile ions): two full PIC_loop.step() calls -- the second one takes the one-pass
    particle side -- against oracle.pic_step on the reference's kernels."""
    from chimeracl_b200.particles import Particles
    from chimeracl_b200.solver import Solver
    from chimeracl_b200.pic_loop import PIC_loop
    M = 1
    K = oracle_kernels(M)
    cfg = {"Xmin": -102.4, "Xmax": 102.4, "Nx": Nx, "Rmin": 0.0, "Rmax": 51.2, "Nr": Nr,
           "M": M, "DampCells": 50 if Nx >= 400 else 8}
    cfg["dt"] = (cfg["Xmax"] - cfg["Xmin"]) / cfg["Nx"]
    arrays = lattice_plasma(cfg, (1, 1, 2), seed=1234, dp=0.01)
    S = Solver(dict(cfg), comm)
    So = O.OracleSolver(dict(cfg), K)
    pcfg = {"Nppc": (1, 1, 2), "dx": So.Args["dx"], "dr": So.Args["dr"], "dt": cfg["dt"],
            "dens": 0.01, "charge": -1}
    P = Particles(dict(pcfg), comm)
    I = Particles(dict(pcfg, charge=1, Immobile=True), comm)
    set_particles(P, arrays)
    set_particles(I, arrays)
    Po = O.OracleParticles(dict(pcfg), K)
    Io = O.OracleParticles(dict(pcfg, charge=1, Immobile=True), K)
    Po.set_particles(**arrays)
    Io.set_particles(**arrays)
    loop = PIC_loop(solvers=[S], species=[P, I])
    for step in range(2):
        loop.step()
        O.pic_step(So, [Po, Io])
        for k in ("x", "y", "z"):
            # coordinates follow the momenta: not bit-exact after the first gather
            assert rel_err(P.DataDev[k].get(), Po.D[k]) < 1e-13, (step, k)
        assert np.array_equal(P.DataDev["sort_indx"].get(), Po.D["sort_indx"]), step
        assert np.array_equal(P.DataDev["cell_offset"].get(), Po.D["cell_offset"]), step
        for k in So.D:
            if "_fb_" in k or k.startswith("G"):
                continue
            e = rel_err(S.DataDev[k].get(), So.D[k])
            assert e < 1e-10, (step, k, e)
        _compare_momenta(P, Po, 1e-10)
    assert loop.fuse_push_sort and P._index_prefilled is False


# ----------------------------------------------------------------------------- cfg1
class ThetaTable:
    def __init__(self):
        self.calls = 0

    def __call__(self, ncells):
        k = self.calls
        self.calls += 1
        i = np.arange(1, ncells + 1, dtype=np.float64)
        return 2 * np.pi * np.mod(i * 0.6180339887498949 + k * 0.41421356237309515, 1.0)


def _build_cfg1(comm, Nx, Nr, prof_start=43.1, Lx=10.0, x0=0.0):
    """examples/lpa_script_small.py:19-45 on both sides, theta offsets from one table."""
    import importlib.util
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location(
        "lpa_small_cfg1", os.path.join(root, "examples", "lpa_script_small.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _, solver, eons, ions, frame, loop = mod.build(Nx=Nx, Nr=Nr, M=1, comm=comm, Lx=Lx, x0=x0,
                                                   profile_start=prof_start)
    tt = ThetaTable()
    for sp in (eons, ions):
        sp._theta_offsets = lambda n, tt=tt: torch.from_numpy(tt(n)).to(comm.device)

    K = oracle_kernels(1)
    grid_in = {"Xmin": -43.0, "Xmax": 43.0, "Nx": Nx, "Rmin": 0.0, "Rmax": 36.0, "Nr": Nr,
               "M": 1, "DampCells": 50}
    grid_in["dt"] = (grid_in["Xmax"] - grid_in["Xmin"]) / grid_in["Nx"]
    So = O.OracleSolver(grid_in, K)
    O.add_gaussian_pulse(So, {"k0": 1.0, "a0": 3, "x0": x0, "Lx": Lx, "R": 12.0,
                              "x_foc": 100.0})
    eons_in = {"Nppc": (2, 2, 4), "dx": So.Args["dx"], "dr": So.Args["dr"],
               "dt": So.Args["dt"], "dens": 7e18 / (1.1e21 / 0.8 ** 2), "charge": -1}
    Eo = O.OracleParticles(dict(eons_in), K)
    Io = O.OracleParticles(dict(eons_in, charge=1, Immobile=True), K)
    Io.Args["InjectorSource"] = Eo
    tto = ThetaTable()
    Fo = O.OracleFrame({"Velocity": 1.0, "dt": So.Args["dt"], "Steps": 20,
                        "DensityProfiles": [{"coord": "x",
                                             "points": [-100, prof_start, 90, 5e5],
                                             "values": [0, 0, 1, 1]}]}, tto)
    return (solver, eons, ions, loop), (So, Eo, Io, Fo)


def _compare_run(gpu, orc, checkpoints, tol_mom=1e-8, tol_fld=1e-8):
    solver, eons, ions, loop = gpu
    So, Eo, Io, Fo = orc
    it = 0
    for stop in checkpoints:
        while it < stop:
            loop.step()
            O.pic_step(So, [Eo, Io], frames=[Fo], it=it)
            it += 1
        assert loop.it == it
        assert int(eons.Args["Np"]) == Eo.Args["Np"], it
        assert int(ions.Args["Np"]) == Io.Args["Np"], it
        assert abs(solver.Args["Xmin"] - So.Args["Xmin"]) < 1e-12
        mg = moments({k: eons.DataDev[k].get() for k in ATTR})
        mo = moments(Eo.D)
        # scales with a floor: before the pulse arrives the momenta are rounding noise
        # of the spectral solve (1e-16 of a0), which no relative bound can hold
        wsum = abs(mo["w"])
        scale = abs(mo["w_abs_p"]) + 1e-3 * wsum
        assert abs(mg["w"] - mo["w"]) <= tol_mom * wsum, (it, mg, mo)
        for k in ("wpx", "wpy", "wpz"):
            assert abs(mg[k] - mo[k]) <= tol_mom * scale, (it, k, mg[k], mo[k], scale)
        assert abs(mg["w_kin"] - mo["w_kin"]) <= tol_mom * (abs(mo["w_kin"]) + 1e-6 * wsum), it
        eg = field_energy({k: solver.DataDev[k].get() for k in solver.DataDev
                           if k[0] in "EB" and "_fb_" not in k and k[1] in "xyz"}, So.Args)
        eo = field_energy({k: v for k, v in So.D.items()
                           if k[0] in "EB" and "_fb_" not in k and k[1] in "xyz"}, So.Args)
        assert abs(eg - eo) <= tol_fld * eo, (it, eg, eo)
        # fields, each family on ITS scale (Ex_m0 etc. are rounding noise next to the
        # laser's Ez_m0, rho_m1 next to rho_m0)
        for fam in ("E", "B", "rho"):
            keys = [k for k in So.D if k.startswith(fam) and "_fb_" not in k]
            scale = max(np.abs(So.D[k]).max() for k in keys)
            if fam == "rho":        # neutral plasma: electrons and ions cancel to noise
                scale = max(scale, 1e-3 * Eo.Args["dens"])
            for k in keys:
                e = np.abs(solver.DataDev[k].get() - So.D[k]).max() / scale
                assert e < tol_fld, (it, k, e)
    return it


def test_cfg1_lwfa_moving_window_against_oracle(comm, Nx=900, Nr=90, checkpoints=(1, 20, 21, 200)):
    """BASELINE configs[0] = examples/lpa_script_small.py verbatim (Nx = 900 -> Bluestein
    FFT, Nr = 90, M = 1, a0 = 3, window every 20 steps, plasma ramp from x = 43.1):
    fields, particle counts and moments after 1, 20, 21 (first step after the second
    injection) and 200 steps against oracle.pic_step with the OracleFrame restatement
    of frame.py:22-64 (SURVEY 8d role of cfg1)."""
    gpu, orc = _build_cfg1(comm, Nx, Nr)
    _compare_run(gpu, orc, checkpoints)
    assert orc[1].Args["Np"] > (100000 if Nx == 900 else 0)


def test_lwfa_wake_against_oracle(comm):
    """Same setup with a shorter pulse placed at x0 = 22 and the plasma ramp starting at
    x = 30, so that within 300 steps the plasma runs through the a0 = 3 peak: electrons
    are strongly driven (relativistic momenta, cells emptied and overfilled, particles
    leaving through the trash bin), on a power-of-two grid."""
    gpu, orc = _build_cfg1(comm, 1024, 64, prof_start=30.0, Lx=5.0, x0=22.0)
    _compare_run(gpu, orc, (100, 300), tol_mom=1e-7, tol_fld=1e-7)
    px = gpu[1].DataDev["px"].get()
    assert np.abs(px).max() > 0.5          # the pulse really drives the electrons
