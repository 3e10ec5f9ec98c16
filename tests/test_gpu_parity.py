"""GPU parity tests: the CUDA path (through the C ABI, via the reference-shaped Python
API) against the oracle on identical seeded inputs, and against the committed golden
vectors produced from the reference's own kernels.

Tolerances (stated per check): integer products and the coordinate push are
bit-exact; gather + Boris <= 1e-13 given identical fields; depositions
<= 1e-12 * max|field| (summation order differs); single transforms / spectral
operators <= 1e-12 relative to max; full steps <= 1e-10.
"""
import numpy as np
import pytest

from oracle import orchestration as O
from oracle.np_kernels import NumpyKernels

from helpers import ATTR, golden_cfgs, load_golden, oracle_case_from_golden, rel_err

pytestmark = pytest.mark.gpu

INT_KEYS = ("indx_in_cell", "sum_in_cell", "cell_offset", "sort_indx")


@pytest.fixture(scope="module")
def comm():
    from chimeracl_b200.methods.generic_methods_cl import Communicator
    return Communicator(answers=[0, 0], seed=7)


def gpu_case_from_golden(G, comm):
    from chimeracl_b200.particles import Particles
    from chimeracl_b200.solver import Solver
    cfg, pcfg = golden_cfgs(G)
    S = Solver(dict(cfg), comm)
    for k in G.files:
        if k.startswith("in/S/"):
            S.DataDev[k[5:]][:] = G[k]
    P = Particles(dict(pcfg), comm)
    I = Particles(dict(pcfg, charge=1, Immobile=True), comm)
    set_particles(P, {a: G["in/P/" + a] for a in ATTR})
    set_particles(I, {a: G["in/P/" + a] for a in ("x", "y", "z", "w")})
    return S, P, I


def set_particles(P, arrays):
    from chimeracl_b200.devarray import DevArray
    for a in P._attr_names():
        P.DataDev[a] = DevArray.from_numpy(np.ascontiguousarray(arrays[a], dtype=np.float64),
                                           P.comm.device)
    P.reset_num_parts()
    P.flag_sorted = False


def check_sort_products(P, Po):
    assert int(P.Args["Np_stay"]) == Po.Args["Np_stay"]
    for k in INT_KEYS:
        assert np.array_equal(P.DataDev[k].get(), Po.D[k]), k


# ----------------------------------------------------------------------------- particles
@pytest.mark.parametrize("M", [0, 1])
@pytest.mark.parametrize("fused", [False, True])
def test_push_and_sort_bit_exact(comm, M, fused):
    G = load_golden(M)
    S, P, I = gpu_case_from_golden(G, comm)
    So, Po, Io = oracle_case_from_golden(G, NumpyKernels(M))
    for p, po in ((P, Po), (I, Io)):
        if fused:
            p.push_and_sort(S, "half")
        else:
            p.push_coords("half")
            p.sort_parts(S)
        po.push_coords("half")
        po.sort_parts(So)
    for k in ("x", "y", "z"):
        assert np.array_equal(P.DataDev[k].get(), Po.D[k]), k
    check_sort_products(P, Po)
    check_sort_products(I, Io)
    # golden vectors (reference kernels): same products
    for k in INT_KEYS:
        assert np.array_equal(P.DataDev[k].get(), G["sort1/P/" + k]), k
    # align: gather by the permutation, drop the trash tail
    P.align_parts()
    Po.align_parts()
    assert P.Args["Np"] == Po.Args["Np"]
    for k in ATTR:
        assert np.array_equal(P.DataDev[k].get(), Po.D[k]), k
    assert np.array_equal(P.DataDev["sort_indx"].get(), Po.D["sort_indx"])


def _random_species(comm, n, box, seed, spread=1.0):
    from chimeracl_b200.particles import Particles
    rng = np.random.default_rng(seed)
    arrays = {"x": rng.uniform(box[0], box[1], n), "y": rng.normal(0, spread, n),
              "z": rng.normal(0, spread, n), "px": rng.normal(0, 1, n),
              "py": rng.normal(0, 1, n), "pz": rng.normal(0, 1, n),
              "w": rng.uniform(0.5, 1.5, n)}
    arrays["g_inv"] = 1 / np.sqrt(1 + arrays["px"] ** 2 + arrays["py"] ** 2 + arrays["pz"] ** 2)
    P = Particles({"charge": -1, "dt": 0.01}, comm)
    set_particles(P, arrays)
    Po = O.OracleParticles({"charge": -1, "dt": 0.01}, NumpyKernels(1))
    Po.set_particles(**arrays)
    return P, Po


@pytest.mark.parametrize("shape,n", [((8, 6), 300000),     # ~10^4 per cell: giant segments
                                     ((64, 40), 400000),   # ~160 per cell: bitonic path
                                     ((256, 128), 500000), # ~15 per cell: insertion path
                                     ((16, 8), 37)])       # fewer particles than a warp
def test_sort_random_order_stable(comm, shape, n):
    """Storage in random order (worst case for the run-aggregated scatter): the
    permutation must still be the stable one, including the trash-bin tail."""
    from chimeracl_b200.solver import Solver
    Nx, Nr = shape
    cfg = {"Xmin": -1.0, "Xmax": 1.0, "Nx": Nx, "Rmin": 0.0, "Rmax": 1.0, "Nr": Nr, "M": 1}
    S = Solver(dict(cfg), comm)
    So = O.OracleSolver(dict(cfg), NumpyKernels(1))
    P, Po = _random_species(comm, n, (-1.2, 1.2), seed=n, spread=0.5)
    P.sort_parts(S)
    Po.sort_parts(So)
    check_sort_products(P, Po)
    assert Po.D["sum_in_cell"][-1] > 0          # the trash bin is populated
    # idempotence: sorting the aligned particles gives the identity permutation
    P.align_parts()
    P.flag_sorted = False
    P.sort_parts(S)
    assert np.array_equal(P.DataDev["sort_indx"].get(), np.arange(P.Args["Np"], dtype=np.uint32))


def test_sort_edge_cases(comm):
    from chimeracl_b200.particles import Particles
    from chimeracl_b200.solver import Solver
    cfg = {"Xmin": 0.0, "Xmax": 1.0, "Nx": 8, "Rmin": 0.0, "Rmax": 1.0, "Nr": 6, "M": 1}
    S = Solver(dict(cfg), comm)
    P = Particles({"charge": -1}, comm)
    P.sort_parts(S)                        # empty species: no-op
    assert P.flag_sorted and "sort_indx" not in P.DataDev
    n = 50
    far = np.full(n, 10.0)
    set_particles(P, {a: far for a in ATTR})
    P.sort_parts(S)                        # everything in the trash bin
    assert P.Args["Np_stay"] == 0
    assert np.array_equal(P.DataDev["sort_indx"].get(), np.arange(n))
    P.align_parts()
    assert P.Args["Np"] == 0 and P.DataDev["x"].size == 0
    nan = np.full(4, np.nan)
    set_particles(P, {a: nan for a in ATTR})
    P.sort_parts(S)                        # NaN coordinates -> trash bin
    assert P.Args["Np_stay"] == 0


# ----------------------------------------------------------------------------- grid
@pytest.mark.parametrize("M", [0, 1])
def test_deposit_and_gather(comm, M):
    G = load_golden(M)
    S, P, I = gpu_case_from_golden(G, comm)
    So, Po, Io = oracle_case_from_golden(G, NumpyKernels(M))
    rng = np.random.default_rng(11)
    for k in sorted(So.D):
        if k[0] in "EB" and "_fb_" not in k:
            a = rng.normal(size=So.D[k].shape)
            if So.D[k].dtype == np.complex128:
                a = a + 1j * rng.normal(size=a.shape)
            So.D[k][...] = a
            S.DataDev[k][:] = a
    for p, po, s, so in ((P, Po, S, So), (I, Io, S, So)):
        p.push_coords("half")
        p.sort_parts(s)
        po.push_coords("half")
        po.sort_parts(so)
    S.depose_currents([P, I])
    S.depose_charge([P, I])
    So.depose_currents([Po, Io])
    So.depose_charge([Po, Io])
    for k in So.D:
        if k.startswith(("rho_m", "Jx_m", "Jy_m", "Jz_m")):
            assert rel_err(S.DataDev[k].get(), So.D[k]) < 1e-12, k
    S.gather_and_push([P, I])
    So.gather_and_push([Po, Io])
    for k in So.D:
        if k[0] in "EB" and "_fb_" not in k:     # ghost rows written by warp_axis
            assert np.array_equal(S.DataDev[k].get(), So.D[k]), k
    for k in ("px", "py", "pz", "g_inv"):          # FMA-contracted mode sum: few ulp
        got, ref = P.DataDev[k].get(), Po.D[k]
        assert rel_err(got, ref) < 1e-13, (k, np.abs(got - ref).max())


def test_deposit_dense_cells(comm):
    """Many particles per cell (batches of the staged deposit wrap several times)."""
    from chimeracl_b200.solver import Solver
    cfg = {"Xmin": -1.0, "Xmax": 1.0, "Nx": 12, "Rmin": 0.0, "Rmax": 1.0, "Nr": 9, "M": 1}
    S = Solver(dict(cfg), comm)
    So = O.OracleSolver(dict(cfg), NumpyKernels(1))
    P, Po = _random_species(comm, 200000, (-1.1, 1.1), seed=3, spread=0.4)
    P.sort_parts(S)
    Po.sort_parts(So)
    S.depose_currents([P])
    S.depose_charge([P])
    So.depose_currents([Po])
    So.depose_charge([Po])
    for k in So.D:
        if k.startswith(("rho_m", "Jx_m", "Jy_m", "Jz_m")):
            assert rel_err(S.DataDev[k].get(), So.D[k]) < 1e-12, k


# ----------------------------------------------------------------------------- spectral
@pytest.mark.parametrize("n", [8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384,
                               30, 100, 900, 1000, 1537])
def test_fft_against_numpy(comm, n):
    from chimeracl_b200.devarray import DevArray
    from chimeracl_b200.solver import Solver
    cfg = {"Xmin": -1.0, "Xmax": 1.3, "Nx": n, "Rmin": 0.0, "Rmax": 1.0, "Nr": 6, "M": 1}
    S = Solver(dict(cfg), comm)
    rng = np.random.default_rng(n)
    a = rng.normal(size=(5, n)) + 1j * rng.normal(size=(5, n))
    src = DevArray.from_numpy(a, comm.device)
    dst = DevArray.zeros((5, n), np.complex128, comm.device)
    S._fft(dst, src, 0)
    assert rel_err(dst.get(), np.fft.fft(a, axis=1)) < 5e-14
    S._fft(dst, src, 1)
    assert rel_err(dst.get(), np.fft.ifft(a, axis=1)) < 5e-14
    S._fft(src, src, 0)      # in place
    assert rel_err(src.get(), np.fft.fft(a, axis=1)) < 5e-14


@pytest.mark.parametrize("K,N", [(5, 8), (89, 900), (127, 130), (128, 64), (255, 2048), (511, 333)])
def test_dht_contraction_against_numpy(comm, K, N):
    from chimeracl_b200.devarray import DevArray
    from chimeracl_b200.solver import Solver
    cfg = {"Xmin": -1.0, "Xmax": 1.0, "Nx": 16, "Rmin": 0.0, "Rmax": 1.0, "Nr": 6, "M": 1}
    S = Solver(dict(cfg), comm)
    rng = np.random.default_rng(K * 1000 + N)
    A = rng.normal(size=(K, K))
    Br = rng.normal(size=(K + 1, N))
    Bc = rng.normal(size=(K + 1, N)) + 1j * rng.normal(size=(K + 1, N))
    dA = DevArray.from_numpy(A, comm.device)
    for B in (Br, Bc):
        dB = DevArray.from_numpy(B, comm.device)
        dC = DevArray.zeros((K + 1, N), B.dtype, comm.device)
        S._dot(dC[1:], dA, dB[1:])                 # row-offset views on both sides
        ref = A.dot(B[1:])
        assert rel_err(dC.get()[1:], ref) < 1e-13
        assert not dC.get()[0].any()
    # fused two-output epilogue: y = -b, z += -i b
    dB = DevArray.from_numpy(Bc[1:].copy(), comm.device)
    y = DevArray.zeros((K, N), np.complex128, comm.device)
    z0 = rng.normal(size=(K, N)) + 1j * rng.normal(size=(K, N))
    z = DevArray.from_numpy(z0, comm.device)
    S._cdot2(dA, dB, y, -1.0, False, z, -1j, True)
    b = A.dot(Bc[1:])
    assert rel_err(y.get(), -b) < 1e-13
    assert rel_err(z.get(), z0 - 1j * b) < 1e-13


@pytest.mark.parametrize("M", [0, 1])
def test_spectral_pipeline_against_oracle(comm, M):
    """fb_transform fwd, smoothing, grad, PSATD advance, damping, rot + Poisson,
    fb_transform bwd -- each compared with the oracle on identical inputs."""
    G = load_golden(M)
    S, _, _ = gpu_case_from_golden(G, comm)
    So, _, _ = oracle_case_from_golden(G, NumpyKernels(M))
    rng = np.random.default_rng(21)
    for k in sorted(So.D):
        if k.startswith(("rho_m", "Jx_m", "Jy_m", "Jz_m")) or k.startswith(("dN1", "dN0")):
            a = rng.normal(size=So.D[k].shape)
            if So.D[k].dtype == np.complex128:
                a = a + 1j * rng.normal(size=a.shape)
            So.D[k][...] = a
            S.DataDev[k][:] = a

    def compare(keys, tol, what):
        for k in So.D:
            if k.startswith(keys):
                e = rel_err(S.DataDev[k].get(), So.D[k])
                assert e < tol, (what, k, e)

    for s in (S, So):
        s.fb_transform(scals=["rho"], vects=["J"], dir=0)
    compare(("rho_fb", "Jx_fb", "Jy_fb", "Jz_fb"), 1e-12, "forward")
    for s in (S, So):
        s.fields_smooth(["rho", "Jx", "Jy", "Jz"])
        s.field_grad("rho", "dN1")
    compare(("dN1",), 1e-12, "grad")
    for s in (S, So):
        s.push_fields()
    compare(("Ex_fb", "Ey_fb", "Ez_fb", "Gx_fb", "Gy_fb", "Gz_fb"), 1e-12, "psatd")
    for s in (S, So):
        s.damp_fields()
    compare(("Ex_fb", "Ey_fb", "Ez_fb", "Gx_fb", "Gy_fb", "Gz_fb"), 1e-12, "damp")
    for s in (S, So):
        s.restore_B_fb()
    compare(("Bx_fb", "By_fb", "Bz_fb"), 1e-12, "rot+poisson")
    for s in (S, So):
        s.fb_transform(vects=["E", "B"], dir=1)
    compare(("Ex_m", "Ey_m", "Ez_m", "Bx_m", "By_m", "Bz_m"), 1e-12, "backward")
    if M == 1:
        for s in (S, So):
            s.field_div("E", "rho")
        compare(("rho_fb",), 1e-12, "div")


# ----------------------------------------------------------------------------- full step
@pytest.mark.parametrize("M", [0, 1])
def test_pic_steps_against_golden(comm, M):
    """Two full PIC_loop.step() calls against the reference-kernel golden run."""
    from chimeracl_b200.pic_loop import PIC_loop
    G = load_golden(M)
    S, P, I = gpu_case_from_golden(G, comm)
    loop = PIC_loop(solvers=[S], species=[P, I], frames=[], diags=[])
    loop.step()
    for k in G.files:
        if k.startswith("step1/S/"):
            assert rel_err(S.DataDev[k[8:]].get(), G[k]) < 1e-10, k
        elif k.startswith("step1/P/"):
            assert rel_err(P.DataDev[k[8:]].get(), G[k]) < 1e-10, k
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


def test_transformer_round_trip_cfg2(comm):
    """examples/test_transformer.py at BASELINE config 2 (Nx=2048, Nr=256, M=1): beam
    -> sort -> align -> depose_charge -> forward -> zero -> backward; error formula
    of test_transformer.py:40-43, stated tolerance 1e-12."""
    from chimeracl_b200.particles import Particles
    from chimeracl_b200.solver import Solver
    grid_in = {"Xmin": -1., "Xmax": 1., "Nx": 2048, "Rmin": 0, "Rmax": 1., "Nr": 256, "M": 1}
    parts = Particles(grid_in, comm)
    grid = Solver(grid_in, comm)
    beam_in = {"Np": int(7e6), "FullCharge": 1, "x_c": 0., "Lx": 0.3, "y_c": 0.2, "Ly": 0.3,
               "z_c": 0.2, "Lz": 0.3}
    parts.add_particles(beam_in=beam_in)
    parts.sort_parts(grid=grid)
    parts.align_parts()
    grid.depose_charge([parts, ])
    tmp0 = grid.DataDev["rho_m0"].get().copy()
    tmp1 = grid.DataDev["rho_m1"].get().copy()
    grid.fb_transform(scals=["rho", ], dir=0)
    grid.set_to(grid.DataDev["rho_m0"], 0)
    grid.set_to(grid.DataDev["rho_m1"], 0)
    grid.fb_transform(scals=["rho", ], dir=1)
    err = (np.abs(grid.DataDev["rho_m0"].get() - tmp0)[1:] / np.abs(tmp0[1:]).max()
           + np.abs(grid.DataDev["rho_m1"].get() - tmp1)[1:] / np.abs(tmp1[1:]).max()).max()
    assert err < 1e-12, err


def test_laser_group_velocity(comm):
    """examples/test_laser_veloc.py (M=0 vacuum) at reduced size: same property and
    bound as tests/test_oracle.py::test_laser_group_velocity_property."""
    from chimeracl_b200.laser import add_gausian_pulse
    from chimeracl_b200.solver import Solver
    cfg = {"Xmin": -40.0, "Xmax": 40.0, "Rmax": 40.0, "M": 0}
    dx, dr, cfg["dt"] = 0.1, 0.5, 0.1
    cfg["Nx"] = int((cfg["Xmax"] - cfg["Xmin"]) / dx) // 2 * 2
    cfg["Nr"] = int(cfg["Rmax"] / dr) // 2 * 2 + 1
    solver = Solver(cfg, comm)
    laser = {"k0": 1.0, "a0": 1.0, "x0": 0, "Lx": 8.0, "R": 8.0, "x_foc": 10.0}
    add_gausian_pulse(solver, laser)
    xc = []
    for _ in range(40):
        solver.push_fields()
        solver.fb_transform(scals=["Ez"], dir=1)
        v = solver.DataDev["Ez_m0"].get()
        Px = (solver.Args["Rgrid"][1:, None] * v[1:, :] ** 2).sum(0)
        xc.append((solver.Args["Xgrid"] * Px).sum() / Px.sum())
    xc = np.array(xc)
    veloc = 1 - (xc[1:] - xc[:-1]) / solver.Args["dt"]
    theory = (2.0 * np.pi * laser["R"]) ** -2
    assert abs(veloc.mean() - theory) / theory < 0.1


@pytest.mark.parametrize("M,Nx", [(0, 256), (1, 900)])
def test_laser_initialiser_matches_oracle(comm, M, Nx):
    """add_gausian_pulse (reference laser.py:3-37) on the CUDA path against
    oracle.orchestration.add_gaussian_pulse: Ez_fb_m0, Gz_fb_m0, the restored B spectra and
    all E / B grids <= 1e-12 of the pulse amplitude (Nx = 900: cfg1's Bluestein FFT)."""
    from chimeracl_b200.laser import add_gausian_pulse
    from chimeracl_b200.solver import Solver
    cfg = {"Xmin": -43.0, "Xmax": 43.0, "Nx": Nx, "Rmin": 0.0, "Rmax": 36.0, "Nr": 90, "M": M,
           "DampCells": 50}
    cfg["dt"] = (cfg["Xmax"] - cfg["Xmin"]) / cfg["Nx"]
    laser = {"k0": 1.0, "a0": 3.0, "x0": 2.0, "Lx": 10.0, "R": 12.0, "x_foc": 100.0}
    S = Solver(dict(cfg), comm)
    add_gausian_pulse(S, dict(laser))
    So = O.OracleSolver(dict(cfg), NumpyKernels(M))
    O.add_gaussian_pulse(So, dict(laser))
    assert np.abs(So.D["Ez_m0"]).max() > 1.0
    for k in ("Ez_fb_m0", "Gz_fb_m0", "Bx_fb_m0", "By_fb_m0", "Bz_fb_m0"):
        scale = np.abs(So.D["Ez_fb_m0" if k[0] == "E" else
                            "Gz_fb_m0" if k[0] == "G" else "By_fb_m0"]).max()
        assert np.abs(S.DataDev[k].get() - So.D[k]).max() / scale < 1e-12, k
    for f in "EB":
        scale = max(np.abs(So.D["%s%s_m0" % (f, c)]).max() for c in "xyz")
        for c in "xyz":
            for m in range(M + 1):
                k = "%s%s_m%d" % (f, c, m)
                assert np.abs(S.DataDev[k].get()[1:] - So.D[k][1:]).max() / scale < 1e-12, k


@pytest.mark.parametrize("M", [0, 1, 2])
def test_field_poiss_scl(comm, M):
    """field_poiss_scl (reference transformer_methods_cl.py:73-77): spectrum times
    Poiss_m = 1 / w_m^2, per mode; one multiplication per element -> bit-exact."""
    from chimeracl_b200.solver import Solver
    cfg = {"Xmin": -2.0, "Xmax": 3.0, "Nx": 96, "Rmin": 0.0, "Rmax": 4.0, "Nr": 37, "M": M}
    S = Solver(dict(cfg), comm)
    So = O.OracleSolver(dict(cfg), NumpyKernels(M))
    rng = np.random.default_rng(40 + M)
    for m in range(M + 1):
        k = "rho_fb_m%d" % m
        a = rng.normal(size=So.D[k].shape) + 1j * rng.normal(size=So.D[k].shape)
        So.D[k][...] = a
        S.DataDev[k][:] = a
    S.field_poiss_scl("rho")
    So.field_poiss_scl("rho")
    for m in range(M + 1):
        k = "rho_fb_m%d" % m
        assert np.array_equal(S.DataDev[k].get(), So.D[k]), k


def test_phase_timer_keys(comm):
    """PIC_loop(timit=True) (reference pic_loop.py:5-8, 42-55): the Timer dict carries
    exactly the reference's phase keys, in seconds, and the phases add up to the step."""
    import torch
    from chimeracl_b200.pic_loop import PIC_loop, loop_steps
    assert loop_steps == ["frame", "push-x", "sort", "depose", "transform", "smooth",
                          "data_copy", "grad", "push-eb", "damp-eb", "restore_B",
                          "gather + push-p"]
    G = load_golden(1)
    S, P, I = gpu_case_from_golden(G, comm)
    loop = PIC_loop(solvers=[S], species=[P, I], timit=True)
    assert list(loop.Timer) == loop_steps and not any(loop.Timer.values())
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(3):
        loop.step()
    t1.record()
    T = loop.timer_collect()
    assert list(T) == loop_steps
    total = t0.elapsed_time(t1) * 1e-3
    assert all(v >= 0 for v in T.values())
    for k in ("sort", "depose", "transform", "grad", "push-eb", "damp-eb", "restore_B",
              "gather + push-p"):
        assert T[k] > 0, k
    assert 0.3 * total < sum(T.values()) <= total * 1.001


# ----------------------------------------------------------------------------- moving window
def test_lwfa_moving_window_run(comm):
    """examples/lpa_script_small.py at reduced size for 45 steps: laser initialiser,
    frame shift + plasma injection (steps 0, 20, 40), immobile ions copied from the
    electrons (InjectorSource), sort + align after every injection.  Checks the
    bookkeeping invariants of reference frame.py:32-64 and that the physics stays sane."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "examples", "lpa_script_small.py")
    spec = importlib.util.spec_from_file_location("lpa_small", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _, solver, eons, ions, frame, loop = mod.build(Nx=300, Nr=48, M=1, comm=comm)
    Ez0 = solver.DataDev["Ez_m0"].get()
    assert np.abs(Ez0).max() > 1.0            # a0 = 3 pulse is on the grid
    xmin0 = solver.Args["Xmin"]
    counts = []
    for _ in range(45):
        loop.step()
        counts.append(int(eons.Args["Np"]))
    # three injections happened, each adds a slab of 20 cells x (Nr-2) rows x 16 ppc
    dx, dt = solver.Args["dx"], solver.Args["dt"]
    assert abs(solver.Args["Xmin"] - (xmin0 + 3 * 20 * dt)) < 1e-9
    assert abs(float(solver.DataDev["Xmin"].get()[0]) - solver.Args["Xmin"]) < 1e-9
    assert counts[0] > 0 and counts[20] > counts[19] and counts[40] > counts[39]
    assert int(ions.Args["Np"]) > 0
    # aligned storage after the last injection: cell indices are non-decreasing
    eons.flag_sorted = False
    eons.sort_parts(solver)
    idx = eons.DataDev["indx_in_cell"].get()
    srt = eons.DataDev["sort_indx"].get()
    assert (np.diff(idx[srt].astype(np.int64)) >= 0).all()
    for k in ("x", "px", "g_inv"):
        v = eons.DataDev[k].get()
        assert np.isfinite(v).all(), k
    for k in ("Ez_m0", "Ex_m1", "Bz_m0", "rho_m0"):
        assert np.isfinite(solver.DataDev[k].get()).all(), k
    # the plasma is quasi-neutral where the laser has not arrived: |rho| small vs n_e
    g = eons.DataDev["g_inv"].get()
    assert g.min() > 0 and g.max() <= 1.0 + 1e-12


# ----------------------------------------------------------------------------- extensions
def test_mode2_particle_kernels(comm):
    """M=2 (BASELINE config 5's mode count): no reference particle kernels exist, so
    the CUDA path is compared with the generalised restatement (parity unpinned) --
    deposit <= 1e-12, gather <= 1e-13."""
    from chimeracl_b200.solver import Solver
    cfg = {"Xmin": -1.0, "Xmax": 1.0, "Nx": 40, "Rmin": 0.0, "Rmax": 1.0, "Nr": 20, "M": 2}
    S = Solver(dict(cfg), comm)
    So = O.OracleSolver(dict(cfg), NumpyKernels(2))
    P, Po = _random_species(comm, 60000, (-1.1, 1.1), seed=12, spread=0.4)
    Po.K = NumpyKernels(2)
    rng = np.random.default_rng(13)
    for k in sorted(So.D):
        if k[0] in "EB" and "_fb_" not in k:
            a = rng.normal(size=So.D[k].shape)
            if So.D[k].dtype == np.complex128:
                a = a + 1j * rng.normal(size=a.shape)
            So.D[k][...] = a
            S.DataDev[k][:] = a
    P.sort_parts(S)
    Po.sort_parts(So)
    check_sort_products(P, Po)
    S.depose_currents([P])
    S.depose_charge([P])
    So.depose_currents([Po])
    So.depose_charge([Po])
    for k in So.D:
        if k.startswith(("rho_m", "Jx_m", "Jy_m", "Jz_m")):
            assert rel_err(S.DataDev[k].get(), So.D[k]) < 1e-12, k
    S.gather_and_push([P])
    So.gather_and_push([Po])
    for k in ("px", "py", "pz", "g_inv"):
        assert rel_err(P.DataDev[k].get(), Po.D[k]) < 1e-13, k


def test_particle_creation_matches_oracle(comm):
    """make_new_domain's lattice (fill_grid layout) and dens_profile against the
    restated reference kernels (kernels/particles_generic.cl:6-84)."""
    import torch
    from chimeracl_b200.particles import Particles
    P = Particles({"Nppc": (2, 3, 4), "dx": 0.25, "dr": 0.1, "charge": -1}, comm)
    rng = np.random.default_rng(3)
    xg = -2.0 + 0.25 * np.arange(9)
    rg = 0.1 * np.arange(7)
    th = rng.uniform(0, 2 * np.pi, (xg.size - 1) * (rg.size - 1))
    P._fill_grid(torch.from_numpy(th).to(comm.device), xg, rg, (2, 3, 4))
    ref = NumpyKernels(1).fill_grid(th, xg, rg, (2, 3, 4))
    for name, r in zip(("x_new", "y_new", "z_new", "w_new"), ref):
        assert rel_err(P.DataDev[name].get(), r) < 1e-14, name
    P.dens_profile([-3.0, -1.0, 0.5, 3.0], [0.0, 0.0, 1.0, 2.0], -2.0, 0.0,
                   coord="x_new", weight="w_new")
    x_loc = np.array([-3.0, -1.0, 0.5, 3.0])
    f_loc = np.array([0.0, 0.0, 1.0, 2.0])
    w = ref[3].copy()
    NumpyKernels(1).profile_by_interpolant(ref[0], w, x_loc, f_loc, 1.0 / np.diff(x_loc))
    assert rel_err(P.DataDev["w_new"].get(), w) < 1e-14


@pytest.mark.parametrize("M", [0, 1])
def test_fused_push_deposit_equals_push_sort_deposit(comm, M):
    """chb_push_depose_vector (half push + current deposit along the PREVIOUS sort)
    against push_coords -> sort_parts -> depose_currents, with fast particles so that
    many change cell, leave the box or come back from the trash bin."""
    from chimeracl_b200.solver import Solver
    from chimeracl_b200.particles import Particles
    cfg = {"Xmin": -1.0, "Xmax": 1.0, "Nx": 48, "Rmin": 0.0, "Rmax": 1.0, "Nr": 24, "M": M}
    S1, S2 = Solver(dict(cfg), comm), Solver(dict(cfg), comm)
    rng = np.random.default_rng(77)
    n = 120000
    arrays = {"x": rng.uniform(-1.15, 1.15, n), "y": rng.normal(0, 0.45, n),
              "z": rng.normal(0, 0.45, n), "px": rng.normal(0, 2, n), "py": rng.normal(0, 2, n),
              "pz": rng.normal(0, 2, n), "w": rng.uniform(0.5, 1.5, n)}
    arrays["g_inv"] = 1 / np.sqrt(1 + arrays["px"] ** 2 + arrays["py"] ** 2 + arrays["pz"] ** 2)
    pcfg = {"charge": -1, "dt": 0.08}
    P1, P2 = Particles(dict(pcfg), comm), Particles(dict(pcfg), comm)
    for P, S in ((P1, S1), (P2, S2)):
        set_particles(P, arrays)
        P.sort_parts(S)                       # the "previous step's" sort
    assert P2.traversal_order_valid(S2)
    # reference sequence
    P1.push_coords("half")
    P1.sort_parts(S1)
    S1.depose_currents([P1])
    # fused
    S2.depose_currents([P2], push_mode="half")
    assert P2.flag_sorted is False
    for k in ("x", "y", "z"):
        assert np.array_equal(P1.DataDev[k].get(), P2.DataDev[k].get()), k
    moved = (P1.DataDev["indx_in_cell"].get() != P2.DataDev["indx_in_cell"].get()).mean()
    assert moved > 0.2                        # the exception path is really exercised
    for k in S1.DataDev:
        if k.startswith(("Jx_m", "Jy_m", "Jz_m")):
            assert rel_err(S2.DataDev[k].get(), S1.DataDev[k].get()) < 1e-12, k
    # and the next sort is the same
    P2.sort_parts(S2)
    check_equal = [np.array_equal(P1.DataDev[k].get(), P2.DataDev[k].get()) for k in INT_KEYS]
    assert all(check_equal)


@pytest.mark.parametrize("M", [0, 1, 2])
def test_one_pass_particle_side_equals_reference_sequence(comm, M):
    """chb_push_depose_push_index (half push + J deposit + second half push + cell
    index / histogram in one pass, then scan + scatter) against the reference sequence
    pic_loop.py:70-76: push_coords, sort_parts, depose_currents, push_coords,
    sort_parts.  Coordinates and every sort product bit-exact, J <= 1e-12; fast
    particles so that the cell-changer queue, its overflow path and the trash bin
    are all exercised."""
    from chimeracl_b200.solver import Solver
    from chimeracl_b200.particles import Particles
    cfg = {"Xmin": -1.0, "Xmax": 1.0, "Nx": 48, "Rmin": 0.0, "Rmax": 1.0, "Nr": 24, "M": M}
    S1, S2 = Solver(dict(cfg), comm), Solver(dict(cfg), comm)
    rng = np.random.default_rng(78 + M)
    n = 150001
    arrays = {"x": rng.uniform(-1.15, 1.15, n), "y": rng.normal(0, 0.45, n),
              "z": rng.normal(0, 0.45, n), "px": rng.normal(0, 2, n), "py": rng.normal(0, 2, n),
              "pz": rng.normal(0, 2, n), "w": rng.uniform(0.5, 1.5, n)}
    arrays["y"][:7] = 0.0
    arrays["z"][:7] = 0.0                     # on-axis particles: guarded 1/r
    arrays["g_inv"] = 1 / np.sqrt(1 + arrays["px"] ** 2 + arrays["py"] ** 2 + arrays["pz"] ** 2)
    pcfg = {"charge": -1, "dt": 0.08}
    P1, P2 = Particles(dict(pcfg), comm), Particles(dict(pcfg), comm)
    for P, S in ((P1, S1), (P2, S2)):
        set_particles(P, arrays)
        P.sort_parts(S)                       # the "previous step's" sort
    # reference sequence
    P1.push_coords("half")
    P1.sort_parts(S1)
    S1.depose_currents([P1])
    P1.push_coords("half")
    P1.sort_parts(S1)
    # one pass + scan/scatter
    S2.depose_currents([P2], push_mode="half+half")
    assert P2.flag_sorted is False and P2._index_prefilled is True
    P2.push_and_sort(S2, mode="half")
    assert P2._index_prefilled is False and P2.flag_sorted is True
    for k in ("x", "y", "z"):
        assert np.array_equal(P1.DataDev[k].get(), P2.DataDev[k].get()), k
    for k in INT_KEYS:
        assert np.array_equal(P1.DataDev[k].get(), P2.DataDev[k].get()), k
    assert int(P1.Args["Np_stay"]) == int(P2.Args["Np_stay"])
    for k in S1.DataDev:
        if k.startswith(("Jx_m", "Jy_m", "Jz_m")):
            assert rel_err(S2.DataDev[k].get(), S1.DataDev[k].get()) < 1e-12, k


def test_diagnostics_records(comm, tmp_path, monkeypatch):
    """Diagnostics.make_record (reference diagnostics.py:57-141) hooked into
    PIC_loop.step(): record layout /data/{info,fields,species}, modes stacked as
    [m0, Re m1, Im m1], particle selection windows and the w2pC weight scaling."""
    import importlib.util
    import os
    from chimeracl_b200.diagnostics import Diagnostics
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "examples", "lpa_script_small.py")
    spec = importlib.util.spec_from_file_location("lpa_small_diag", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _, solver, eons, ions, frame, loop = mod.build(Nx=200, Nr=40, M=1, comm=comm)
    monkeypatch.chdir(tmp_path)
    diag_in = {"Interval": 2, "ScalarFields": ["rho", "Ez"], "VectorFields": ["B"], "w2pC": 2.5,
               "Species": {"Components": ["x", "w", "px"], "Selections": [["px", -0.5, None]]}}
    diag = Diagnostics(diag_in, solver, species=[eons, ions][:1], frame=frame)
    loop.diags = [diag]
    for _ in range(3):
        loop.step()
    files = sorted(os.listdir(tmp_path / "diags"))
    assert files == ["000000000.npz", "000000002.npz"]      # it = 0 and it = 2
    rec = np.load(tmp_path / "diags" / files[1])
    assert int(rec["/data/info/iteration"]) == 2
    assert float(rec["/data/info/FrameVelocity"]) == frame.Args["Velocity"]
    assert rec["/data/info/Xgrid"].shape == (solver.Args["Nx"],)
    Nr, Nx = solver.Args["Nr"], solver.Args["Nx"]
    for name in ("rho", "Ez", "Bx", "By", "Bz"):
        fld = rec["/data/fields/" + name]
        assert fld.shape == (3, Nr, Nx) and fld.dtype == np.float32, name
        assert np.isfinite(fld).all(), name
    assert np.abs(rec["/data/fields/Ez"]).max() > 0.1        # the laser is in the box
    px = rec["/data/species/species_0/px"]
    assert px.dtype == np.float64 and (px > -0.5).all()
    assert rec["/data/species/species_0/x"].shape == px.shape
    # selection and weight scaling against the device data of the same iteration:
    # a record written now (no step in between) must reproduce the current arrays
    diag.Args["Interval"] = 1
    diag.make_record(loop.it)
    now = np.load(tmp_path / "diags" / ("%09d.npz" % loop.it))
    keep = eons.DataDev["px"].get() > -0.5
    assert np.array_equal(now["/data/species/species_0/x"], eons.DataDev["x"].get()[keep])
    assert np.allclose(now["/data/species/species_0/w"], 2.5 * eons.DataDev["w"].get()[keep],
                       rtol=1e-15)
    Ez = solver.DataDev["Ez_m1"].get()
    assert np.array_equal(now["/data/fields/Ez"][1], Ez.real.astype(np.float32))
    assert np.array_equal(now["/data/fields/Ez"][2], Ez.imag.astype(np.float32))


@pytest.mark.parametrize("M,Nx", [(0, 256), (1, 512), (2, 1024)])
def test_fused_damp_fields_equals_three_calls(comm, M, Nx):
    """Solver.damp_fields through chb_fft_damp_x_batched (inverse FFT -> edge profile ->
    FFT, in place and on chip) against the reference's three calls (solver.py:32-35):
    every element sees the same operations, so the spectra must be bit-identical; a
    non-zero Xmin exercises both x-phase tables."""
    from chimeracl_b200.solver import Solver
    cfg = {"Xmin": -3.7, "Xmax": 9.1, "Nx": Nx, "Rmin": 0.0, "Rmax": 5.0, "Nr": 33, "M": M,
           "DampCells": 30}
    S1, S2 = Solver(dict(cfg), comm), Solver(dict(cfg), comm)
    rng = np.random.default_rng(5 + M)
    for k in sorted(S1.DataDev):
        if k[0] in "EG" and "_fb_m" in k:
            a = rng.normal(size=S1.DataDev[k].shape) + 1j * rng.normal(size=S1.DataDev[k].shape)
            S1.DataDev[k][:] = a
            S2.DataDev[k][:] = a
    assert S1.damp_fields_fused(["E", "G"]) is True
    S2.fb_transform(vects=["E", "G"], dir=1, mode="half")
    S2.profile_edges(["E", "G"])
    S2.fb_transform(vects=["E", "G"], dir=0, mode="half")
    checked = 0
    for k in sorted(S1.DataDev):
        if k[0] in "EG" and "_fb_m" in k:
            a, b = S1.DataDev[k].get(), S2.DataDev[k].get()
            assert np.array_equal(a, b), (k, np.abs(a - b).max())
            checked += 1
    assert checked == 6 * (M + 1)
    # and the damping really acted: the x-space edge columns are attenuated
    S2.fb_transform(vects=["E"], dir=1, mode="half")
    ex = np.abs(S2.DataDev["Ex_m0"].get()[1:])
    assert ex[:, 0].max() < 1e-6 * ex[:, Nx // 2].max()


@pytest.mark.parametrize("overlap", [False, True])
def test_host_buffer_step(comm, overlap):
    """host_api.step_from_host (the e2e path of bench.py): attributes uploaded from pinned
    host memory, one PIC step, results downloaded -- with and without the early
    coordinate download on the second stream; the host copies must equal the device
    arrays, and both variants must give the same step."""
    import torch
    from chimeracl_b200 import host_api
    from chimeracl_b200.solver import Solver
    from chimeracl_b200.pic_loop import PIC_loop
    cfg = {"Xmin": -1.0, "Xmax": 1.0, "Nx": 64, "Rmin": 0.0, "Rmax": 1.0, "Nr": 32, "M": 1,
           "DampCells": 8}
    S = Solver(dict(cfg), comm)
    P, _ = _random_species(comm, 50000, (-0.9, 0.9), seed=21, spread=0.3)
    P.sort_parts(S)
    loop = PIC_loop(solvers=[S], species=[P])
    loop.step()
    host_in, host_out = host_api.make_host_buffers(P, S)
    x_before = host_in["x"].numpy().copy()
    h2d, d2h = host_api.step_from_host(loop, P, host_in, host_out, overlap=overlap)
    torch.cuda.synchronize()
    assert h2d == 8 * 8 * P.Args["Np"] and d2h == 7 * 8 * P.Args["Np"] + 8 * 64 * 32
    for a in host_api.ATTRS_OUT:
        assert np.array_equal(host_out[a].numpy(), P.DataDev[a].get()), a
    assert np.array_equal(host_out["rho_m0"].numpy(), S.DataDev["rho_m0"].get())
    assert not np.array_equal(host_out["x"].numpy(), x_before)      # the step moved them
    assert loop.on_coordinates_final is None


def test_host_buffer_pipeline_equals_single_steps(comm):
    """host_api.HostStepPipeline (upload of step k+1 and download of step k overlapped with
    the compute of step k on two device buffer sets): every step's host results must be
    the same as for the same steps run one by one through step_from_host
    (coordinates bit-exact, momenta / rho to the rounding of the atomics' order)."""
    import torch
    from chimeracl_b200 import host_api
    from chimeracl_b200.solver import Solver
    from chimeracl_b200.pic_loop import PIC_loop
    cfg = {"Xmin": -1.0, "Xmax": 1.0, "Nx": 256, "Rmin": 0.0, "Rmax": 1.0, "Nr": 64, "M": 1,
           "DampCells": 8}
    nsteps = 5
    rng = np.random.default_rng(5)

    def fresh():
        S = Solver(dict(cfg), comm)
        P, _ = _random_species(comm, 400000, (-0.9, 0.9), seed=22, spread=0.3)
        P.sort_parts(S)
        loop = PIC_loop(solvers=[S], species=[P])
        loop.step()
        return S, P, loop

    S, P, loop = fresh()
    host_in, out0 = host_api.make_host_buffers(P, S)
    # a different batch every step (the pipeline must not mix them up)
    batches = []
    for k in range(nsteps):
        b = {a: host_in[a].clone().pin_memory() for a in host_api.ATTRS_IN}
        b["px"] += torch.from_numpy(rng.normal(0, 0.01, b["px"].numel()))
        batches.append(b)
    ref = []
    for k in range(nsteps):
        host_api.step_from_host(loop, P, batches[k], out0)
        torch.cuda.synchronize()
        ref.append({a: out0[a].clone() for a in out0})
    S, P, loop = fresh()
    outs = [{a: torch.empty_like(out0[a]).pin_memory() for a in out0} for _ in range(nsteps)]
    pipe = host_api.HostStepPipeline(loop, P)
    for k in range(nsteps):
        h2d, d2h = pipe.submit(batches[k], outs[k])
    pipe.drain()
    torch.cuda.synchronize()
    assert h2d == 8 * 8 * P.Args["Np"] and d2h == 7 * 8 * P.Args["Np"] + 8 * 256 * 64
    for k in range(nsteps):
        for a in ("x", "y", "z"):          # pushed from the uploaded batch: bit-exact
            assert torch.equal(outs[k][a], ref[k][a]), (k, a)
        for a in ("px", "py", "pz", "g_inv", "rho_m0"):   # through the deposits' atomics
            assert rel_err(outs[k][a].numpy(), ref[k][a].numpy()) < 1e-11, (k, a)


def test_cell_changer_queue_overflow_is_detected(comm):
    """When the worst-case queue would take more than a quarter of the device memory the
    fused pass gets a quarter-size queue and the device counter is checked in the SAME
    step, before the deposited current is used: an overflow (incomplete J) must raise
    instead of passing silently; a sufficient queue must not."""
    from chimeracl_b200.solver import Solver
    from chimeracl_b200.particles import Particles
    cfg = {"Xmin": -1.0, "Xmax": 1.0, "Nx": 48, "Rmin": 0.0, "Rmax": 1.0, "Nr": 24, "M": 1}
    rng = np.random.default_rng(91)
    n = 120000

    def make(pscale):
        S = Solver(dict(cfg), comm)
        arrays = {"x": rng.uniform(-0.9, 0.9, n), "y": rng.normal(0, 0.3, n),
                  "z": rng.normal(0, 0.3, n), "px": rng.normal(0, pscale, n),
                  "py": rng.normal(0, pscale, n), "pz": rng.normal(0, pscale, n),
                  "w": rng.uniform(0.5, 1.5, n)}
        arrays["g_inv"] = 1 / np.sqrt(1 + arrays["px"] ** 2 + arrays["py"] ** 2 + arrays["pz"] ** 2)
        P = Particles({"charge": -1, "dt": 0.08}, comm)
        set_particles(P, arrays)
        P._exc_full_limit = 0                  # force the bounded queue
        P.sort_parts(S)
        return S, P

    S, P = make(2.0)                           # ~40 % change cell: 48 k > n/4 + 4096
    with pytest.raises(RuntimeError, match="changed cell"):
        S.depose_currents([P], push_mode="half")      # checked before J is post-processed
    S, P = make(2.0)
    S.depose_currents([P], push_mode="half", defer=True)
    with pytest.raises(RuntimeError, match="changed cell"):
        S.finish_currents()                    # deferred (PIC_loop): checked before J is used
    S, P = make(0.02)                          # slow particles: a few per cent
    S.depose_currents([P], push_mode="half")
    P.exception_workspace()                    # no overflow -> no exception


def test_reference_named_transform_helpers(comm):
    """The reference's per-component helpers (_transform_forward / _backward and the
    half variants, transformer_methods_cl.py:290-455) exist with its signature and give
    what transform_field gives; unknown array namings are refused loudly."""
    from chimeracl_b200.solver import Solver
    cfg = {"Xmin": -1.0, "Xmax": 1.0, "Nx": 64, "Rmin": 0.0, "Rmax": 1.0, "Nr": 24, "M": 1}
    S1, S2 = Solver(dict(cfg), comm), Solver(dict(cfg), comm)
    rng = np.random.default_rng(31)
    a0 = rng.normal(size=(24, 64))
    a1 = rng.normal(size=(24, 64)) + 1j * rng.normal(size=(24, 64))
    for S in (S1, S2):
        S.DataDev["rho_m0"][:] = a0
        S.DataDev["rho_m1"][:] = a1
    S1.transform_field("rho", 0, "full")
    S2._phase(0)
    S2._transform_forward("DHT_m", "rho_m", "rho_fb_m", S2.DataDev["phs_shft"])
    for m in "01":
        assert np.array_equal(S1.DataDev["rho_fb_m" + m].get(), S2.DataDev["rho_fb_m" + m].get())
    S1.transform_field("rho", 1, "half")
    S2._half_transform_backward("DHT_inv_m", "rho_fb_m", "rho_m", None)
    for m in "01":
        assert np.array_equal(S1.DataDev["rho_m" + m].get(), S2.DataDev["rho_m" + m].get())
    with pytest.raises(ValueError):
        S2._transform_forward("DHT_m", "rho_m", "Jx_fb_m", None)


def test_hermitian_contraction_matches_full(comm):
    """chb_dht2_hermitian (columns k <= Nx/2 contracted, the rest mirrored) against
    chb_dht2 on the spectrum of a real field, with complex alphas, accumulate on one
    output and not on the other."""
    from chimeracl_b200.solver import Solver
    cfg = {"Xmin": -1.0, "Xmax": 1.0, "Nx": 256, "Rmin": 0.0, "Rmax": 1.0, "Nr": 41, "M": 1}
    S = Solver(dict(cfg), comm)
    rng = np.random.default_rng(17)
    K, Nx = 40, 256
    spec = np.fft.fft(rng.normal(size=(K, Nx)), axis=1)            # Hermitian in kx
    b = S.DataDev["rho_fb_m0"]
    b[:] = spec
    outs = {}
    for herm in (False, True):
        c1, c2 = S.DataDev["dN1y_fb_m1"], S.DataDev["dN1z_fb_m1"]
        c1[:] = np.zeros((K, Nx), complex)
        c2[:] = np.full((K, Nx), 0.5 - 0.25j)
        S._cdot2(S.DataDev["dDHT_minus_m1"], b, c1, -1.0, False, c2, -1.0j, True, hermitian=herm)
        outs[herm] = (c1.get().copy(), c2.get().copy())
    for k in range(2):
        full, half = outs[False][k], outs[True][k]
        assert rel_err(half, full) < 1e-13, k
    # the mirrored half really is the conjugate image (before alpha): c1 = -A.b
    c1 = outs[True][0]
    assert np.allclose(c1[:, 1:Nx // 2], np.conj(c1[:, :Nx // 2:-1]), rtol=0, atol=1e-12 * np.abs(c1).max())


# ----------------------------------------------------------------------------- kr-row sharded field solve
@pytest.mark.parametrize("M,world", [(0, 2), (1, 2), (1, 3), (1, 0)])
def test_sharded_field_solve_virtual_shards_against_golden(comm, M, world):
    """Two full PIC steps with the field solve split into `world` kr-row shards run one
    after the other in this process (Solver.enable_spectral_sharding(emulate=True)):
    row-sliced operator matrices / FFT batches / PSATD ranges and the partial-sum
    backward transform must reproduce the golden run (K = 13: shards of 8 and 5 rows,
    world = 3 has an empty shard; world = 0: the non-emulated code path of a single rank,
    which owns every row)."""
    from chimeracl_b200.pic_loop import PIC_loop
    G = load_golden(M)
    S, P, I = gpu_case_from_golden(G, comm)
    if world:
        S.enable_spectral_sharding(world=world, emulate=True)
    else:
        S.enable_spectral_sharding()
    loop = PIC_loop(solvers=[S], species=[P, I], frames=[], diags=[])
    loop.step()
    for k in G.files:
        if k.startswith("step1/S/"):
            assert rel_err(S.DataDev[k[8:]].get(), G[k]) < 1e-10, k
        elif k.startswith("step1/P/"):
            assert rel_err(P.DataDev[k[8:]].get(), G[k]) < 1e-10, k
    loop.step()
    P.align_parts()
    for k in G.files:
        if k.startswith("step2_aligned/S/"):
            assert rel_err(S.DataDev[k[16:]].get(), G[k]) < 1e-10, k


@pytest.mark.parametrize("Nx,Nr,M,world", [(512, 258, 1, 8),     # K = 257, R = 40, one empty shard
                                           (300, 48, 2, 4),      # Bluestein FFT, three-call damping
                                           (4096, 512, 1, 8)])   # cfg3 shape: 7 x 64 + 63 rows
def test_sharded_field_solve_equals_replicated(comm, Nx, Nr, M, world):
    """Field solve only (deposited grids -> E, B grids), seeded random state, two steps:
    virtual kr-row shards against the unsharded solve on the same device.  Covers the
    contraction shapes the sharding produces at scale (64-row outputs, 64-deep
    contractions with lda = 512, operands at row / column offsets)."""
    from chimeracl_b200.solver import Solver
    from test_sharded_solve_cpu import _cfg, _random_state, _Loop, _results
    outs = []
    for sharded in (False, True):
        S = Solver(dict(_cfg(Nx, Nr, M)), comm)
        _random_state(S, 23)
        if sharded:
            S.enable_spectral_sharding(world=world, emulate=True)
        for _ in range(2):
            if sharded:
                _Loop().solve_sharded(S)
            else:
                _Loop().solve_plain(S)
        outs.append(_results(S))
        del S
    for k in outs[0]:
        assert rel_err(outs[1][k], outs[0][k]) < 1e-11, k


def test_cpu_binding_is_best_effort(comm):
    """parallel.bind_to_gpu_cpus (multi-rank runs: NUMA-local pinned buffers) never raises
    and never leaves the process without CPUs."""
    import os
    from chimeracl_b200.parallel import bind_to_gpu_cpus
    before = os.sched_getaffinity(0)
    try:
        ok = bind_to_gpu_cpus(comm.device)
        assert ok in (True, False)
        assert len(os.sched_getaffinity(0)) >= 1
    finally:
        os.sched_setaffinity(0, before)


# ----------------------------------------------------------------------------- CUDA graph
def test_cuda_graph_replay_matches_eager(comm):
    """PIC_loop(use_cuda_graph=True): two captured steps replayed alternately (the dN0/dN1
    swap has period 2) give what eager steps give -- compared after 7 steps (odd: both
    captured graphs and the host-side swap bookkeeping are exercised) on the golden case,
    and against the oracle.  Tolerance 1e-10 (the full-step tolerance: deposit REDs are
    unordered in both runs)."""
    from chimeracl_b200.pic_loop import PIC_loop
    G = load_golden(1)
    runs = []
    for graph in (False, True):
        S, P, I = gpu_case_from_golden(G, comm)
        loop = PIC_loop(solvers=[S], species=[P, I], use_cuda_graph=graph)
        for _ in range(7):
            loop.step()
        comm.synchronize()
        runs.append((S, P, loop))
    (S0, P0, _), (S1, P1, loop1) = runs
    assert loop1.graph_captures == 1 and loop1.graph_replays == 5   # steps 0, 1 ran eagerly
    assert int(P1.Args["Np_stay"]) == int(P0.Args["Np_stay"])
    assert np.array_equal(P1.DataDev["sort_indx"].get(), P0.DataDev["sort_indx"].get())
    for k in ("Ex_m0", "Ez_m1", "Bz_m1", "rho_m0", "Jx_m1", "Ex_fb_m1", "dN0x_fb_m1", "dN1x_fb_m0"):
        assert rel_err(S1.DataDev[k].get(), S0.DataDev[k].get()) < 1e-10, k
    for k in ("x", "y", "z", "px", "py", "pz", "g_inv"):
        assert rel_err(P1.DataDev[k].get(), P0.DataDev[k].get()) < 1e-10, k
    So, Po, Io = oracle_case_from_golden(G, NumpyKernels(1))
    for _ in range(7):
        O.pic_step(So, [Po, Io])
    for k in ("Ex_m0", "Bz_m1", "rho_m0"):
        assert rel_err(S1.DataDev[k].get(), So.D[k]) < 1e-9, k


def test_cuda_graph_with_moving_window(comm):
    """Graph replay across plasma injections (reduced lpa_script_small, injections at steps
    0, 20, ..., 120): every injection changes particle counts and array addresses, so that step
    and the next run eagerly and the pair of graphs is re-captured."""
    import importlib.util
    import os
    from chimeracl_b200.methods.generic_methods_cl import Communicator
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "examples", "lpa_script_small.py")
    spec = importlib.util.spec_from_file_location("lpa_small_g", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    res = []
    for graph in (False, True):
        c = Communicator(answers=[0, 0], seed=11)     # same theta draws in both runs
        _, solver, eons, ions, frame, loop = mod.build(Nx=300, Nr=48, M=1, comm=c)
        loop.use_cuda_graph = graph
        # 140 steps: the plasma has reached the laser; later this reduced run amplifies
        # rounding noise by ~100x per 10 steps (two eager runs diverge the same way:
        # tools/graph_lockstep.py with BOTH_EAGER=1)
        for _ in range(140):
            loop.step()
        c.synchronize()
        res.append((solver, eons, loop))
    (s0, e0, _), (s1, e1, l1) = res
    assert l1.graph_captures == 7 and l1.graph_replays == 140 - 7
    assert int(e1.Args["Np"]) == int(e0.Args["Np"])
    # scale: the laser field (rho and the m = 1 longitudinal fields are rounding noise while
    # the plasma that has entered still has ~zero weight, so their own maximum is no scale)
    names = [c + a + m for c in "EB" for a in "xyz" for m in ("_m0", "_m1")]
    scale = max(float(np.abs(s0.DataDev[k].get()).max()) for k in names)
    assert scale > 1.0
    for k in names:
        d = np.abs(s1.DataDev[k].get() - s0.DataDev[k].get()).max()
        assert d / scale < 1e-8, (k, d / scale)
    assert rel_err(e1.DataDev["x"].get(), e0.DataDev["x"].get()) < 1e-9
    # momenta of near-axis particles amplify the rounding noise of the m = 1 fields (1/r)
    for k in ("px", "py", "pz", "g_inv"):
        assert rel_err(e1.DataDev[k].get(), e0.DataDev[k].get()) < 1e-6, k


# ----------------------------------------------------------------------------- periodic align
def test_align_every_matches_oracle(comm):
    """PIC_loop(align_every=2): the loop calls sort_parts + align_parts itself at steps 2 and
    4; the oracle gets the same calls at the same places.  Integer products bit-exact,
    fields / momenta to the full-step tolerance 1e-10."""
    from chimeracl_b200.pic_loop import PIC_loop
    G = load_golden(1)
    S, P, I = gpu_case_from_golden(G, comm)
    loop = PIC_loop(solvers=[S], species=[P, I], align_every=2)
    So, Po, Io = oracle_case_from_golden(G, NumpyKernels(1))
    for it in range(5):
        loop.step()
        if it > 0 and it % 2 == 0:
            Po.sort_parts(So)
            Po.align_parts()
        O.pic_step(So, [Po, Io])
    comm.synchronize()
    assert int(P.Args["Np"]) == Po.Args["Np"]
    assert np.array_equal(P.DataDev["sort_indx"].get(), Po.D["sort_indx"])
    for k in ("Ex_m0", "Ez_m1", "Bz_m1", "rho_m0", "Jx_m1"):
        assert rel_err(S.DataDev[k].get(), So.D[k]) < 1e-10, k
    for k in ("x", "y", "z", "px", "py", "pz", "g_inv", "w"):
        assert rel_err(P.DataDev[k].get(), Po.D[k]) < 1e-10, k
