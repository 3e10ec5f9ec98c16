"""CPU tests of the oracle itself: the NumPy restatement against (a) the committed
golden vectors produced from the reference's own kernels and (b) oracle/_ref when
that library is present, plus the two property checks the reference ships as
scripts (examples/test_transformer.py, examples/test_laser_veloc.py)."""
import numpy as np
import pytest

from oracle import orchestration as O
from oracle.np_kernels import NumpyKernels
from oracle.ref_kernels import RefKernels, ref_available

from helpers import ATTR, load_golden, oracle_case_from_golden, rel_err

INT_KEYS = ("indx_in_cell", "sum_in_cell", "cell_offset", "sort_indx")


def _check_stage(G, stage, S, P, tol):
    for k in G.files:
        if not k.startswith(stage + "/"):
            continue
        parts = k.split("/")
        if parts[1] == "S":
            got = S.D[parts[2]]
        elif parts[1] == "P":
            got = P.D[parts[2]]
        else:
            assert int(G[k]) == P.Args["Np_stay"], k
            continue
        if parts[2] in INT_KEYS:
            assert np.array_equal(got, G[k]), k
        else:
            assert rel_err(got, G[k]) <= tol, (k, rel_err(got, G[k]))


@pytest.mark.parametrize("M", [0, 1])
def test_restatement_matches_golden_step(M):
    """One full PIC step, stage by stage (pic_loop.py:57-142)."""
    G = load_golden(M)
    S, P, I = oracle_case_from_golden(G, NumpyKernels(M))
    for p in (P, I):
        p.push_coords("half")
        p.sort_parts(S)
    _check_stage(G, "sort1", S, P, 0.0)      # push is bit-exact, indices exact
    S.depose_currents([P, I])
    for p in (P, I):
        p.push_coords("half")
        p.sort_parts(S)
    S.depose_charge([P, I])
    _check_stage(G, "depose", S, P, 1e-13)
    S.fb_transform(scals=["rho"], vects=["J"], dir=0)
    S.fields_smooth(["rho", "Jx", "Jy", "Jz"])
    for m in range(S.M + 1):
        for c in "xyz":
            S.D["dN0%s_fb_m%d" % (c, m)][...] = S.D["dN1%s_fb_m%d" % (c, m)]
    S.field_grad("rho", "dN1")
    _check_stage(G, "grad", S, P, 1e-12)
    S.push_fields()
    S.damp_fields()
    S.restore_B_fb()
    S.fb_transform(vects=["E", "B"], dir=1)
    S.gather_and_push([P, I])
    _check_stage(G, "step1", S, P, 1e-12)
    O.pic_step(S, [P, I])
    P.align_parts()
    _check_stage(G, "step2_aligned", S, P, 1e-11)


@pytest.mark.skipif(not ref_available(1), reason="oracle/_ref not built")
@pytest.mark.parametrize("M", [0, 1])
def test_restatement_matches_reference_kernels(M):
    """Kernel by kernel against the reference's own OpenCL source run on the host."""
    G = load_golden(M)
    Kn, Kr = NumpyKernels(M), RefKernels(M)
    out = []
    for K in (Kn, Kr):
        S, P, I = oracle_case_from_golden(G, K)
        rng = np.random.default_rng(5)
        for k in sorted(S.D):
            if k[0] in "EB" and "_fb_" not in k:
                a = rng.normal(size=S.D[k].shape)
                S.D[k][...] = a if S.D[k].dtype == np.float64 else a + 1j * rng.normal(size=a.shape)
        P.push_coords("half")
        P.sort_parts(S)
        I.sort_parts(S)
        S.depose_currents([P, I])
        S.depose_charge([P, I])
        S.gather_and_push([P, I])
        out.append((S, P))
    (Sn, Pn), (Sr, Pr) = out
    for k in INT_KEYS:
        assert np.array_equal(Pn.D[k], Pr.D[k]), k
    for k in ("x", "y", "z"):
        assert np.array_equal(Pn.D[k], Pr.D[k]), k      # push_xyz bit-exact
    for k in ("px", "py", "pz", "g_inv"):
        assert np.array_equal(Pn.D[k], Pr.D[k]), k      # gather + Boris bit-exact
    for k in Sn.D:
        if k.startswith(("rho_m", "J")) and "_fb_" not in k:
            assert rel_err(Sn.D[k], Sr.D[k]) < 1e-14, k


def test_sort_edge_cases():
    """Empty species, all-trash species, single cell crowding."""
    K = NumpyKernels(1)
    cfg = {"Xmin": 0.0, "Xmax": 1.0, "Nx": 8, "Rmin": 0.0, "Rmax": 1.0, "Nr": 6, "M": 1}
    S = O.OracleSolver(cfg, K)
    P = O.OracleParticles({"charge": -1}, K)
    P.sort_parts(S)                       # Np == 0: no-op (particles.py:24-25)
    assert P.flag_sorted and "sort_indx" not in P.D
    n = 50
    far = np.full(n, 10.0)
    P.set_particles(x=far, y=far, z=far, px=far, py=far, pz=far, w=far, g_inv=far)
    P.sort_parts(S)
    assert P.Args["Np_stay"] == 0
    assert np.array_equal(P.D["sort_indx"], np.arange(n))
    assert P.D["sum_in_cell"][-1] == n
    one = np.full(n, 0.5)
    P.set_particles(x=one, y=one * 0.3, z=one * 0.3, px=one, py=one, pz=one, w=one, g_inv=one)
    P.sort_parts(S)
    assert P.Args["Np_stay"] == n and P.D["sum_in_cell"].max() == n


def test_transformer_round_trip_property():
    """examples/test_transformer.py:27-43 at reduced size: rho deposited from a
    Gaussian beam survives forward+backward Fourier-Bessel transform."""
    M = 1
    K = NumpyKernels(M)
    cfg = {"Xmin": -1.0, "Xmax": 1.0, "Nx": 256, "Rmin": 0.0, "Rmax": 1.0, "Nr": 64, "M": M}
    S = O.OracleSolver(dict(cfg), K)
    P = O.OracleParticles(dict(cfg), K)
    rng = np.random.default_rng(0)
    n = 200000
    x, y, z = rng.normal(0, 0.3, n), rng.normal(0.2, 0.3, n), rng.normal(0.2, 0.3, n)
    zero = np.zeros(n)
    P.set_particles(x=x, y=y, z=z, px=zero, py=zero, pz=zero, w=np.full(n, 1.0 / n),
                    g_inv=np.ones(n))
    P.sort_parts(S)
    P.align_parts()
    S.depose_charge([P])
    t0, t1 = S.D["rho_m0"].copy(), S.D["rho_m1"].copy()
    S.fb_transform(scals=["rho"], dir=0)
    S.D["rho_m0"][...] = 0
    S.D["rho_m1"][...] = 0
    S.fb_transform(scals=["rho"], dir=1)
    err = (np.abs(S.D["rho_m0"] - t0)[1:] / np.abs(t0[1:]).max()
           + np.abs(S.D["rho_m1"] - t1)[1:] / np.abs(t1[1:]).max()).max()
    assert err < 1e-12, err


def test_laser_group_velocity_property():
    """examples/test_laser_veloc.py:9-59 at reduced size (M=0, vacuum): centroid
    velocity deficit of a Gaussian pulse ~ (2 pi R)^-2."""
    K = NumpyKernels(0)
    cfg = {"Xmin": -40.0, "Xmax": 40.0, "Rmax": 40.0, "M": 0}
    dx, dr, cfg["dt"] = 0.1, 0.5, 0.1
    cfg["Nx"] = int((cfg["Xmax"] - cfg["Xmin"]) / dx) // 2 * 2
    cfg["Nr"] = int(cfg["Rmax"] / dr) // 2 * 2 + 1
    S = O.OracleSolver(cfg, K)
    laser = {"k0": 1.0, "a0": 1.0, "x0": 0, "Lx": 8.0, "R": 8.0, "x_foc": 10.0}
    O.add_gaussian_pulse(S, laser)
    xc = []
    for _ in range(40):
        S.push_fields()
        S.fb_transform(scals=["Ez"], dir=1)
        v = S.D["Ez_m0"]
        Px = (S.Args["Rgrid"][1:, None] * v[1:, :] ** 2).sum(0)
        xc.append((S.Args["Xgrid"] * Px).sum() / Px.sum())
    xc = np.array(xc)
    veloc = 1 - (xc[1:] - xc[:-1]) / S.Args["dt"]
    theory = (2.0 * np.pi * laser["R"]) ** -2
    assert abs(veloc.mean() - theory) / theory < 0.1, (veloc.mean(), theory)


@pytest.mark.skipif(not ref_available(1), reason="oracle/_ref not built")
def test_particle_creation_kernels_match_reference():
    """fill_grid and profile_by_interpolant (init / injection path,
    kernels/particles_generic.cl:6-84) against the reference kernels."""
    Kn, Kr = NumpyKernels(1), RefKernels(1)
    rng = np.random.default_rng(3)
    xg = -2.0 + 0.25 * np.arange(9)
    rg = 0.1 * np.arange(7)
    th = rng.uniform(0, 2 * np.pi, (xg.size - 1) * (rg.size - 1))
    a = Kn.fill_grid(th, xg, rg, (2, 3, 4))
    b = Kr.fill_grid(th, xg, rg, (2, 3, 4))
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    x = a[0].copy()
    x_loc = np.array([-3.0, -1.0, 0.5, 3.0])
    f_loc = np.array([0.0, 0.0, 1.0, 2.0])
    dxm1 = 1.0 / (x_loc[1:] - x_loc[:-1])
    w1, w2 = a[3].copy(), a[3].copy()
    Kn.profile_by_interpolant(x, w1, x_loc, f_loc, dxm1)
    Kr.profile_by_interpolant(x, w2, x_loc, f_loc, dxm1)
    assert np.array_equal(w1, w2)


def test_mode2_generalisation_reduces_to_mode1():
    """M=2 particle terms have no reference source (parity unpinned): the
    generalisation must reduce to the M=1 restatement when nothing lives in m=2."""
    cfg = {"Xmin": -1.0, "Xmax": 1.0, "Nx": 20, "Rmin": 0.0, "Rmax": 1.0, "Nr": 12}
    rng = np.random.default_rng(8)
    n = 4000
    arrays = dict(x=rng.uniform(-1, 1, n), y=rng.normal(0, 0.4, n), z=rng.normal(0, 0.4, n),
                  px=rng.normal(0, 1, n), py=rng.normal(0, 1, n), pz=rng.normal(0, 1, n),
                  w=rng.uniform(0.5, 1, n))
    arrays["g_inv"] = 1 / np.sqrt(1 + arrays["px"] ** 2 + arrays["py"] ** 2 + arrays["pz"] ** 2)
    res = {}
    for M in (1, 2):
        K = NumpyKernels(M)
        S = O.OracleSolver(dict(cfg, M=M), K)
        P = O.OracleParticles({"charge": -1, "dt": 0.05}, K)
        P.set_particles(**arrays)
        P.sort_parts(S)
        S.depose_currents([P])
        S.depose_charge([P])
        r2 = np.random.default_rng(9)
        for k in sorted(S.D):
            if k[0] in "EB" and "_fb_" not in k and not k.endswith("_m2"):
                a = r2.normal(size=S.D[k].shape)
                S.D[k][...] = a if S.D[k].dtype == np.float64 else a + 1j * r2.normal(size=a.shape)
        S.gather_and_push([P])
        res[M] = (S, P)
    (S1, P1), (S2, P2) = res[1], res[2]
    for k in S1.D:
        if k.startswith(("rho_m", "J")) and "_fb_" not in k:
            assert np.array_equal(S1.D[k], S2.D[k]), k
    assert np.abs(S2.D["rho_m2"]).max() > 0
    for k in ("px", "py", "pz", "g_inv"):
        assert np.array_equal(P1.D[k], P2.D[k]), k
