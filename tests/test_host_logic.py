"""CPU tests (no GPU): the C-ABI library loads and exports every symbol declared in
include/chimera_b200.h; host-side math of the wrappers (geometry, spectral axes, DHT
matrices, PSATD coefficients, damping profile, FFT plans) agrees with the oracle's
restatement of the reference formulas; lazy Args; sharding helpers."""
import os
import re

import numpy as np
import pytest

from oracle import orchestration as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from chimeracl_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "chimera_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(chb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = _lib.load()                      # raises if the .so is missing
    bound = set(_lib.SIGNATURES)
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert getattr(lib._cdll, name) is not None
    assert lib.chb_version() >= 100
    assert b"invalid argument" in lib.chb_error_string(-1)
    # pure host-side helper entry points (no device needed)
    assert lib.chb_cell_offsets_workspace_bytes(10 ** 6) >= 4 * (10 ** 6 // 4096)
    assert lib.chb_sort_workspace_bytes(1000, 100) >= 4 * 1000
    assert lib.chb_fft_max_pow2() == 16384


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from chimeracl_b200.methods.generic_methods_cl import Communicator
    with pytest.raises(RuntimeError):
        Communicator(answers=[0, 0])


@pytest.mark.parametrize("M", [0, 1, 2])
def test_host_math_matches_reference_formulas(M):
    from chimeracl_b200.grid import grid_geometry
    from chimeracl_b200.transformer import spectral_axes, hankel_matrices
    from chimeracl_b200.solver import psatd_coefficients
    from chimeracl_b200.methods.solver_methods_cl import damping_profile
    cfg = {"Xmin": -3.0, "Xmax": 5.0, "Nx": 48, "Rmin": 0.0, "Rmax": 2.5, "Nr": 21, "M": M,
           "DampCells": 7}
    cfg["dt"] = (cfg["Xmax"] - cfg["Xmin"]) / cfg["Nx"]
    A = psatd_coefficients(hankel_matrices(spectral_axes(grid_geometry(dict(cfg)))))
    R = O.spectral_args(O.grid_args(dict(cfg)))
    for k, v in R.items():
        if k == "DampProfile":
            continue                      # built by init_solver_methods, checked below
        if isinstance(v, np.ndarray):
            assert k in A, k
            assert np.array_equal(A[k], v), k
        elif isinstance(v, (int, float)) and k in A:
            assert A[k] == v, k
    assert np.array_equal(damping_profile(7), R["DampProfile"])
    # DHT_inv . DHT = I (what the reference's round-trip check pins)
    for m in range(M + 1):
        eye = A["DHT_inv_m%d" % m].dot(A["DHT_m%d" % m])
        assert np.abs(eye - np.eye(eye.shape[0])).max() < 1e-9


@pytest.mark.parametrize("n", [8, 64, 900, 30, 1537])
def test_fft_plan_tables(n):
    """The Bluestein tables really evaluate a DFT (checked with NumPy)."""
    from chimeracl_b200.methods.transformer_methods_cl import fft_plan_tables
    L, tw, chirp, bfft = fft_plan_tables(n)
    assert L >= 8 and L & (L - 1) == 0
    assert np.allclose(tw, np.exp(-2j * np.pi * np.arange(L) / L))
    L2, tw2, _, _ = fft_plan_tables(16384)
    assert L2 == 16384 and tw2.size == 16384
    assert np.allclose(tw2[8192:], np.exp(-2j * np.pi * np.arange(8192) / 16384))
    if chirp is None:
        assert L == n
        return
    rng = np.random.default_rng(n)
    x = rng.normal(size=n) + 1j * rng.normal(size=n)
    a = np.zeros(L, dtype=complex)
    a[:n] = x * chirp
    conv = np.fft.ifft(np.fft.fft(a) * (bfft * L))
    assert np.abs(conv[:n] * chirp - np.fft.fft(x)).max() < 1e-10 * np.abs(x).sum()


def test_particle_defaults_and_lazy_args():
    from chimeracl_b200.methods.generic_methods_cl import ArgsDict
    a = ArgsDict({"x": 1})
    calls = []
    a.set_lazy("Np_stay", lambda: calls.append(1) or 42)
    assert not calls
    assert a["Np_stay"] == 42 and a["Np_stay"] == 42 and len(calls) == 1
    assert a.get("Np_stay") == 42 and a.get("nope", 7) == 7
    R = O.particle_args({"Nppc": (2, 2, 4), "dx": 0.1, "dr": 0.2, "dt": 0.05, "dens": 0.3,
                         "charge": -1})
    assert R["FactorPush"] == 2 * np.pi * 0.05 * -1 / 1.0
    assert R["w0"] == 2 * np.pi * 0.1 * 0.2 * 0.3 / 16


def test_shard_range_partitions_everything():
    from chimeracl_b200.parallel import shard_range
    for n in (0, 1, 7, 33480720):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a1 - a0 >= b1 - b0 >= 0


def test_dht_tile_width_fills_whole_waves():
    """chb_dht_tile_columns (host-only): the wide contraction kernel picks, per launch,
    the tile width whose tile count wastes the least of the 148-SM waves.  cfg3 (Nr=512,
    Nx=4096): 112 columns -> exactly 148 tiles (real) / 296 (complex) per contraction."""
    from chimeracl_b200 import _lib
    lib = _lib.load()
    rows = 511

    def waste(width, n, nbatch):
        tiles = -(-n // width) * -(-rows // 128) * nbatch
        waves = -(-tiles // 148)
        return waves * width

    for n, nbatch in ((4096, 1), (8192, 1), (4096, 4), (8192, 6), (32768, 3), (1800, 1)):
        w = lib.chb_dht_tile_columns(rows, n, nbatch)
        assert w in (64, 80, 96, 112, 128)
        assert waste(w, n, nbatch) == min(waste(c, n, nbatch) for c in (64, 80, 96, 112, 128))
    assert lib.chb_dht_tile_columns(rows, 4096, 1) == 112
    assert lib.chb_dht_tile_columns(rows, 8192, 1) == 112
    assert -(-8192 // 112) * 4 == 296          # two full waves
    # the worst-case queue of the one-pass particle side: one 64-byte record per particle
    assert lib.chb_push_depose_workspace_bytes(1000) == 16 + 64 * 1000


def test_injected_slab_is_dealt_to_ranks_by_radial_bands():
    """Multi-GPU moving window: make_new_domain(..., r_shard=(rank, world)) creates one
    band of radial cell rows per rank; the union of the bands is exactly the lattice a
    single rank creates (x, r, weight; the per-cell theta offsets are random), with
    equal particle counts per row (reference lattice: kernels/particles_generic.cl:33-84)."""
    import torch
    from cabi_emulator import EmulatedComm
    from chimeracl_b200.particles import Particles

    def make(shard):
        P = Particles({'Nppc': (2, 2, 4), 'dx': 0.25, 'dr': 0.125, 'dt': 0.25, 'dens': 0.02,
                       'charge': -1}, EmulatedComm())
        dom = {'Xmin': 3.0, 'Xmax': 5.0, 'Rmin': 0.0, 'Rmax': 2.5, 'dpx': 0.01}
        if shard is not None:
            dom['r_shard'] = shard
        P.make_new_domain(dom, density_profiles=[{'coord': 'x', 'points': [0, 3.5, 4.5, 9],
                                                  'values': [0, 0, 1, 1]}])
        D = P.DataDev
        r = torch.sqrt(D['y_new'].t ** 2 + D['z_new'].t ** 2)
        assert D['px_new'].size == D['x_new'].size == D['g_inv_new'].size
        return D['x_new'].t.numpy(), r.numpy(), D['w_new'].t.numpy(), P.Args['right_lim']

    x, r, w, lim = make(None)
    world = 3
    parts = [make((rank, world)) for rank in range(world)]
    assert all(p[3] == lim for p in parts)
    counts = [p[0].size for p in parts]
    assert sum(counts) == x.size and max(counts) - min(counts) <= 9 * 16   # one cell row
    xs, rs, ws = (np.concatenate([p[i] for p in parts]) for i in range(3))
    key = lambda a, b: np.lexsort((np.round(b, 9), np.round(a, 9)))        # noqa: E731
    i0, i1 = key(x, r), key(xs, rs)
    assert np.array_equal(x[i0], xs[i1])
    assert np.allclose(r[i0], rs[i1], rtol=1e-13, atol=1e-15)
    assert np.allclose(w[i0], ws[i1], rtol=1e-13, atol=0)


def test_argsdict_never_leaks_pending_values():
    """Args['Np_stay'] is produced lazily (asynchronous read-back); every read path of the
    dict -- items(), values(), copy(), dict(), ** -- must hand out the number."""
    from chimeracl_b200.methods.generic_methods_cl import ArgsDict
    calls = []

    def make(v):
        def fn():
            calls.append(v)
            return v
        return fn
    a = ArgsDict({"Np": 10})
    a.set_lazy("Np_stay", make(7))
    assert dict(a) == {"Np": 10, "Np_stay": 7}
    a.set_lazy("Np_stay", make(8))
    assert sorted(a.items()) == [("Np", 10), ("Np_stay", 8)]
    a.set_lazy("Np_stay", make(9))
    assert 9 in list(a.values())
    a.set_lazy("Np_stay", make(11))
    assert {**a}["Np_stay"] == 11
    a.set_lazy("Np_stay", make(12))
    assert a.copy()["Np_stay"] == 12 and a.pop("Np_stay") == 12
    assert calls == [7, 8, 9, 11, 12]            # each resolved exactly once
    assert a.get("missing", 3) == 3
