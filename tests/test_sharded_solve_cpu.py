"""CPU tests of the HOST orchestration of the field solve (transformer / solver mixins,
PIC_loop's phase sequence, the kr-row sharding and its collectives) with the C-ABI calls
served by tests/cabi_emulator.py (NumPy on host memory; test infrastructure only):

  * unsharded host path == golden vectors generated from oracle/_ref (pins the emulator),
  * virtual shards in one process == unsharded (row arithmetic, empty shards, both
    damp_fields flavours, M = 0, 1, 2),
  * two real ranks over gloo == unsharded (the all-gathers and the partial-sum exchange).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import load_golden, golden_cfgs, rel_err
from cabi_emulator import make_solver


class _Loop:
    """The two PIC_loop members the field solve uses, without species."""

    def __init__(self):
        from chimeracl_b200.pic_loop import PIC_loop
        self.loop = PIC_loop.__new__(PIC_loop)
        self.loop.timit = False
        self.loop.real_m0_symmetry = True

    def solve_sharded(self, S):
        self.loop._solve_fields_sharded(S)

    def solve_plain(self, S, real_m0=True):
        """The unsharded phase sequence of PIC_loop.step() between the deposits and the
        gather."""
        S.fb_transform(scals=['rho'], vects=['J'], dir=0, smooth=True)
        for m in range(S.Args['M'] + 1):
            for c in 'xyz':
                k = c + '_fb_m' + str(m)
                S.DataDev['dN0' + k], S.DataDev['dN1' + k] = S.DataDev['dN1' + k], S.DataDev['dN0' + k]
        S.m0_spectra_of_real_fields = real_m0
        S.field_grad('rho', 'dN1')
        S.push_fields()
        S.damp_fields()
        S.restore_B_fb()
        S.m0_spectra_of_real_fields = False
        S.fb_transform(vects=['E', 'B'], dir=1)


def _load_golden_state(S, G):
    for k in G.files:
        if k.startswith("in/S/") or k.startswith("depose/S/"):
            S.DataDev[k.split("/S/")[1]][:] = G[k]


def _random_state(S, seed):
    """Real m = 0 / complex m > 0 grids for rho, J and previous-step spectra for E, G, dN1
    that are spectra of real fields for m = 0 (what PIC_loop maintains)."""
    rng = np.random.default_rng(seed)
    Nr, Nx, M = S.Args['Nr'], S.Args['Nx'], S.Args['M']
    for name in ['rho'] + ['J' + c for c in 'xyz']:
        for m in range(M + 1):
            a = rng.standard_normal((Nr, Nx))
            if m:
                a = a + 1j * rng.standard_normal((Nr, Nx))
            S.DataDev['%s_m%d' % (name, m)][:] = a
    for name in [f + c for f in ('E', 'G', 'dN1') for c in 'xyz']:
        for m in range(M + 1):
            a = rng.standard_normal((Nr - 1, Nx))
            if m:
                a = a + 1j * rng.standard_normal((Nr - 1, Nx))
            S.DataDev['%s_fb_m%d' % (name, m)][:] = np.fft.fft(a, axis=1)


def _results(S):
    out = {}
    for f in 'EB':
        for c in 'xyz':
            for m in range(S.Args['M'] + 1):
                k = '%s%s_m%d' % (f, c, m)
                out[k] = S.DataDev[k].get()[1:].copy()
    for f in ('E', 'G', 'B', 'dN1'):
        for c in 'xyz':
            for m in range(S.Args['M'] + 1):
                k = '%s%s_fb_m%d' % (f, c, m)
                out[k] = S.DataDev[k].get().copy()
    return out


def test_host_field_solve_matches_golden():
    G = load_golden(1)
    cfg, _ = golden_cfgs(G)
    S = make_solver(cfg)
    _load_golden_state(S, G)
    _Loop().solve_plain(S)
    n = 0
    for k in G.files:
        if k.startswith("step1/S/") and k[8] in "EB" and "_fb_" not in k:
            assert rel_err(S.DataDev[k[8:]].get()[1:], G[k][1:]) < 1e-11, k
            n += 1
    assert n == 12


CASES = [
    # Nx, Nr, M, world   (Nx < 256: three-call damp_fields; Nx = 256: the fused pass)
    (32, 14, 1, 2),
    (32, 14, 1, 3),      # K = 13, R = 8: the third shard is empty
    (48, 30, 0, 2),      # Bluestein plan, M = 0
    (256, 42, 1, 4),     # K = 41, R = 16: [0,16) [16,32) [32,41) and an empty one
    (32, 19, 2, 2),
    (32, 14, 1, 0),      # world 0: the non-emulated path of a single rank owning every row
]


def _cfg(Nx, Nr, M):
    cfg = {'Xmin': -3.0, 'Xmax': 3.5, 'Nx': Nx, 'Rmin': 0.0, 'Rmax': 2.0, 'Nr': Nr, 'M': M,
           'DampCells': 5}
    cfg['dt'] = (cfg['Xmax'] - cfg['Xmin']) / Nx
    return cfg


@pytest.mark.parametrize("Nx,Nr,M,world", CASES)
def test_virtual_shards_equal_unsharded(Nx, Nr, M, world):
    ref = make_solver(_cfg(Nx, Nr, M))
    _random_state(ref, 7)
    _Loop().solve_plain(ref)
    want = _results(ref)

    S = make_solver(_cfg(Nx, Nr, M))
    _random_state(S, 7)
    if world:
        S.enable_spectral_sharding(world=world, emulate=True)
    else:
        S.enable_spectral_sharding()
    _Loop().solve_sharded(S)
    got = _results(S)
    for k in want:
        assert rel_err(got[k], want[k]) < 1e-12, k


def test_spectral_rows_partition():
    from chimeracl_b200.parallel import spectral_rows
    for K in (1, 13, 41, 63, 251, 511, 1023):
        for world in (1, 2, 3, 4, 8):
            R = spectral_rows(K, 0, world)[2]
            assert R % 8 == 0 and R * world >= K
            cover = []
            for r in range(world):
                lo, hi, R2 = spectral_rows(K, r, world)
                assert R2 == R and lo % 8 == 0 or lo == K
                assert hi - lo <= R
                cover += list(range(lo, hi))
            assert cover == list(range(K))
    assert spectral_rows(511, 7, 8) == (448, 511, 64)


# ------------------------------------------------------------------ two real ranks (gloo)
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, Nx, Nr, M):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    from chimeracl_b200.parallel import init_distributed
    pg = init_distributed(backend="gloo")
    S = make_solver(_cfg(Nx, Nr, M), process_group=pg)
    _random_state(S, 11)             # rho / J after their all-reduce: equal on all ranks
    S.enable_spectral_sharding()
    lo, hi = S._shard.lo, S._shard.hi
    # rows a rank does not own may hold anything before the step
    for name in [f + c for f in ('E', 'G', 'dN1') for c in 'xyz']:
        for m in range(M + 1):
            a = S.DataDev['%s_fb_m%d' % (name, m)].t
            a[:lo] = 777.0
            a[hi:] = -777.0
    for step in range(2):
        _Loop().solve_sharded(S)
    res = _results(S)
    # a direct caller (Diagnostics.add_field does this) gets complete grids from the
    # backward transform: forward + backward of rho is the identity up to rounding
    rho_before = [S.DataDev['rho_m%d' % m].get()[1:].copy() for m in range(M + 1)]
    S.fb_transform(scals=['rho'], dir=0)
    for m in range(M + 1):
        S.DataDev['rho_m%d' % m].t[1:] = 0
    S.fb_transform(scals=['rho'], dir=1)
    for m in range(M + 1):
        assert rel_err(S.DataDev['rho_m%d' % m].get()[1:], rho_before[m]) < 1e-9, m
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lo=lo, hi=hi, **res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("Nx,Nr,M,world", [(32, 14, 1, 2), (256, 26, 1, 2),
                                           (32, 14, 1, 4)])    # K = 13: ranks 2, 3 own no rows
def test_ranks_gloo_equal_unsharded(tmp_path, Nx, Nr, M, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), Nx, Nr, M), nprocs=world,
             join=True)
    ref = make_solver(_cfg(Nx, Nr, M))
    _random_state(ref, 11)
    for step in range(2):
        _Loop().solve_plain(ref)
    want = _results(ref)
    for rank in range(world):
        got = np.load(tmp_path / ("rank%d.npz" % rank))
        lo, hi = int(got["lo"]), int(got["hi"])
        for k in want:
            if "_fb_" in k:          # spectral state: the owned rows
                assert rel_err(got[k][lo:hi], want[k][lo:hi]) < 1e-12, (rank, k)
            else:                    # E and B grids: complete on every rank
                assert rel_err(got[k], want[k]) < 1e-12, (rank, k)
