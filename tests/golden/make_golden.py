"""Generates tests/golden/*.npz from oracle/_ref, i.e. from the reference's own
kernel source (chimeraCL/kernels/*.cl) host-compiled by oracle/Makefile, driven by
oracle/orchestration.py.  Run where /root/reference exists:

    make -C oracle && python tests/golden/make_golden.py

The reference repository ships no golden vectors, so these
fixtures are the committed record of what its kernels compute on seeded inputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orchestration as O  # noqa: E402
from oracle.np_kernels import NumpyKernels  # noqa: E402
from oracle.ref_kernels import RefKernels  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make_case(M, K, seed, Nx=32, Nr=14, ppc=(2, 2, 4), shuffle=True):
    cfg = {"Xmin": -3.0, "Xmax": 3.5, "Nx": Nx, "Rmin": 0.0, "Rmax": 2.0, "Nr": Nr,
           "M": M, "DampCells": 5}
    cfg["dt"] = (cfg["Xmax"] - cfg["Xmin"]) / cfg["Nx"]
    S = O.OracleSolver(cfg, K)
    A = S.Args
    rng = np.random.default_rng(seed)
    # lattice one cell wider than the box on every side: populates the trash bin
    xg = A["Xmin"] + A["dx"] * np.arange(-1, Nx + 1)
    rg = A["dr"] * np.arange(0, Nr)
    th = rng.uniform(0, 2 * np.pi, (xg.size - 1) * (rg.size - 1))
    x, y, z, w = NumpyKernels(M).fill_grid(th, xg, rg, ppc)
    n = x.size
    perm = rng.permutation(n) if shuffle else np.arange(n)
    pc = {"Nppc": ppc, "dx": A["dx"], "dr": A["dr"], "dt": A["dt"], "dens": 0.02,
          "charge": -1}
    P = O.OracleParticles(pc, K)
    px, py, pz = (rng.normal(0, 0.4, n) for _ in range(3))
    wgt = (w * P.Args["w0"])[perm]
    P.set_particles(x=x[perm], y=y[perm], z=z[perm], px=px, py=py, pz=pz, w=wgt,
                    g_inv=1 / np.sqrt(1 + px * px + py * py + pz * pz))
    pi = dict(pc, charge=1, Immobile=True)
    I = O.OracleParticles(pi, K)
    I.set_particles(x=x[perm], y=y[perm], z=z[perm], w=wgt)
    for k in S.D:
        if k[0] in "EG" and "_fb_" in k:
            S.D[k][...] = 0.05 * (rng.normal(size=S.D[k].shape)
                                  + 1j * rng.normal(size=S.D[k].shape))
    return cfg, pc, S, P, I


def snapshot(prefix, S, P, out, skeys=(), pkeys=()):
    """Store the solver arrays whose key starts with one of `skeys` and the
    particle arrays named in `pkeys` (fixtures are kept small on purpose)."""
    for k, v in S.D.items():
        if k.startswith(tuple(skeys)) and skeys:
            out["%s/S/%s" % (prefix, k)] = v.copy()
    for k, v in P.D.items():
        if k in pkeys:
            out["%s/P/%s" % (prefix, k)] = v.copy()
    out["%s/Np_stay" % prefix] = np.array(P.Args["Np_stay"])


ATTR = ("x", "y", "z", "px", "py", "pz", "w", "g_inv")
SORT = ("indx_in_cell", "sum_in_cell", "cell_offset", "sort_indx")


def main():
    for M in (0, 1):
        K = RefKernels(M)
        cfg, pc, S, P, I = make_case(M, K, seed=100 + M)
        out = {}
        for k, v in cfg.items():
            out["cfg/" + k] = np.array(v)
        for k, v in pc.items():
            out["pcfg/" + k] = np.array(v)
        snapshot("in", S, P, out, ("E", "G"), ATTR)
        # one full PIC step, with the intermediate products of every phase
        for p in (P, I):
            p.push_coords("half")
            p.sort_parts(S)
        snapshot("sort1", S, P, out, (), ("x", "y", "z") + SORT)
        S.depose_currents([P, I])
        for p in (P, I):
            p.push_coords("half")
            p.sort_parts(S)
        S.depose_charge([P, I])
        snapshot("depose", S, P, out, ("rho_m", "Jx_m", "Jy_m", "Jz_m"), SORT)
        S.fb_transform(scals=["rho"], vects=["J"], dir=0)
        S.fields_smooth(["rho", "Jx", "Jy", "Jz"])
        for m in range(S.M + 1):
            for c in "xyz":
                S.D["dN0%s_fb_m%d" % (c, m)][...] = S.D["dN1%s_fb_m%d" % (c, m)]
        S.field_grad("rho", "dN1")
        snapshot("grad", S, P, out, ("rho_fb", "Jx_fb", "Jy_fb", "Jz_fb", "dN1"), ())
        S.push_fields()
        S.damp_fields()
        S.restore_B_fb()
        S.fb_transform(vects=["E", "B"], dir=1)
        S.gather_and_push([P, I])
        snapshot("step1", S, P, out, ("E", "B", "G"), ("px", "py", "pz", "g_inv"))
        O.pic_step(S, [P, I])
        P.align_parts()
        snapshot("step2_aligned", S, P, out, ("Ex_m", "Bz_m", "rho_m"), ATTR + ("sort_indx",))
        path = os.path.join(HERE, "pic_step_m%d.npz" % M)
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) // 1024, "KiB", "Np", P.Args["Np"])


if __name__ == "__main__":
    main()
