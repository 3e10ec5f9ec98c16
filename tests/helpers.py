"""Shared helpers for the parity tests: rebuild the seeded cases stored in
tests/golden/*.npz on the oracle side (NumPy) so that the CUDA path and the oracle
start from identical inputs."""
import os

import numpy as np

from oracle import orchestration as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ATTR = ("x", "y", "z", "px", "py", "pz", "w", "g_inv")


def load_golden(M):
    return np.load(os.path.join(GOLDEN_DIR, "pic_step_m%d.npz" % M))


def golden_cfgs(G):
    cfg = {}
    for k in G.files:
        if k.startswith("cfg/"):
            v = G[k]
            cfg[k[4:]] = int(v) if v.dtype.kind in "iu" else float(v)
    pcfg = {}
    for k in G.files:
        if k.startswith("pcfg/"):
            v = G[k]
            name = k[5:]
            if name == "Nppc":
                pcfg[name] = tuple(int(t) for t in v)
            elif name == "charge":
                pcfg[name] = int(v)
            else:
                pcfg[name] = float(v)
    return cfg, pcfg


def oracle_case_from_golden(G, K):
    """(solver, electrons, ions) on the oracle side, loaded with the 'in/' state."""
    cfg, pcfg = golden_cfgs(G)
    S = O.OracleSolver(cfg, K)
    for k in G.files:
        if k.startswith("in/S/"):
            S.D[k[5:]][...] = G[k]
    P = O.OracleParticles(pcfg, K)
    P.set_particles(**{a: G["in/P/" + a] for a in ATTR})
    I = O.OracleParticles(dict(pcfg, charge=1, Immobile=True), K)
    I.set_particles(**{a: G["in/P/" + a] for a in ("x", "y", "z", "w")})
    return S, P, I


def rel_err(a, b):
    """max|a-b| / max|b| (0 if both vanish)."""
    a = np.asarray(a)
    b = np.asarray(b)
    s = np.abs(b).max() if b.size else 0.0
    d = np.abs(a - b).max() if b.size else 0.0
    return 0.0 if d == 0 else d / max(s, 1e-300)
