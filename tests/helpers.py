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


def per_particle_rel_err(got, ref):
    """Worst PER-PARTICLE relative error of the momentum vector, |dp| / |p_ref| of that
    particle, and of g_inv (not normalised by an array-wide maximum)."""
    if ref["px"].size == 0:
        return 0.0, 0.0
    d2 = sum((np.asarray(got[k]) - ref[k]) ** 2 for k in ("px", "py", "pz"))
    n2 = sum(ref[k] ** 2 for k in ("px", "py", "pz"))
    ep = float(np.sqrt(d2 / np.maximum(n2, 1e-300)).max())
    eg = float((np.abs(np.asarray(got["g_inv"]) - ref["g_inv"]) / np.abs(ref["g_inv"])).max())
    return ep, eg


def moments(D):
    """Particle moments compared over N-step runs: sum w, sum w p_k,
    sum w (gamma - 1); w_abs_p = sum |w| |p| is the scale the momentum sums are
    compared on (they cancel to ~0 in a symmetric plasma)."""
    w = D["w"]
    out = {"w": float(w.sum())}
    p2 = 0
    for k in ("px", "py", "pz"):
        out["w" + k] = float((w * D[k]).sum())
        p2 = p2 + D[k] ** 2
    out["w_abs_p"] = float((np.abs(w) * np.sqrt(p2)).sum())
    out["w_kin"] = float((w * (np.sqrt(1 + p2) - 1)).sum())
    return out


def field_energy(F, A):
    """sum over modes of |E|^2 + |B|^2, weighted by the ring volume r dr dx (m >= 1
    modes count twice: +m and -m)."""
    r = A["Rgrid"][1:, None]
    tot = 0.0
    for k, v in F.items():
        m = int(k.split("_m")[1])
        tot += (1 if m == 0 else 2) * float((np.abs(v[1:]) ** 2 * r).sum())
    return tot * A["dx"] * A["dr"] * 2 * np.pi
