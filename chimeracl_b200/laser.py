"""Gaussian laser pulse initialiser (API of the reference's chimeraCL/laser.py:3-37):
one-off host NumPy math on the spectral arrays around fb_transform / restore_B_fb."""
import numpy as np


def add_gausian_pulse(solver, laser):
    k0 = 2 * np.pi * laser['k0']
    a0, Lx, R, x0 = laser['a0'], laser['Lx'], laser['R'], laser['x0']
    X_focus = x0 - laser['x_foc']
    X = solver.Args['Xgrid'][None, :] - x0
    r = solver.Args['Rgrid'][1:, None]
    kx, w = solver.Args['kx'][None, :], solver.Args['w_m0']

    envelope = np.exp(-X ** 2 / Lx ** 2 - r ** 2 / R ** 2) * (abs(r) < 3.5 * R) * (abs(X) < 3.5 * Lx)
    solver.DataDev['Ez_m0'][1:] = a0 * np.sin(k0 * X) * envelope

    solver.fb_transform(scals=['Ez', ], dir=0)
    EE = solver.DataDev['Ez_fb_m0'].get()

    # forward-propagating pulse: G = -i w sign(kx) E, then vacuum propagation to focus
    GG = -1.j * w * np.sign(kx + (kx == 0)) * EE
    cs, sn = np.cos(w * X_focus), np.sin(w * X_focus)
    EE, GG = cs * EE + sn / w * GG, -w * sn * EE + cs * GG
    shift = np.exp(1.j * kx * X_focus)
    solver.DataDev['Ez_fb_m0'][:] = EE * shift
    solver.DataDev['Gz_fb_m0'][:] = GG * shift

    solver.restore_B_fb()
    solver.fb_transform(vects=['B', 'E'], dir=1)
