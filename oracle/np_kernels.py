"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

NumPy restatement of chimeraCL's device kernels for the per-step PIC hot path.
Every function follows one OpenCL kernel of the reference (file:line cited) and
keeps its operation order, so that for per-particle arithmetic the results are
bit-identical to the reference kernels compiled without FMA contraction
(oracle/_ref, see oracle/Makefile); depositions differ only by summation order.

Parity pin: tests/test_oracle.py checks this restatement against oracle/_ref
(the reference's own kernel source executed on the host) whenever that library
is present, and against tests/golden/*.npz, which were generated from oracle/_ref
by tests/golden/make_golden.py.  The reference ships no golden vectors of its own
(SURVEY.md section 4).
"""
import numpy as np


class NumpyKernels:
    """M in {0, 1}: restatement of the reference kernels.  M >= 2 has NO reference
    particle kernels (grid_deposit_m2.cl does not exist); for it the
    visible m=1 pattern is generalised -- weights ((y+iz)/r)^m on deposit,
    2*Re(F_m e^{-i m theta}) on gather -- and parity is UNPINNED."""
    kind = "port"

    def __init__(self, M):
        self.M = M

    @staticmethod
    def _powers(e0, e1, M):
        """[(Re, Im) of (e0 + i e1)^m for m = 1..M], by the recurrence the CUDA
        kernels use."""
        out = [(e0, e1)]
        for _ in range(1, M):
            r, i = out[-1]
            out.append((r * e0 - i * e1, r * e1 + i * e0))
        return out

    # ------------------------------------------------------------------ particles
    def push_xyz(self, x, y, z, px, py, pz, g_inv, dt):
        """particles_generic.cl:129-153: dt_g = dt*g_inv; x = x + px*dt_g (in place)."""
        dt_g = dt * g_inv
        x += px * dt_g
        y += py * dt_g
        z += pz * dt_g

    @staticmethod
    def cell_coords(x, y, z, g):
        """particles_generic.cl:102-107 (also grid_deposit_m1.cl:380-384):
        r = sqrt(y*y+z*z); ix = floor((x-xmin)*dx_inv); ir = floor((r-rmin)*dr_inv)."""
        r = np.sqrt(y * y + z * z)
        ix = np.floor((x - g["Xmin"]) * g["dx_inv"]).astype(np.int64)
        ir = np.floor((r - g["Rmin"]) * g["dr_inv"]).astype(np.int64)
        return r, ix, ir

    def index_and_sum(self, x, y, z, g):
        """particles_generic.cl:88-126: cell = ix + ir*(Nx-1) if 0<ix<Nx-2 and
        0<=ir<Nr-2 else the trash bin (Nr-1)(Nx-1); histogram with Ncells+1 bins."""
        Nx_loc, Nr_loc = g["Nx"] - 1, g["Nr"] - 1
        _, ix, ir = self.cell_coords(x, y, z, g)
        ok = (ix > 0) & (ix < Nx_loc - 1) & (ir < Nr_loc - 1) & (ir >= 0)
        indx = np.where(ok, ix + ir * Nx_loc, Nr_loc * Nx_loc).astype(np.uint32)
        summ = np.bincount(indx, minlength=Nr_loc * Nx_loc + 1).astype(np.uint32)
        return indx, summ

    def sort_scatter(self, cell_offset, indx):
        """particles_generic.cl:186-201 executed serially: slot = cell_offset[c] +
        (number of earlier particles of cell c) -> the stable counting sort."""
        out = np.argsort(indx, kind="stable").astype(np.uint32)
        counters = np.bincount(indx, minlength=cell_offset.size - 1).astype(np.uint32)
        return out, counters

    def align(self, arr, sort_indx, n_stay):
        """particles_generic.cl:156-169: x_new[ip] = x[sort_indx[ip]], ip < Np_stay."""
        return arr[sort_indx[:n_stay]].copy()

    def fill_grid(self, theta_var, xgrid, rgrid, nppc):
        """particles_generic.cl:33-84: regular (x, r, theta) lattice per cell."""
        npx, npr, npt = (int(v) for v in nppc)
        Nx_cell = xgrid.size - 1
        ncells = Nx_cell * (rgrid.size - 1)
        ic = np.arange(ncells)
        ir = ic // Nx_cell
        ix = ic - Nx_cell * ir
        xmin, rmin = xgrid[ix], rgrid[ir]
        Lx = xgrid[ix + 1] - xgrid[ix]
        Lr = rgrid[ir + 1] - rgrid[ir]
        dx, dr = 1.0 / npx, 1.0 / npr
        dth = 2 * np.pi / npt
        x = np.empty((ncells, npt, npr, npx))
        y, z, w = np.empty_like(x), np.empty_like(x), np.empty_like(x)
        for it in range(npt):
            th = theta_var + it * dth
            s, c = np.sin(th), np.cos(th)
            for jr in range(npr):
                rp = rmin + (0.5 + jr) * dr * Lr
                for kx in range(npx):
                    x[:, it, jr, kx] = xmin + (0.5 + kx) * dx * Lx
                    y[:, it, jr, kx] = rp * s
                    z[:, it, jr, kx] = rp * c
                    w[:, it, jr, kx] = rp
        return x.ravel(), y.ravel(), z.ravel(), w.ravel()

    def profile_by_interpolant(self, x, w, x_loc, f_loc, dxm1_loc):
        """particles_generic.cl:6-30: piecewise-linear profile multiplies w.  The
        reference's search loop leaves ix = Nx_loc-1 when no interval matches (and
        then reads one past the tables); inputs used here always match."""
        n = x_loc.size
        ix = np.full(x.shape, n - 1, dtype=np.int64)
        for k in range(n - 2, -1, -1):
            hit = (x > x_loc[k]) & (x <= x_loc[k + 1])
            ix = np.where(hit, k, ix)
        assert (ix < n - 1).all(), "particle outside the profile table"
        f_minus = f_loc[ix] * dxm1_loc[ix]
        f_plus = f_loc[ix + 1] * dxm1_loc[ix]
        w *= f_minus * (x_loc[ix + 1] - x) + f_plus * (x - x_loc[ix])

    # ------------------------------------------------------------------ deposition
    def _valid_sorted(self, sort_indx, cell_offset, g):
        """Particles the four colour passes visit: every particle of every cell
        with 0<ix<Nx-2, ir<Nr-2 (grid_deposit_m1.cl:65), in (cell, sorted) order."""
        ncells = g["Nxm1Nrm1"]
        n_stay = int(cell_offset[ncells])  # == cell_offset[-2]
        ips = sort_indx[:n_stay].astype(np.int64)
        counts = np.diff(cell_offset[: ncells + 1].astype(np.int64))
        cell = np.repeat(np.arange(ncells, dtype=np.int64), counts)
        Nx_cell = g["Nx"] - 1
        ir = cell // Nx_cell
        ix = cell - ir * Nx_cell
        ok = (ix > 0) & (ix < Nx_cell - 1) & (ir < g["Nr"] - 2)
        return ips[ok], ix[ok], ir[ok]

    @staticmethod
    def _shape(xp, rp, ix, ir, wp, g):
        """grid_deposit_m1.cl:119-130: sX1=(xp-xmin)*dx_inv-ix, sX0=1-sX1, same
        in r; sX0,sX1 *= wp; C[i][j] = sR_i*sX_j."""
        sX1 = (xp - g["Xmin"]) * g["dx_inv"] - ix
        sX0 = 1.0 - sX1
        sR1 = (rp - g["Rmin"]) * g["dr_inv"] - ir
        sR0 = 1.0 - sR1
        if wp is not None:
            sX0 = sX0 * wp
            sX1 = sX1 * wp
        return ((sR0 * sX0, sR0 * sX1), (sR1 * sX0, sR1 * sX1))

    @staticmethod
    def _scatter(fld, node, vals):
        flat = fld.reshape(-1)
        if fld.dtype == np.complex128:
            n = flat.size
            flat += (np.bincount(node, weights=vals[0], minlength=n)
                     + 1j * np.bincount(node, weights=vals[1], minlength=n))
        else:
            flat += np.bincount(node, weights=vals, minlength=flat.size)

    def depose_scalar(self, sort_indx, x, y, z, w, cell_offset, charge, g, flds):
        """grid_deposit_m0.cl:20-129 / grid_deposit_m1.cl:20-152 (all 4 colours).
        m=1 uses the unguarded 1/r (grid_deposit_m1.cl:115)."""
        ips, ix, ir = self._valid_sorted(sort_indx, cell_offset, g)
        if ips.size == 0:
            return
        xp, yp, zp = x[ips], y[ips], z[ips]
        wp = w[ips] * np.int8(charge)
        rp = np.sqrt(yp * yp + zp * zp)
        C = self._shape(xp, rp, ix, ir, wp, g)
        Nx = g["Nx"]
        if self.M >= 1:
            with np.errstate(divide="ignore", invalid="ignore"):
                rp_inv = 1.0 / rp
                e0, e1 = yp * rp_inv, zp * rp_inv
        pw = self._powers(e0, e1, self.M) if self.M >= 1 else []
        for i in range(2):
            for j in range(2):
                node = ix + j + (ir + i) * Nx
                self._scatter(flds[0], node, C[i][j])
                for m, (er, ei) in enumerate(pw):
                    self._scatter(flds[1 + m], node, (C[i][j] * er, C[i][j] * ei))

    def depose_vector(self, sort_indx, x, y, z, px, py, pz, g_inv, w,
                      cell_offset, charge, g, flds):
        """grid_deposit_m0.cl:148-277 / grid_deposit_m1.cl:171-327 (all 4 colours).
        wp = w*g_inv*charge; rp_inv guarded (grid_deposit_m1.cl:280-281);
        m1 += (C*jp_k)*exp_m1."""
        ips, ix, ir = self._valid_sorted(sort_indx, cell_offset, g)
        if ips.size == 0:
            return
        xp, yp, zp = x[ips], y[ips], z[ips]
        jp = (px[ips], py[ips], pz[ips])
        wp = w[ips] * g_inv[ips] * np.int8(charge)
        rp = np.sqrt(yp * yp + zp * zp)
        C = self._shape(xp, rp, ix, ir, wp, g)
        Nx = g["Nx"]
        if self.M >= 1:
            rp_inv = np.zeros_like(rp)
            np.divide(1.0, rp, out=rp_inv, where=rp > 0)
            e0, e1 = yp * rp_inv, zp * rp_inv
        pw = self._powers(e0, e1, self.M) if self.M >= 1 else []
        for k in range(3):
            for i in range(2):
                for j in range(2):
                    node = ix + j + (ir + i) * Nx
                    jp_proj = C[i][j] * jp[k]
                    self._scatter(flds[k], node, jp_proj)
                    for m, (er, ei) in enumerate(pw):
                        self._scatter(flds[3 * (m + 1) + k], node, (jp_proj * er, jp_proj * ei))

    def treat_axis(self, arr, Nx):
        """grid_generic.cl:37-59: row1 -= row0."""
        arr[1] -= arr[0]

    def divide_by_dv(self, arr, g, dV_inv):
        """grid_generic.cl:4-34: arr[ir,:] *= dV_inv[ir]."""
        arr *= dV_inv[:, None]

    def warp_axis(self, arr, Nx):
        """grid_generic.cl:63-86: row0 = +row1 (m=0, real) / -row1 (m>=1, complex)."""
        arr[0] = -arr[1] if arr.dtype == np.complex128 else arr[1]

    # ------------------------------------------------------------------ gather + Boris
    def gather_and_push(self, x, y, z, px, py, pz, g_inv, sort_indx, cell_offset,
                        factor_push, Np, Np_stay, g, flds):
        """grid_deposit_m0.cl:280-427 / grid_deposit_m1.cl:330-511.  Gate on the
        STORAGE index (sort_indx[ip] < Np_stay, :367-368) and on the recomputed
        cell (no ir>=0 test); m=1 terms carry the factor 2 (:435); dt_2 =
        0.5*FactorPush (:392)."""
        s = sort_indx[:Np].astype(np.int64)
        s = s[s < Np_stay]
        xp, yp, zp = x[s], y[s], z[s]
        rp, ix, ir = self.cell_coords(xp, yp, zp, g)
        Nx_grid = g["Nx"]
        ok = (ix > 0) & (ix < Nx_grid - 2) & (ir < g["Nr"] - 2)
        s, xp, yp, zp, rp, ix, ir = (a[ok] for a in (s, xp, yp, zp, rp, ix, ir))
        if s.size == 0:
            return
        u_p = [px[s], py[s], pz[s]]
        dt_2 = 0.5 * factor_push
        C = self._shape(xp, rp, ix, ir, None, g)
        if self.M >= 1:
            with np.errstate(divide="ignore", invalid="ignore"):
                rp_inv = 1.0 / rp
            e0 = yp * rp_inv
            e1 = -zp * rp_inv
        e_p = [np.zeros_like(xp) for _ in range(3)]
        b_p = [np.zeros_like(xp) for _ in range(3)]
        pw = self._powers(e0, e1, self.M) if self.M >= 1 else []   # e^{-i m theta}
        for k in range(3):
            for i in range(2):
                for j in range(2):
                    node = ix + j + (ir + i) * Nx_grid
                    c = C[i][j]
                    e_p[k] = e_p[k] + c * flds[k].reshape(-1)[node]
                    b_p[k] = b_p[k] + c * flds[3 + k].reshape(-1)[node]
                    for m, (er, ei) in enumerate(pw):
                        em = flds[6 * (m + 1) + k].reshape(-1)[node]
                        bm = flds[6 * (m + 1) + 3 + k].reshape(-1)[node]
                        e_p[k] = e_p[k] + c * (2 * em.real) * er
                        e_p[k] = e_p[k] - c * (2 * em.imag) * ei
                        b_p[k] = b_p[k] + c * (2 * bm.real) * er
                        b_p[k] = b_p[k] - c * (2 * bm.imag) * ei
        um = [u_p[k] + dt_2 * e_p[k] for k in range(3)]
        g_p_inv = 1.0 / np.sqrt(1.0 + um[0] * um[0] + um[1] * um[1] + um[2] * um[2])
        t = [dt_2 * b_p[k] * g_p_inv for k in range(3)]
        t2 = 2.0 / (1.0 + t[0] * t[0] + t[1] * t[1] + t[2] * t[2])
        sv = [t[k] * t2 for k in range(3)]
        u0 = [um[0] + um[1] * t[2] - um[2] * t[1],
              um[1] - um[0] * t[2] + um[2] * t[0],
              um[2] + um[0] * t[1] - um[1] * t[0]]
        up = [um[0] + u0[1] * sv[2] - u0[2] * sv[1],
              um[1] - u0[0] * sv[2] + u0[2] * sv[0],
              um[2] + u0[0] * sv[1] - u0[1] * sv[0]]
        un = [up[k] + dt_2 * e_p[k] for k in range(3)]
        px[s], py[s], pz[s] = un
        g_inv[s] = 1.0 / np.sqrt(1.0 + un[0] * un[0] + un[1] * un[1] + un[2] * un[2])

    # ------------------------------------------------------------------ generic.cl
    def append_c2c(self, base, add):
        base += add

    def zpaxz(self, z, a, x):
        """generic.cl:33-45: z = z + a*x."""
        z += complex(a) * x

    def mult_elementwise(self, x, z):
        """generic.cl:47-58: z = x*z, x real."""
        z *= x

    def axpbyz(self, a, x, b, y, z):
        """generic.cl:60-77."""
        z[...] = complex(a) * x + complex(b) * y

    def ab_dot_x(self, a, b, x, z, NxNrm1, Nx):
        """generic.cl:80-98: z[ir,ix] = b[ix]*(a*x[ir,ix])."""
        z[...] = b[None, :] * (complex(a) * x)

    def cast_c2d(self, arr_in, arr_out):
        """generic.cl:100-112: real part."""
        arr_out[...] = arr_in.real

    # ------------------------------------------------------------------ transformer_generic.cl
    def get_m1(self, dst, src, Nx, NxNrm1):
        """transformer_generic.cl:3-24: F_{-1}(ix) = -conj(F_1((Nx-ix) mod Nx))."""
        idx = (Nx - np.arange(Nx)) % Nx
        dst[...] = -np.conj(src[:, idx])

    def get_phase(self, kx, x0, direction):
        """transformer_generic.cl:28-55: exp(+-i*x0*kx); dir 0 -> minus."""
        sgn = 1.0 if direction == 1 else -1.0
        return np.cos(x0 * kx) + 1j * (sgn * np.sin(x0 * kx))

    def multiply_by_phase(self, arr, phs, Nx):
        """transformer_generic.cl:58-80."""
        arr *= phs[None, :]

    # ------------------------------------------------------------------ solver_ms_pic.cl
    def profile_edges(self, arr, prof, Nx, Nf):
        """solver_ms_pic.cl:5-55: ix<Nf: *= f[ix]; ix>Nx-Nf: *= f[Nx-ix]."""
        ix = np.arange(Nx)
        fac = np.ones(Nx)
        lo = ix < Nf
        fac[lo] *= prof[ix[lo]]
        hi = ix > Nx - Nf
        fac[hi] *= prof[Nx - ix[hi]]
        arr *= fac[None, :]

    def advance_e_g(self, n, dt_inv, c1, c2, c3, f):
        """solver_ms_pic.cl:57-143 (PSATD update of E and G, per mode)."""
        pi2 = 2 * np.pi
        for k in range(3):
            e0, g0 = f[k], f[3 + k]
            j0, n0, n1 = f[6 + k] * pi2, f[9 + k] * pi2, f[12 + k] * pi2
            e1 = c1 * e0 + c2 * c3 * (g0 - j0) + c3 * (
                c1 * n0 - n1 - (n0 - n1) * dt_inv * c2 * c3)
            g1 = -c2 * e0 + c1 * (g0 - j0) + j0 + c3 * (
                dt_inv * (1.0 - c1) * (n0 - n1) - c2 * n0)
            f[k][...] = e1
            f[3 + k][...] = g1
