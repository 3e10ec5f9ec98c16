"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

NumPy restatement of chimeraCL's host-side orchestration for the per-step PIC hot
path: the wrapper classes (particles.py, grid.py, transformer.py, solver.py), the
methods/ mixins (which kernel runs on which DataDev keys, in which order) and
pic_loop.py's step.  The device kernels come from a backend object with the
interface of oracle/np_kernels.NumpyKernels (restatement) or
oracle/ref_kernels.RefKernels (the reference's own kernels, host-compiled).

Third-party arithmetic the reference delegates to libraries that are not under
/root/reference (all unpinned, requirements.txt:1-6), restated as SURVEY.md 8(c):
  Reikna MatrixMul -> np.dot (the reference's own CPU-device path,
                      methods/transformer_methods_cl.py:474-480)
  Reikna FFT       -> np.fft.fft / np.fft.ifft along axis 1 (normalised inverse)
  pyopencl cumsum  -> np.cumsum (uint32)
  scipy jn/jn_zeros, np.linalg.pinv -> the same calls

Parity pin: see oracle/np_kernels.py header.
"""
import numpy as np
from scipy.special import jn, jn_zeros

# Threads for the x-FFTs.  None: np.fft (single-threaded; what the parity tests use).
# bench.py's CPU legs set it to the core count (scipy.fft, the same pocketfft algorithm run
# over the rows in parallel), so that the timed baseline uses all host cores in every phase.
FFT_WORKERS = None


def _fft(a, inverse=False):
    if FFT_WORKERS is None:
        return np.fft.ifft(a, axis=1) if inverse else np.fft.fft(a, axis=1)
    import scipy.fft
    f = scipy.fft.ifft if inverse else scipy.fft.fft
    return f(a, axis=1, workers=FFT_WORKERS)


# ----------------------------------------------------------------------------- configs
def grid_args(cfg):
    """grid.py:68-103."""
    A = dict(cfg)
    A.setdefault("M", 0)
    A["dx"] = (A["Xmax"] - A["Xmin"]) / (A["Nx"] - 1)
    A["dx_inv"] = 1.0 / A["dx"]
    A["dr"] = A["Rmax"] / (A["Nr"] - 1.5)
    A["dr_inv"] = 1.0 / A["dr"]
    A.setdefault("dt", A["dx"])
    A["dt_inv"] = 1.0 / A["dt"]
    A["Xgrid"] = A["Xmin"] + A["dx"] * np.arange(A["Nx"])
    A["Rmin"] = -0.5 * A["dr"]
    A["Rgrid"] = A["Rmin"] + A["dr"] * np.arange(A["Nr"])
    A["Rmax"] = A["Rgrid"].max()
    with np.errstate(divide="ignore"):
        A["dV_inv"] = (A["Rgrid"] > 0) / (2 * np.pi * A["dx"] * A["dr"] * A["Rgrid"])
    A["NxNr"] = A["Nr"] * A["Nx"]
    A["Nxm1Nrm1"] = (A["Nr"] - 1) * (A["Nx"] - 1)
    A["NxNrm1"] = (A["Nr"] - 1) * A["Nx"]
    A["NxNr_4"] = (A["Nr"]) // 2 * (A["Nx"]) // 2
    A["vec_comps"] = ["x", "y", "z"]
    return A


def particle_args(cfg):
    """particles.py:53-100."""
    A = dict(cfg)
    A["Np"] = 0
    A["Np_stay"] = 0
    for k, v in (("dt", 1.0), ("dx", 1.0), ("dr", 1.0), ("charge", -1.0),
                 ("mass", 1.0), ("dens", 1.0)):
        A.setdefault(k, v)
    A["dt_2"] = 0.5 * A["dt"]
    if "Nppc" in A:
        A["Nppc"] = np.array(A["Nppc"], dtype=np.uint32)
        A["w0"] = 2 * np.pi * A["dx"] * A["dr"] * A["dens"] / np.prod(A["Nppc"])
        A["ddx"] = A["dx"] / A["Nppc"][0]
    else:
        A["ddx"] = 1.0
    A["FactorPush"] = 2 * np.pi * A["dt"] * A["charge"] / A["mass"]
    A["right_lim"] = 0.0
    return A


def spectral_args(A):
    """transformer.py:28-97 (spectral axes, filters, DHT and dDHT matrices) and
    solver.py:41-52 (PSATD coefficients), solver_methods_cl.py:32-41 (damping)."""
    dx, Nx, dr, Nr, M = A["dx"], A["Nx"], A["dr"], A["Nr"], A["M"]
    kx = 2 * np.pi * np.fft.fftfreq(Nx, dx)
    R_period = A["Rgrid"][-1] + 0.5 * dr
    A["kx"] = kx
    for m in range(M + 2):
        A["kr_m%d" % m] = jn_zeros(m, Nr - 1) / R_period
    for m in range(M + 1):
        kr = A["kr_m%d" % m]
        A["w_m%d" % m] = np.sqrt(kx[None, :] ** 2 + kr[:, None] ** 2)
        A["Poiss_m%d" % m] = 1.0 / A["w_m%d" % m] ** 2
        A["SmoothingFilter_m%d" % m] = \
            (1 - np.sin(0.5 * np.pi * kx[None, :] / kx.max()) ** 2) \
            * (1 - np.sin(0.5 * np.pi * kr[:, None] / kr.max()) ** 2)
    Rgrid = A["Rgrid"][1:, None]
    R_period = Rgrid[-1] + 0.5 * dr
    for m in range(M + 1):
        kr_0 = jn_zeros(m, Nr - 1) / R_period
        kr_p = jn_zeros(m + 1, Nr - 1) / R_period
        kr_m = jn_zeros(m - 1, Nr - 1) / R_period
        A["DHT_inv_m%d" % m] = jn(m, Rgrid * kr_0)
        A["DHT_m%d" % m] = np.linalg.pinv(A["DHT_inv_m%d" % m])
        A["dDHT_plus_m%d" % m] = A["DHT_m%d" % m].dot(0.5 * kr_p * jn(m, Rgrid * kr_p))
        A["dDHT_minus_m%d" % m] = A["DHT_m%d" % m].dot(0.5 * kr_m * jn(m, Rgrid * kr_m))
    for m in range(M + 1):
        w, dt = A["w_m%d" % m], A["dt"]
        A["MxSlv_cos(wdt)_m%d" % m] = np.cos(w * dt)
        A["MxSlv_sin(wdt)*w_m%d" % m] = np.sin(w * dt) * w
        A["MxSlv_1/w**2_m%d" % m] = 1.0 / w ** 2
    if "DampCells" in A:
        N = int(A["DampCells"])
        z = np.arange(2 * N)
        z_shft = 3.0 * (z - N + 1) / (N + 1)
        A["DampProfile"] = ((z < 4.0 * N / 3) * (z >= N) * np.sin(0.5 * np.pi * z_shft) ** 2
                            + (z >= 4.0 * N / 3)).astype(np.float64)
    return A


# ----------------------------------------------------------------------------- particles
class OracleParticles:
    """particles.py + methods/particles_methods_cl.py on NumPy arrays."""

    ATTRS = ["x", "y", "z", "px", "py", "pz", "w", "g_inv"]

    def __init__(self, cfg, kernels):
        self.K = kernels
        self.Args = particle_args(cfg)
        self.immobile = "Immobile" in self.Args
        self.attrs = ["x", "y", "z", "w"] if self.immobile else list(self.ATTRS)
        self.D = {a: np.zeros(0) for a in self.attrs}
        self.flag_sorted = False

    def set_particles(self, **arrays):
        for a in self.attrs:
            self.D[a] = np.ascontiguousarray(arrays[a], dtype=np.float64).copy()
        self.reset_num_parts()
        self.flag_sorted = False

    def reset_num_parts(self):
        """particles_methods_cl.py:288-293."""
        n = self.D["x"].size
        self.Args["Np"] = n
        self.Args["Np_stay"] = n

    def push_coords(self, mode="half"):
        """particles_methods_cl.py:206-223."""
        if self.Args["Np"] == 0 or self.immobile:
            return
        dt = self.Args["dt_2"] if mode == "half" else self.Args["dt"]
        D = self.D
        self.K.push_xyz(D["x"], D["y"], D["z"], D["px"], D["py"], D["pz"], D["g_inv"], dt)
        self.flag_sorted = False

    def sort_parts(self, grid):
        """particles.py:23-29 -> particles_methods_cl.py:225-261."""
        if self.Args["Np"] == 0:
            self.flag_sorted = True
        if not self.flag_sorted:
            D = self.D
            D["indx_in_cell"], D["sum_in_cell"] = self.K.index_and_sum(
                D["x"], D["y"], D["z"], grid.Args)
            # _cumsum (:303-311): inclusive u32 scan with a 0 prepended
            D["cell_offset"] = np.concatenate(
                ([0], np.cumsum(D["sum_in_cell"], dtype=np.uint32))).astype(np.uint32)
            self.Args["Np_stay"] = int(D["cell_offset"][-2])
            D["sort_indx"], D["sum_in_cell"] = self.K.sort_scatter(
                D["cell_offset"], D["indx_in_cell"])
            self.flag_sorted = True

    def align_parts(self):
        """particles.py:42-51 -> particles_methods_cl.py:263-286."""
        if self.Args["Np"] == 0:
            return
        n_stay = self.Args["Np_stay"]
        comps = ["x", "y", "z", "w"] if self.immobile else \
            ["x", "y", "z", "px", "py", "pz", "g_inv", "w"]
        if n_stay == 0:
            for c in comps:
                self.D[c] = np.zeros(0)
            self.D["sort_indx"] = np.zeros(0, dtype=np.uint32)
        else:
            for c in comps:
                self.D[c] = self.K.align(self.D[c], self.D["sort_indx"], n_stay)
            self.D["sort_indx"] = np.arange(n_stay, dtype=np.uint32)
        self.reset_num_parts()


    # ---- particle creation (init / injection path)
    def make_new_domain(self, parts_in, density_profiles=None, theta_source=None):
        """particles_methods_cl.py:66-147.  The reference draws the per-cell theta
        offsets from pyopencl's Threefry stream (:80-82), which is not reproducible
        (parity unpinned); theta_source(ncells) -> float64[ncells] in
        [0, 2 pi) injects the table instead, so that both sides of a parity test
        create the same particles.  Thermal momenta (dpx/dpy/dpz != 0) come from the
        same stream and are refused here for the same reason."""
        A = self.Args
        xmin, xmax, rmin, rmax = (parts_in[k] for k in ("Xmin", "Xmax", "Rmin", "Rmax"))
        Nx_loc = int(np.ceil((xmax - xmin) / A["dx"]) + 1)
        Nr_loc = int(np.round((rmax - rmin) / A["dr"]) + 1)
        Xgrid_loc = xmin + A["dx"] * np.arange(Nx_loc)
        Rgrid_loc = rmin + A["dr"] * np.arange(Nr_loc)
        A["right_lim"] = Xgrid_loc[-1]
        ncells = (Nx_loc - 1) * (Nr_loc - 1)
        theta = np.ascontiguousarray(theta_source(ncells), dtype=np.float64)
        assert theta.shape == (ncells,)
        x, y, z, w = self.K.fill_grid(theta, Xgrid_loc, Rgrid_loc, A["Nppc"])
        w *= A["w0"]
        N = self.new = {"x": x, "y": y, "z": z, "w": w}
        if density_profiles is not None:
            for profile in density_profiles:
                if profile["coord"] != "x":
                    continue
                self.dens_profile(profile["points"], profile["values"],
                                  parts_in["Xmin"], parts_in["Xmax"], N["x"], N["w"])
        if not self.immobile:
            for a in ("px", "py", "pz"):
                assert parts_in.get("d" + a, 0) == 0, "thermal momenta: RNG stream unpinned"
                N[a] = np.full(x.size, float(parts_in.get(a + "_c", 0)))
            N["g_inv"] = 1.0 / np.sqrt(1 + N["px"] ** 2 + N["py"] ** 2 + N["pz"] ** 2)

    def dens_profile(self, x_prf, f_prf, xmin, xmax, x, w):
        """particles_methods_cl.py:179-204."""
        x_prf = np.array(x_prf, dtype=np.double)
        f_prf = np.array(f_prf, dtype=np.double)
        i_start = (x_prf < xmin).sum() - 1
        i_stop = (x_prf < xmax).sum() + 1
        x_loc = np.ascontiguousarray(x_prf[i_start:i_stop])
        f_loc = np.ascontiguousarray(f_prf[i_start:i_stop])
        dxm1_loc = 1.0 / (x_loc[1:] - x_loc[:-1])
        self.K.profile_by_interpolant(x, w, x_loc, f_loc, dxm1_loc)

    def add_new_particles(self, source=None):
        """particles_methods_cl.py:40-64: append the *_new arrays (the ions copy the
        electrons' through InjectorSource)."""
        src = self.new if source is None else source.new
        for a in self.attrs:
            self.D[a] = np.concatenate((self.D[a], src[a]))
        self.reset_num_parts()
        self.flag_sorted = False

    def free_added(self):
        """particles_methods_cl.py:318-325."""
        self.new = None


# ----------------------------------------------------------------------------- solver
class OracleSolver:
    """grid.py + transformer.py + solver.py and their methods/ mixins."""

    def __init__(self, cfg, kernels):
        self.K = kernels
        self.Args = spectral_args(grid_args(cfg))
        A = self.Args
        self.M = A["M"]
        Nr, Nx = A["Nr"], A["Nx"]
        self.D = {}
        comps = [f + c for f in ("E", "B", "J", "G") for c in "xyz"] + ["rho"]
        for name in comps:  # grid.py:106-127
            self.D[name + "_m0"] = np.zeros((Nr, Nx))
            for m in range(1, self.M + 1):
                self.D["%s_m%d" % (name, m)] = np.zeros((Nr, Nx), dtype=np.complex128)
        sp = ["rho"] + [f + c for f in ("E", "B", "G", "J", "dN0", "dN1") for c in "xyz"]
        for name in sp:  # transformer.py:100-132
            for m in range(self.M + 1):
                self.D["%s_fb_m%d" % (name, m)] = np.zeros((Nr - 1, Nx), dtype=np.complex128)
        for c in "xyz":
            self.D["buff_fb_m-1_" + c] = np.zeros((Nr - 1, Nx), dtype=np.complex128)

    # ---- deposition (grid.py:23-54, grid_methods_cl.py:44-151)
    def _flds(self, name):
        return [self.D["%s_m%d" % (name, m)] for m in range(self.M + 1)]

    def depose_charge(self, species):
        for f in self._flds("rho"):
            f[...] = 0
        for p in species:
            if p.Args["Np"] <= 0:
                continue
            D = p.D
            self.K.depose_scalar(D["sort_indx"], D["x"], D["y"], D["z"], D["w"],
                                 D["cell_offset"], p.Args["charge"], self.Args,
                                 self._flds("rho"))
        self.postproc_depose("rho")

    def depose_currents(self, species):
        flds = [self.D["J%s_m%d" % (c, m)] for m in range(self.M + 1) for c in "xyz"]
        for f in flds:
            f[...] = 0
        for p in species:
            if p.immobile:
                continue
            D = p.D
            self.K.depose_vector(D["sort_indx"], D["x"], D["y"], D["z"], D["px"],
                                 D["py"], D["pz"], D["g_inv"], D["w"],
                                 D["cell_offset"], p.Args["charge"], self.Args, flds)
        for c in "xyz":
            self.postproc_depose("J" + c)

    def postproc_depose(self, name):
        """grid_methods_cl.py:98-151: treat_axis then divide_by_dv, per mode."""
        for f in self._flds(name):
            self.K.treat_axis(f, self.Args["Nx"])
        for f in self._flds(name):
            self.K.divide_by_dv(f, self.Args, self.Args["dV_inv"])

    def gather_and_push(self, species):
        """grid.py:56-66, grid_methods_cl.py:153-192."""
        for fld in ("E", "B"):
            for c in "xyz":
                for f in self._flds(fld + c):
                    self.K.warp_axis(f, self.Args["Nx"])
        flds = [self.D["%s%s_m%d" % (f, c, m)] for m in range(self.M + 1)
                for f in ("E", "B") for c in "xyz"]
        for p in species:
            if p.immobile:
                continue
            D = p.D
            self.K.gather_and_push(D["x"], D["y"], D["z"], D["px"], D["py"], D["pz"],
                                   D["g_inv"], D["sort_indx"], D["cell_offset"],
                                   p.Args["FactorPush"], p.Args["Np"],
                                   p.Args["Np_stay"], self.Args, flds)

    # ---- Fourier-Bessel transforms (transformer_methods_cl.py:38-64, 290-455)
    def fb_transform(self, scals=(), vects=(), dir=0, mode="full"):
        for s in scals:
            self.transform_field(s, dir, mode)
        for v in vects:
            for c in "xyz":
                self.transform_field(v + c, dir, mode)

    def transform_field(self, name, dir, mode):
        A, D, K = self.Args, self.D, self.K
        Nx = A["Nx"]
        phs = K.get_phase(A["kx"], A["Xmin"], dir)
        for m in range(self.M + 1):
            if dir == 0:
                src = D["%s_m%d" % (name, m)][1:]
                buf = np.ascontiguousarray(src)
                if mode == "full":
                    buf = np.dot(A["DHT_m%d" % m], buf)
                out = _fft(buf.astype(np.complex128))
                out = np.ascontiguousarray(out)
                K.multiply_by_phase(out, phs, Nx)
                D["%s_fb_m%d" % (name, m)][...] = out
            else:
                buf = D["%s_fb_m%d" % (name, m)].copy()
                K.multiply_by_phase(buf, phs, Nx)
                buf = np.ascontiguousarray(_fft(buf, inverse=True))
                if m == 0:
                    tmp = np.empty(buf.shape)
                    K.cast_c2d(buf, tmp)
                    buf = tmp
                if mode == "full":
                    buf = np.dot(A["DHT_inv_m%d" % m], buf)
                D["%s_m%d" % (name, m)][1:] = buf

    # ---- spectral operators (transformer_methods_cl.py:66-263)
    def fields_smooth(self, flds):
        for m in range(self.M + 1):
            for f in flds:
                self.K.mult_elementwise(self.Args["SmoothingFilter_m%d" % m],
                                        self.D["%s_fb_m%d" % (f, m)])

    def field_poiss_vec(self, fld):
        for m in range(self.M + 1):
            for c in "xyz":
                self.K.mult_elementwise(self.Args["Poiss_m%d" % m],
                                        self.D["%s%s_fb_m%d" % (fld, c, m)])

    def field_poiss_scl(self, fld):
        """transformer_methods_cl.py:73-77."""
        for m in range(self.M + 1):
            self.K.mult_elementwise(self.Args["Poiss_m%d" % m],
                                    self.D["%s_fb_m%d" % (fld, m)])

    def _get_mm1(self, src_name, comp):
        if self.M == 0:
            return
        A = self.Args
        self.K.get_m1(self.D["buff_fb_m-1_" + comp], self.D[src_name + "_fb_m1"],
                      A["Nx"], A["NxNrm1"])

    def field_grad(self, scl, vec):
        """transformer_methods_cl.py:87-133."""
        A, D, K, M = self.Args, self.D, self.K, self.M
        self._get_mm1(scl, "x")
        for m in range(M + 1):
            ox, oy, oz = (D["%s%s_fb_m%d" % (vec, c, m)] for c in "xyz")
            for o in (ox, oy, oz):
                o[...] = 0
            K.ab_dot_x(1j, A["kx"], D["%s_fb_m%d" % (scl, m)], ox, A["NxNrm1"], A["Nx"])
            if m > 0:
                src = D["%s_fb_m%d" % (scl, m - 1)]
            elif M > 0:
                src = D["buff_fb_m-1_x"]
            else:
                continue
            b = np.ascontiguousarray(np.dot(A["dDHT_minus_m%d" % m], src))
            K.zpaxz(oy, -1.0, b)
            K.zpaxz(oz, -1j, b)
            if m < M:
                b = np.ascontiguousarray(
                    np.dot(A["dDHT_plus_m%d" % m], D["%s_fb_m%d" % (scl, m + 1)]))
                K.append_c2c(oy, b)
                K.zpaxz(oz, -1j, b)

    def field_div(self, vin, sout):
        """transformer_methods_cl.py:135-183 (unused by the loop, part of the API)."""
        A, D, K, M = self.Args, self.D, self.K, self.M
        for c in "yz":
            self._get_mm1(vin + c, c)
        b0 = np.empty((A["Nr"] - 1, A["Nx"]), dtype=np.complex128)
        for m in range(M + 1):
            out = D["%s_fb_m%d" % (sout, m)]
            out[...] = 0
            K.ab_dot_x(1j, A["kx"], D["%sx_fb_m%d" % (vin, m)], out, A["NxNrm1"], A["Nx"])
            if m > 0:
                fy, fz = (D["%s%s_fb_m%d" % (vin, c, m - 1)] for c in "yz")
            elif M > 0:
                fy, fz = D["buff_fb_m-1_y"], D["buff_fb_m-1_z"]
            else:
                continue
            K.axpbyz(-1j, fz, -1.0, fy, b0)
            K.append_c2c(out, np.ascontiguousarray(np.dot(A["dDHT_minus_m%d" % m], b0)))
            if m < M:
                fy, fz = (D["%s%s_fb_m%d" % (vin, c, m + 1)] for c in "yz")
                K.axpbyz(-1j, fz, 1.0, fy, b0)
                K.append_c2c(out, np.ascontiguousarray(np.dot(A["dDHT_plus_m%d" % m], b0)))

    def field_rot(self, fin, fout):
        """transformer_methods_cl.py:185-263."""
        A, D, K, M = self.Args, self.D, self.K, self.M
        for c in "xyz":
            self._get_mm1(fin + c, c)
        b0 = np.empty((A["Nr"] - 1, A["Nx"]), dtype=np.complex128)
        for m in range(M + 1):
            ox, oy, oz = (D["%s%s_fb_m%d" % (fout, c, m)] for c in "xyz")
            for o in (ox, oy, oz):
                o[...] = 0
            K.ab_dot_x(-1j, A["kx"], D["%sz_fb_m%d" % (fin, m)], oy, A["NxNrm1"], A["Nx"])
            K.ab_dot_x(1j, A["kx"], D["%sy_fb_m%d" % (fin, m)], oz, A["NxNrm1"], A["Nx"])
            if m > 0:
                fx, fy, fz = (D["%s%s_fb_m%d" % (fin, c, m - 1)] for c in "xyz")
            elif M > 0:
                fx, fy, fz = (D["buff_fb_m-1_" + c] for c in "xyz")
            else:
                continue
            K.axpbyz(-1, fz, 1j, fy, b0)
            b1 = np.ascontiguousarray(np.dot(A["dDHT_minus_m%d" % m], b0))
            K.append_c2c(ox, b1)
            b1 = np.ascontiguousarray(np.dot(A["dDHT_minus_m%d" % m], fx))
            K.zpaxz(oy, -1j, b1)
            K.zpaxz(oz, 1, b1)
            if m < M:
                fx, fy, fz = (D["%s%s_fb_m%d" % (fin, c, m + 1)] for c in "xyz")
                K.axpbyz(1, fz, 1j, fy, b0)
                b1 = np.ascontiguousarray(np.dot(A["dDHT_plus_m%d" % m], b0))
                K.append_c2c(ox, b1)
                b1 = np.ascontiguousarray(np.dot(A["dDHT_plus_m%d" % m], fx))
                K.zpaxz(oy, -1j, b1)
                K.zpaxz(oz, -1, b1)

    # ---- Maxwell solver (solver.py:29-39, solver_methods_cl.py:43-83)
    def push_fields(self):
        A, D = self.Args, self.D
        for m in range(self.M + 1):
            f = [D["%s%s_fb_m%d" % (v, c, m)] for v in ("E", "G", "J", "dN0", "dN1")
                 for c in "xyz"]
            self.K.advance_e_g(A["NxNrm1"], A["dt_inv"], A["MxSlv_cos(wdt)_m%d" % m],
                               A["MxSlv_sin(wdt)*w_m%d" % m], A["MxSlv_1/w**2_m%d" % m], f)

    def profile_edges(self, flds):
        A = self.Args
        for fld in flds:
            for c in "xyz":
                for m in range(self.M + 1):
                    self.K.profile_edges(self.D["%s%s_m%d" % (fld, c, m)],
                                         A["DampProfile"], A["Nx"], 2 * A["DampCells"])

    def damp_fields(self):
        self.fb_transform(vects=["E", "G"], dir=1, mode="half")
        self.profile_edges(["E", "G"])
        self.fb_transform(vects=["E", "G"], dir=0, mode="half")

    def restore_B_fb(self):
        self.field_rot("G", "B")
        self.field_poiss_vec("B")


class OracleFrame:
    """frame.py:4-64: moving window + plasma injector."""

    def __init__(self, cfg, theta_source):
        self.Args = dict(cfg)
        for k, v in (("Steps", 1.0), ("Velocity", 0.0), ("dt", 1), ("DensityProfiles", None)):
            self.Args.setdefault(k, v)
        self.theta_source = theta_source

    def shift_grids(self, grids, steps=None):
        """frame.py:22-30 (host Args and the device scalars are one dict here)."""
        if steps is None:
            steps = self.Args["Steps"]
        x_shift = steps * self.Args["dt"] * self.Args["Velocity"]
        for grid in grids:
            for arg in ("Xmax", "Xmin"):
                grid.Args[arg] += x_shift
            grid.Args["Xgrid"] = grid.Args["Xgrid"] + x_shift

    def inject_plasma(self, species, grid, steps=None):
        """frame.py:32-64."""
        if steps is None:
            steps = self.Args["Steps"]
        x_shift = steps * self.Args["dt"] * self.Args["Velocity"]
        for sp in species:
            if sp.Args["Np"] == 0:
                sp.Args["right_lim"] = grid.Args["Xmax"] - x_shift
            dom = {"Xmin": sp.Args["right_lim"]}
            dom["Xmax"] = dom["Xmin"] + x_shift
            dom["Rmin"] = grid.Args["Rmin"] * (grid.Args["Rmin"] > 0)
            dom["Rmax"] = grid.Args["Rmax"]
            sp.make_new_domain(dom, density_profiles=self.Args["DensityProfiles"],
                               theta_source=self.theta_source)
            sp.add_new_particles(sp.Args.get("InjectorSource"))
        for sp in species:
            sp.free_added()
            sp.sort_parts(grid)
            sp.align_parts()
            num_ppc = int(np.prod(sp.Args["Nppc"]) + 1)
            x_max = sp.D["x"][-num_ppc:].max()
            sp.Args["right_lim"] = x_max + 0.5 * sp.Args["ddx"]


def pic_step(solver, species, frames=(), it=0):
    """pic_loop.py:57-142 (diagnostics are out of scope here); `it` is the loop
    counter before the step, which decides whether the frames fire (:63-66)."""
    for frame in frames:
        if np.mod(it, frame.Args["Steps"]) == 0:
            frame.shift_grids([solver])
            frame.inject_plasma(species, solver)
    for p in species:
        p.push_coords("half")
        p.sort_parts(solver)
    solver.depose_currents(species)
    for p in species:
        p.push_coords("half")
        p.sort_parts(solver)
    solver.depose_charge(species)
    solver.fb_transform(scals=["rho"], vects=["J"], dir=0)
    solver.fields_smooth(["rho", "Jx", "Jy", "Jz"])
    for m in range(solver.M + 1):
        for c in "xyz":
            solver.D["dN0%s_fb_m%d" % (c, m)][...] = solver.D["dN1%s_fb_m%d" % (c, m)]
    solver.field_grad("rho", "dN1")
    solver.push_fields()
    if "DampCells" in solver.Args:
        solver.damp_fields()
    solver.restore_B_fb()
    solver.fb_transform(vects=["E", "B"], dir=1)
    solver.gather_and_push(species)


def add_gaussian_pulse(solver, laser):
    """laser.py:3-37 (host NumPy math on the spectral arrays)."""
    A, D = solver.Args, solver.D
    k0 = 2 * np.pi * laser["k0"]
    a0, Lx, R, x0 = laser["a0"], laser["Lx"], laser["R"], laser["x0"]
    X_focus = x0 - laser["x_foc"]
    Xgrid, Rgrid = A["Xgrid"], A["Rgrid"]
    kx, w = A["kx"][None, :], A["w_m0"]
    D["Ez_m0"][1:] = a0 * np.sin(k0 * (Xgrid[None, :] - x0)) \
        * np.exp(-(Xgrid[None, :] - x0) ** 2 / Lx ** 2 - Rgrid[1:, None] ** 2 / R ** 2) \
        * (abs(Rgrid[1:, None]) < 3.5 * R) * (abs(Xgrid[None, :] - x0) < 3.5 * Lx)
    solver.fb_transform(scals=["Ez"], dir=0)
    EE = D["Ez_fb_m0"].copy()
    DT = -1.j * w * np.sign(kx + (kx == 0))
    GG = DT * EE
    EE_tmp = np.cos(w * X_focus) * EE + np.sin(w * X_focus) / w * GG
    GG = -w * np.sin(w * X_focus) * EE + np.cos(w * X_focus) * GG
    EE = EE_tmp
    EE *= np.exp(1.j * kx * X_focus)
    GG *= np.exp(1.j * kx * X_focus)
    D["Ez_fb_m0"][...] = EE
    D["Gz_fb_m0"][...] = GG
    solver.restore_B_fb()
    solver.fb_transform(vects=["B", "E"], dir=1)
