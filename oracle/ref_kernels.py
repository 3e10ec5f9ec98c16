"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

ctypes binding to oracle/_ref/libchimera_ref_m{0,1}.so, i.e. the reference's own
OpenCL C kernels (chimeraCL/kernels/*.cl) compiled unmodified for the host by
oracle/Makefile.  Exposes the same kernel-level interface as
oracle/np_kernels.py so that oracle/orchestration.py can run on either.

Each method cites the reference launch site it stands in for.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")

# type codes: p = pointer to numpy data, u = uint32 by value, d = double by
# value, c = signed char by value
_SIGS = {
    "push_xyz": "ppppppppp",
    "index_and_sum_in_cell": "pppppppppppp",
    "sort": "ppppu",
    "data_align_dbl": "pppu",
    "fill_grid": "pppppppuuuuu",
    "profile_by_interpolant": "ppupppu",
    "divide_by_dv_d": "pppp",
    "divide_by_dv_c": "pppp",
    "treat_axis_d": "pu",
    "treat_axis_c": "pu",
    "warp_axis_m0_d": "pu",
    "warp_axis_m1plus_c": "pu",
    "set_cdouble_to": "pddu",
    "append_c2c": "ppu",
    "zpaxz_c2c": "ddppu",
    "mult_elementwise_d2c": "ppu",
    "axpbyz_c2c": "ddpddppu",
    "ab_dot_x": "ddpppuu",
    "cast_array_d2c": "ppu",
    "get_m1": "pppp",
    "get_phase_plus": "ppdu",
    "get_phase_minus": "ppdu",
    "multiply_by_phase": "pppp",
    "profile_edges_c": "ppuuu",
    "profile_edges_d": "ppuuu",
    "advance_e_g_m": "p" * 20,
}
_SIGS_M = {
    0: {"depose_scalar": "upppppp" + "c" + "ppppppp" + "p",
        "depose_vector": "uppppppppp" + "p" + "c" + "ppppppp" + "ppp",
        "gather_and_push": "ppppppppp" + "p" + "uu" + "ppppppp" + "pppppp"},
    1: {"depose_scalar": "upppppp" + "c" + "ppppppp" + "pp",
        "depose_vector": "uppppppppp" + "p" + "c" + "ppppppp" + "pppppp",
        "gather_and_push": "ppppppppp" + "p" + "uu" + "ppppppp" + "pppppppppppp"},
}
_CT = {"p": ctypes.c_void_p, "u": ctypes.c_uint32, "d": ctypes.c_double,
       "c": ctypes.c_byte}


def ref_available(M=1):
    return os.path.exists(os.path.join(_REF_DIR, "libchimera_ref_m%d.so" % M))


def _u32(v):
    return np.array([v], dtype=np.uint32)


def _f64(v):
    return np.array([v], dtype=np.float64)


class RefKernels:
    """The reference kernels themselves, run serially (parallel=False, the
    deterministic meaning) or with OpenMP over work-items (parallel=True, used
    only when timing the CPU baseline; `sort` always runs serially so that the
    permutation is the stable one)."""

    kind = "reference"

    def __init__(self, M, parallel=False):
        if M not in (0, 1):
            raise ValueError("the reference ships particle kernels for M in {0,1} only "
                             "(methods/grid_methods_cl.py:21-23)")
        path = os.path.join(_REF_DIR, "libchimera_ref_m%d.so" % M)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " missing: run `make -C oracle` where /root/reference exists")
        self.M = M
        self.par = 1 if parallel else 0
        self.lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
        assert self.lib.nd_ref_mode_count() == M
        sigs = dict(_SIGS)
        sigs.update(_SIGS_M[M])
        self._fn = {}
        for name, sig in sigs.items():
            f = getattr(self.lib, "nd_" + name)
            f.restype = None
            f.argtypes = [ctypes.c_size_t, ctypes.c_int] + [_CT[c] for c in sig]
            self._fn[name] = (f, sig)

    def _call(self, name, n, *args, par=None):
        f, sig = self._fn[name]
        assert len(args) == len(sig), (name, len(args), len(sig))
        conv = []
        keep = []
        for a, c in zip(args, sig):
            if c == "p":
                assert isinstance(a, np.ndarray) and a.flags.c_contiguous, name
                keep.append(a)
                conv.append(a.ctypes.data)
            elif c == "u":
                conv.append(int(a))
            elif c == "d":
                conv.append(float(a))
            else:
                conv.append(int(a))
        f(int(n), self.par if par is None else par, *conv)

    # ---- particles (methods/particles_methods_cl.py) ----
    def push_xyz(self, x, y, z, px, py, pz, g_inv, dt):
        """particles_methods_cl.py:206-223 -> particles_generic.cl:129-153"""
        n = x.size
        self._call("push_xyz", n, x, y, z, px, py, pz, g_inv, _f64(dt), _u32(n))

    def index_and_sum(self, x, y, z, g):
        """particles_methods_cl.py:225-243 -> particles_generic.cl:88-126"""
        n = x.size
        indx = np.empty(n, dtype=np.uint32)
        summ = np.zeros(g["Nxm1Nrm1"] + 1, dtype=np.uint32)
        self._call("index_and_sum_in_cell", n, x, y, z, summ, _u32(n), indx,
                   _u32(g["Nx"]), _f64(g["Xmin"]), _f64(g["dx_inv"]),
                   _u32(g["Nr"]), _f64(g["Rmin"]), _f64(g["dr_inv"]))
        return indx, summ

    def sort_scatter(self, cell_offset, indx):
        """particles_methods_cl.py:252-261 -> particles_generic.cl:186-201
        (serial => stable order)."""
        n = indx.size
        counters = np.zeros(cell_offset.size - 1, dtype=np.uint32)
        out = np.empty(n, dtype=np.uint32)
        self._call("sort", n, cell_offset, indx, counters, out, n, par=0)
        return out, counters

    def align(self, arr, sort_indx, n_stay):
        """particles_methods_cl.py:263-286 -> particles_generic.cl:156-169"""
        out = np.empty(n_stay, dtype=np.float64)
        self._call("data_align_dbl", n_stay, arr, out, sort_indx, n_stay)
        return out

    def fill_grid(self, theta_var, xgrid, rgrid, nppc):
        """particles_methods_cl.py:66-96 -> particles_generic.cl:33-84"""
        Nx_loc, Nr_loc = xgrid.size, rgrid.size
        ncells = (Nx_loc - 1) * (Nr_loc - 1)
        Np = ncells * int(np.prod(nppc))
        x, y, z, w = (np.empty(Np) for _ in range(4))
        self._call("fill_grid", ncells, x, y, z, w, theta_var, xgrid, rgrid,
                   Nx_loc, ncells, int(nppc[0]), int(nppc[1]), int(nppc[2]))
        return x, y, z, w

    def profile_by_interpolant(self, x, w, x_loc, f_loc, dxm1_loc):
        """particles_methods_cl.py:179-204 -> particles_generic.cl:6-30"""
        self._call("profile_by_interpolant", x.size, x, w, x.size, x_loc, f_loc,
                   dxm1_loc, x_loc.size)

    # ---- grid (methods/grid_methods_cl.py) ----
    def _grid_ptrs(self, g):
        return (_u32(g["Nx"]), _f64(g["Xmin"]), _f64(g["dx_inv"]),
                _u32(g["Nr"]), _f64(g["Rmin"]), _f64(g["dr_inv"]))

    def depose_scalar(self, sort_indx, x, y, z, w, cell_offset, charge, g, flds):
        """grid_methods_cl.py:44-69 -> grid_deposit_m{0,1}.cl depose_scalar,
        four colour passes i_off = 0..3."""
        n4 = g["NxNr_4"]
        gp = self._grid_ptrs(g)
        for i_off in range(4):
            self._call("depose_scalar", n4, i_off, sort_indx, x, y, z, w,
                       cell_offset, np.int8(charge), *gp, _u32(n4), *flds)

    def depose_vector(self, sort_indx, x, y, z, px, py, pz, g_inv, w,
                      cell_offset, charge, g, flds):
        """grid_methods_cl.py:71-95 -> depose_vector, four colour passes."""
        n4 = g["NxNr_4"]
        gp = self._grid_ptrs(g)
        for i_off in range(4):
            self._call("depose_vector", n4, i_off, sort_indx, x, y, z, px, py,
                       pz, g_inv, w, cell_offset, np.int8(charge), *gp,
                       _u32(n4), *flds)

    def treat_axis(self, arr, Nx):
        """grid_generic.cl:37-59"""
        self._call("treat_axis_c" if arr.dtype == np.complex128 else "treat_axis_d",
                   Nx, arr, Nx)

    def divide_by_dv(self, arr, g, dV_inv):
        """grid_generic.cl:4-34"""
        self._call("divide_by_dv_c" if arr.dtype == np.complex128 else "divide_by_dv_d",
                   g["NxNr"], arr, _u32(g["NxNr"]), _u32(g["Nx"]), dV_inv)

    def warp_axis(self, arr, Nx):
        """grid_generic.cl:63-86"""
        self._call("warp_axis_m1plus_c" if arr.dtype == np.complex128 else "warp_axis_m0_d",
                   Nx, arr, Nx)

    def gather_and_push(self, x, y, z, px, py, pz, g_inv, sort_indx, cell_offset,
                        factor_push, Np, Np_stay, g, flds):
        """grid_methods_cl.py:168-192 -> gather_and_push."""
        gp = self._grid_ptrs(g)
        self._call("gather_and_push", Np, x, y, z, px, py, pz, g_inv, sort_indx,
                   cell_offset, _f64(factor_push), Np, Np_stay, *gp,
                   _u32(g["Nxm1Nrm1"]), *flds)

    # ---- generic (methods/generic_methods_cl.py) ----
    def append_c2c(self, base, add):
        self._call("append_c2c", base.size, base, add, base.size)

    def zpaxz(self, z, a, x):
        a = complex(a)
        self._call("zpaxz_c2c", x.size, a.real, a.imag, x, z, x.size)

    def mult_elementwise(self, x, z):
        self._call("mult_elementwise_d2c", x.size, x, z, x.size)

    def axpbyz(self, a, x, b, y, z):
        a, b = complex(a), complex(b)
        self._call("axpbyz_c2c", x.size, a.real, a.imag, x, b.real, b.imag, y, z, x.size)

    def ab_dot_x(self, a, b, x, z, NxNrm1, Nx):
        a = complex(a)
        self._call("ab_dot_x", x.size, a.real, a.imag, b, x, z, NxNrm1, Nx)

    def cast_c2d(self, arr_in, arr_out):
        self._call("cast_array_d2c", arr_in.size, arr_in, arr_out, arr_in.size)

    # ---- transformer (methods/transformer_methods_cl.py) ----
    def get_m1(self, dst, src, Nx, NxNrm1):
        self._call("get_m1", NxNrm1, dst, src, _u32(Nx), _u32(NxNrm1))

    def get_phase(self, kx, x0, direction):
        """transformer_methods_cl.py:38-44; dir 0 -> minus, 1 -> plus."""
        out = np.zeros(kx.size, dtype=np.complex128)
        self._call("get_phase_plus" if direction == 1 else "get_phase_minus",
                   kx.size, out, kx, x0, kx.size)
        return out

    def multiply_by_phase(self, arr, phs, Nx):
        self._call("multiply_by_phase", arr.size, arr, _u32(arr.size), _u32(Nx), phs)

    # ---- solver (methods/solver_methods_cl.py) ----
    def profile_edges(self, arr, prof, Nx, Nf):
        self._call("profile_edges_c" if arr.dtype == np.complex128 else "profile_edges_d",
                   arr.size, arr, prof, arr.size, Nx, Nf)

    def advance_e_g(self, n, dt_inv, c1, c2, c3, flds15):
        """solver_methods_cl.py:43-62 -> solver_ms_pic.cl:57-143"""
        self._call("advance_e_g_m", n, _u32(n), _f64(dt_inv), c1, c2, c3, *flds15)
