"""ctypes binding of libchimera_b200.so (the C ABI declared in include/chimera_b200.h).

There is no fallback: if the CUDA library is missing the import of any compute
method fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CHB_LIB") or os.path.join(_HERE, "libchimera_b200.so")

_vp, _u32, _i32, _sz, _dbl = (ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int,
                              ctypes.c_size_t, ctypes.c_double)

# name -> (restype, argtypes); must list every symbol of include/chimera_b200.h
SIGNATURES = {
    "chb_version": (_i32, []),
    "chb_error_string": (ctypes.c_char_p, [_i32]),
    "chb_push_xyz": (_i32, [_vp] * 8 + [_u32, _vp]),
    "chb_index_and_sum": (_i32, [_vp] * 5 + [_u32, _u32, _u32] + [_vp] * 4 + [_vp]),
    "chb_push_index": (_i32, [_vp] * 10 + [_u32, _u32, _u32] + [_vp] * 4 + [_vp]),
    "chb_cell_offsets_workspace_bytes": (_sz, [_u32]),
    "chb_cell_offsets": (_i32, [_vp, _u32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "chb_sort_workspace_bytes": (_sz, [_u32, _u32]),
    "chb_sort_scatter_stable": (_i32, [_vp, _vp, _vp, _vp, _u32, _u32, _vp, _sz, _vp]),
    "chb_align": (_i32, [_vp, _vp, _i32, _vp, _u32, _vp, _vp]),
    "chb_depose_scalar": (_i32, [_i32] + [_vp] * 6 + [_i32, _u32, _u32] + [_vp] * 4 + [_vp, _vp]),
    "chb_depose_vector": (_i32, [_i32] + [_vp] * 10 + [_i32, _u32, _u32] + [_vp] * 4 + [_vp, _vp]),
    "chb_push_depose_vector": (_i32, [_i32] + [_vp] * 11 + [_u32, _i32, _u32, _u32] + [_vp] * 4 + [_vp, _vp, _sz, _vp]),
    "chb_fft_damp_x_batched": (_i32, [_vp, _vp, _i32, _u32, _u32, _sz, _vp, _vp, _vp, _u32, _vp, _vp]),
    "chb_dht_tile_columns": (_i32, [_u32, _u32, _i32]),
    "chb_dmma_peak": (_i32, [_vp, _sz, _i32, _vp, _vp]),
    "chb_push_depose_workspace_bytes": (_sz, [_u32]),
    "chb_push_depose_push_index": (_i32, [_i32] + [_vp] * 11 + [_u32, _i32, _u32, _u32] + [_vp] * 4 + [_vp, _vp, _vp, _vp, _sz, _vp]),
    "chb_postproc_depose": (_i32, [_vp, _vp, _i32, _u32, _u32, _vp, _vp]),
    "chb_warp_axis": (_i32, [_vp, _vp, _i32, _u32, _vp]),
    "chb_gather_push": (_i32, [_i32] + [_vp] * 10 + [_u32, _vp, _u32, _u32] + [_vp] * 4 + [_vp, _vp]),
    "chb_cast_c2d": (_i32, [_vp, _vp, _sz, _vp]),
    "chb_cast_d2c": (_i32, [_vp, _vp, _sz, _vp]),
    "chb_append_c2c": (_i32, [_vp, _vp, _sz, _vp]),
    "chb_zpaxz_c2c": (_i32, [_dbl, _dbl, _vp, _vp, _sz, _vp]),
    "chb_mult_elementwise_d2c": (_i32, [_vp, _vp, _sz, _vp]),
    "chb_axpbyz_c2c": (_i32, [_dbl, _dbl, _vp, _dbl, _dbl, _vp, _vp, _sz, _vp]),
    "chb_ab_dot_x": (_i32, [_dbl, _dbl, _vp, _vp, _vp, _sz, _u32, _vp]),
    "chb_get_m1": (_i32, [_vp, _vp, _sz, _u32, _vp]),
    "chb_get_phase": (_i32, [_vp, _vp, _dbl, _i32, _u32, _vp]),
    "chb_multiply_by_phase": (_i32, [_vp, _vp, _sz, _u32, _vp]),
    "chb_profile_edges": (_i32, [_vp, _vp, _i32, _vp, _u32, _u32, _u32, _vp]),
    "chb_psatd_advance": (_i32, [_sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "chb_dht": (_i32, [_vp, _u32, _vp, _u32, _vp, _u32, _u32, _u32, _u32, _i32, _dbl, _dbl, _i32, _vp]),
    "chb_dht2": (_i32, [_vp, _u32, _vp, _u32, _vp, _dbl, _dbl, _i32, _vp, _dbl, _dbl, _i32,
                        _u32, _u32, _u32, _u32, _i32, _vp]),
    "chb_dht2_hermitian": (_i32, [_vp, _u32, _vp, _u32, _vp, _dbl, _dbl, _i32, _vp, _dbl, _dbl, _i32,
                                  _u32, _u32, _u32, _u32, _vp]),
    "chb_dht_batched": (_i32, [_vp, _u32, _vp, _vp, _i32, _u32, _u32, _u32, _u32, _u32, _i32, _vp]),
    "chb_fft_x_batched": (_i32, [_vp, _vp, _i32, _u32, _u32, _sz, _sz, _i32, _i32, _i32, _vp, _i32,
                                 _vp, _u32, _vp, _vp, _vp, _vp]),
    "chb_mirror_axpy": (_i32, [_vp, _vp, _dbl, _dbl, _dbl, _dbl, _i32, _sz, _u32, _vp]),
    "chb_peer_allreduce_f64": (_i32, [_vp, _i32, _i32, ctypes.c_uint64, _sz, _vp]),
    "chb_peer_allgather_f64": (_i32, [_vp, _i32, _i32, ctypes.c_uint64, _sz, _sz, _vp]),
    "chb_fft_max_pow2": (_i32, []),
    "chb_fft_x": (_i32, [_vp, _vp, _u32, _u32, _sz, _sz, _i32, _i32, _i32, _vp, _i32, _vp, _u32, _vp, _vp, _vp]),
}

_lib = None


class _LibProxy:
    """Attribute access returns the ctypes entry points; enable_profiling() swaps in
    wrappers that bracket every call with CUDA events on the current stream (used by
    bench.py to time individual kernels inside a real step)."""

    def __init__(self, cdll):
        self._cdll = cdll
        self._raw = {}
        self._events = []

    def _install(self, name, fn):
        self._raw[name] = fn
        object.__setattr__(self, name, fn)

    def enable_profiling(self, names=None):
        import torch
        for name, fn in self._raw.items():
            if not name.startswith("chb_") or fn.restype is not _i32 or not fn.argtypes \
                    or fn.argtypes[-1] is not _vp:
                continue
            if names is not None and name not in names:
                continue

            def timed(*args, _fn=fn, _name=name):
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = _fn(*args)
                e1.record()
                self._events.append((_name, e0, e1))
                return rc
            object.__setattr__(self, name, timed)

    def disable_profiling(self):
        for name, fn in self._raw.items():
            object.__setattr__(self, name, fn)

    def profile_report(self):
        """{name: (calls, total_ms)}; synchronises."""
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in self._events:
            n, t = out.get(name, (0, 0.0))
            out[name] = (n + 1, t + e0.elapsed_time(e1))
        self._events = []
        return out


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "chimera_b200: %s is missing -- build it with `python -m chimeracl_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    cdll = ctypes.CDLL(LIB_PATH)
    lib = _LibProxy(cdll)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(cdll, name)     # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
        lib._install(name, fn)
    _lib = lib
    return lib


# kernel launches behind each C-ABI call (bench.py reports the total as gpu_launches)
KERNELS_PER_CALL = {"chb_cell_offsets": 3, "chb_sort_scatter_stable": 3, "chb_align": 2,
                    "chb_push_depose_vector": 2, "chb_push_depose_push_index": 2}
CALL_COUNTS = {}


def launches():
    return sum(n * KERNELS_PER_CALL.get(k, 1) for k, n in CALL_COUNTS.items())


def check(rc, what=""):
    CALL_COUNTS[what] = CALL_COUNTS.get(what, 0) + 1
    if rc != 0:
        msg = load().chb_error_string(int(rc)).decode()
        raise RuntimeError("chimera_b200 %s failed: %s (code %d)" % (what, msg, rc))


def ptr_array(ptrs):
    """Host array of device pointers."""
    return (ctypes.c_void_p * len(ptrs))(*ptrs)


def int_array(vals):
    return (ctypes.c_int * len(vals))(*vals)
