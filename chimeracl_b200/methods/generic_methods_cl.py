"""Runtime holder (`Communicator`) and the generic device helpers mixin.

Mirrors the interface of the reference's chimeraCL/methods/generic_methods_cl.py
(class and method names, argument meaning) so that wrapper classes and user scripts
written against chimeraCL run unchanged; the implementation is a CUDA stream on one
B200 per process plus ctypes calls into libchimera_b200.so.

  Communicator                  reference :152-195  (OpenCL ctx/queue/Reikna thread)
  GenericMethodsCL.send_args_to_dev   :56-76
  GenericMethodsCL.dev_arr            :78-87
  cast_array_c2d/set_to/append_c2c/mult_elementwise/axpbyz/zpaxz/ab_dot_x  :89-142
"""
import os

import numpy as np
import torch

from .. import _lib
from ..devarray import DevArray


class _Queue:
    def __init__(self, comm):
        self._comm = comm

    def finish(self):
        self._comm.synchronize()


class _Thread:
    """Stand-in for the Reikna Thread the reference exposes as `comm.thr`."""

    def __init__(self, comm):
        self._comm = comm

    def synchronize(self):
        self._comm.synchronize()

    def to_device(self, arr):
        return DevArray.from_numpy(arr, self._comm.device)


class Communicator:
    """One process drives one B200.  `answers=[platform, device]` is accepted for
    script compatibility (reference :153-166); the device index comes from
    LOCAL_RANK when launched with torchrun, else from answers[1], else 0.
    There is no CPU mode: without CUDA the constructor raises."""

    def __init__(self, answers=None, device=None, seed=None, verbose=False, **_ignored):
        if not torch.cuda.is_available():
            raise RuntimeError("chimera_b200: no CUDA device visible (this build has no CPU path)")
        self.lib = _lib.load()
        self.rank = int(os.environ.get("RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        if device is None:
            if "LOCAL_RANK" in os.environ:
                device = int(os.environ["LOCAL_RANK"])
            elif answers is not None and len(answers) > 1:
                device = int(answers[1]) % torch.cuda.device_count()
            else:
                device = 0
        self.device = torch.device("cuda", int(device))
        torch.cuda.set_device(self.device)
        self.generator = torch.Generator(device=self.device)
        self.generator.manual_seed(1234 + self.rank if seed is None else int(seed))
        # names the reference exposes
        self.ctx = self
        self.queue = _Queue(self)
        self.thr = _Thread(self)
        self.dev_type = "GPU"
        self.dev_name = torch.cuda.get_device_name(self.device)
        self.device_memory_bytes = int(torch.cuda.get_device_properties(self.device).total_memory)
        self.plat_name = "NVIDIA"
        self.ocl_version = "CUDA sm_%d%d" % torch.cuda.get_device_capability(self.device)
        self.fft_method = "chimera_b200"
        self.dot_method = "chimera_b200"
        self.process_group = None       # set by parallel.init_distributed()
        if verbose:
            print("chimera_b200 on %s (%s), rank %d/%d" %
                  (self.dev_name, self.ocl_version, self.rank, self.world_size))

    @property
    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def synchronize(self):
        torch.cuda.synchronize(self.device)


class ArgsDict(dict):
    """Plain dict (the reference's `Args`) whose values may be produced lazily: a
    callable stored through set_lazy() is evaluated on first read.  Used for
    Args['Np_stay'], which the reference reads back from the device after every
    sort (particles_methods_cl.py:250) -- here the read-back is asynchronous and only
    waited for if somebody actually looks at the number."""

    def set_lazy(self, key, fn):
        dict.__setitem__(self, key, _Lazy(fn))

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if isinstance(v, _Lazy):
            v = v.fn()
            dict.__setitem__(self, key, v)
        return v

    def get(self, key, default=None):
        return self[key] if key in self else default

    # every other read path resolves pending values too, so that a _Lazy never leaks
    def __iter__(self):            # (own tp_iter: dict(Args) then goes through __getitem__)
        return dict.__iter__(self)

    def _resolve_all(self):
        for key in [k for k, v in dict.items(self) if isinstance(v, _Lazy)]:
            self[key]

    def items(self):
        self._resolve_all()
        return dict.items(self)

    def values(self):
        self._resolve_all()
        return dict.values(self)

    def copy(self):
        self._resolve_all()
        return ArgsDict(dict.copy(self))

    def pop(self, key, *default):
        if key in self:
            self[key]
        return dict.pop(self, key, *default)


class _Lazy:
    __slots__ = ("fn",)

    def __init__(self, fn):
        self.fn = fn


class GenericMethodsCL:
    def init_generic_methods(self):
        self.set_global_working_group_size()
        self._lib = _lib.load()

    def set_global_working_group_size(self):
        # kept for interface compatibility (reference :40-46); launch shapes are
        # chosen inside the library
        self.WGS = 256
        self.block_def_str = ""

    def get_wgs(self, Nelem):
        if Nelem <= self.WGS:
            return Nelem, Nelem
        return self.WGS, int(np.ceil(1. * Nelem / self.WGS)) * self.WGS

    # ------------------------------------------------------------------ plumbing
    def import_comm(self, comm):
        self.comm = comm
        self.queue = comm.queue
        self.ctx = comm.ctx
        self.thr = comm.thr
        self.dev_type = comm.dev_type
        self.plat_name = comm.plat_name
        self._lib = comm.lib

    @property
    def _stream(self):
        return self.comm.stream

    def _call(self, name, *args):
        _lib.check(getattr(self._lib, name)(*args, self._stream), name)

    def send_args_to_dev(self):
        """Every int -> uint32[1], float -> float64[1], ndarray -> device array under
        the same key, honouring dont_send / dont_keep (reference :56-76)."""
        skip = self.Args['dont_send'] if 'dont_send' in self.Args else []
        drop = self.Args['dont_keep'] if 'dont_keep' in self.Args else []
        for arg in list(self.Args.keys()):
            if arg in skip:
                continue
            val = self.Args[arg]
            kind = type(val)
            if kind is int:
                dtype = np.uint32
            elif kind is float:
                dtype = np.double
            elif kind is np.ndarray:
                dtype = val.dtype
            else:
                continue
            self.DataDev[arg] = self.dev_arr(val, dtype=dtype)
            if arg in drop:
                self.Args.pop(arg)

    def dev_arr(self, val=None, shape=(1, ), dtype=np.double, allocator=None):
        dev = self.comm.device
        if type(val) is np.ndarray:
            return DevArray.from_numpy(val, dev)
        if val is not None and val == 0:
            return DevArray.zeros(shape, dtype, dev)
        arr = DevArray.empty(shape, dtype, dev)
        if val is not None:
            self.set_to(arr, val)
        return arr

    # ------------------------------------------------------------------ element-wise
    def set_to(self, arr, val):
        arr.fill(val)

    def cast_array_c2d(self, arr_in, arr_out):
        self._call('chb_cast_c2d', arr_in.ptr, arr_out.ptr, arr_in.size)

    def append_c2c(self, arr_base, arr_add):
        self._call('chb_append_c2c', arr_base.ptr, arr_add.ptr, arr_base.size)

    def mult_elementwise(self, x, z):
        self._call('chb_mult_elementwise_d2c', x.ptr, z.ptr, x.size)

    def axpbyz(self, a, x, b, y, z):
        a, b = complex(a), complex(b)
        self._call('chb_axpbyz_c2c', a.real, a.imag, x.ptr, b.real, b.imag, y.ptr, z.ptr, x.size)

    def zpaxz(self, z, a, x):
        a = complex(a)
        self._call('chb_zpaxz_c2c', a.real, a.imag, x.ptr, z.ptr, x.size)

    def ab_dot_x(self, a, b, x, z):
        a = complex(a)
        self._call('chb_ab_dot_x', a.real, a.imag, b.ptr, x.ptr, z.ptr, x.size,
                   int(self.Args['Nx']))
