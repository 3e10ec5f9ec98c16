"""Device-array duck type the reference's callers rely on (pyopencl.array.Array):
.data, .get(), .size, .shape, .dtype, .fill(v), slicing / slice assignment from
device or NumPy data, scalar += / *=, .astype, .map_to_host().  Storage is a torch
CUDA tensor (torch is the allocator and stream plumbing, nothing more)."""
import numpy as np
import torch

_NP2T = {np.dtype(np.float64): torch.float64, np.dtype(np.complex128): torch.complex128,
         np.dtype(np.uint32): torch.int32, np.dtype(np.int32): torch.int32,
         np.dtype(np.int64): torch.int64, np.dtype(np.float32): torch.float32,
         np.dtype(np.uint8): torch.uint8, np.dtype(np.int8): torch.int8}


def torch_dtype(dtype):
    return _NP2T[np.dtype(dtype)]


class DevArray:
    __slots__ = ("t", "dtype")
    __array_priority__ = 1000

    def __init__(self, tensor, dtype=None):
        self.t = tensor
        if dtype is None:
            dtype = {torch.float64: np.float64, torch.complex128: np.complex128,
                     torch.int32: np.int32, torch.int64: np.int64,
                     torch.float32: np.float32, torch.uint8: np.uint8,
                     torch.int8: np.int8}[tensor.dtype]
        self.dtype = np.dtype(dtype)

    # ---- construction helpers
    @staticmethod
    def empty(shape, dtype, device):
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),)
        return DevArray(torch.empty(tuple(int(s) for s in shape), dtype=torch_dtype(dtype),
                                    device=device), dtype)

    @staticmethod
    def zeros(shape, dtype, device):
        a = DevArray.empty(shape, dtype, device)
        a.t.zero_()
        return a

    @staticmethod
    def from_numpy(arr, device):
        arr = np.ascontiguousarray(arr)
        dt = arr.dtype
        src = arr.view(np.int32) if dt == np.uint32 else arr
        return DevArray(torch.from_numpy(src.copy()).to(device), dt)

    # ---- pyopencl-like surface
    @property
    def data(self):
        return self

    @property
    def ptr(self):
        return self.t.data_ptr()

    @property
    def size(self):
        return self.t.numel()

    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def nbytes(self):
        return self.t.numel() * self.t.element_size()

    def get(self):
        out = self.t.detach().cpu().numpy()
        if self.dtype == np.uint32:
            out = out.view(np.uint32)
        return out

    map_to_host = get

    def item(self):
        return self.get().item()

    def fill(self, val):
        self.t.fill_(val)
        return self

    def astype(self, dtype):
        return DevArray(self.t.to(torch_dtype(dtype)), dtype)

    def copy(self):
        return DevArray(self.t.clone(), self.dtype)

    def __len__(self):
        return self.t.shape[0]

    def __getitem__(self, key):
        return DevArray(self.t[key], self.dtype)

    def _coerce(self, val):
        if isinstance(val, DevArray):
            return val.t
        if isinstance(val, np.ndarray):
            src = val.view(np.int32) if val.dtype == np.uint32 else val
            return torch.from_numpy(np.ascontiguousarray(src)).to(self.t.device)
        return val

    def __setitem__(self, key, val):
        v = self._coerce(val)
        if isinstance(v, torch.Tensor) and v.dtype != self.t.dtype:
            if v.is_complex() and not self.t.is_complex():
                v = v.real
            v = v.to(self.t.dtype)
        self.t[key] = v

    def _bin(self, other, op):
        return DevArray(op(self.t, self._coerce(other)))

    def __add__(self, o): return self._bin(o, torch.add)
    __radd__ = __add__
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return DevArray(self._coerce(o) - self.t)
    def __mul__(self, o): return self._bin(o, torch.mul)
    __rmul__ = __mul__
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __rtruediv__(self, o): return DevArray(self._coerce(o) / self.t)
    def __neg__(self): return DevArray(-self.t, self.dtype)

    def __iadd__(self, o):
        self.t += self._coerce(o)
        return self

    def __isub__(self, o):
        self.t -= self._coerce(o)
        return self

    def __imul__(self, o):
        self.t *= self._coerce(o)
        return self

    def __repr__(self):
        return "DevArray(shape=%s, dtype=%s, device=%s)" % (self.shape, self.dtype, self.t.device)


def sqrt(a):
    """pyopencl.clmath.sqrt stand-in (used by particle creation only)."""
    return DevArray(torch.sqrt(a.t))
