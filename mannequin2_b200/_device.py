"""Device memory, streams and host<->device copies (PyTorch is plumbing only).

Every numeric operation of this package runs in libmq_b200.so on the current
CUDA device.  There is no CPU execution path: `device()` raises when CUDA is
not available.
"""
import numpy as np
import torch

from . import _lib

_F = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
      np.dtype(np.int32): torch.int32, np.dtype(np.uint8): torch.uint8,
      np.dtype(np.bool_): torch.uint8, np.dtype(np.int64): torch.int64}


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("mannequin2_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream():
    """cudaStream_t of torch's current stream, as an int for ctypes."""
    return torch.cuda.current_stream().cuda_stream


def synchronize():
    torch.cuda.current_stream().synchronize()


def empty(n, dtype=torch.float32):
    return torch.empty(int(n), dtype=dtype, device=device())


def zeros(n, dtype=torch.float32):
    return torch.zeros(int(n), dtype=dtype, device=device())


def from_host(a, dtype=None, pinned=False):
    """Copy a host array to the device (flattening is the caller's business)."""
    a = np.ascontiguousarray(a if dtype is None else np.asarray(a, dtype=dtype))
    if a.dtype == np.bool_:
        a = a.view(np.uint8)
    t = torch.from_numpy(a)
    if pinned:
        t = t.pin_memory()
    return t.to(device(), non_blocking=pinned)


def to_host(t):
    """Device tensor -> fresh NumPy array (synchronises the current stream)."""
    return t.detach().to("cpu").numpy()


def ptr(t):
    return None if t is None else t.data_ptr()


def is_dev(x):
    return isinstance(x, torch.Tensor)


def as_dev_f32(x):
    """Device float32 tensor from a tensor, a DeviceVector or anything array-like."""
    if isinstance(x, DeviceVector):
        return x.dev(torch.float32)
    if isinstance(x, torch.Tensor):
        return x if x.dtype == torch.float32 else x.to(torch.float32)
    return from_host(x, np.float32)


_ws = {}


def workspace(key, nbytes):
    """A cached scratch buffer per (key, current device, current stream); grows, never shrinks.

    Keyed by stream so that independent learners driven on different streams never share scratch."""
    if torch.cuda.is_available():
        k = (key, torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
    else:
        k = (key, -1, 0)
    t = _ws.get(k)
    if t is None or t.numel() < nbytes:
        t = torch.empty(int(max(nbytes, 256)), dtype=torch.uint8, device=device())
        _ws[k] = t
    return t


def call(name, *args):
    """Invoke a C-ABI entry point and raise the mapped exception on failure."""
    lib = _lib.lib()
    _lib.check(getattr(lib, name)(*args), lib)


class DeviceVector(object):
    """A device-resident array that behaves like a read-only ndarray on demand.

    Returned by `backprop()` and `Adam.get_value()` so that the usual script
    sequence  backprop -> apply_gradient -> get_value -> load_params  never
    leaves the GPU.  Any NumPy use (`np.asarray`, arithmetic, printing,
    indexing) materialises a host copy once and caches it.
    """

    __array_priority__ = 1000

    def __init__(self, tensor, shape=None, writeable=False):
        self._t = tensor
        self._shape = tuple(shape) if shape is not None else (tensor.numel(),)
        self._host = None
        self._writeable = writeable
        self._host_dirty = False          # host copy modified, device stale

    # -- device side
    def dev(self, dtype=None):
        if self._host_dirty:
            self._t = from_host(self._host, self._host.dtype)
            self._host_dirty = False
        if dtype is not None and self._t.dtype != dtype:
            return self._t.to(dtype)
        return self._t

    # -- host side
    def _np(self):
        if self._host is None:
            self._host = to_host(self._t).reshape(self._shape)
            self._host.setflags(write=self._writeable)
        return self._host

    def __array__(self, dtype=None, copy=None):
        a = self._np()
        if dtype is not None and a.dtype != dtype:
            return a.astype(dtype)
        return a

    @property
    def shape(self):
        return self._shape

    @property
    def dtype(self):
        return np.dtype(str(self._t.dtype).replace("torch.", ""))

    @property
    def size(self):
        return int(np.prod(self._shape))

    ndim = property(lambda self: len(self._shape))

    def __len__(self):
        return self._shape[0]

    def reshape(self, *shape):
        return self._np().reshape(*shape)

    def astype(self, dtype):
        return self._np().astype(dtype)

    def copy(self):
        return np.array(self._np())

    def setflags(self, write=False):
        self._writeable = bool(write)
        if self._host is not None:
            self._host.setflags(write=write)

    def __getitem__(self, i):
        return self._np()[i]

    def __setitem__(self, i, v):
        if not self._writeable:
            raise ValueError("assignment destination is read-only")
        self._np()[i] = v
        self._host_dirty = True

    def __iter__(self):
        return iter(self._np())

    def __str__(self):
        return str(self._np())

    def __repr__(self):
        return repr(self._np())

    def __neg__(self):
        return -self._np()

    def __abs__(self):
        return abs(self._np())


def _binop(name):
    def f(self, other):
        return getattr(self._np(), name)(np.asarray(other))
    return f


for _n in ("add", "sub", "mul", "truediv", "pow", "radd", "rsub", "rmul", "rtruediv", "lt", "le",
           "gt", "ge", "eq", "ne", "matmul", "rmatmul"):
    setattr(DeviceVector, "__%s__" % _n, _binop("__%s__" % _n))


class DeviceRNG(object):
    """Unit normals generated on the device (counter-based Philox4x32-10, csrc/mq_rollout.cu) with the interface piece
    `Gauss(..., rng=...)` looks for (`randn_dev`).  The stream position lives in device memory and is advanced by a
    device op, so a CUDA graph that contains the draw produces fresh noise at every replay."""

    def __init__(self, seed=0):
        self.seed = int(seed)
        self._counter = torch.zeros(1, dtype=torch.int32, device=device())
        self._draws = 0

    def randn_dev(self, shape):
        out = torch.empty(tuple(shape), dtype=torch.float32, device=device())
        call("mq_randn_philox", self.seed, 0x5eed, 0, ptr(self._counter), out.numel(), ptr(out), stream())
        self._counter.add_(1)
        return out

    def randn(self, *shape):                 # host interface of np.random.RandomState, for code that wants arrays
        return to_host(self.randn_dev(shape)).astype(np.float64)
