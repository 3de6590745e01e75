"""RunningMean / RunningNormalize (mannequin/_normalize.py:4-81) on the device.

The statistics (fp64 mean / variance vectors) live in device memory and every
pass over the data is a kernel from csrc/mq_norm.cu / mq_adam.cu.  Only the
data-independent scalars -- total weights, sample count, which branch of the
update applies -- are tracked on the host, exactly as the reference computes
them (they depend on batch sizes and weights alone).
"""
import numpy as np
import torch

from . import _device as D
from . import dist as mqdist
from ._lib import MQ_F32, MQ_F64


def _code(t):
    return MQ_F64 if t.dtype == torch.float64 else MQ_F32


def _dev_in(values):
    """float32 device data stays float32 (exactly representable in fp64)."""
    if isinstance(values, D.DeviceVector):
        values = values.dev()
    if D.is_dev(values):
        return values if values.dtype in (torch.float32, torch.float64) else values.to(torch.float64)
    a = np.asarray(values)
    return D.from_host(a.reshape(-1), np.float32 if a.dtype == np.float32 else np.float64)


class RunningMean(object):
    def __init__(self, shape=(), *, horizon=None):
        self._shape = tuple(np.zeros(shape).shape)
        self._n = int(np.prod(self._shape)) if self._shape else 1
        self._mean = D.zeros(self._n, torch.float64)
        self._tw = 0.0
        self._horizon = horizon

    def _coef(self, weight):
        weight = float(weight)
        assert weight >= 0.0
        if self._horizon is not None:
            self._tw *= np.power(1.0 - 1.0 / float(self._horizon), weight)
        self._tw += weight
        return weight / self._tw

    def update(self, values, *, weight=1.0, _old_out=None):
        x = _dev_in(values)
        assert x.numel() == self._n
        coef = self._coef(weight)
        D.call("mq_running_mean_update", D.ptr(self._mean), D.ptr(x), _code(x), self._n, coef,
               D.ptr(_old_out), D.stream())

    def get_dev(self):
        return self._mean

    def get(self):
        return D.to_host(self._mean).reshape(self._shape)

    def __call__(self, values, **kwargs):
        self.update(values, **kwargs)
        return self.get()


class RunningNormalize(object):
    def __init__(self, shape=(), *, horizon=None):
        self._shape = tuple(np.zeros(shape).shape)
        self._d = int(np.prod(self._shape)) if self._shape else 1
        self._mean = RunningMean(self._shape, horizon=horizon)
        self._var = RunningMean(self._shape, horizon=horizon)
        self._stats = D.zeros(2 * self._d, torch.float64)
        self._n_samples = 0

    # -- device-level API (used by the fused PPO / ES drivers) -------------------
    def update_dev(self, x, rows, sharded=False):
        """x: device float32/float64 tensor holding rows*D values.  With sharded=True every rank
        passes its equally sized shard of one global batch: the two batch statistics are
        all-reduced (2 x D doubles), everything else is replicated."""
        lib = D._lib.lib()
        d = self._d
        w = mqdist.world() if sharded else 1
        # ONE pass over the data: s1 = mean(x - m), s2 = mean((x - m)^2) with m = the running mean before this update
        # (csrc/mq_norm.cu); the mean and variance updates of _normalize.py:44-57 follow from them on the device
        wsb = lib.mq_norm_stats_ws_bytes(rows, d)
        ws = D.workspace("norm", wsb)
        stats = self._stats
        D.call("mq_norm_stats", D.ptr(x), _code(x), rows, d, D.ptr(self._mean.get_dev()), 1.0 / w, D.ptr(stats),
               D.ptr(ws), wsb, D.stream())
        if w > 1:
            mqdist.allreduce_sum_(stats)                                  # 2 x D doubles, once per update
        coef_mean = self._mean._coef(1.0)                                 # one weight-1 update per batch
        rows = rows * w
        coef_var, scale = 0.0, 1.0
        if self._n_samples >= 1:
            coef_var = self._var._coef(1.0)
        elif rows >= 2:
            weight = (rows - 1.0) / rows                                  # _normalize.py:50-55
            scale = 1.0 / weight
            coef_var = self._var._coef(weight)
        D.call("mq_norm_update_from_stats", D.ptr(self._mean.get_dev()), D.ptr(self._var.get_dev()), D.ptr(stats), d,
               coef_mean, coef_var, scale, D.stream())
        self._n_samples += rows

    def apply_dev(self, x, rows, out=None, fuse_tanh=False, out_dtype=torch.float64):
        if out is None:
            out = torch.empty(rows * self._d, dtype=out_dtype, device=x.device)
        if self._n_samples < 2:
            out.zero_()
            return out
        D.call("mq_norm_apply", D.ptr(x), _code(x), rows, self._d, D.ptr(self._mean.get_dev()),
               D.ptr(self._var.get_dev()), D.ptr(out), _code(out), 1 if fuse_tanh else 0, D.stream())
        return out

    # -- reference API -----------------------------------------------------------
    def update(self, value):
        a = np.asarray(value) if not D.is_dev(value) else None
        x = _dev_in(value)
        assert x.numel() % self._d == 0
        self.update_dev(x, x.numel() // self._d)

    def apply(self, value):
        if self._n_samples < 2:
            return np.zeros_like(value, dtype=np.float64)
        shape = np.shape(value)
        x = _dev_in(value)
        assert x.numel() % self._d == 0
        out = self.apply_dev(x, x.numel() // self._d)
        return D.to_host(out).reshape(shape)

    def get_mean(self):
        return self._mean.get()

    def get_var(self):
        if self._n_samples < 2:
            return np.ones(self._shape, dtype=np.float64)
        return self._var.get()

    def get_std(self):
        return np.sqrt(self.get_var())

    def __call__(self, value):
        self.update(value)
        return self.apply(value)
