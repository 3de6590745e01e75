"""Adam / Adams (mannequin/_adam.py:7-64): gradient ASCENT with two running moments.

State (value, first and second moment) is device-resident; one launch of the
fused kernel (csrc/mq_adam.cu) replaces RunningMean.update x2 + update_rule +
`value + add` + the fp64->fp32 `load_params` cast.  By default the state is
fp64 and follows the reference's operation order bit for bit; pass
`state_dtype="float32"` for the 24 B/param fast variant.
"""
import weakref

import numpy as np
import torch

from . import _device as D
from ._lib import MQ_F32, MQ_F64


class BaseTwoMoments(object):
    _adams = 0
    _out_ref = None

    def __init__(self, value, *, horizon, var_horizon=100, print_norm=False, state_dtype="float64",
                 **global_params):
        host = np.asarray(value.__array__() if isinstance(value, D.DeviceVector) else value,
                          dtype=np.float64)
        self._shape = host.shape
        self._n = host.size
        self._tdtype = torch.float64 if str(state_dtype) in ("float64", "f64") else torch.float32
        self._code = MQ_F64 if self._tdtype == torch.float64 else MQ_F32
        self._value = D.from_host(host.reshape(-1), np.float64).to(self._tdtype)
        self._m = torch.zeros_like(self._value)
        self._v = torch.zeros_like(self._value)
        self._theta32 = self._value.to(torch.float32)
        self._b1 = 1.0 - 1.0 / float(horizon)
        self._b2 = 1.0 - 1.0 / float(var_horizon)
        self._tw = [0.0, 0.0]
        self._print_norm = print_norm
        self._global = dict(global_params)
        self._out = None

    def _params(self, local):
        p = dict(self._global, **local)
        if "lr" not in p:
            raise TypeError("update_rule() missing 1 required keyword-only argument: 'lr'")
        extra = set(p) - {"lr", "epsilon"}
        if extra:
            raise TypeError("update_rule() got an unexpected keyword argument '%s'" % sorted(extra)[0])
        return float(p["lr"]), float(p.get("epsilon", 1e-8))

    def next_coefs(self):
        """Advance the two total weights (RunningMean with weight 1) -> (c1, c2)."""
        self._tw[0] = self._tw[0] * self._b1 + 1.0
        self._tw[1] = self._tw[1] * self._b2 + 1.0
        return 1.0 / self._tw[0], 1.0 / self._tw[1]

    def apply_gradient(self, grad, **local_params):
        lr, eps = self._params(local_params)
        if isinstance(grad, D.DeviceVector):
            g = grad.dev()
        elif D.is_dev(grad):
            g = grad
        else:
            g = np.asarray(grad, dtype=np.float64)
            assert g.shape == self._shape
            g = D.from_host(g.reshape(-1), np.float64 if self._code == MQ_F64 else np.float32)
        assert g.numel() == self._n
        if self._code == MQ_F32 and g.dtype != torch.float32:
            g = g.to(torch.float32)
        gcode = MQ_F64 if g.dtype == torch.float64 else MQ_F32
        prev = self._value.clone() if self._print_norm else None
        c1, c2 = self.next_coefs()
        self._out = None                          # (detaches a handle somebody still holds BEFORE the in-place update)
        D.call("mq_adam_step", D.ptr(self._value), D.ptr(self._m), D.ptr(self._v), D.ptr(self._theta32),
               D.ptr(g), self._n, self._code, gcode, c1, c2, lr, eps, self._adams, D.stream())
        if self._print_norm:
            add = D.to_host(self._value).astype(np.float64) - D.to_host(prev).astype(np.float64)
            print("Update norm: %10.4f" % np.sqrt(np.sum(np.square(add))))

    def get_value(self):
        """The optimiser's value (read-only, like the reference's array); device backed.  The reference returns an immutable
        SNAPSHOT (value = value + add, _adam.py:47-53): a handle the caller still holds when the next step is applied is
        detached onto a copy first (`_out = None`), so `old = get_value(); apply_gradient(g); old` still reads the old value."""
        out = self._out
        if out is None:
            out = D.DeviceVector(self._value, self._shape)
            out.f32 = self._theta32               # float32 mirror written by the same kernel
            self._out_ref = weakref.ref(out)
        return out

    @property
    def _out(self):
        return self._out_ref() if self._out_ref is not None else None

    @_out.setter
    def _out(self, handle):
        assert handle is None                     # the drivers invalidate the cached handle before they update in place
        old = self._out_ref() if self._out_ref is not None else None
        if old is not None and old._t is self._value:        # somebody still holds it: give it its own storage
            old._t = self._value.clone()
            old.f32 = self._theta32.clone()
        self._out_ref = None

    # device handles for the fused drivers
    def state(self):
        return self._value, self._m, self._v, self._theta32


class Adam(BaseTwoMoments):
    _adams = 0


class Adams(BaseTwoMoments):
    _adams = 1
