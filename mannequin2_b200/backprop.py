"""Primitive forward/backward ops (mannequin/backprop.py:5-60) as device kernels.

Each function returns `(value, backward)` exactly like the reference; values are
device tensors inside the layer graph.  Called with NumPy arrays (user code), the
same kernels run and NumPy arrays come back.
"""
import functools

import numpy as np
import torch

from . import _device as D
from ._lib import ACT_LRELU, ACT_TANH
from ._utils import endswith


def _hostable(fn):
    """Let a device op accept NumPy inputs: upload, run the kernels, download."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        if any(D.is_dev(a) for a in args):
            return fn(*args, **kwargs)
        dev_args = [D.from_host(np.asarray(a, dtype=np.float32)) for a in args]
        val, bpp = fn(*dev_args, **kwargs)

        def host_bpp(g):
            gs = bpp(D.from_host(np.asarray(g, dtype=np.float32)))
            return tuple(D.to_host(x) for x in gs)
        return D.to_host(val), host_bpp
    wrapper._mq_dev = True
    return wrapper


def _new(shape, like):
    return torch.empty(tuple(shape), dtype=torch.float32, device=like.device)


def _binary(a, b, op):
    """a (op) b with b broadcast over leading axes of a (b.shape is a suffix of a.shape)."""
    a, b = a.contiguous(), b.contiguous()
    out = _new(a.shape, a)
    D.call("mq_binary", D.ptr(a), D.ptr(b), D.ptr(out), a.numel(), max(1, b.numel()), op, D.stream())
    return out


def _reduce_leading(g, a, shape):
    """sum over the leading axes of g (times a, if given) down to `shape`."""
    cols = int(np.prod(shape)) if len(shape) else 1
    rows = g.numel() // cols
    out = _new(shape, g)
    D.call("mq_col_sum", D.ptr(g.contiguous()), D.ptr(a.contiguous()) if a is not None else None,
           rows, cols, 1.0, D.ptr(out), D.stream())
    return out


def autograd(f):
    """Bridge to the third-party `autograd` package (mannequin/backprop.py:5-9).

    User-supplied differentiable functions run on the host exactly as in the
    reference; the built-in Gaussian / KL heads do not use this bridge (they
    have closed-form device kernels)."""
    import autograd as _ag
    import autograd.numpy as anp
    df = _ag.grad(lambda a, ka, g: anp.sum(f(*a, **ka) * g))
    return lambda *a, **ka: (f(*a, **ka), lambda g: df(a, ka, g))


@_hostable
def add(a, b):
    ash, bsh = tuple(a.shape), tuple(b.shape)
    if ash == bsh:
        return _binary(a, b, 0), lambda g: (g, g)
    elif endswith(ash, bsh):
        return _binary(a, b, 0), lambda g: (g, _reduce_leading(g, None, bsh))
    elif endswith(bsh, ash):
        return _binary(b, a, 0), lambda g: (_reduce_leading(g, None, ash), g)
    raise ValueError("Invalid shapes: %s, %s" % (ash, bsh))


@_hostable
def multiply(a, b):
    ash, bsh = tuple(a.shape), tuple(b.shape)
    if ash == bsh:
        return _binary(a, b, 1), lambda g: (_binary(g, b, 1), _binary(g, a, 1))
    elif endswith(ash, bsh):
        return _binary(a, b, 1), lambda g: (_binary(g, b, 1), _reduce_leading(g, a, bsh))
    elif endswith(bsh, ash):
        # (a of lower rank: d/da = sum over the leading axes of g * b, d/db = g * a.  The reference, backprop.py:35-39, returns
        # (sum(g * a), g * b) here -- the two factors swapped, which is not the derivative; this branch follows the mathematics
        # and tests/test_host_logic.py pins the difference)
        return _binary(b, a, 1), lambda g: (_reduce_leading(g, b, ash), _binary(g, a, 1))
    raise ValueError("Invalid shapes: %s, %s" % (ash, bsh))


# Operand format of the LARGE matmul products (the TMA-fed tcgen05 GEMM): 0 = three bf16 planes (exact, any range); trainers
# whose operands are bounded (images, tanh activations, clipped log-deviations) select "fp16x2" for their own steps -- the
# same fp32 bar in half the tensor work, operands must stay below 65504 (include/mq_b200.h, MQ_TC_FP16).
_GEMM_TC = [0]


class gemm_mode(object):
    """with gemm_mode("fp16x2"): ... -- operand format of the large matmul products evaluated inside the block."""

    def __init__(self, mode):
        self.tc = D._lib.TC_MODES[mode] if mode != "exact" else 0

    def __enter__(self):
        self.prev, _GEMM_TC[0] = _GEMM_TC[0], self.tc

    def __exit__(self, *exc):
        _GEMM_TC[0] = self.prev


def _gemm(a, a_rs, a_cs, b, b_rs, b_cs, m, n, k, like):
    c = _new((m, n), like)
    if m * n * k >= (1 << 26) or (k >= 1024 and m * n <= (1 << 22)):
        # large products: TMA-fed tcgen05 kernel (needs room for the packed bf16 planes); few output tiles with a long
        # contraction (dW = x^T g over a large batch): split along K
        wsb = int(D._lib.lib().mq_gemm_ws_bytes(m, n, k))
        ws = D.workspace("gemm", wsb)
        D.call("mq_gemm_ex", D.ptr(a), a_rs, a_cs, D.ptr(b), b_rs, b_cs, D.ptr(c), m, n, k, 1.0, _GEMM_TC[0], D.ptr(ws), wsb,
               D.stream())
    else:
        D.call("mq_gemm", D.ptr(a), a_rs, a_cs, D.ptr(b), b_rs, b_cs, D.ptr(c), m, n, k, 1.0, D.stream())
    return c


@_hostable
def matmul(a, b):
    assert len(a.shape) in (1, 2)
    assert len(b.shape) == 2
    a, b = a.contiguous(), b.contiguous()
    k, n = b.shape
    m = a.shape[0] if a.dim() == 2 else 1
    assert a.shape[-1] == k
    out = _gemm(a, k, 1, b, n, 1, m, n, k, a)
    batched = a.dim() == 2

    def bpp(g):
        g = g.contiguous()
        ga = _gemm(g, n, 1, b, 1, n, m, k, n, g)            # g @ b.T
        gb = _gemm(a, 1, k, g, n, 1, k, n, m, g)            # a.T @ g  (outer product if m == 1)
        return (ga if batched else ga.view(k)), gb
    return (out if batched else out.view(n)), bpp


def _act(a, code, leak):
    a = a.contiguous()
    y = _new(a.shape, a)
    D.call("mq_act_forward", D.ptr(a), D.ptr(y), a.numel(), code, leak, D.stream())

    def bpp(g):
        out = _new(a.shape, a)
        D.call("mq_act_backward", D.ptr(y), D.ptr(g.contiguous()), D.ptr(out), a.numel(), code, leak,
               D.stream())
        return (out,)
    return y, bpp


@_hostable
def relu(a, *, leak=0.0):
    return _act(a, ACT_LRELU, float(leak))


@_hostable
def tanh(a):
    return _act(a, ACT_TANH, 0.0)


add._mq_op = ("add",)
multiply._mq_op = ("multiply",)
matmul._mq_op = ("matmul",)
