"""Layer graph with the API of mannequin/basicnet.py, executed on the B200.

Graph construction (Input -> Affine -> LReLU ...) is plain Python as in the
reference.  Parameters live in a device arena so that the flat parameter vector
of a net (DFS post-order of Params nodes, basicnet.py:13-23,35-51) is ONE
contiguous device slice.  At the first `evaluate` a root is pattern-matched:

  * Input -> [Reshape] -> (Affine -> LReLU|Tanh)* -> Affine [-> act]  becomes an
    MLP plan: one fused launch when the weights fit in shared memory
    (csrc/mq_fused.cu), else one GEMM launch per layer (csrc/mq_gemm.cu);
  * anything else runs node by node through the device ops of
    mannequin2_b200.backprop; user-supplied Python `f`s receive NumPy arrays,
    exactly like in the reference (they are host code by definition).

`evaluate(x)` returns a read-only ndarray and a `backprop`; `backprop(dy)`
returns a device-backed flat float32 vector (DeviceVector) holding batch-MEAN
parameter gradients (basicnet.py:242-249).
"""

import weakref

import numpy as np
import torch

from . import _device as D
from . import backprop
from ._lib import ACT_LRELU, ACT_TANH
from ._utils import endswith


# ---------------------------------------------------------------- parameter arena
class _Arena(object):
    """Device memory for `Params`, handed out in order so that the parameters of a net built in one go are ONE slice (the
    flat vector of get_params / load_params).  Only the chunk being filled is held strongly; a full chunk lives as long as
    a `Params` created in it does (each keeps its chunk alive), so nets that are dropped give their memory back."""
    CHUNK = 1 << 22                      # floats per chunk (16 MB)

    def __init__(self):
        self.chunks = []                 # [tensor (the last chunk) or weakref to it (full chunks), floats used]

    def alloc(self, n):
        if not self.chunks or self.chunks[-1][1] + n > self.chunks[-1][0].numel():
            if self.chunks:
                self.chunks[-1][0] = weakref.ref(self.chunks[-1][0])
            self.chunks.append([D.zeros(max(self.CHUNK, n)), 0])
        c = self.chunks[-1]
        cid, off = len(self.chunks) - 1, c[1]
        c[1] += n
        return cid, off, c[0][off:off + n], c[0]

    def view(self, cid, off, n):
        t = self.chunks[cid][0]
        if isinstance(t, weakref.ref):
            t = t()
            assert t is not None, "parameter chunk already released"
        return t[off:off + n]


_arenas = {}


def _arena():
    key = torch.cuda.current_device() if torch.cuda.is_available() else -1
    if key not in _arenas:
        _arenas[key] = _Arena()
    return _arenas[key]


# ---------------------------------------------------------------------- base class
class Layer(object):
    def __init__(self, *inputs, evaluate, shape, n_local_params=0, get_params=None, load_params=None):
        inputs = tuple(inputs)
        shape = tuple(max(1, int(s)) for s in shape)
        n_local_params = max(0, int(n_local_params))

        if n_local_params > 0:
            assert len(inputs) == 0          # composite layers cannot have parameters
            assert get_params is not None and load_params is not None
            prerequisites = []
            n_params = n_local_params
        else:
            prerequisites, seen = [], set()

            def visit(node):                 # DFS post-order, each node once
                if node in seen:
                    return
                seen.add(node)
                for i in node.inputs:
                    visit(i)
                prerequisites.append(node)
            for i in inputs:
                visit(i)
            params = [l for l in prerequisites if l.n_local_params >= 1]
            n_params = sum(l.n_local_params for l in params)
            flat = _flat_slice(params)

            def get_params(*, output=None):
                if output is None:
                    output = np.zeros(n_params, dtype=np.float32)
                assert output.shape == (n_params,)
                if flat is not None:
                    output[:] = D.to_host(flat)
                else:
                    pos = 0
                    for l in params:
                        l.get_params(output=output[pos:pos + l.n_local_params])
                        pos += l.n_local_params
                return output

            def load_params(new_value):
                if isinstance(new_value, D.DeviceVector):
                    src = getattr(new_value, "f32", None)
                    src = src if src is not None else new_value.dev(torch.float32)
                    assert src.numel() == n_params
                else:
                    host = np.asarray(new_value, dtype=np.float32)
                    assert host.shape == (n_params,)
                    src = D.from_host(host)
                if flat is not None:
                    flat.copy_(src)                       # one device-to-device copy
                else:
                    pos = 0
                    for l in params:
                        l.dev.copy_(src[pos:pos + l.n_local_params].view(l.dev.shape))
                        pos += l.n_local_params
            self._flat = flat
            self._params = params

        self.inputs = inputs
        self.evaluate = evaluate
        self.prerequisites = tuple(prerequisites)
        self.shape = shape
        self.n_params = n_params
        self.n_local_params = n_local_params
        self.get_params = get_params
        self.load_params = load_params

    def __call__(self, *args, **kwargs):
        outs, _ = self.evaluate(*args, **kwargs)
        return outs

    def __add__(self, other):
        return Add(self, other)

    def __mul__(self, other):
        return Multiply(self, other)

    def __str__(self):
        return "<%s: %s>" % (self.__class__.__name__, self.shape)

    __repr__ = __str__


def _flat_slice(params):
    """The arena slice covering `params` if they are contiguous and in order."""
    if not params:
        return None
    if not all(isinstance(p, Params) for p in params):
        return None
    cid, off = params[0]._cid, params[0]._off
    pos = off
    for p in params:
        if p._cid != cid or p._off != pos:
            return None
        pos += p.n_local_params
    return _arena().view(cid, off, pos - off)


def _readonly(a):
    a.setflags(write=False)
    return a


class Input(Layer):
    def __init__(self, *shape):
        self._grad_dev = None
        self._grad_fn = None

        def backprop(grad, output=None):
            assert endswith(tuple(grad.shape), self.shape)
            self._grad_dev, self._grad_fn = grad, None
            if output is None:
                return []
            assert len(output) == 0
            return output

        def evaluate(array):
            array = np.asarray(array, dtype=np.float32)
            if not endswith(array.shape, self.shape):
                raise ValueError("Input doesn't match shape " + str(self.shape))
            return array, backprop

        super().__init__(evaluate=evaluate, shape=shape)
        self._dev_backprop = backprop

    @property
    def last_gradient(self):
        """Gradient w.r.t. the input of the latest backprop (basicnet.py:94-99)."""
        if self._grad_fn is not None:
            self._grad_dev, self._grad_fn = self._grad_fn(), None
        if self._grad_dev is None:
            return None
        if not D.is_dev(self._grad_dev):                 # a bare Input back-propagated with a host array: grad[:] (basicnet.py:94-99)
            return _readonly(np.array(self._grad_dev, dtype=np.float32))
        return _readonly(D.to_host(self._grad_dev))


class Const(Layer):
    def __init__(self, value):
        value = np.array(value, dtype=np.float32)
        assert len(value.shape) >= 1
        value.setflags(write=False)
        self._host = value
        self._devval = None

        def backprop(grad, output=None):
            assert endswith(tuple(grad.shape), value.shape)
            if output is None:
                return []
            assert len(output) == 0
            return output

        super().__init__(evaluate=lambda array: (value[:], backprop), shape=value.shape)
        self._dev_backprop = backprop

    @property
    def dev(self):
        if self._devval is None:
            self._devval = D.from_host(self._host)
        return self._devval


class Params(Layer):
    def __init__(self, *shape, init=lambda *s: np.zeros(s)):
        shape = tuple(max(1, int(s)) for s in shape)
        assert len(shape) >= 1
        value = np.array(init(*shape), dtype=np.float32).reshape(shape)
        self._cid, self._off, flat, self._chunk = _arena().alloc(value.size)     # (_chunk: keeps the arena chunk alive)
        flat.copy_(D.from_host(value.reshape(-1)))
        self._flatdev = flat

        def backprop(grad, output=None):
            if output is None:
                return grad.reshape(value.size)
            output[:] = grad.reshape(value.size)
            return output

        def get(*, output=None):
            host = D.to_host(flat)
            if output is None:
                return _readonly(host)
            output[:] = host
            return output

        def load(new_value):
            if D.is_dev(new_value):
                flat.copy_(new_value.reshape(-1))
            else:
                flat.copy_(D.from_host(np.asarray(new_value, dtype=np.float32).reshape(-1)))

        super().__init__(evaluate=lambda array: (_readonly(D.to_host(flat).reshape(shape)), backprop),
                         shape=shape, n_local_params=value.size, get_params=get, load_params=load)

    @property
    def dev(self):
        return self._flatdev.view(self.shape)


class Function(Layer):
    def __init__(self, *args, f, shape=None):
        assert len(args) >= 1
        if shape is None:
            shape = args[0].shape
        self._f = f
        self._plan = False                 # False: not compiled yet; None: no fast plan
        super().__init__(*args, evaluate=self._evaluate, shape=shape)

    # -- one node, device values in / out ------------------------------------------
    def local_evaluate(self, inps, **kwargs):
        f = self._f
        inps = [inps[a] for a in self.inputs]
        for i, a in zip(inps, self.inputs):
            assert endswith(tuple(i.shape), a.shape)
        if getattr(f, "_mq_dev", False):
            val, bpp = f(*inps, **kwargs)
        else:
            # user-supplied Python function: NumPy in, NumPy out, as in the reference
            host = [_readonly(D.to_host(i)) for i in inps]
            hval, hbpp = f(*host, **kwargs)
            hval = np.asarray(hval)
            if not endswith(hval.shape, self.shape):
                raise ValueError("Output of '%s' doesn't match shape %s"
                                 % (getattr(f, "__name__", "f"), self.shape))
            val = D.from_host(hval, np.float32)

            def bpp(g, _h=hbpp):
                gs = _h(D.to_host(g))
                assert isinstance(gs, tuple)
                return tuple(D.from_host(np.asarray(x), np.float32) for x in gs)
        if not endswith(tuple(val.shape), self.shape):
            raise ValueError("Output of '%s' doesn't match shape %s"
                             % (getattr(f, "__name__", "f"), self.shape))
        return val, bpp

    # -- root evaluation --------------------------------------------------------------
    def _evaluate(self, array, **kwargs):
        array = np.asarray(array, dtype=np.float32)
        if self._plan is False:
            from . import _plan
            self._plan = _plan.compile_root(self)
        if self._plan is not None:
            return self._plan.evaluate(array, **kwargs)
        return self._evaluate_graph(array, **kwargs)

    def evaluate_dev(self, x_dev, **kwargs):
        """Device tensors in, device tensors out: (output, backprop) with backprop(grad_dev) -> DeviceVector.  Always the
        node-by-node path; nothing synchronises with the host, so a whole step can be captured in a CUDA graph."""
        return self._evaluate_graph(None, x_dev=x_dev, **kwargs)

    def _evaluate_graph(self, array, x_dev=None, **kwargs):
        on_dev = x_dev is not None
        if on_dev:
            assert D.is_dev(x_dev) and x_dev.dtype == torch.float32
            x_dev = x_dev.contiguous()
            array = x_dev                       # only its .shape is used below
        else:
            x_dev = D.from_host(array)
        outs, bpps = {}, {}
        for layer in self.prerequisites:
            if isinstance(layer, Function):
                out, bpp = layer.local_evaluate(outs)
            elif isinstance(layer, Input):
                if not endswith(array.shape, layer.shape):
                    raise ValueError("Input doesn't match shape " + str(layer.shape))
                out, bpp = x_dev, layer._dev_backprop
            elif isinstance(layer, (Params, Const)):
                out, bpp = layer.dev, None
            else:                          # foreign Layer subclass: host evaluate
                if on_dev:
                    raise NotImplementedError("evaluate_dev: %r has no device evaluation" % (layer,))
                hout, hbpp = layer.evaluate(array)
                out, bpp = D.from_host(np.asarray(hout), np.float32), hbpp
            outs[layer], bpps[layer] = out, bpp
        out, bpp = self.local_evaluate(outs, **kwargs)
        outs[self], bpps[self] = out, bpp
        out_shape = tuple(out.shape)
        shape, n_params = self.shape, self.n_params
        exec_order = self.prerequisites + (self,)

        def backprop(grad, *, output=None):
            if not (D.is_dev(grad) or isinstance(grad, D.DeviceVector)):
                grad = np.asarray(grad, dtype=np.float32)
            assert tuple(grad.shape) == out_shape
            g = D.as_dev_f32(grad).reshape(out_shape)
            batch_shape = out_shape[:len(out_shape) - len(shape)]
            grads = {self: g}
            flat = D.zeros(n_params)
            pos = n_params
            for layer in reversed(exec_order):
                if not isinstance(layer, Function):
                    pos -= layer.n_local_params
                    lg = grads.pop(layer, None)
                    if isinstance(layer, Params):
                        if lg is not None:
                            flat[pos:pos + layer.n_local_params].copy_(lg.reshape(-1))
                    elif isinstance(layer, (Input, Const)):
                        if lg is not None:
                            layer._dev_backprop(lg)
                    elif lg is not None:
                        bpps[layer](D.to_host(lg))
                    continue
                if layer not in grads:
                    continue
                inp_grads = bpps[layer](grads.pop(layer))
                assert isinstance(inp_grads, tuple)
                assert len(inp_grads) == len(layer.inputs)
                for gi, i in zip(inp_grads, layer.inputs):
                    assert tuple(gi.shape) == tuple(outs[i].shape)
                    g_batch = tuple(gi.shape)[:gi.dim() - len(i.shape)]
                    assert endswith(batch_shape, g_batch)
                    if len(g_batch) < len(batch_shape):
                        # batch average (basicnet.py:242-249)
                        div = float(np.prod(batch_shape[:len(batch_shape) - len(g_batch)]))
                        gi = _scaled(gi, 1.0 / div)
                    grads[i] = backprop_add(grads[i], gi) if i in grads else gi
            assert pos == 0
            if output is not None:
                assert output.shape == (n_params,)
                output[:] = D.to_host(flat)
                return output
            return D.DeviceVector(flat, (n_params,), writeable=True)

        if on_dev:
            return out, backprop
        return _readonly(D.to_host(out)), backprop


def _scaled(t, alpha):
    s = torch.full((1,), alpha, dtype=torch.float32, device=t.device)
    out = torch.empty_like(t)
    D.call("mq_binary", D.ptr(t.contiguous()), D.ptr(s), D.ptr(out), t.numel(), 1, 1, D.stream())
    return out


def backprop_add(a, b):
    out = torch.empty_like(a)
    D.call("mq_binary", D.ptr(a.contiguous()), D.ptr(b.contiguous()), D.ptr(out), a.numel(), a.numel(),
           0, D.stream())
    return out


# ------------------------------------------------------------------- constructors
def normed_columns(inps, outs):
    """randn(in,out) with unit-L2 columns (basicnet.py:265-267); host RNG so that
    np.random.seed reproduces the reference's initial weights."""
    m = np.random.randn(inps, outs).astype(np.float32)
    return m / np.sqrt(np.sum(np.square(m), axis=0))


def Add(a, b):
    return Function(a, b, f=backprop.add)


def Multiply(a, b):
    return Function(a, b, f=backprop.multiply)


def Reshape(inner, out_shape):
    out_shape = np.zeros(inner.shape).reshape(out_shape).shape

    def f(v):
        assert endswith(tuple(v.shape), inner.shape)
        batch_shape = tuple(v.shape)[:len(v.shape) - len(inner.shape)]
        in_shape = tuple(v.shape)
        return v.reshape(batch_shape + out_shape), lambda g: (g.reshape(in_shape),)
    f._mq_dev = True
    f._mq_op = ("reshape",)
    return Function(inner, f=f, shape=out_shape)


def Linear(inner, *shape, init=normed_columns):
    shape = tuple(max(1, int(s)) for s in shape)
    if len(shape) > 1:
        return Reshape(Linear(inner, int(np.prod(shape)), init=init), shape)
    if len(inner.shape) > 1:
        inner = Reshape(inner, -1)
    if not callable(init):
        init = (lambda m: lambda *s: normed_columns(*s) * m)(float(init))
    return Function(inner, Params(inner.shape[0], shape[0], init=init), f=backprop.matmul, shape=shape)


def Bias(inner, *, init=lambda *s: np.zeros(s)):
    return Add(inner, Params(*inner.shape, init=init))


def LReLU(inner, *, leak=0.1):
    f = lambda a: backprop.relu(a, leak=leak)
    f._mq_dev = True
    f._mq_op = ("act", ACT_LRELU, float(leak))
    return Function(inner, f=f)


def Tanh(inner):
    f = lambda a: backprop.tanh(a)
    f._mq_dev = True
    f._mq_op = ("act", ACT_TANH, 0.0)
    return Function(inner, f=f)


def Affine(*args, **kwargs):
    return Bias(Linear(*args, **kwargs))
