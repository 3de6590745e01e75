"""Plan compiler: map a layer graph root onto the MLP kernels.

Recognised roots (everything else runs node by node, basicnet.Function):

  chain  = Input -> [Reshape(-1)] -> (Linear -> Bias -> [LReLU|Tanh])+
  logp   = Gauss.logprob over `chain` with a learnt Params(A) logstd
           (mannequin/distrib.py:61-68,73)
  logp   = Discrete.logprob over `chain` (mannequin/distrib.py:24-32,35)

A root is only compiled when its Params are one contiguous arena slice in the
reference's flat order (W1, b1, W2, b2, ..., logstd), so the kernels can address
the parameters through a single pointer.
"""
import ctypes

import numpy as np

from . import _device as D
from . import _lib
from ._lib import ACT_NONE, HEAD_DISCRETE_LOGP, HEAD_DY, HEAD_GAUSS_LOGP, MqHead
from ._utils import endswith


def _op(node):
    return getattr(getattr(node, "_f", None), "_mq_op", None)


def _match_chain(root):
    """-> (input_node, [(W, b, act, leak), ...]) or None."""
    from .basicnet import Function, Input, Params
    layers = []
    node = root
    while True:
        act, leak = ACT_NONE, 0.0
        op = _op(node)
        if op and op[0] == "act":
            act, leak = op[1], op[2]
            node = node.inputs[0]
            op = _op(node)
        if not (op and op[0] == "add" and len(node.inputs) == 2):
            return None
        lin, b = node.inputs
        if not isinstance(b, Params) or len(b.shape) != 1:
            return None
        if not (_op(lin) and _op(lin)[0] == "matmul"):
            return None
        inner, w = lin.inputs
        if not isinstance(w, Params) or len(w.shape) != 2 or w.shape[1] != b.shape[0]:
            return None
        layers.append((w, b, act, leak))
        node = inner
        op = _op(node)
        if op and op[0] == "reshape" and isinstance(node.inputs[0], Input) and len(node.shape) == 1:
            node = node.inputs[0]
        if isinstance(node, Input):
            break
        if not isinstance(node, Function):
            return None
    layers.reverse()
    k = int(np.prod(node.shape))
    for w, b, _, _ in layers:
        if w.shape[0] != k:
            return None
        k = w.shape[1]
    return node, layers


def compile_root(root):
    from .basicnet import Params
    op = _op(root)
    if op and op[0] == "gauss_logprob":
        mean, logstd = root.inputs
        if not isinstance(logstd, Params) or len(logstd.shape) != 1:
            return None
        m = _match_chain(mean)
        if m is None or len(m[1]) > _lib.MAX_LAYERS:
            return None
        inp, layers = m
        order = [p for l in layers for p in (l[0], l[1])] + [logstd]
        if list(root._params) != order or root._flat is None:
            return None
        return MLPPlan(root, inp, layers, gauss=True)
    if op and op[0] == "discrete_logprob":
        m = _match_chain(root.inputs[0])
        if m is None or len(m[1]) > _lib.MAX_LAYERS:
            return None
        inp, layers = m
        order = [p for l in layers for p in (l[0], l[1])]
        if list(root._params) != order or root._flat is None:
            return None
        return MLPPlan(root, inp, layers, gauss=False, discrete=True)
    m = _match_chain(root)
    if m is None or len(m[1]) > _lib.MAX_LAYERS:
        return None
    inp, layers = m
    order = [p for l in layers for p in (l[0], l[1])]
    if list(root._params) != order or root._flat is None:
        return None
    return MLPPlan(root, inp, layers, gauss=False)


class MLPPlan(object):
    def __init__(self, root, inp, layers, gauss, discrete=False):
        self.root, self.inp, self.gauss, self.discrete = root, inp, gauss, discrete
        self.n_in = int(np.prod(inp.shape))
        self.net = _lib.make_net(self.n_in, [l[0].shape[1] for l in layers], [l[2] for l in layers],
                                 leak=[l[3] for l in layers], logstd=gauss)
        assert self.net.n_params == root.n_params
        self.theta = root._flat
        lib = _lib.lib()
        self.fused_fwd = lib.mq_fused_tile_rows(ctypes.byref(self.net), 0) > 0
        self.fused_bwd = lib.mq_fused_tile_rows(ctypes.byref(self.net), 1) > 0

    # ------------------------------------------------------------------ helpers
    def _batch(self, array):
        if not endswith(array.shape, self.inp.shape):
            raise ValueError("Input doesn't match shape " + str(self.inp.shape))
        batch_shape = array.shape[:len(array.shape) - len(self.inp.shape)]
        return batch_shape, int(np.prod(batch_shape)) if batch_shape else 1

    def _forward_layered(self, x, batch):
        lib, net = _lib.lib(), self.net
        acts = D.empty(lib.mq_mlp_acts_floats(ctypes.byref(net), batch))
        D.call("mq_mlp_forward", ctypes.byref(net), D.ptr(self.theta), D.ptr(x), None, batch, D.ptr(acts),
               D.stream())
        n_last = net.n_out * batch
        return acts, acts[acts.numel() - n_last:]

    def _backward_layered(self, x, acts, dy, batch, want_dx):
        lib, net = _lib.lib(), self.net
        wsb = lib.mq_mlp_ws_bytes(ctypes.byref(net), batch)
        ws = D.workspace("mlp", wsb)
        grad = D.zeros(net.n_params)
        dx = D.empty(batch * self.n_in) if want_dx else None
        D.call("mq_mlp_backward", ctypes.byref(net), D.ptr(self.theta), D.ptr(x), None, D.ptr(acts),
               D.ptr(dy), batch, 1.0 / batch, D.ptr(grad), D.ptr(dx), D.ptr(ws), wsb, D.stream())
        return grad, dx

    # ------------------------------------------------------------------ evaluate
    def evaluate(self, array, **kwargs):
        if self.gauss or self.discrete:
            return self._evaluate_logprob(array, **kwargs)
        if kwargs:
            raise TypeError("unexpected keyword arguments %s" % sorted(kwargs))
        batch_shape, batch = self._batch(array)
        net, root = self.net, self.root
        x = D.from_host(array.reshape(batch, self.n_in))
        acts = None
        if self.fused_fwd:
            out = D.empty(batch * net.n_out)
            D.call("mq_fused_forward", ctypes.byref(net), D.ptr(self.theta), D.ptr(x), None, batch,
                   D.ptr(out), None, None, D.stream())
        else:
            acts, out = self._forward_layered(x, batch)
        out_shape = batch_shape + root.shape
        host_out = D.to_host(out).reshape(out_shape)
        host_out.setflags(write=False)
        state = {"acts": acts}

        def input_grad():
            if state["acts"] is None:
                state["acts"], _ = self._forward_layered(x, batch)
            _, dx = self._backward_layered(x, state["acts"], state["dy"].clone(), batch, True)
            return dx.reshape(batch_shape + self.inp.shape)

        def backprop(grad, *, output=None):
            if not (D.is_dev(grad) or isinstance(grad, D.DeviceVector)):
                grad = np.asarray(grad, dtype=np.float32)
            assert tuple(grad.shape) == out_shape
            dy = D.as_dev_f32(grad).reshape(-1)
            state["dy"] = dy
            if self.fused_bwd:
                g = self._fused_grad(x, batch, HEAD_DY, dy=dy)
            else:
                if state["acts"] is None:
                    state["acts"], _ = self._forward_layered(x, batch)
                g, _ = self._backward_layered(x, state["acts"], dy.clone(), batch, False)
            self.inp._grad_dev, self.inp._grad_fn = None, input_grad      # lazy last_gradient
            return _finish(g, root.n_params, output)

        return host_out, backprop

    def _fused_grad(self, x, batch, kind, dy=None, target=None, logp=None):
        lib, net = _lib.lib(), self.net
        head = MqHead()
        head.kind = kind
        head.dy, head.target = D.ptr(dy), D.ptr(target)
        wsb = lib.mq_fused_ws_bytes(ctypes.byref(net), batch)
        ws = D.workspace("fused", wsb)
        grad = D.empty(net.n_params)
        D.call("mq_fused_grad", ctypes.byref(net), D.ptr(self.theta), D.ptr(x), None, batch,
               ctypes.byref(head), 1.0 / batch, D.ptr(grad), None, D.ptr(logp), D.ptr(ws), wsb, D.stream())
        return grad

    def _evaluate_logprob(self, array, *, sample):
        batch_shape, batch = self._batch(array)
        net = self.net
        if not self.fused_bwd:
            return self.root._evaluate_graph(array, sample=sample)
        x = D.from_host(array.reshape(batch, self.n_in))
        if self.discrete:                    # the categorical heads take the action as a one-hot target row
            a = D.from_host(np.asarray(sample, dtype=np.int32).reshape(batch), np.int32)
            s = D.empty(batch * net.n_out)
            D.call("mq_onehot", D.ptr(a), batch, net.n_out, D.ptr(s), D.stream())
        else:
            s = D.from_host(np.asarray(sample, dtype=np.float32).reshape(batch, net.n_out))
        logp = D.empty(batch)
        D.call("mq_fused_forward", ctypes.byref(net), D.ptr(self.theta), D.ptr(x), None, batch, None,
               D.ptr(s), D.ptr(logp), D.stream())
        host = D.to_host(logp).reshape(batch_shape)
        host.setflags(write=False)

        def backprop(grad, *, output=None):
            if not (D.is_dev(grad) or isinstance(grad, D.DeviceVector)):
                grad = np.asarray(grad, dtype=np.float32)
            assert tuple(grad.shape) == batch_shape
            w = D.as_dev_f32(grad).reshape(-1)
            g = self._fused_grad(x, batch, HEAD_DISCRETE_LOGP if self.discrete else HEAD_GAUSS_LOGP, dy=w, target=s)
            return _finish(g, self.root.n_params, output)

        return host, backprop


def _finish(grad_dev, n_params, output):
    if output is not None:
        assert output.shape == (n_params,)
        output[:] = D.to_host(grad_dev)
        return output
    return D.DeviceVector(grad_dev, (n_params,), writeable=True)
