"""Fused training loop of mnist/mlp.py:24-43: Input(28,28) -> 2 x LReLU(Affine 128) -> Affine(10),
dy = onehot - softmax(out) - 0.01*out, Adam(horizon 10, lr 1e-3), batch 128 drawn by index.

Two paths.  At the reference's batch size (128; hidden width 128) a block of steps is ONE persistent
launch (csrc/mq_mlp_fused.cu, `fused=True`, the default when the net fits): the 784x128 first layer
(401 KB, more than one SM's shared memory) stays split over 128 CTAs with its Adam moments, the
K-slice partial sums meet through distributed shared memory, two grid-wide barriers per step.
Any other batch runs on the layered path: one GEMM launch per layer with fused bias/activation
epilogues (csrc/mq_gemm.cu, the TMA-fed tcgen05 GEMM for large batches), the softmax head, the
backward GEMMs and the fused Adam step, a whole block of steps captured ONCE in a CUDA graph
(Adam coefficients and minibatch indices are read from device tables).
"""
import ctypes
import os

import numpy as np
import torch

from . import _device as D
from . import _lib
from ._adam import Adam
from ._lib import ACT_LRELU, ACT_NONE, MQ_F32
from .ppo import init_mlp


class MLPTrainer(object):
    def __init__(self, n_in=784, widths=(128, 128, 10), acts=(ACT_LRELU, ACT_LRELU, ACT_NONE), *,
                 leak=0.1, batch=128, horizon=10, lr=1e-3, l2=0.01, params=None, state_dtype="float32",
                 graph_steps=32, fused=True, gemm_mode="fp16x2"):
        # gemm_mode: operand format of the layers that run on the tcgen05 GEMM (large batches).  "fp16x2" meets the same fp32
        # bar as "exact" (three bf16 planes) in half the tensor work and 2/3 of the packed bytes; its range precondition
        # (|operand| < 65504) holds for images, LReLU activations of a normed-column net and their batch-mean gradients.
        # (+ MQ_TC_KEEP_X: a step's forward and backward pass share `ws`, so the gathered minibatch is packed once per step)
        self.tc = _lib.TC_MODES[gemm_mode] | _lib.TC_KEEP_X
        self.net = _lib.make_net(n_in, list(widths), list(acts), leak=leak)
        if params is None:
            params = init_mlp(n_in, list(widths), 0.1)                # mlp.py:24-27 (init=0.1 head)
        assert len(params) == self.net.n_params
        self.batch, self.lr, self.l2, self.n_in, self.n_out = batch, lr, l2, n_in, widths[-1]
        self.theta = D.from_host(np.asarray(params, np.float32))
        self.opt = Adam(params, horizon=horizon, state_dtype=state_dtype)
        lib = _lib.lib()
        nb = ctypes.byref(self.net)
        self.acts = D.empty(lib.mq_mlp_acts_floats(nb, batch))
        self.dy = D.empty(batch * self.n_out)
        self.grad = D.zeros(self.net.n_params)
        self.wsb = lib.mq_mlp_ws_bytes(nb, batch)
        self.ws = torch.empty(self.wsb, dtype=torch.uint8, device=D.device())
        self.graph_steps = graph_steps
        # one persistent launch per block of steps when the net fits (batch 128, hidden width 128)
        self._fused_wsb = int(lib.mq_mlp_train_block_ws_bytes(nb, batch)) if fused else 0
        self.fused = self._fused_wsb > 0
        self._fused_ws = torch.empty(max(self._fused_wsb, 16), dtype=torch.uint8, device=D.device())
        self.idx = torch.zeros(graph_steps * batch, dtype=torch.int32, device=D.device())
        self.coefs = torch.zeros(graph_steps * 4, dtype=torch.float64, device=D.device())
        self._coef_pin = torch.zeros(graph_steps * 4, dtype=torch.float64).pin_memory()      # staging for the coefficient table
        self._coef_ev = torch.cuda.Event()
        self._coef_ev.record()
        self._graph, self._graph_data = None, None
        # parameter-gradient products on a side stream, overlapping the chain of activation gradients
        self._side = torch.cuda.Stream()
        self._events = [torch.cuda.Event() for _ in range(self.net.n_layers + 1)]
        for ev in self._events:
            ev.record()                                                # creates the CUDA event behind the handle
        self._event_ptrs = (ctypes.c_void_p * len(self._events))(*[ev.cuda_event for ev in self._events])

    # one step, all pointers fixed except through the idx / coef tables ---------------------
    def _enqueue_step(self, x, y, s, idx_ptr=None, coef_ptr=None):
        nb, B = ctypes.byref(self.net), self.batch
        idx_ptr = idx_ptr if idx_ptr is not None else D.ptr(self.idx) + s * B * 4
        coef_ptr = coef_ptr if coef_ptr is not None else D.ptr(self.coefs) + s * 32
        value, m, v, _ = self.opt.state()
        D.call("mq_mlp_forward_tc", nb, D.ptr(self.theta), D.ptr(x), idx_ptr, B, D.ptr(self.acts), self.tc, D.ptr(self.ws), self.wsb,
               D.stream())
        logits = self.acts[self.acts.numel() - B * self.n_out:]
        D.call("mq_softmax_ce_grad", D.ptr(logits), D.ptr(y), idx_ptr, B, self.n_out, self.l2,
               D.ptr(self.dy), D.stream())
        D.call("mq_mlp_backward_tc", nb, D.ptr(self.theta), D.ptr(x), idx_ptr, D.ptr(self.acts),
               D.ptr(self.dy), B, 1.0 / B, D.ptr(self.grad), None, self.tc, D.ptr(self.ws), self.wsb, D.stream(),
               None if os.environ.get("MQ_MLP_NO_OVERLAP") else self._side.cuda_stream, self._event_ptrs)
        D.call("mq_adam_step_dev", D.ptr(value), D.ptr(m), D.ptr(v), D.ptr(self.theta), D.ptr(self.grad),
               self.net.n_params, self.opt._code, MQ_F32, coef_ptr, 0, D.stream())

    def train_block(self, x, y, idx_table):
        """x f32[n,784] and y f32[n,10] device tensors (whole data set), idx_table host or device
        int32 [steps, batch] with steps <= graph_steps.  Runs the steps as one graph replay."""
        steps = idx_table.shape[0]
        assert steps <= self.graph_steps and idx_table.shape[1] == self.batch
        idx_d = idx_table if D.is_dev(idx_table) else D.from_host(np.ascontiguousarray(idx_table, np.int32))
        self.idx[:steps * self.batch].copy_(idx_d.reshape(-1), non_blocking=True)
        self._coef_ev.synchronize()                        # the previous block's upload has left the staging buffer
        co = self._coef_pin.numpy().reshape(self.graph_steps, 4)
        for s in range(steps):
            c1, c2 = self.opt.next_coefs()
            co[s] = (c1, c2, self.lr, 1e-8)
        self.coefs.copy_(self._coef_pin, non_blocking=True)
        self._coef_ev.record()
        self.opt._out = None
        if self.fused:
            self._fused_steps(x, y, D.ptr(self.idx), D.ptr(self.coefs), steps)
            return
        key = (D.ptr(x), D.ptr(y), steps)
        if self._graph is None or self._graph_data != key:
            value, m, v, _ = self.opt.state()
            snap = [t.clone() for t in (self.theta, value, m, v)]
            self._enqueue_step(x, y, 0)                    # warm-up outside capture (sets kernel attributes)
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for s in range(steps):
                    self._enqueue_step(x, y, s)
            for t, c in zip((self.theta, value, m, v), snap):   # undo the warm-up step
                t.copy_(c)
            self._graph, self._graph_data = g, key
        self._graph.replay()

    def _fused_steps(self, x, y, idx_ptr, coef_ptr, steps):
        value, m, v, _ = self.opt.state()
        D.call("mq_mlp_train_block", ctypes.byref(self.net), D.ptr(self.theta), D.ptr(value), D.ptr(m), D.ptr(v), self.opt._code,
               D.ptr(x), D.ptr(y), idx_ptr, self.batch, steps, coef_ptr, self.l2, D.ptr(self._fused_ws), self._fused_wsb,
               D.stream())

    def _sgd_setup(self):
        """Staging for sgd_step: double-buffered pinned host buffers (batch, labels, Adam coefficients), fixed device
        buffers, and the step captured once as a CUDA graph on them."""
        B, dev = self.batch, D.device()
        st = self._sgd = dict(slot=0)
        st["pin"] = [(torch.empty(B, self.n_in).pin_memory(), torch.empty(B, self.n_out).pin_memory(),
                      torch.empty(4, dtype=torch.float64).pin_memory()) for _ in range(2)]
        st["np"] = [tuple(t.numpy() for t in trio) for trio in st["pin"]]
        st["ev"] = [torch.cuda.Event() for _ in range(2)]
        st["x"] = torch.empty(B, self.n_in, dtype=torch.float32, device=dev)
        st["y"] = torch.empty(B, self.n_out, dtype=torch.float32, device=dev)
        st["coef"] = torch.zeros(4, dtype=torch.float64, device=dev)
        st["idx"] = torch.arange(B, dtype=torch.int32, device=dev)
        st["graph"] = None
        return st

    def sgd_step(self, inps, lbls):
        """mlp.py:31-35 with host arrays: the batch goes through pinned staging (one async H2D copy each for images,
        labels and the step's Adam coefficients), the step itself is one CUDA-graph replay; nothing is copied back."""
        st = getattr(self, "_sgd", None) or self._sgd_setup()
        B = self.batch
        slot = st["slot"] = st["slot"] ^ 1
        st["ev"][slot].synchronize()                       # the upload that last used this staging slot has finished
        nx, ny, nc = st["np"][slot]
        np.copyto(nx, np.asarray(inps, np.float32).reshape(B, self.n_in))
        np.copyto(ny, np.asarray(lbls, np.float32).reshape(B, self.n_out))
        c1, c2 = self.opt.next_coefs()
        nc[:] = (c1, c2, self.lr, 1e-8)
        px, py, pc = st["pin"][slot]
        st["x"].copy_(px, non_blocking=True)
        st["y"].copy_(py, non_blocking=True)
        st["coef"].copy_(pc, non_blocking=True)
        st["ev"][slot].record()
        self.opt._out = None
        # (one step per call stays on the layered graph: a fused launch loads and stores the optimiser state of the whole net,
        # which pays off over a block of steps, not over one -- 153 against 89 us per host-fed step)
        if st["graph"] is None:
            self._enqueue_step(st["x"], st["y"], 0, D.ptr(st["idx"]), D.ptr(st["coef"]))   # eager once (kernel attributes)
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):                              # capturing runs nothing: this call's step was the eager one
                self._enqueue_step(st["x"], st["y"], 0, D.ptr(st["idx"]), D.ptr(st["coef"]))
            st["graph"] = g
            return
        st["graph"].replay()

    def forward(self, inps):
        x = np.asarray(inps, np.float32).reshape(-1, self.n_in)
        xd = D.from_host(x)
        acts = D.empty(_lib.lib().mq_mlp_acts_floats(ctypes.byref(self.net), len(x)))
        D.call("mq_mlp_forward", ctypes.byref(self.net), D.ptr(self.theta), D.ptr(xd), None, len(x),
               D.ptr(acts), D.stream())
        return D.to_host(acts[acts.numel() - len(x) * self.n_out:]).reshape(len(x), self.n_out)

    def get_params(self):
        return D.to_host(self.theta)
