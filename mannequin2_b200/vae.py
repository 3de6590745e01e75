"""The VAE step of mnist/vae.py on the layer graph: DKLUninormal and Clip as device ops, and the training loop.

mnist/vae.py defines its two extra nodes inside the script (vae.py:13-34):

    DKLUninormal(mean, logstd) = 0.5 * (sum(exp(logstd) - logstd + mean^2, axis=-1) - D)     (autograd)
    Clip(v, a, b)              = np.clip forward; backward zeroes g where (v < a and g < 0) or (v > b and g > 0)

Here both are `Function`s whose `f` runs the closed-form kernels of csrc/mq_ops.cu (mq_dkl_*, mq_clip_*), so the
encoder / decoder graph of vae.py:39-55 evaluates node by node on the device; its Affine layers are `matmul` nodes,
i.e. mq_gemm, which sends batch-4096 layers (BASELINE configs[4]) to the tcgen05 GEMM (csrc/mq_gemm_tc.cu).
`VAETrainer.step` is vae.py:57-71: two evaluations of the shared graph, clipped-momentum ascent.
`VAETrainer.step_fused` is the same iteration as ONE plan (csrc/mq_vae.cu, `mq_vae_step`): one forward and one backward
pass, ~30 launches instead of ~110, every product on the TMA-fed tcgen05 GEMM with its operands kept as packed planes.
"""
import ctypes

import numpy as np
import torch

from . import _device as D
from . import _lib
from .backprop import gemm_mode
from .basicnet import Affine, Function, Input, Tanh
from .distrib import Gauss


def DKLUninormal(*, mean, logstd):
    """KL(N(mean, exp(logstd)) || N(0, 1)) as written in mnist/vae.py:13-24 (note exp(logstd), not exp(2 logstd))."""
    def dkl(mean_v, logstd_v):
        mean_v, logstd_v = mean_v.contiguous(), logstd_v.contiguous()
        d = int(mean_v.shape[-1])
        rows = mean_v.numel() // d
        out = torch.empty(tuple(mean_v.shape)[:-1], dtype=torch.float32, device=mean_v.device)
        D.call("mq_dkl_forward", D.ptr(mean_v), D.ptr(logstd_v), rows, d, D.ptr(out), D.stream())

        def bpp(g):
            d_mean, d_ls = torch.empty_like(mean_v), torch.empty_like(mean_v)
            D.call("mq_dkl_backward", D.ptr(mean_v), D.ptr(logstd_v), D.ptr(g.contiguous()), rows, d, D.ptr(d_mean),
                   D.ptr(d_ls), D.stream())
            return (d_mean, d_ls)
        return out, bpp
    dkl._mq_dev = True
    dkl._mq_op = ("dkl_uninormal",)
    return Function(mean, logstd, f=dkl, shape=())


def Clip(inner, a, b):
    """mnist/vae.py:26-34."""
    a, b = float(a), float(b)

    def clip(v):
        v = v.contiguous()
        y = torch.empty_like(v)
        D.call("mq_clip_forward", D.ptr(v), D.ptr(y), v.numel(), a, b, D.stream())

        def bpp(g):
            out = torch.empty_like(v)
            D.call("mq_clip_backward", D.ptr(v), D.ptr(g.contiguous()), D.ptr(out), v.numel(), a, b, D.stream())
            return (out,)
        return y, bpp
    clip._mq_dev = True
    clip._mq_op = ("clip", a, b)
    return Function(inner, f=clip)


class VAETrainer(object):
    """mnist/vae.py:39-71 with the layer sizes as arguments (script: image (28, 28), hidden 256, latent 3;
    BASELINE configs[4]: hidden 400, latent 20, batch 4096)."""

    def __init__(self, image_shape=(28, 28), hidden=256, latent=3, n_hidden=2, lr=1e-3, rng=None, gemm_mode="fp16x2"):
        # gemm_mode: operand format of the batch-4096-sized Affine products on the tcgen05 GEMM.  "fp16x2" meets the same fp32
        # bar as "exact" (three bf16 planes) in half the tensor work; its range precondition (|operand| < 65504) holds for
        # images, tanh activations, clipped log-deviations and their gradients.
        self.gemm_mode = gemm_mode
        enc = Input(*image_shape)
        for _ in range(n_hidden):
            enc = Tanh(Affine(enc, hidden))
        self.encoder = Gauss(mean=Affine(enc, latent, init=0.1), logstd=Clip(Affine(enc, latent), -6.0, 0.0), rng=rng)
        self.dkl = DKLUninormal(mean=self.encoder.mean, logstd=self.encoder.logstd)
        dec = self.encoder.sample
        for _ in range(n_hidden):
            dec = Tanh(Affine(dec, hidden))
        self.decoder = Gauss(mean=Affine(dec, *image_shape, init=0.1), logstd=Clip(Affine(dec, *image_shape), -6.0, 0.0))
        self.lr = float(lr)
        self.momentum = None
        self.n_params = self.decoder.n_params
        self.image_shape = tuple(int(s) for s in image_shape)
        self.hidden, self.latent, self.n_hidden = int(hidden), int(latent), int(n_hidden)
        self._rng = rng
        self._fused = None

    def step(self, inps):
        """One iteration of vae.py:57-71 on a batch of images; returns (mean logprob, mean DKL)."""
        n = len(inps)
        with gemm_mode(self.gemm_mode):
            logprob, backprop = self.decoder.logprob.evaluate(inps, sample=inps)
            grad1 = D.as_dev_f32(backprop(np.ones(n, np.float32)))
            dkl_value, backprop = self.dkl.evaluate(inps)
            grad2 = D.as_dev_f32(backprop(-np.ones(n, np.float32)))
        flat = self.decoder.logprob._flat                              # encoder parameters first (vae.py:66)
        if self.momentum is None:
            self.momentum = D.zeros(self.n_params)
        if flat is not None:
            D.call("mq_momentum_step", D.ptr(flat), D.ptr(self.momentum), D.ptr(grad1), D.ptr(grad2), grad2.numel(),
                   self.n_params, 0.9, 1.0, self.lr, D.stream())
        else:                                                          # parameters not in one arena slice
            theta = D.as_dev_f32(self.decoder.get_params()).clone()
            D.call("mq_momentum_step", D.ptr(theta), D.ptr(self.momentum), D.ptr(grad1), D.ptr(grad2), grad2.numel(),
                   self.n_params, 0.9, 1.0, self.lr, D.stream())
            self.decoder.load_params(D.DeviceVector(theta, (self.n_params,)))
        return float(np.mean(logprob)), float(np.mean(dkl_value))

    # -- device-resident step: no host round trip, capturable ----------------------------------------------------
    def step_dev(self, x_dev):
        """The same iteration with the batch already on the device (float32 [B, *image_shape]) and the sampling noise
        drawn on the device (construct the trainer with rng=DeviceRNG(seed)).  Returns device scalars
        (mean logprob, mean DKL); nothing synchronises with the host."""
        n = x_dev.shape[0]
        if getattr(self, "_ones", None) is None or self._ones.numel() != n:
            self._ones = torch.ones(n, dtype=torch.float32, device=x_dev.device)
            self._minus = -self._ones
        with gemm_mode(self.gemm_mode):
            logprob, backprop = self.decoder.logprob.evaluate_dev(x_dev, sample=x_dev)
            grad1 = D.as_dev_f32(backprop(self._ones))
            dkl_value, backprop = self.dkl.evaluate_dev(x_dev)
            grad2 = D.as_dev_f32(backprop(self._minus))
        flat = self.decoder.logprob._flat
        assert flat is not None, "step_dev needs the parameters in one arena slice"
        if self.momentum is None:
            self.momentum = D.zeros(self.n_params)
        D.call("mq_momentum_step", D.ptr(flat), D.ptr(self.momentum), D.ptr(grad1), D.ptr(grad2), grad2.numel(),
               self.n_params, 0.9, 1.0, self.lr, D.stream())
        return logprob.mean(), dkl_value.mean()

    # -- fused plan: one forward + one backward pass in one C call --------------------------------------------------
    def fused_supported(self):
        """The plan of csrc/mq_vae.cu covers this net: hidden % 8 == 0, at most 4 hidden layers per side, the parameters
        in one arena slice in the order of basicnet.py:66-77, and a device generator for the latent noise."""
        return (self.hidden % 8 == 0 and 1 <= self.n_hidden <= 4 and self.decoder.logprob._flat is not None
                and hasattr(self._rng, "randn_dev"))

    def _fused_plan(self, n):
        f = self._fused
        if f is not None and f["desc"].batch == n:
            return f
        assert self.fused_supported(), "the fused VAE plan does not cover this net (see fused_supported)"
        desc = _lib.MqVae(batch=int(n), d_img=int(np.prod(self.image_shape)), hidden=self.hidden, latent=self.latent,
                          n_hidden=self.n_hidden, tc=_lib.TC_MODES[self.gemm_mode], clip_lo=-6.0, clip_hi=0.0)
        lib = _lib.lib()
        assert int(lib.mq_vae_n_params(ctypes.byref(desc))) == self.n_params, "parameter layout differs from the plan's"
        d_img, h, l = int(np.prod(self.image_shape)), self.hidden, self.latent
        want = []
        for side_in, head_out, head_b in ((d_img, l, (l,)), (l, d_img, self.image_shape)):
            for i in range(self.n_hidden):
                want += [(side_in if i == 0 else h, h), (h,)]
            want += [(h, head_out), head_b, (h, head_out), head_b]
        assert [tuple(p.shape) for p in self.decoder.logprob._params] == want, "parameter order differs from the plan's"
        wsb = int(lib.mq_vae_ws_bytes(ctypes.byref(desc)))
        ws = torch.empty(wsb, dtype=torch.uint8, device=D.device())
        D.call("mq_vae_ws_init", ctypes.byref(desc), D.ptr(ws), wsb, D.stream())
        f = dict(desc=desc, ws=ws, wsb=wsb, grad=D.zeros(self.n_params), logprob=D.zeros(n), dkl=D.zeros(n))
        self._fused = f
        return f

    def grad_fused(self, x_dev, unit):
        """grad1 with grad2 added onto the encoder slice (vae.py:60-66), log-probs and DKLs of the batch, from the fused
        plan; `unit` is the N(0,1) draw of the latent sample, [batch, latent] on the device."""
        n = int(x_dev.shape[0])
        f = self._fused_plan(n)
        x = x_dev.to(torch.float32).contiguous()
        D.call("mq_vae_grad", ctypes.byref(f["desc"]), D.ptr(self.decoder.logprob._flat), D.ptr(x), D.ptr(unit.contiguous()),
               D.ptr(f["grad"]), D.ptr(f["logprob"]), D.ptr(f["dkl"]), D.ptr(f["ws"]), f["wsb"], D.stream())
        return f["grad"], f["logprob"], f["dkl"]

    def step_fused(self, x_dev):
        """One iteration of vae.py:57-71 as ONE C call (`mq_vae_step`); returns device scalars (mean logprob, mean DKL).
        The latent noise comes from the trainer's device generator, exactly as in `step_dev`."""
        n = int(x_dev.shape[0])
        f = self._fused_plan(n)
        if self.momentum is None:
            self.momentum = D.zeros(self.n_params)
        x = x_dev.to(torch.float32).contiguous()
        unit = self._rng.randn_dev((n, self.latent))
        D.call("mq_vae_step", ctypes.byref(f["desc"]), D.ptr(self.decoder.logprob._flat), D.ptr(self.momentum), D.ptr(x),
               D.ptr(unit), D.ptr(f["grad"]), D.ptr(f["logprob"]), D.ptr(f["dkl"]), 0.9, 1.0, self.lr, D.ptr(f["ws"]), f["wsb"],
               D.stream())
        return f["logprob"].mean(), f["dkl"].mean()

    def step_graph(self, x_dev, fused=None):
        """One step replayed as ONE CUDA graph: the fused plan (`step_fused`, default when it covers the net) or the
        node-by-node `step_dev` (~110 kernel launches).  The first call runs eagerly and captures; later calls copy the batch
        into the captured input buffer and replay.  Returns the device scalars of the replayed step."""
        fused = self.fused_supported() if fused is None else bool(fused)
        step = self.step_fused if fused else self.step_dev
        g = getattr(self, "_graph", None)
        if g is None or self._graph_in.shape != x_dev.shape or self._graph_fused != fused:
            if self.momentum is None:
                self.momentum = D.zeros(self.n_params)
            self._graph_in = x_dev.clone()
            self._graph_fused = fused
            out = step(self._graph_in)                     # eager: sets kernel attributes, sizes every workspace
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._graph_out = step(self._graph_in)
            self._graph = g
            return out
        self._graph_in.copy_(x_dev, non_blocking=True)
        g.replay()
        return self._graph_out
