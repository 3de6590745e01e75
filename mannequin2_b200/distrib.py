"""Distributions with the API of mannequin/distrib.py.

Gauss (distrib.py:41-76): diagonal Gaussian with a learnt state-independent
logstd `Params(A)` (or a logstd sub-network).  `logprob` uses the closed-form
derivative of distrib.py:61-68 in device kernels instead of the `autograd`
package; `sample` is reparameterised with host-drawn N(0,1) noise (the reference
draws from an OS-seeded RandomState, distrib.py:46, so the draw itself is an
input of the kernel).

Discrete (distrib.py:8-39): softmax categorical over a logits layer.
"""
import numpy as np
import torch

from . import _device as D
from ._utils import endswith
from .basicnet import Function, Params


def _period(mean_v, logstd_v):
    if tuple(mean_v.shape) == tuple(logstd_v.shape):
        return mean_v.numel()
    if endswith(tuple(mean_v.shape), tuple(logstd_v.shape)):
        return logstd_v.numel()
    raise ValueError("Invalid shapes: %s, %s" % (tuple(mean_v.shape), tuple(logstd_v.shape)))


def _reduce_to(g, shape):
    cols = int(np.prod(shape))
    out = torch.empty(tuple(shape), dtype=torch.float32, device=g.device)
    D.call("mq_col_sum", D.ptr(g.contiguous()), None, g.numel() // cols, cols, 1.0, D.ptr(out), D.stream())
    return out


class Gauss(object):
    def __init__(self, *, mean, logstd=None, rng=None):
        """`rng` (extension): the RandomState `sample` draws from; the reference always uses an OS-seeded one."""
        logstd = logstd or Params(*mean.shape)
        assert logstd.shape == mean.shape
        rng = rng or np.random.RandomState()
        n_act = int(np.prod(mean.shape))

        def sample(mean_v, logstd_v):
            mean_v, logstd_v = mean_v.contiguous(), logstd_v.contiguous()
            period = _period(mean_v, logstd_v)
            if hasattr(rng, "randn_dev"):            # device generator (no host round trip; capturable)
                unit = rng.randn_dev(tuple(mean_v.shape))
            else:
                unit = D.from_host(rng.randn(*mean_v.shape), np.float32)
            out, noise = torch.empty_like(mean_v), torch.empty_like(mean_v)
            D.call("mq_gauss_sample", D.ptr(mean_v), D.ptr(logstd_v), period, D.ptr(unit),
                   mean_v.numel() // n_act, n_act, D.ptr(out), D.ptr(noise), D.stream())
            same = tuple(mean_v.shape) == tuple(logstd_v.shape)

            def bpp(g):
                gn = torch.empty_like(g)
                D.call("mq_binary", D.ptr(g.contiguous()), D.ptr(noise), D.ptr(gn), g.numel(), g.numel(),
                       1, D.stream())
                return (g, gn if same else _reduce_to(gn, logstd_v.shape))
            return out, bpp
        sample._mq_dev = True
        sample._mq_op = ("gauss_sample",)

        def logprob(mean_v, logstd_v, *, sample):
            mean_v, logstd_v = mean_v.contiguous(), logstd_v.contiguous()
            period = _period(mean_v, logstd_v)
            rows = mean_v.numel() // n_act
            batch_shape = tuple(mean_v.shape)[:mean_v.dim() - len(mean.shape)]
            if D.is_dev(sample):
                s = sample.to(torch.float32).contiguous().reshape(-1)
            else:
                s = D.from_host(np.asarray(sample, dtype=np.float32).reshape(-1))
            assert s.numel() == mean_v.numel()
            logp = torch.empty(batch_shape, dtype=torch.float32, device=mean_v.device)
            D.call("mq_gauss_logprob", D.ptr(mean_v), D.ptr(logstd_v), period, D.ptr(s), rows, n_act,
                   D.ptr(logp), D.stream())
            same = tuple(mean_v.shape) == tuple(logstd_v.shape)

            def bpp(g):
                d_mu, d_ls = torch.empty_like(mean_v), torch.empty_like(mean_v)
                D.call("mq_gauss_logprob_bwd", D.ptr(mean_v), D.ptr(logstd_v), period, D.ptr(s),
                       D.ptr(g.contiguous()), rows, n_act, D.ptr(d_mu), D.ptr(d_ls), D.stream())
                return (d_mu, d_ls if same else _reduce_to(d_ls, logstd_v.shape))
            return logp, bpp
        logprob._mq_dev = True
        logprob._mq_op = ("gauss_logprob",)

        def entropy(logstd_v):
            # extension: the reference's Gauss has no entropy; H = sum(logstd) + A/2 (1 + ln 2 pi)
            logstd_v = logstd_v.contiguous()
            rows = logstd_v.numel() // n_act
            batch_shape = tuple(logstd_v.shape)[:logstd_v.dim() - len(mean.shape)]
            h = torch.empty(batch_shape, dtype=torch.float32, device=logstd_v.device)
            D.call("mq_gauss_entropy", D.ptr(logstd_v), None, rows, n_act, D.ptr(h), None, D.stream())

            def bpp(g):
                d_ls = torch.empty_like(logstd_v)
                D.call("mq_gauss_entropy", None, D.ptr(g.contiguous()), rows, n_act, None, D.ptr(d_ls), D.stream())
                return (d_ls,)
            return h, bpp
        entropy._mq_dev = True
        entropy._mq_op = ("gauss_entropy",)

        self.mean = mean
        self.logstd = logstd
        self.entropy = Function(logstd, f=entropy, shape=())
        self.sample = Function(mean, logstd, f=sample, shape=mean.shape)
        self.logprob = Function(mean, logstd, f=logprob, shape=())
        self.n_params = self.logprob.n_params
        self.get_params = self.logprob.get_params
        self.load_params = self.logprob.load_params


class Discrete(object):
    """Softmax categorical head (mannequin/distrib.py:8-39).

    `logprob(logits; sample)` = log softmax(logits)[sample] with the reference's backward rule
    g * (onehot(sample) - softmax(logits)) (distrib.py:24-32) in a device kernel; over an MLP chain the
    plan compiler turns it into the fused categorical head.  `sample(inps)` draws on the host from an
    OS-seeded RandomState exactly like the reference (distrib.py:21-22): one action per call."""

    def __init__(self, *, logits):
        assert len(logits.shape) == 1
        n = logits.shape[0]
        rng = np.random.RandomState()

        def softmax(v):
            v = np.asarray(v, dtype=np.float64).T
            v = np.exp(v - np.amax(v, axis=0))
            v /= np.sum(v, axis=0)
            return v.T

        def sample(inps):
            return rng.choice(n, p=softmax(logits(inps)))

        def logprob(logits_v, *, sample):
            s_host = np.asarray(sample, dtype=np.int32).reshape(-1)
            rows = len(s_host)
            lg = logits_v.contiguous().reshape(rows, n)
            s = D.from_host(s_host, np.int32)
            logp = torch.empty(rows, dtype=torch.float32, device=lg.device)
            D.call("mq_discrete_logprob", D.ptr(lg), D.ptr(s), None, rows, n, D.ptr(logp), None, D.stream())

            def bpp(g):
                d = torch.empty_like(lg)
                D.call("mq_discrete_logprob", D.ptr(lg), D.ptr(s), D.ptr(g.contiguous().reshape(-1)), rows, n,
                       None, D.ptr(d), D.stream())
                return (d.reshape(logits_v.shape),)
            return logp.reshape(tuple(logits_v.shape)[:-1]), bpp
        logprob._mq_dev = True
        logprob._mq_op = ("discrete_logprob",)

        self.logits = logits
        self.sample = sample
        self.logprob = Function(logits, f=logprob, shape=())
        self.n_params = self.logprob.n_params
        self.get_params = self.logprob.get_params
        self.load_params = self.logprob.load_params
