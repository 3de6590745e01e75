"""Evolution strategies, population form of benchmarks/mutation.py:33-37.

Each member p evaluates the deterministic policy theta + noise[p] on a shared observation
batch (one CTA per member, weights built in shared memory: csrc/mq_fused.cu es_eval_kernel),
the returns are normalised as ONE batch by RunningNormalize(horizon 10) and Adam(lr 0.01)
ascends mean_p noise_p * r_p.  Noise is an input (host-drawn for parity).

Multi-GPU: members are sharded across ranks with no communication during evaluation; the
returns are all-gathered (4 B per member) and the weighted noise sum is all-reduced once.
"""
import ctypes

import numpy as np
import torch

from . import _device as D
from . import _lib
from . import dist as mqdist
from ._adam import Adam
from ._lib import ACT_NONE, ACT_TANH
from ._normalize import RunningNormalize
from .ppo import init_mlp


class ESPopulation(object):
    def __init__(self, n_obs, n_act, *, hid=64, n_hid=2, horizon=10, lr=0.01, norm_horizon=10, params=None):
        widths = [hid] * n_hid + [n_act]
        self.net = _lib.make_net(n_obs, widths, [ACT_TANH] * n_hid + [ACT_NONE])
        if params is None:
            params = init_mlp(n_obs, widths, 1.0)              # mutation.py:7-15 (no init= on the head)
        self.opt = Adam(params, horizon=horizon, lr=lr)        # fp64 state like the reference
        self.normalize = RunningNormalize(horizon=norm_horizon)
        self.n_params = self.net.n_params

    def generation(self, noise, obs, ret_coef):
        """noise f32[M_local, P] (this rank's members, already scaled by 0.1), obs f32[T, n_obs],
        ret_coef f32[n_act]: synthetic return R_p = sum_t sum_a act[p,t,a]*ret_coef[a].
        Returns the device tensor of this rank's returns."""
        lib = _lib.lib()
        value, _, _, theta = self.opt.state()
        m_local, t = noise.shape[0], obs.shape[0]
        rets = torch.empty(m_local, dtype=torch.float32, device=noise.device)
        D.call("mq_es_eval", ctypes.byref(self.net), D.ptr(theta), D.ptr(noise), m_local, D.ptr(obs), t,
               D.ptr(ret_coef), D.ptr(rets), D.stream())
        all_rets = mqdist.allgather_concat(rets)                   # [M_global]
        m_global = all_rets.numel()
        self.normalize.update_dev(all_rets, m_global)              # one batch of M returns (a13)
        r = self.normalize.apply_dev(all_rets, m_global, out_dtype=torch.float64)
        lo = mqdist.rank() * m_local
        r_local = r[lo:lo + m_local].contiguous()
        wsb = lib.mq_es_ws_bytes(m_local, self.n_params)
        ws = D.workspace("es", wsb)
        grad = torch.empty(self.n_params, dtype=torch.float64, device=noise.device)
        D.call("mq_es_weighted_sum", D.ptr(noise), D.ptr(r_local), m_local, self.n_params, 1.0 / m_global,
               D.ptr(grad), D.ptr(ws), wsb, D.stream())
        mqdist.allreduce_sum_(grad)
        self.opt.apply_gradient(grad)
        return rets

    def get_params(self):
        return np.asarray(self.opt.get_value())
