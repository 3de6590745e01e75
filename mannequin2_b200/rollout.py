"""Rollout side of the PPO loop on the device.

The reference steps ONE gym environment from Python (mannequin/gym.py:115-137): per step a
`NormalizedObservations` update (gym.py:61-84), one batch-1 policy evaluation and one `env.step`.
Here E environments advance in lock step, three kernel launches per time step and no host work:

    NormalizedObservations   gym.py:61-84 with the E observations of a time step as one batch of
                             RunningNormalize (mannequin/_normalize.py:34-81); statistics in device memory
    BatchedActor.collect     one_step of gym.py:115-122 for all environments, T steps; the records are laid
                             out environment-major (segment e = rows e*T .. e*T+T-1), which is the layout the
                             GAE scan wants, and the first observation of the NEXT chunk is carried over like
                             benchmarks/gae.py:64 (`hist = hist[length:]`)
    SyntheticEnv             NOT the reference's environments (gym is out of scope and not shippable): a
                             damped linear system x' = x + dt (-damp x + B a) + sigma noise with reward
                             -|x'|^2/d - ctrl |a|^2/A, fixed-length or randomly ending episodes, observed
                             through a per-dimension affine map so that the normaliser has work to do.

`PPOUpdater.update_rollout` (ppo.py) consumes a collected chunk: per-segment bootstrap
(mq_bootstrap_segments = the chunk-end case of gae.py:52-61), then the unchanged update path.
The noise is a counter-based Philox stream keyed by (seed, step, environment): a rollout is
reproducible and, with `use_graph=True`, the T steps are replayed as ONE CUDA graph.
"""
import ctypes

import numpy as np
import torch

from . import _device as D
from ._lib import MqEnv


class NormalizedObservations(object):
    """Batched, device-resident counterpart of mannequin/gym.py:61-84 (wrapper around an observation stream
    rather than around a gym.Env): `__call__(raw)` = clip(normalize(raw), -clip, clip), one RunningNormalize
    batch update per call."""

    def __init__(self, shape, *, horizon=None, clip=5.0):
        self._shape = tuple(np.zeros(shape).shape)
        self._d = int(np.prod(self._shape)) if self._shape else 1
        assert 1 <= self._d <= 64
        self._horizon = float(horizon) if horizon is not None else 0.0
        self._clip = float(clip)
        self._state = D.zeros(2 * self._d + 3, torch.float64)

    def step_dev(self, raw, rows, out):
        D.call("mq_obsnorm_step", D.ptr(raw), rows, self._d, self._horizon, D.ptr(self._state), self._clip,
               D.ptr(out), D.stream())
        return out

    def __call__(self, raw):
        raw = np.asarray(raw, dtype=np.float32)
        x = D.from_host(raw.reshape(-1, self._d))
        out = torch.empty_like(x)
        self.step_dev(x, x.shape[0], out)
        return D.to_host(out).reshape(raw.shape)

    def _host_state(self):
        return D.to_host(self._state)

    def get_mean(self):
        return self._host_state()[:self._d].reshape(self._shape)

    def get_var(self):
        s = self._host_state()
        if s[2 * self._d + 2] < 2:
            return np.ones(self._shape, dtype=np.float64)
        return s[self._d:2 * self._d].reshape(self._shape)

    def get_std(self):
        return np.sqrt(self.get_var())


class SyntheticEnv(object):
    """E copies of a damped linear system with quadratic cost (see the module docstring)."""

    def __init__(self, n_env, n_obs=11, n_act=3, *, horizon=200, dt=0.1, damp=0.5, sigma=0.05, ctrl=0.1, p_end=0.0,
                 seed=0):
        rs = np.random.RandomState(seed)
        self.n_env, self.n_obs, self.n_act = int(n_env), int(n_obs), int(n_act)
        self.params = dict(dt=float(dt), damp=float(damp), sigma=float(sigma), ctrl=float(ctrl), p_end=float(p_end),
                           horizon=int(horizon),
                           b=(rs.randn(n_obs, n_act) / np.sqrt(n_act)).astype(np.float32),
                           obs_scale=np.exp(rs.randn(n_obs) * 0.5).astype(np.float32),
                           obs_shift=(rs.randn(n_obs) * 2.0).astype(np.float32))
        self._b = D.from_host(self.params["b"])
        self._scale = D.from_host(self.params["obs_scale"])
        self._shift = D.from_host(self.params["obs_shift"])
        self.seed = int(seed)
        self.x = torch.empty(self.n_env, self.n_obs, dtype=torch.float32, device=D.device())
        self.ep_step = torch.empty(self.n_env, dtype=torch.int32, device=D.device())
        self.raw = torch.empty(self.n_env, self.n_obs, dtype=torch.float32, device=D.device())
        self.reset()

    def descriptor(self):
        e = MqEnv()
        e.d, e.a, e.horizon = self.n_obs, self.n_act, self.params["horizon"]
        e.dt, e.damp, e.sigma = self.params["dt"], self.params["damp"], self.params["sigma"]
        e.ctrl, e.p_end = self.params["ctrl"], self.params["p_end"]
        e.b, e.obs_scale, e.obs_shift = D.ptr(self._b), D.ptr(self._scale), D.ptr(self._shift)
        return e

    def reset(self):
        env = self.descriptor()
        D.call("mq_env_reset", ctypes.byref(env), self.n_env, self.seed, D.ptr(self.x), D.ptr(self.ep_step),
               D.ptr(self.raw), D.stream())


class BatchedActor(object):
    """The Gaussian policy `net` (an object with `.net` = MqNet incl. logstd and `.theta` = float32 device
    parameters, e.g. PPOUpdater.policy) acting in all environments of `env` at once."""

    def __init__(self, env, net, *, normalizer=None, seed=1, use_graph=False):
        assert net.net.logstd_off >= 0 and net.net.n_in == env.n_obs and net.net.n_out == env.n_act
        self.env, self.net = env, net
        self.normalizer = normalizer if normalizer is not None else NormalizedObservations(env.n_obs)
        self.seed = int(seed)
        self.use_graph = bool(use_graph)
        dev = D.device()
        self._obs_n = torch.empty(env.n_env, env.n_obs, dtype=torch.float32, device=dev)
        self._mean = torch.empty(env.n_env, env.n_act, dtype=torch.float32, device=dev)
        self._step_base = torch.zeros(1, dtype=torch.int32, device=dev)     # steps taken so far (noise stream index)
        self._have_obs = False
        self._bufs, self._graphs = {}, {}
        self.steps_done = 0

    def _buffers(self, t_len):
        b = self._bufs.get(t_len)
        if b is None:
            e, dev = self.env, D.device()
            n = e.n_env * t_len
            b = dict(obs=torch.zeros(n + 1, e.n_obs, dtype=torch.float32, device=dev),      # + the unused bootstrap row
                     act=torch.empty(n, e.n_act, dtype=torch.float32, device=dev),
                     rew=torch.empty(n, dtype=torch.float32, device=dev),
                     done=torch.empty(n, dtype=torch.uint8, device=dev),
                     last_obs=torch.empty(e.n_env, e.n_obs, dtype=torch.float32, device=dev))
            self._bufs[t_len] = b
        return b

    def _steps(self, t_len, b):
        """T x (policy mean, sample + record + environment, normalise the new observations)."""
        e, net = self.env, self.net
        envd = e.descriptor()
        logstd = net.theta[net.net.logstd_off:net.net.logstd_off + e.n_act]
        for t in range(t_len):
            D.call("mq_fused_forward", ctypes.byref(net.net), D.ptr(net.theta), D.ptr(self._obs_n), None, e.n_env,
                   D.ptr(self._mean), None, None, D.stream())
            D.call("mq_actor_env_step", ctypes.byref(envd), e.n_env, t, t_len, self.seed, t, D.ptr(self._step_base),
                   D.ptr(self._obs_n), D.ptr(self._mean), D.ptr(logstd), D.ptr(e.x), D.ptr(e.ep_step), D.ptr(b["obs"]),
                   D.ptr(b["act"]), D.ptr(b["rew"]), D.ptr(b["done"]), D.ptr(e.raw), D.stream())
            self.normalizer.step_dev(e.raw, e.n_env, self._obs_n)
        b["last_obs"].copy_(self._obs_n)
        self._step_base.add_(t_len)

    def collect(self, t_len):
        """Advance every environment by t_len steps.  Returns device tensors: obs f32[E*T+1, D] (normalised
        observations, environment-major; the last row is padding), act f32[E*T, A], rew f32[E*T], done u8[E*T],
        last_obs f32[E, D] (the observation after the chunk: bootstrap state and first state of the next chunk).
        The tensors are reused by the next call with the same t_len."""
        b = self._buffers(t_len)
        if not self._have_obs:                          # very first observation (gym.py:76-79: reset)
            self.normalizer.step_dev(self.env.raw, self.env.n_env, self._obs_n)
            self._have_obs = True
        if not self.use_graph:
            self._steps(t_len, b)
        else:
            g = self._graphs.get(t_len)
            if g is None:
                self._steps(t_len, b)                   # eager once: attributes set, workspaces allocated
                self.steps_done += t_len
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._steps(t_len, b)
                self._graphs[t_len] = g
                # the capture did not run anything: this call's chunk is the eager one above
                return b
            g.replay()
        self.steps_done += t_len
        return b
