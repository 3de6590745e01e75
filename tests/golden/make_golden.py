#!/usr/bin/env python3
"""Generate golden fixtures by running the REAL reference (build container only).

    python tests/golden/make_golden.py         # needs /root/reference

The reference (squirrelinhell/mannequin2) is pure Python, so it is imported
unmodified from /root/reference.  Three third-party modules it needs are not
installed here; they are replaced by minimal stand-ins that contain no
hot-path arithmetic of their own:

  * ``autograd``  -- ``autograd.grad`` is replaced by a complex-step derivative
    (exact to fp64 rounding for the analytic expressions in
    mannequin/distrib.py:61-68), ``autograd.numpy`` by numpy.  The forward
    expression that runs is the reference's own code.
  * ``gym``       -- only ``gym.Wrapper`` / ``gym.spaces.Box`` are needed by
    mannequin/gym.py and benchmarks/policy.py; a synthetic environment drives
    the rollout.
  * ``_env``      -- benchmarks/_env.py (env factory + progress) is replaced
    so that benchmarks/ppo.py:run() executes exactly two chunks.

Unseeded ``np.random.RandomState()`` instances (distrib.py:13,46, gae.py:28)
are made reproducible by patching the constructor; every random index table
the reference draws is recorded so the oracle / CUDA path can replay it.

Outputs: tests/golden/*.npz (committed).  Nothing here runs on the GPU box.
"""

import os
import sys
import types

import numpy as np

REF = os.environ.get("MQ_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


# ------------------------------------------------------------------ stand-ins
def install_fake_autograd():
    ag = types.ModuleType("autograd")

    def grad(fun):
        def dfun(a, ka, g):
            res = []
            for i in range(len(a)):
                base = [np.asarray(v, dtype=np.complex128) for v in a]
                d = np.zeros(base[i].shape, dtype=np.float64)
                for idx in np.ndindex(*base[i].shape):
                    pert = [v.copy() for v in base]
                    pert[i][idx] += 1e-30j
                    d[idx] = np.imag(fun(tuple(pert), ka, g)) / 1e-30
                res.append(d)
            return tuple(res)
        return dfun

    ag.grad = grad
    ag.numpy = np
    sys.modules["autograd"] = ag
    sys.modules["autograd.numpy"] = np


def install_fake_gym():
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")

    class Box:
        def __init__(self, low, high, shape=None):
            shape = shape if shape is not None else np.shape(low)
            self.low = np.full(shape, low, dtype=np.float32)
            self.high = np.full(shape, high, dtype=np.float32)
            self.shape = tuple(shape)

    class Discrete:
        def __init__(self, n):
            self.n = n

    class Wrapper:
        def __init__(self, env):
            self.env = env
            self.observation_space = env.observation_space
            self.action_space = env.action_space

        def step(self, action):
            return self._step(action)

        def reset(self):
            return self._reset()

        def _step(self, action):
            return self.env.step(action)

        def _reset(self):
            return self.env.reset()

    spaces.Box, spaces.Discrete = Box, Discrete
    gym.spaces, gym.Wrapper = spaces, Wrapper
    sys.modules["gym"] = gym
    sys.modules["gym.spaces"] = spaces
    return gym


class SyntheticEnv:
    """Hopper-shaped (obs 11, act 3) random walk; episodes ~200 steps."""

    def __init__(self, gym, seed):
        self.rs = np.random.RandomState(seed)
        self.observation_space = gym.spaces.Box(-np.inf, np.inf, (11,))
        self.action_space = gym.spaces.Box(-1.0, 1.0, (3,))

    def _step(self, action):
        return self.step(action)

    def _reset(self):
        return self.reset()

    def reset(self):
        self.state = self.rs.randn(11)
        return self.state.copy()

    def step(self, action):
        action = np.asarray(action, dtype=np.float64).reshape(-1)
        self.state = 0.9 * self.state + 0.5 * self.rs.randn(11)
        self.state[:3] += 0.1 * action
        rew = float(self.rs.randn() + 0.3 * np.sum(action) - 0.05 * np.sum(action ** 2))
        done = bool(self.rs.rand() < 1.0 / 200.0)
        return self.state.copy(), rew, done, {}


# ------------------------------------------------------------------- fixtures
def gen_nets(out):
    import mannequin.basicnet as bn
    rs = np.random.RandomState(11)

    # LReLU MLP (mnist/mlp.py structure at reduced size) with a (4,5) image input
    np.random.seed(3)
    m = bn.Input(4, 5)
    for _ in range(2):
        m = bn.LReLU(bn.Affine(m, 16))
    m = bn.Affine(m, 5, init=0.1)
    theta = m.get_params() + rs.randn(m.n_params).astype(np.float32) * 0.05
    m.load_params(theta)
    x = rs.rand(7, 4, 5).astype(np.float32)
    dy = rs.randn(7, 5).astype(np.float32)
    y, bpp = m.evaluate(x)
    out["lrelu_theta"], out["lrelu_x"], out["lrelu_dy"] = m.get_params(), x, dy
    out["lrelu_y"], out["lrelu_grad"] = np.asarray(y), bpp(dy)
    out["lrelu_dx"] = np.asarray(m.prerequisites[0].last_gradient)

    # Tanh policy trunk 11-64-64-3 (benchmarks/mutation.py:7-15), batch 64
    np.random.seed(4)
    p = bn.Input(11)
    for _ in range(2):
        p = bn.Tanh(bn.Affine(p, 64))
    p = bn.Affine(p, 3)
    theta = p.get_params()
    theta[-3:] = rs.randn(3) * 0.1
    p.load_params(theta)
    x = np.clip(rs.randn(64, 11), -5, 5).astype(np.float32)
    dy = rs.randn(64, 3).astype(np.float32)
    y, bpp = p.evaluate(x)
    out["tanh_theta"], out["tanh_x"], out["tanh_dy"] = p.get_params(), x, dy
    out["tanh_y"], out["tanh_grad"] = np.asarray(y), bpp(dy)
    # unbatched call (basicnet.py:242-249: no batch -> no division)
    y1, bpp1 = p.evaluate(x[0])
    out["tanh_y1"], out["tanh_grad1"] = np.asarray(y1), bpp1(dy[0])


def gen_adam(out):
    from mannequin import Adam, Adams
    rs = np.random.RandomState(5)
    for name, cls in (("adam", Adam), ("adams", Adams)):
        opt = cls(rs.randn(37), horizon=10, lr=0.003)
        grads = rs.randn(25, 37).astype(np.float32)
        grads[3:6, 4] = 0.0
        lrs = [None] * 20 + [0.01, 0.02, 0.001, 0.5, 0.003]
        vals = [opt.get_value().copy()]
        for g, lr in zip(grads, lrs):
            if lr is None:
                opt.apply_gradient(g)
            else:
                opt.apply_gradient(g, lr=lr, epsilon=1e-6)
            vals.append(opt.get_value().copy())
        out[name + "_grads"] = grads
        out[name + "_lrs"] = np.array([np.nan if l is None else l for l in lrs])
        out[name + "_values"] = np.array(vals)
    # the two sequences of tests/adam.py at full precision
    opt = Adam([10.0, -5.0], lr=2.0, horizon=3)
    seq = []
    for _ in range(10):
        opt.apply_gradient(-opt.get_value())
        seq.append(opt.get_value().copy())
    out["adam_kat1"] = np.array(seq)
    opt = Adam([0.0, 0.0, 0.0], horizon=1)
    seq = []
    for i in range(10):
        opt.apply_gradient([0.1, 1.0, 10.0] - opt.get_value(), lr=7.0 / (i + 1))
        seq.append(opt.get_value().copy())
    out["adam_kat2"] = np.array(seq)


def gen_norm(out):
    from mannequin import RunningMean, RunningNormalize, discounted
    rs = np.random.RandomState(6)
    # scalar stream in batches of varying size, horizon 2 (benchmarks/ppo.py:19,24)
    n = RunningNormalize(horizon=2)
    sizes = [2048, 1, 5, 300]
    data = [rs.randn(s) * 3.0 + 1.5 for s in sizes]
    res = []
    for d in data:
        res.append(np.concatenate([n(d), [n.get_mean(), n.get_var()]]))
    out["norm_scalar_in"] = np.concatenate(data)
    out["norm_scalar_sizes"] = np.array(sizes)
    out["norm_scalar_out"] = np.concatenate(res)
    # first batch of a single sample, then more (zeros until 2 samples)
    n = RunningNormalize(horizon=10)
    seq = rs.randn(6) * 2.0
    out["norm_single_in"] = seq
    out["norm_single_out"] = np.array([n(v) for v in seq])
    # vector stream [11], no horizon (mannequin/gym.py:66-78)
    n = RunningNormalize((11,))
    obs = rs.randn(9, 11) * np.arange(1, 12)
    o = [n(v) for v in obs]
    out["norm_vec_in"], out["norm_vec_out"] = obs, np.array(o)
    out["norm_vec_batch_apply"] = n.apply(obs)
    out["norm_vec_mean"], out["norm_vec_var"] = n.get_mean(), n.get_var()
    # running mean with weights and horizon
    m = RunningMean((7,), horizon=3)
    xs, ws, ms = rs.randn(8, 7), rs.rand(8) * 2.0, []
    for x, w in zip(xs, ws):
        ms.append(m(x, weight=w))
    out["rmean_x"], out["rmean_w"], out["rmean_out"] = xs, ws, np.array(ms)
    r = rs.randn(700)
    out["disc_in"] = r
    out["disc_h500"] = discounted(r, horizon=500)
    out["disc_h3"] = discounted(r, horizon=3)


def gen_trajectory(out):
    from mannequin import Trajectory
    rs = np.random.RandomState(7)
    o = rs.randn(40, 11)
    a = rs.randn(40, 3)
    r = rs.randn(40)
    t = Trajectory(o, a, r)
    idx = rs.randint(40, size=64)
    b = t[idx]
    out["traj_o"], out["traj_a"], out["traj_r"], out["traj_idx"] = o, a, r, idx
    out["traj_take_o"], out["traj_take_a"], out["traj_take_r"] = b.o, b.a, b.r
    s = t[5:33:3]
    out["traj_slice_o"], out["traj_slice_r"] = s.o, s.r
    d = t.discounted(horizon=20)
    out["traj_disc_r"] = d.r
    ti = Trajectory([3, 1, 2], [[1], [0], [1]])
    out["traj_int_o"], out["traj_int_a"], out["traj_int_r"] = ti.o, ti.a, ti.r
    j = Trajectory.joined(t[:3], t[10:12])
    out["traj_join_r"] = j.r


def gen_gauss(out):
    import mannequin.basicnet as bn
    from mannequin.distrib import Gauss
    rs = np.random.RandomState(8)
    np.random.seed(9)
    p = bn.Input(11)
    for _ in range(2):
        p = bn.Tanh(bn.Affine(p, 64))
    p = bn.Affine(p, 3, init=0.1)
    pol = Gauss(mean=p)
    theta = pol.get_params()
    theta[-3:] = [-0.3, 0.1, 0.4]
    pol.load_params(theta)
    obs = np.clip(rs.randn(64, 11), -5, 5).astype(np.float32)
    act = rs.randn(64, 3).astype(np.float32)
    g = rs.randn(64).astype(np.float32)
    logp, bpp = pol.logprob.evaluate(obs, sample=act)
    out["gauss_theta"], out["gauss_obs"], out["gauss_act"], out["gauss_g"] = \
        pol.get_params(), obs, act, g
    out["gauss_logp"], out["gauss_grad"] = np.asarray(logp), bpp(g)
    # sample path: mean + randn*exp(logstd) with backward (distrib.py:50-59)
    smp, bpp = pol.sample.evaluate(obs[:5])
    mu = pol.mean(obs[:5])
    out["gauss_sample_unit"] = (np.asarray(smp) - mu) / np.exp(theta[-3:])
    out["gauss_sample_out"] = np.asarray(smp)
    gs = rs.randn(5, 3).astype(np.float32)
    out["gauss_sample_g"], out["gauss_sample_grad"] = gs, bpp(gs)


def gen_discrete(out):
    """mannequin.distrib.Discrete over the cartpole-shaped policy of benchmarks/policy.py:14-25 (no autograd needed:
    the reference's own forward and backward rule, distrib.py:24-32)."""
    import mannequin.basicnet as bn
    from mannequin.distrib import Discrete
    rs = np.random.RandomState(18)
    np.random.seed(19)
    p = bn.Input(4)
    for _ in range(2):
        p = bn.Tanh(bn.Affine(p, 64))
    p = bn.Affine(p, 3, init=0.1)
    pol = Discrete(logits=p)
    obs = rs.randn(64, 4).astype(np.float32)
    act = rs.randint(3, size=64).astype(np.int32)
    g = rs.randn(64).astype(np.float32)
    logp, bpp = pol.logprob.evaluate(obs, sample=act)
    out["disc_theta"], out["disc_obs"], out["disc_act"], out["disc_g"] = pol.get_params(), obs, act, g
    out["disc_logits"] = np.asarray(pol.logits(obs))
    out["disc_logp"], out["disc_grad"] = np.asarray(logp), bpp(g)


def gen_obsnorm(out, gym):
    """The REAL mannequin.gym.NormalizedObservations (gym.py:61-84) around a recording environment: the raw observation
    stream (scaled and shifted so that the statistics matter) and what the wrapper returns, 40 steps incl. a reset; and
    RunningNormalize fed whole batches (the lock-step form: one update per time step of E environments)."""
    import mannequin
    import mannequin.gym as mgym

    class Recorder(SyntheticEnv):
        def __init__(self, gym, seed):
            super().__init__(gym, seed)
            self.raw = []
            self.scale = np.exp(np.random.RandomState(5).randn(11))
            self.shift = 3.0 * np.random.RandomState(6).randn(11)

        def reset(self):
            o = super().reset() * self.scale + self.shift
            self.raw.append(o.copy())
            return o

        def step(self, action):
            o, r, d, i = super().step(action)
            o = o * self.scale + self.shift
            self.raw.append(o.copy())
            return o, r, d, i

    inner = Recorder(gym, seed=31)
    env = mgym.NormalizedObservations(inner)
    outs = [env.reset()]
    rs = np.random.RandomState(32)
    for t in range(39):
        if t == 17:
            outs.append(env.reset())
        else:
            outs.append(env.step(rs.randn(3))[0])
    out["on_raw"], out["on_out"] = np.array(inner.raw), np.array(outs)
    out["on_mean"], out["on_var"] = env.get_mean(), env.get_var()
    norm = mannequin.RunningNormalize((11,))
    batches = (rs.randn(6, 7, 11) * inner.scale + inner.shift).astype(np.float32)
    out["on_batches"] = batches
    out["on_batch_out"] = np.array([np.clip(norm(b), -5.0, 5.0) for b in batches])
    out["on_batch_mean"], out["on_batch_var"] = norm.get_mean(), norm.get_var()


def gen_vae(out):
    """mnist/vae.py:13-34 itself: DKLUninormal and Clip are module-level constructors there, so the module is imported
    (matplotlib replaced by an empty stand-in; the autograd stand-in differentiates the reference's own dkl expression)
    and the two are evaluated inside a small encoder graph: forward values and the flat parameter gradient."""
    import importlib.util
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    spec = importlib.util.spec_from_file_location("ref_vae", os.path.join(REF, "mnist", "vae.py"))
    vae = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vae)
    import mannequin.basicnet as bn
    rs = np.random.RandomState(41)
    np.random.seed(42)
    x = bn.Input(5)
    h = bn.Tanh(bn.Affine(x, 8))
    mean = bn.Affine(h, 3, init=0.1)
    logstd = vae.Clip(bn.Affine(h, 3), -6.0, 0.0)
    dkl = vae.DKLUninormal(mean=mean, logstd=logstd)
    theta = dkl.get_params() + rs.randn(dkl.n_params).astype(np.float32) * 0.3
    dkl.load_params(theta)
    inp = rs.randn(7, 5).astype(np.float32)
    val, bpp = dkl.evaluate(inp)
    g = rs.randn(7).astype(np.float32)
    out["vae_theta"], out["vae_x"], out["vae_g"] = theta, inp, g
    out["vae_dkl"], out["vae_grad"] = np.asarray(val), bpp(g)
    out["vae_mean"], out["vae_logstd"] = np.asarray(mean(inp)), np.asarray(logstd(inp))
    # Clip on its own: values on both sides of the range, gradients of both signs
    c_in = bn.Input(6)
    clip = vae.Clip(c_in, -2.5, 2.5)
    v = (rs.randn(4, 6) * 3).astype(np.float32)
    cg = rs.randn(4, 6).astype(np.float32)
    y, cb = clip.evaluate(v)
    cb(cg)
    out["clip_v"], out["clip_g"], out["clip_y"], out["clip_dx"] = v, cg, np.asarray(y), np.asarray(c_in.last_gradient)


def gen_vae_step(out):
    """The graph of mnist/vae.py:39-55 built from the reference's own classes (small sizes, two hidden layers per side
    as in the script) and ONE iteration of its loop body, vae.py:60-69, statement by statement: both evaluations, both
    backward passes, the clipped-momentum update.  The OS-seeded RandomState inside Gauss (distrib.py:46) is replaced by
    a seeded one while the encoder is constructed, so that the N(0,1) draw of the latent sample is a recorded input."""
    import importlib.util
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    spec = importlib.util.spec_from_file_location("ref_vae", os.path.join(REF, "mnist", "vae.py"))
    vae = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vae)
    from mannequin.basicnet import Input, Affine, Tanh
    from mannequin.distrib import Gauss
    shape, hidden, latent, batch = (4, 3), 8, 3, 6
    np.random.seed(142)
    real_rs = np.random.RandomState
    encoder = Input(*shape)
    encoder = Tanh(Affine(encoder, hidden))
    encoder = Tanh(Affine(encoder, hidden))
    np.random.RandomState = lambda *a: real_rs(977)
    try:
        encoder = Gauss(mean=Affine(encoder, latent, init=0.1), logstd=vae.Clip(Affine(encoder, latent), -6.0, 0.0))
    finally:
        np.random.RandomState = real_rs
    dkl = vae.DKLUninormal(mean=encoder.mean, logstd=encoder.logstd)
    decoder = encoder.sample
    decoder = Tanh(Affine(decoder, hidden))
    decoder = Tanh(Affine(decoder, hidden))
    decoder = Gauss(mean=Affine(decoder, *shape, init=0.1), logstd=vae.Clip(Affine(decoder, *shape), -6.0, 0.0))
    rs = real_rs(143)
    # perturbed parameters: log-deviations on both sides of Clip's range, so that its backward rule is exercised
    theta = decoder.get_params() + rs.randn(decoder.n_params).astype(np.float32) * 0.6
    decoder.load_params(theta)
    theta = decoder.get_params().copy()
    inps = rs.rand(batch, *shape).astype(np.float32)
    momentum = (rs.randn(decoder.n_params) * 0.05).astype(np.float32)

    logprob, backprop = decoder.logprob.evaluate(inps, sample=inps)                 # vae.py:60
    grad1 = backprop(np.ones(batch))
    dkl_value, backprop = dkl.evaluate(inps)                                          # vae.py:63
    grad2 = backprop(-np.ones(batch))
    out["vstep_grad1"], out["vstep_grad2"] = np.array(grad1), np.array(grad2)
    grad1 = np.array(grad1)
    grad1[:len(grad2)] += grad2                                                       # vae.py:66
    mom = momentum * 0.9 + grad1 * 0.1
    mom = np.clip(mom, -1.0, 1.0)
    decoder.load_params(decoder.get_params() + 0.001 * mom)                           # vae.py:69
    out["vstep_dims"] = np.array([shape[0] * shape[1], hidden, latent, 2, batch])
    out["vstep_theta"], out["vstep_x"], out["vstep_momentum"] = theta, inps, momentum
    out["vstep_unit"] = real_rs(977).randn(batch, latent)                              # the draw of distrib.py:49
    out["vstep_logprob"], out["vstep_dkl"] = np.array(logprob), np.array(dkl_value)
    out["vstep_grad"], out["vstep_momentum_new"] = grad1, mom
    out["vstep_theta_new"] = decoder.get_params().copy()
    out["vstep_logstd_e"], out["vstep_logstd_d"] = np.array(encoder.logstd(inps)), np.array(decoder.logstd(inps))


def gen_ppo(out, gym):
    """Run benchmarks/ppo.py:run() itself for two chunks and record everything."""
    import mannequin
    import mannequin.gym as mgym

    bench_dir = os.path.join(REF, "benchmarks")
    sys.path.insert(0, bench_dir)

    env_mod = types.ModuleType("_env")
    calls = {"n": 0}

    def get_progress():
        calls["n"] += 1
        return 0.0 if calls["n"] <= 2 else 1.0

    env_mod.build_env = lambda: SyntheticEnv(gym, seed=21)
    env_mod.get_progress = get_progress
    sys.modules["_env"] = env_mod

    # record every Adam the run constructs (policy first, value predictor second)
    adams = []
    RealAdam = mannequin.Adam

    class RecordingAdam(RealAdam):
        def __init__(self, value, **kw):
            super().__init__(value, **kw)
            self.init_value = np.array(value, dtype=np.float64)
            self.trace = []
            inner = self.apply_gradient

            def apply(grad, **lk):
                inner(grad, **lk)
                self.trace.append(self.get_value().astype(np.float32))
            self.apply_gradient = apply
            adams.append(self)

    mannequin.Adam = RecordingAdam

    # seeded stand-ins for RandomState() and recording of index draws
    RealRS = np.random.RandomState
    rs_seed = {"n": 100}
    idx_log = {"value": [], "policy": []}

    class SeededRS(RealRS):
        def __new__(cls, *a, **k):
            return RealRS.__new__(cls)

        def __init__(self, seed=None):
            if seed is None:
                rs_seed["n"] += 1
                seed = rs_seed["n"]
            super().__init__(seed)

        def randint(self, *a, **k):
            r = super().randint(*a, **k)
            if k.get("size") == 64:
                idx_log["value"].append(np.array(r))
            return r

    np.random.RandomState = SeededRS
    real_randint = np.random.randint

    def rec_randint(*a, **k):
        r = real_randint(*a, **k)
        if k.get("size") == 64:
            idx_log["policy"].append(np.array(r))
        return r

    np.random.randint = rec_randint

    hist = []
    real_one_step = mgym.one_step

    def rec_one_step(env, policy):
        s = real_one_step(env, policy)
        hist.append(s)
        return s

    mgym.one_step = rec_one_step

    try:
        np.random.seed(31)
        import ppo
        ppo.run()
    finally:
        np.random.RandomState = RealRS
        np.random.randint = real_randint
        mgym.one_step = real_one_step
        mannequin.Adam = RealAdam
        sys.path.remove(bench_dir)

    pol, val = adams[0], adams[1]
    assert len(pol.trace) == 600 and len(val.trace) == 600, (len(pol.trace), len(val.trace))
    ck = [0, 1, 2, 4, 9, 19, 49, 99, 199, 299, 300, 301, 309, 399, 599]
    out["ppo_obs"] = np.array([h[0] for h in hist], dtype=np.float32)
    out["ppo_act"] = np.array([h[1] for h in hist], dtype=np.float32)
    out["ppo_rew"] = np.array([h[2] for h in hist], dtype=np.float64)
    out["ppo_done"] = np.array([h[3] for h in hist], dtype=np.bool_)
    out["ppo_val_idx"] = np.array(idx_log["value"], dtype=np.int32)
    out["ppo_pol_idx"] = np.array(idx_log["policy"], dtype=np.int32)
    out["ppo_pol_init"] = pol.init_value.astype(np.float32)
    out["ppo_val_init"] = val.init_value.astype(np.float32)
    out["ppo_ck_steps"] = np.array(ck)
    out["ppo_pol_ck"] = np.array([pol.trace[i] for i in ck])
    out["ppo_val_ck"] = np.array([val.trace[i] for i in ck])


def gen_gae(out, gym):
    """benchmarks/gae.py:gae() driven directly: advantages of two chunks."""
    import mannequin.gym as mgym
    bench_dir = os.path.join(REF, "benchmarks")
    sys.path.insert(0, bench_dir)
    try:
        import gae as gae_mod
        env = mgym.NormalizedObservations(SyntheticEnv(gym, seed=77))
        rs = np.random.RandomState(78)
        hist = []
        real_one_step = mgym.one_step

        def rec_one_step(e, p):
            s = real_one_step(e, p)
            hist.append(s)
            return s
        mgym.one_step = rec_one_step
        np.random.seed(79)
        get_chunk = gae_mod.gae(env, lambda o: rs.randn(3) * 0.5, gam=0.97, lam=0.9)
        t1 = get_chunk()
        t2 = get_chunk()
        mgym.one_step = real_one_step
    finally:
        sys.path.remove(bench_dir)
    out["gae_obs"] = np.array([h[0] for h in hist], dtype=np.float32)
    out["gae_rew"] = np.array([h[2] for h in hist], dtype=np.float64)
    out["gae_done"] = np.array([h[3] for h in hist], dtype=np.bool_)
    out["gae_adv1"], out["gae_adv2"] = t1.r, t2.r
    out["gae_o2_first"] = t2.o[0]


def gen_es(out, gym):
    """benchmarks/mutation.py:run() itself for 12 generations (fake env factory; episodes cut after 40 steps): every
    perturbation, episode return and the optimiser value after each step are recorded."""
    import mannequin
    bench_dir = os.path.join(REF, "benchmarks")
    sys.path.insert(0, bench_dir)
    n_gen = 12

    class ShortEnv(SyntheticEnv):
        def __init__(self, gym, seed):
            super().__init__(gym, seed)
            self.t = 0

        def reset(self):
            self.t = 0
            return super().reset()

        def step(self, action):
            o, r, d, i = super().step(action)
            self.t += 1
            return o, r, d or self.t >= 40, i

    env_mod = types.ModuleType("_env")
    calls = {"n": 0}

    def get_progress():
        calls["n"] += 1
        return 0.0 if calls["n"] <= n_gen else 1.0

    env_mod.build_env = lambda: ShortEnv(gym, seed=91)
    env_mod.get_progress = get_progress
    saved_env = sys.modules.get("_env")
    sys.modules["_env"] = env_mod
    adams, rets = [], []
    RealAdam, RealNorm = mannequin.Adam, mannequin.RunningNormalize

    def rec_adam(*a, **k):
        opt = RealAdam(*a, **k)
        adams.append(opt)
        real_apply = opt.apply_gradient
        opt.trace = [np.array(opt.get_value())]

        def apply(grad, **kw):
            opt.grads = getattr(opt, "grads", []) + [np.array(grad)]
            real_apply(grad, **kw)
            opt.trace.append(np.array(opt.get_value()))
        opt.apply_gradient = apply
        return opt

    class RecNorm(RealNorm):
        def __call__(self, value):
            rets.append(float(value))
            return super().__call__(value)
    mannequin.Adam, mannequin.RunningNormalize = rec_adam, RecNorm
    try:
        import importlib
        np.random.seed(93)
        mut = importlib.import_module("mutation")
        mut.run()
    finally:
        mannequin.Adam, mannequin.RunningNormalize = RealAdam, RealNorm
        sys.path.remove(bench_dir)
        if saved_env is not None:
            sys.modules["_env"] = saved_env
        else:
            del sys.modules["_env"]
    opt = adams[0]
    # the perturbations are np.random.randn(P) * 0.1 drawn from the global generator seeded above, after the three
    # normed_columns draws of build_policy: the test replays them instead of storing 12 x 5123 doubles
    keep = [0, 1, 2, 3, 6, n_gen]
    out["es_seed"] = np.array([93])
    out["es_keep"] = np.array(keep)
    out["es_trace"] = np.array([opt.trace[k] for k in keep])    # optimiser value (fp64) after `keep` generations
    out["es_grad_norms"] = np.array([np.linalg.norm(x) for x in opt.grads])
    out["es_returns"] = np.array(rets)                          # episode returns fed to the normaliser


def gen_reinforce(out, gym):
    """benchmarks/policy.py:run() itself (plain REINFORCE, policy.py:29-47) for 6 episodes of <= 30 steps on a stand-in
    environment with a Box action space (Gauss policy): the trajectories it sampled and the optimiser values."""
    import mannequin
    import mannequin.gym as mgym
    bench_dir = os.path.join(REF, "benchmarks")
    sys.path.insert(0, bench_dir)
    n_ep = 6

    class ShortEnv(SyntheticEnv):
        def __init__(self, gym, seed):
            super().__init__(gym, seed)
            self.t = 0

        def reset(self):
            self.t = 0
            return super().reset()

        def step(self, action):
            o, r, d, i = super().step(action)
            self.t += 1
            return o, r, d or self.t >= 30, i

    env_mod = types.ModuleType("_env")
    calls = {"n": 0}

    def get_progress():
        calls["n"] += 1
        return 0.0 if calls["n"] <= n_ep else 1.0

    env_mod.build_env = lambda: ShortEnv(gym, seed=61)
    env_mod.get_progress = get_progress
    saved_env = sys.modules.get("_env")
    sys.modules["_env"] = env_mod
    adams, trajs = [], []
    RealAdam = mannequin.Adam

    def rec_adam(*a, **k):
        opt = RealAdam(*a, **k)
        adams.append(opt)
        real_apply = opt.apply_gradient
        opt.trace = [np.array(opt.get_value())]

        def apply(grad, **kw):
            real_apply(grad, **kw)
            opt.trace.append(np.array(opt.get_value()))
        opt.apply_gradient = apply
        return opt

    real_episode = mgym.episode

    def rec_episode(env, policy, **kw):
        t = real_episode(env, policy, **kw)
        trajs.append((np.array(t.o), np.array(t.a), np.array(t.r)))
        return t
    RealRS = np.random.RandomState
    seeds = {"n": 300}

    class SeededRS(RealRS):
        def __new__(cls, *a, **k):
            return RealRS.__new__(cls)

        def __init__(self, seed=None):
            if seed is None:
                seeds["n"] += 1
                seed = seeds["n"]
            super().__init__(seed)
    mannequin.Adam, mgym.episode, np.random.RandomState = rec_adam, rec_episode, SeededRS
    try:
        import importlib
        np.random.seed(63)
        pol = importlib.import_module("policy")
        importlib.reload(pol)
        pol.run()
    finally:
        mannequin.Adam, mgym.episode, np.random.RandomState = RealAdam, real_episode, RealRS
        sys.path.remove(bench_dir)
        if saved_env is not None:
            sys.modules["_env"] = saved_env
        else:
            del sys.modules["_env"]
    opt = adams[0]
    out["rf_trace"] = np.array(opt.trace).astype(np.float32)      # [n_ep + 1, 5126] optimiser values
    out["rf_len"] = np.array([len(t[0]) for t in trajs])
    out["rf_obs"] = np.concatenate([t[0] for t in trajs]).astype(np.float32)
    out["rf_act"] = np.concatenate([t[1] for t in trajs]).astype(np.float32)
    out["rf_rew"] = np.concatenate([t[2] for t in trajs]).astype(np.float32)


def mlp_script_data():
    """Synthetic MNIST-shaped data for mnist/mlp.py (regenerated identically by the test)."""
    rs = np.random.RandomState(5)
    train_x = rs.rand(384, 28, 28).astype(np.float32)
    train_y = np.eye(10, dtype=np.float32)[rs.randint(10, size=384)]
    test_x = rs.rand(64, 28, 28).astype(np.float32)
    test_y = np.eye(10, dtype=np.float32)[rs.randint(10, size=64)]
    return dict(train_x=train_x, train_y=train_y, test_x=test_x, test_y=test_y)


def gen_mlp_script(out):
    """mnist/mlp.py:run() itself (30 optimiser steps: 10 "epochs" of 3 on a 384-sample synthetic __mnist.npz in a scratch
    directory): the minibatch index draws and a strided subset of the optimiser value after chosen steps."""
    import contextlib
    import importlib.util
    import io
    import tempfile
    import mannequin
    idx_log, adams = [], []
    RealAdam = mannequin.Adam

    def rec_adam(*a, **k):
        opt = RealAdam(*a, **k)
        adams.append(opt)
        real_apply = opt.apply_gradient
        opt.trace = [np.array(opt.get_value())]

        def apply(grad, **kw):
            real_apply(grad, **kw)
            opt.trace.append(np.array(opt.get_value()))
        opt.apply_gradient = apply
        return opt
    real_randint = np.random.randint

    def rec_randint(*a, **k):
        r = real_randint(*a, **k)
        if k.get("size") == 128:
            idx_log.append(np.array(r))
        return r
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        np.savez(os.path.join(tmp, "__mnist.npz"), **mlp_script_data())
        os.chdir(tmp)
        mannequin.Adam, np.random.randint = rec_adam, rec_randint
        try:
            spec = importlib.util.spec_from_file_location("ref_mlp", os.path.join(REF, "mnist", "mlp.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            np.random.seed(71)
            with contextlib.redirect_stdout(io.StringIO()) as printed:
                mod.run()
        finally:
            mannequin.Adam, np.random.randint = RealAdam, real_randint
            os.chdir(cwd)
    opt = adams[0]
    keep = [0, 1, 2, 3, 6, 10, 15, 20]              # beyond ~20 steps rounding-level differences grow (chaotic on 384 samples)
    out["mlps_seed"] = np.array([71])
    out["mlps_idx"] = np.array(idx_log).astype(np.int16)          # [30, 128]
    out["mlps_keep"] = np.array(keep)
    out["mlps_stride"] = np.array([37])
    out["mlps_trace"] = np.array([opt.trace[k][::37] for k in keep]).astype(np.float32)
    out["mlps_bars"] = np.array(printed.getvalue().splitlines())


def main():
    sys.path.insert(0, REF)
    install_fake_autograd()
    gym = install_fake_gym()
    np.set_printoptions(precision=3, linewidth=100, suppress=True)

    groups = {}
    for name, fn in (("nets", gen_nets), ("adam", gen_adam), ("norm", gen_norm),
                     ("trajectory", gen_trajectory), ("gauss", gen_gauss), ("discrete", gen_discrete)):
        d = {}
        fn(d)
        groups[name] = d
    d = {}
    gen_ppo(d, gym)
    groups["ppo"] = d
    d = {}
    gen_gae(d, gym)
    groups["gae"] = d
    d = {}
    gen_obsnorm(d, gym)
    groups["obsnorm"] = d
    d = {}
    gen_vae(d)
    groups["vae"] = d
    d = {}
    gen_vae_step(d)
    groups["vae_step"] = d
    d = {}
    gen_es(d, gym)
    groups["es"] = d
    d = {}
    gen_reinforce(d, gym)
    groups["reinforce"] = d
    d = {}
    gen_mlp_script(d)
    groups["mlp_script"] = d
    only = sys.argv[1:]
    for name, d in groups.items():
        if only and name not in only:
            continue
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print("%-12s %3d arrays  %7.1f KB" % (name, len(d), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
