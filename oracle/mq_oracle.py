"""CPU oracle for the mannequin2 training hot path.  TEST INFRASTRUCTURE ONLY.

This module is a fresh NumPy restatement of the algorithms on the hot path of
squirrelinhell/mannequin2.  It is the *checker*: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it.  Nothing under
``mannequin2_b200/`` imports it and the product path never falls back to it.

Parity status (see DESIGN.md section "Oracle"):
  * pinned against the reference's own goldens (tests/adam.out, basicnet.out,
    graph.out, reshape.out, trajectory.out, running_mean.py, running_norm.py)
    and against outputs of the *real* reference imported from /root/reference
    in the build container (tests/golden/make_golden.py -> tests/golden/*.npz);
  * ``gauss_logprob`` / ``dkl_uninormal`` gradients: the reference delegates
    the derivative to the third-party ``autograd`` package (unpinned, absent).
    The forward expressions are pinned by running the reference code with a
    complex-step stand-in for ``autograd`` (make_golden.py); the closed-form
    derivatives are additionally checked by finite differences;
  * ``vae_grads`` / ``vae_step`` against one iteration of mnist/vae.py:57-69 run
    from the reference's own classes (golden/vae_step.npz);
  * ``dkl_uninormal`` / ``clip_*`` against mnist/vae.py imported as a module,
    ``es_apply_returns`` against benchmarks/mutation.py:run() itself,
    ``normalized_observations_step`` against mannequin/gym.py's wrapper, the
    REINFORCE chain (discounted, normalise, logprob gradient, Adam) against
    benchmarks/policy.py:run() (golden/vae.npz, es.npz, obsnorm.npz,
    reinforce.npz);
  * ``gauss_entropy`` has no reference code at all: "parity unpinned".

All citations are relative to /root/reference.
"""

import numpy as np

ACT_NONE, ACT_TANH, ACT_LRELU = 0, 1, 2


# --------------------------------------------------------------------------
# Flat parameter layout (mannequin/basicnet.py:13-23,35-51)
# --------------------------------------------------------------------------

def mlp_spec(n_in, widths, acts, leak=0.1):
    """Describe a stack of Affine(+activation) layers.

    Flat layout is DFS post-order of Params nodes: W1 (in x out, row-major),
    b1, W2, b2, ...   (basicnet.py:303-311,319-320)
    """
    layers, pos, k = [], 0, int(n_in)
    for n, a in zip(widths, acts):
        n = int(n)
        layers.append(dict(n_in=k, n_out=n, act=int(a), leak=float(leak),
                           w_off=pos, b_off=pos + k * n))
        pos += k * n + n
        k = n
    return dict(layers=layers, n_params=pos, n_in=int(n_in), n_out=k)


def normed_columns(rs, n_in, n_out, scale=1.0):
    """Column-normalised Gaussian init (basicnet.py:265-267,300-301)."""
    m = rs.randn(n_in, n_out).astype(np.float32)
    m = m / np.sqrt(np.sum(np.square(m), axis=0))
    return (m * np.float32(scale)).astype(np.float32)


def mlp_init(spec, rs, head_scale=1.0):
    """theta in the reference's layout; biases zero (basicnet.py:310-311)."""
    theta = np.zeros(spec["n_params"], dtype=np.float32)
    last = len(spec["layers"]) - 1
    for i, L in enumerate(spec["layers"]):
        w = normed_columns(rs, L["n_in"], L["n_out"],
                           head_scale if i == last else 1.0)
        theta[L["w_off"]:L["b_off"]] = w.reshape(-1)
    return theta


def _wb(theta, L):
    w = theta[L["w_off"]:L["b_off"]].reshape(L["n_in"], L["n_out"])
    b = theta[L["b_off"]:L["b_off"] + L["n_out"]]
    return w, b


# --------------------------------------------------------------------------
# MLP forward / backward (basicnet.py:192-259, backprop.py:11-60)
# --------------------------------------------------------------------------

def mlp_forward(theta, spec, x, dtype=np.float64):
    """Returns the list of post-activation outputs, one per layer.

    The reference keeps float32 until the first LReLU and float64 after it
    (backprop.py:53-56); the oracle computes in ``dtype``
    (float64 by default, which is at least as accurate as either).
    """
    theta = np.asarray(theta)
    h = np.asarray(x, dtype=dtype).reshape(-1, spec["n_in"])
    outs = []
    for L in spec["layers"]:
        w, b = _wb(theta, L)
        z = h @ w.astype(dtype) + b.astype(dtype)
        if L["act"] == ACT_TANH:
            h = np.tanh(z)
        elif L["act"] == ACT_LRELU:
            h = np.where(z < 0.0, z * dtype(L["leak"]), z)
        else:
            h = z
        outs.append(h)
    return outs


def mlp_backward(theta, spec, x, outs, dy, dtype=np.float64, want_dx=False):
    """Flat parameter gradient, batch MEAN (basicnet.py:242-249)."""
    theta = np.asarray(theta)
    x = np.asarray(x, dtype=dtype).reshape(-1, spec["n_in"])
    g = np.asarray(dy, dtype=dtype).reshape(x.shape[0], spec["n_out"])
    batch = x.shape[0]
    grad = np.zeros(spec["n_params"], dtype=dtype)
    for i in range(len(spec["layers"]) - 1, -1, -1):
        L = spec["layers"][i]
        w, _ = _wb(theta, L)
        y = outs[i]
        if L["act"] == ACT_TANH:
            g = g * (1.0 - np.square(y))                  # backprop.py:58-60
        elif L["act"] == ACT_LRELU:
            g = np.where(y < 0.0, g * dtype(L["leak"]), g)  # backprop.py:53-56
        inp = outs[i - 1] if i > 0 else x
        grad[L["w_off"]:L["b_off"]] = (inp.T @ g).reshape(-1) / batch
        grad[L["b_off"]:L["b_off"] + L["n_out"]] = g.sum(axis=0) / batch
        if i > 0 or want_dx:
            g = g @ w.astype(dtype).T                     # backprop.py:48-51
    if want_dx:
        return grad.astype(np.float32), g
    return grad.astype(np.float32)                        # basicnet.py:218


# --------------------------------------------------------------------------
# RunningMean / RunningNormalize (mannequin/_normalize.py:4-81)
# --------------------------------------------------------------------------

class RunningMeanO:
    def __init__(self, shape=(), horizon=None):
        self.shape = tuple(np.zeros(shape).shape)
        self.mean = np.zeros(self.shape, dtype=np.float64)
        self.tw = 0.0
        self.decay = None if horizon is None else 1.0 - 1.0 / float(horizon)

    def update(self, values, weight=1.0):
        v = np.asarray(values, dtype=np.float64).reshape(self.shape)
        weight = float(weight)
        if self.decay is not None:
            self.tw *= np.power(self.decay, weight)       # _normalize.py:18-22
        self.tw += weight
        self.mean = self.mean + (weight / self.tw) * (v - self.mean)

    def get(self):
        return np.array(self.mean)


class RunningNormalizeO:
    def __init__(self, shape=(), horizon=None):
        self.shape = tuple(np.zeros(shape).shape)
        self.m = RunningMeanO(self.shape, horizon)
        self.v = RunningMeanO(self.shape, horizon)
        self.n = 0

    def update(self, value):
        x = np.array(value, dtype=np.float64).reshape((-1,) + self.shape)
        d_old = x - self.m.get()
        self.m.update(x.mean(axis=0))                     # one weight-1 update per batch
        d_new = x - self.m.get()
        cov = (d_old * d_new).mean(axis=0)
        if self.n >= 1:
            self.v.update(cov)
        elif len(x) >= 2:
            w = (len(x) - 1.0) / len(x)                   # _normalize.py:50-55
            self.v.update(cov / w, weight=w)
        self.n += len(x)

    def get_mean(self):
        return self.m.get()

    def get_var(self):
        return np.ones(self.shape) if self.n < 2 else self.v.get()

    def get_std(self):
        return np.sqrt(self.get_var())

    def apply(self, value):
        if self.n < 2:
            return np.zeros_like(value, dtype=np.float64)
        return (value - self.m.get()) / np.maximum(1e-8, np.sqrt(self.v.get()))

    def __call__(self, value):
        self.update(value)
        return self.apply(value)


# --------------------------------------------------------------------------
# Adam / Adams (mannequin/_adam.py:7-64): gradient ASCENT, fp64 state
# --------------------------------------------------------------------------

class AdamO:
    def __init__(self, value, horizon, var_horizon=100, lr=None, epsilon=1e-8,
                 adams=False):
        self.value = np.array(value, dtype=np.float64)
        self.b1 = 1.0 - 1.0 / float(horizon)
        self.b2 = 1.0 - 1.0 / float(var_horizon)
        self.tw1 = self.tw2 = 0.0
        self.m = np.zeros_like(self.value)
        self.v = np.zeros_like(self.value)
        self.lr, self.eps, self.adams = lr, float(epsilon), bool(adams)

    def apply_gradient(self, grad, lr=None, epsilon=None):
        g = np.asarray(grad, dtype=np.float64)
        assert g.shape == self.value.shape
        lr = float(self.lr if lr is None else lr)
        eps = self.eps if epsilon is None else float(epsilon)
        self.tw1 = self.tw1 * self.b1 + 1.0
        self.tw2 = self.tw2 * self.b2 + 1.0
        self.m = self.m + (1.0 / self.tw1) * (g - self.m)
        self.v = self.v + (1.0 / self.tw2) * (np.square(g) - self.v)
        den = eps + (self.v if self.adams else np.sqrt(self.v))
        self.value = self.value + lr * (self.m / den)     # _adam.py:42,50-53

    def get_value(self):
        return self.value


# --------------------------------------------------------------------------
# Diagonal Gaussian head (mannequin/distrib.py:41-76)
# --------------------------------------------------------------------------

def gauss_logprob(mu, logstd, sample):
    """distrib.py:47,61-68: const - sum(logstd + 0.5*((s-mu)/exp(logstd))^2)."""
    mu = np.asarray(mu, dtype=np.float64)
    ls = np.broadcast_to(np.asarray(logstd, dtype=np.float64), mu.shape)
    s = np.asarray(sample, dtype=np.float64).reshape(mu.shape)
    a = mu.shape[-1]
    z = (s - mu) / np.exp(ls)
    return -0.5 * a * np.log(2.0 * np.pi) - np.sum(ls + 0.5 * z * z, axis=-1)


def gauss_logprob_grads(mu, logstd, sample, g):
    """Closed-form derivative of gauss_logprob weighted by g[B].

    Returns (d_mu[B,A], d_logstd[B,A]); when logstd is a Params(A) the caller
    sums d_logstd over the batch (and basicnet.py:242-249 divides by B).
    """
    mu = np.asarray(mu, dtype=np.float64)
    ls = np.broadcast_to(np.asarray(logstd, dtype=np.float64), mu.shape)
    s = np.asarray(sample, dtype=np.float64).reshape(mu.shape)
    g = np.asarray(g, dtype=np.float64).reshape(mu.shape[:-1] + (1,))
    inv = np.exp(-ls)
    z = (s - mu) * inv
    return g * z * inv, g * (z * z - 1.0)


def gauss_sample(mu, logstd, unit_noise):
    """distrib.py:50-59 with the N(0,1) draw as an input."""
    noise = np.asarray(unit_noise, dtype=np.float64) * np.exp(
        np.asarray(logstd, dtype=np.float64))
    return np.asarray(mu, dtype=np.float64) + noise, noise


def gauss_entropy(logstd, n_act):
    """Not in the reference: H = sum(logstd) + A/2 (1+ln 2pi)."""
    ls = np.asarray(logstd, dtype=np.float64)
    return np.sum(ls, axis=-1) + 0.5 * n_act * (1.0 + np.log(2.0 * np.pi))


def policy_spec(n_obs, n_act, hid=64, n_hid=2):
    """benchmarks/policy.py:7-27: Tanh MLP + Gauss; logstd = last A params."""
    spec = mlp_spec(n_obs, [hid] * n_hid + [n_act],
                    [ACT_TANH] * n_hid + [ACT_NONE])
    spec = dict(spec)
    spec["logstd_off"] = spec["n_params"]
    spec["n_mlp_params"] = spec["n_params"]
    spec["n_params"] = spec["n_params"] + n_act
    return spec


def policy_logprob(theta, spec, obs, act, dtype=np.float64):
    """dtype = np.float32 keeps the Tanh net in float32 like the reference (basicnet.py:193,210): the CPU
    baseline's setting; the parity tests use the float64 default."""
    outs = mlp_forward(theta[:spec["n_mlp_params"]], _mlp_part(spec), obs, dtype=dtype)
    ls = np.asarray(theta[spec["logstd_off"]:], dtype=np.float64)
    return gauss_logprob(outs[-1], ls, act), outs


def _mlp_part(spec):
    s = dict(spec)
    s["n_params"] = spec.get("n_mlp_params", spec["n_params"])
    return s


def policy_logprob_grad(theta, spec, obs, act, g, outs=None, dtype=np.float64):
    """Flat gradient [P] of sum_b g_b * logp_b, batch mean (F4)."""
    if outs is None:
        _, outs = policy_logprob(theta, spec, obs, act, dtype=dtype)
    ls = np.asarray(theta[spec["logstd_off"]:], dtype=np.float64)
    d_mu, d_ls = gauss_logprob_grads(outs[-1], ls, act, g)
    grad = np.zeros(spec["n_params"], dtype=np.float32)
    grad[:spec["n_mlp_params"]] = mlp_backward(
        theta[:spec["n_mlp_params"]], _mlp_part(spec), obs, outs, d_mu, dtype=dtype)
    grad[spec["logstd_off"]:] = (d_ls.sum(axis=0) / len(d_mu)).astype(np.float32)
    return grad


# --------------------------------------------------------------------------
# PPO ratio/clip rule (benchmarks/ppo.py:27,29-40)
# --------------------------------------------------------------------------

def ppo_clip_weight(logp, logp_old, adv, lo=0.8, hi=1.2):
    ratio = np.exp(np.asarray(logp, np.float64) - np.asarray(logp_old, np.float64))
    adv = np.asarray(adv, dtype=np.float64)
    w = ratio.copy()
    w[np.logical_and(ratio > hi, adv > 0.0)] = 0.0
    w[np.logical_and(ratio < lo, adv < 0.0)] = 0.0
    return w * adv


# --------------------------------------------------------------------------
# GAE and discounted (benchmarks/gae.py:43-64, mannequin/_utils.py:9-20)
# --------------------------------------------------------------------------

def gae_advantages(rew, value, done, gam=0.99, lam=0.95, dtype=np.float64):
    """rew[N], value[N+1], done[N] -> (adv[N] float32, ret[N] float32).

    done[i] true means step i is the last of its episode (gae.py:50-52);
    otherwise bootstrap from value[i+1] and chain gam*lam*adv[i+1]; adv[N]=0.
    ``dtype=np.float32`` reproduces NumPy-2 (NEP 50) arithmetic of the
    reference loop; float64 is the NumPy-1 behaviour.
    """
    n = len(rew)
    r = np.asarray(rew, dtype=dtype)
    v = np.asarray(value, dtype=dtype)
    adv = np.zeros(n + 1, dtype=dtype)
    g, gl = dtype(gam), dtype(float(gam) * float(lam))
    for i in range(n - 1, -1, -1):
        if done[i]:
            adv[i] = r[i] - v[i]
        else:
            adv[i] = r[i] + g * v[i + 1] - v[i]
            adv[i] += gl * adv[i + 1]
    adv32 = adv[:n].astype(np.float32)
    ret = (adv.astype(np.float32) + np.asarray(value, np.float32))[:n]  # gae.py:69
    return adv32, ret


def discounted(values, horizon):
    step = 1.0 / float(horizon)
    assert 1e-6 < step < 1.0
    out = np.array(values, dtype=np.float64)
    acc = 0.0
    for i in range(len(out) - 1, -1, -1):
        acc += step * (out[i] - acc)
        out[i] = acc
    return out


# --------------------------------------------------------------------------
# MNIST head (mnist/mlp.py:10-14,31-35) and VAE heads (mnist/vae.py:13-34)
# --------------------------------------------------------------------------

def softmax(v):
    v = np.asarray(v, dtype=np.float64)
    e = np.exp(v - v.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


def softmax_ce_grad(logits, onehot, l2=0.01):
    logits = np.asarray(logits, dtype=np.float64)
    return np.asarray(onehot, np.float64) - softmax(logits) - logits * l2


def dkl_uninormal(mean, logstd):
    """vae.py:13-22 (note exp(logstd), not exp(2 logstd))."""
    mean = np.asarray(mean, np.float64)
    logstd = np.asarray(logstd, np.float64)
    return 0.5 * (np.sum(np.exp(logstd) - logstd + mean * mean, axis=-1)
                  - mean.shape[-1])


def dkl_uninormal_grads(mean, logstd, g):
    g = np.asarray(g, np.float64).reshape(np.shape(mean)[:-1] + (1,))
    return g * np.asarray(mean, np.float64), \
        0.5 * g * (np.exp(np.asarray(logstd, np.float64)) - 1.0)


def clip_forward(v, lo, hi):
    return np.clip(v, lo, hi)


def clip_backward(v, g, lo, hi):
    g = np.array(g, dtype=np.float64)
    g[np.logical_and(v < lo, g < 0.0)] = 0.0              # vae.py:26-33
    g[np.logical_and(v > hi, g > 0.0)] = 0.0
    return g


# --------------------------------------------------------------------------
# The training step of mnist/vae.py:39-71 (two evaluations of the shared graph)
# --------------------------------------------------------------------------

def vae_layout(d_img, hidden, latent, n_hidden):
    """Flat layout of decoder.logprob's parameters (basicnet.py:66-77 order for
    the graph of vae.py:39-55): encoder hidden (W, b)..., W_mean, b_mean,
    W_logstd, b_logstd, decoder hidden (W, b)..., W_mean, b_mean, W_logstd,
    b_logstd.  Returns (encoder spec, decoder spec, offsets, n_params): the
    specs describe the tanh stacks only, the offsets the four heads."""
    enc = mlp_spec(d_img, [hidden] * n_hidden, [ACT_TANH] * n_hidden)
    pos = enc["n_params"]
    off = {}
    for name, n_out in (("e_mean", latent), ("e_logstd", latent)):
        off[name] = (pos, pos + hidden * n_out, n_out)
        pos += hidden * n_out + n_out
    off["dec"] = pos
    dec = mlp_spec(latent, [hidden] * n_hidden, [ACT_TANH] * n_hidden)
    pos += dec["n_params"]
    for name, n_out in (("d_mean", d_img), ("d_logstd", d_img)):
        off[name] = (pos, pos + hidden * n_out, n_out)
        pos += hidden * n_out + n_out
    return enc, dec, off, pos


def vae_grads(theta, d_img, hidden, latent, n_hidden, x, unit, lo=-6.0, hi=0.0):
    """vae.py:57-66 with the N(0,1) draw of the latent sample as an input:
    returns (logprob[B], dkl[B], grad1, grad2) -- grad1 the gradient of
    sum(logprob) w.r.t. all parameters, grad2 of -sum(dkl) w.r.t. the encoder's,
    both batch means (basicnet.py:231-236), computed as the reference does: two
    separate backward passes, the Clip rule (vae.py:26-34) applied in each."""
    theta = np.asarray(theta, np.float64)
    enc, dec, off, n_params = vae_layout(d_img, hidden, latent, n_hidden)
    assert theta.shape == (n_params,)
    x = np.asarray(x, np.float64).reshape(-1, d_img)
    batch = x.shape[0]
    unit = np.asarray(unit, np.float64).reshape(batch, latent)

    def head(name, h):
        w0, b0, n_out = off[name]
        return h @ theta[w0:b0].reshape(hidden, n_out) + theta[b0:b0 + n_out]

    def head_grads(name, h, g, grad):
        w0, b0, n_out = off[name]
        grad[w0:b0] += (h.T @ g).reshape(-1) / batch
        grad[b0:b0 + n_out] += g.sum(axis=0) / batch
        return g @ theta[w0:b0].reshape(hidden, n_out).T

    th_e, th_d = theta[:enc["n_params"]], theta[off["dec"]:off["dec"] + dec["n_params"]]
    e_outs = mlp_forward(th_e, enc, x)
    mean_e, raw_e = head("e_mean", e_outs[-1]), head("e_logstd", e_outs[-1])
    ls_e = clip_forward(raw_e, lo, hi)
    z, noise = gauss_sample(mean_e, ls_e, unit)                    # distrib.py:48-51
    d_outs = mlp_forward(th_d, dec, z)
    mean_d, raw_d = head("d_mean", d_outs[-1]), head("d_logstd", d_outs[-1])
    ls_d = clip_forward(raw_d, lo, hi)
    logprob = gauss_logprob(mean_d, ls_d, x)
    dkl = dkl_uninormal(mean_e, ls_e)

    def encoder_backward(d_mean, d_ls, grad):
        dh = head_grads("e_mean", e_outs[-1], d_mean, grad)
        dh = dh + head_grads("e_logstd", e_outs[-1], clip_backward(raw_e, d_ls, lo, hi), grad)
        grad[:enc["n_params"]] += mlp_backward(th_e, enc, x, e_outs, dh)

    # pass 1: decoder.logprob, upstream +1 (vae.py:60-61)
    grad1 = np.zeros(n_params)
    d_mu, d_ls = gauss_logprob_grads(mean_d, ls_d, x, np.ones(batch))
    dh = head_grads("d_mean", d_outs[-1], d_mu, grad1)
    dh = dh + head_grads("d_logstd", d_outs[-1], clip_backward(raw_d, d_ls, lo, hi), grad1)
    g_dec, dz = mlp_backward(th_d, dec, z, d_outs, dh, want_dx=True)
    grad1[off["dec"]:off["dec"] + dec["n_params"]] += g_dec
    encoder_backward(dz, dz * noise, grad1)                         # sample: (g, g * noise), distrib.py:51
    # pass 2: dkl, upstream -1 (vae.py:63-64)
    grad2 = np.zeros(n_params)
    d_mean, d_ls = dkl_uninormal_grads(mean_e, ls_e, -np.ones(batch))
    encoder_backward(d_mean, d_ls, grad2)
    return logprob, dkl, grad1, grad2[:off["dec"]]


def vae_step(theta, momentum, d_img, hidden, latent, n_hidden, x, unit, lr=0.001):
    """vae.py:57-69: grad1[:len(grad2)] += grad2; momentum = clip(0.9 momentum +
    0.1 grad1, -1, 1); theta += 0.001 momentum.  Returns (theta, momentum,
    logprob, dkl, grad)."""
    logprob, dkl, grad1, grad2 = vae_grads(theta, d_img, hidden, latent, n_hidden, x, unit)
    grad1 = grad1.copy()
    grad1[:len(grad2)] += grad2
    momentum = np.clip(np.asarray(momentum, np.float64) * 0.9 + grad1 * 0.1, -1.0, 1.0)
    return np.asarray(theta, np.float64) + lr * momentum, momentum, logprob, dkl, grad1


# --------------------------------------------------------------------------
# Trajectory dtype / indexing rules (mannequin/_trajectory.py:9-30,79-86)
# --------------------------------------------------------------------------

def standardize_field(values):
    if isinstance(values[0], (int, np.integer)):
        return np.asarray(values, dtype=np.int32)
    return np.asarray(values, dtype=np.float32)


def trajectory_take(obs, act, rew, idx):
    o, a, r = obs[idx], act[idx], rew[idx]
    if np.ndim(r) >= 1:
        return o, a, r
    return o[None], a[None], np.asarray([r], dtype=np.float32)


# --------------------------------------------------------------------------
# Whole-step restatements used as parity targets and CPU baselines
# --------------------------------------------------------------------------

def mlp_sgd_step(theta, spec, opt, x, onehot, l2=0.01):
    """One step of mnist/mlp.py:31-35.  Returns the new float32 theta."""
    outs = mlp_forward(theta, spec, x)
    dy = softmax_ce_grad(outs[-1], onehot, l2)
    opt.apply_gradient(mlp_backward(theta, spec, x, outs, dy))
    return opt.get_value().astype(np.float32)


def value_sgd_step(theta, spec, opt, obs, target, dtype=np.float64):
    """benchmarks/gae.py:20-23: dy = labels - outputs."""
    outs = mlp_forward(theta, spec, obs, dtype=dtype)
    dy = np.asarray(target, dtype).reshape(outs[-1].shape) - outs[-1]
    opt.apply_gradient(mlp_backward(theta, spec, obs, outs, dy, dtype=dtype))
    return opt.get_value().astype(np.float32)


def ppo_policy_step(theta, spec, opt, obs, act, adv, logp_old, lr=3e-4,
                    lo=0.8, hi=1.2, ent_coef=0.0, dtype=np.float64):
    """benchmarks/ppo.py:33-40 on an already gathered minibatch.

    ent_coef (not in the reference): + ent_coef * gauss_entropy
    in the objective, i.e. + ent_coef on every logstd gradient entry."""
    logp, outs = policy_logprob(theta, spec, obs, act, dtype=dtype)
    g = ppo_clip_weight(logp, logp_old, adv, lo, hi).astype(np.float32)
    grad = policy_logprob_grad(theta, spec, obs, act, g, outs, dtype=dtype)
    if ent_coef:
        grad = np.array(grad, dtype=np.float32)
        grad[spec["logstd_off"]:spec["logstd_off"] + act.shape[1]] += np.float32(ent_coef)
    opt.apply_gradient(grad, lr=lr)
    return opt.get_value().astype(np.float32)


def ppo_update(pol_theta, pol_spec, pol_opt, val_theta, val_spec, val_opt,
               norm, obs, act, rew, done, val_idx, pol_idx,
               gam=0.99, lam=0.95, pol_lr=3e-4, ent_coef=0.0, dtype=np.float64):
    """One chunk of benchmarks/ppo.py:22-40 + benchmarks/gae.py:34-75.

    obs[N+1,D], act[N,A], rew[N], done[N]; val_idx/pol_idx are the
    [steps, minibatch] int tables the reference draws with randint.
    """
    n = len(rew)
    value = mlp_forward(val_theta, val_spec, obs, dtype=dtype)[-1].reshape(-1).astype(np.float32)
    adv, ret = gae_advantages(rew, value, done, gam, lam)
    for idx in val_idx:                                   # gae.py:71-73
        val_theta = value_sgd_step(val_theta, val_spec, val_opt,
                                   obs[:n][idx], ret[idx].reshape(-1, 1), dtype=dtype)
    adv_n = np.tanh(norm(adv)).astype(np.float32)         # ppo.py:24-25
    logp_old, _ = policy_logprob(pol_theta, pol_spec, obs[:n], act, dtype=dtype)
    for idx in pol_idx:                                   # ppo.py:29-40
        pol_theta = ppo_policy_step(pol_theta, pol_spec, pol_opt, obs[:n][idx],
                                    act[idx], adv_n[idx], logp_old[idx], pol_lr,
                                    ent_coef=ent_coef, dtype=dtype)
    return pol_theta, val_theta, dict(adv=adv, ret=ret, adv_n=adv_n,
                                      logp_old=logp_old, value=value)


def es_generation(theta, spec, opt, norm, noise, obs, ret_coef, lr=0.01):
    """Population form of benchmarks/mutation.py:33-37.

    noise[M,P] (already scaled by 0.1), shared obs[T,D]; synthetic return
    R_p = sum_t sum_a act[p,t,a]*ret_coef[a]; all returns are fed to the
    normaliser as ONE batch; grad = mean_p noise_p * r_p.
    """
    rets = np.zeros(len(noise), dtype=np.float64)
    for p in range(len(noise)):
        th = (np.asarray(theta, np.float64) + noise[p]).astype(np.float32)
        a = mlp_forward(th, spec, obs, dtype=np.float32)[-1]
        rets[p] = np.sum(a.astype(np.float64) @ np.asarray(ret_coef, np.float64))
    r = es_apply_returns(opt, norm, noise, rets.astype(np.float32), lr)
    return opt.get_value().astype(np.float32), rets, r


def es_apply_returns(opt, norm, noise, rets, lr=0.01):
    """benchmarks/mutation.py:35-37 for a population: the returns go through the
    normaliser as ONE batch, Adam ascends mean_p noise_p * r_p (population 1:
    exactly `r = normalize(R); opt.apply_gradient(diff * r, lr=0.01)`)."""
    r = np.reshape(norm(np.asarray(rets)), -1)
    grad = (np.asarray(noise, np.float64) * r[:, None]).mean(axis=0)
    opt.apply_gradient(grad, lr=lr)
    return r


# ---------------------------------------------------------------------------------------------------
# Categorical policy head: mannequin/distrib.py:8-39
def discrete_logprob(logits, sample):
    """log softmax(logits)[sample] per row and the softmax (distrib.py:15-19,24-30); fp64."""
    y = np.asarray(logits, dtype=np.float64)
    y = y.reshape(-1, y.shape[-1])
    s = np.asarray(sample, dtype=np.int64).reshape(-1)
    z = y - y.max(axis=1, keepdims=True)
    p = np.exp(z)
    p /= p.sum(axis=1, keepdims=True)
    return np.log(p[np.arange(len(s)), s]), p


def discrete_logprob_grad(theta, spec, obs, sample, g, outs=None):
    """d/dtheta mean_b g_b logp_b for logits = mlp(obs): backward rule g*(onehot - p) (distrib.py:31)."""
    outs = mlp_forward(theta, spec, obs) if outs is None else outs
    _, p = discrete_logprob(outs[-1], sample)
    onehot = np.eye(p.shape[1])[np.asarray(sample, dtype=np.int64).reshape(-1)]
    dy = np.asarray(g, dtype=np.float64).reshape(-1, 1) * (onehot - p)
    return mlp_backward(theta, spec, obs, outs, dy)


def ppo_discrete_policy_step(theta, spec, opt, obs, act, adv, logp_old, lr=3e-4, lo=0.8, hi=1.2):
    """benchmarks/ppo.py:33-40 with the categorical policy of benchmarks/policy.py:15-17."""
    outs = mlp_forward(theta, spec, obs)
    logp, _ = discrete_logprob(outs[-1], act)
    g = ppo_clip_weight(logp, logp_old, adv, lo, hi).astype(np.float32)
    opt.apply_gradient(discrete_logprob_grad(theta, spec, obs, act, g, outs), lr=lr)
    return opt.get_value().astype(np.float32)


# ---------------------------------------------------------------------------
# Rollout side: NormalizedObservations, batched actor
# ---------------------------------------------------------------------------
def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011; Random123), vectorised
    over uint32 arrays.  Not part of the reference (its RNG is unseeded
    np.random); the rollout kernels draw their noise from it."""
    m0, m1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    c = [np.asarray(v, dtype=np.uint32) for v in (c0, c1, c2, c3)]
    c = list(np.broadcast_arrays(*c))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(rounds):
        p0 = m0 * c[0].astype(np.uint64)
        p1 = m1 * c[2].astype(np.uint64)
        n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c[1] ^ k0
        n1 = (p1 & mask).astype(np.uint32)
        n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c[3] ^ k1
        n3 = (p0 & mask).astype(np.uint32)
        c = [n0, n1, n2, n3]
        k0 = np.uint32((int(k0) + 0x9E3779B9) & 0xFFFFFFFF)
        k1 = np.uint32((int(k1) + 0xBB67AE85) & 0xFFFFFFFF)
    return c


def philox_normals(seed, a, b, j):
    """Unit normal number(s) `j` of stream (seed, a, b) as mq_rollout.cu draws
    them: Philox block j // 4 -> two Box-Muller pairs, element j % 4."""
    j, a, b = np.broadcast_arrays(np.asarray(j, dtype=np.int64), np.asarray(a, dtype=np.uint32),
                                  np.asarray(b, dtype=np.uint32))
    r = philox4x32((j >> 2).astype(np.uint32), a, b, np.uint32(0),
                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    r = [np.asarray(v) for v in np.broadcast_arrays(*r)]
    z = []
    for h in range(2):
        u1 = ((r[2 * h] >> np.uint32(8)).astype(np.float64) + 1.0) / 16777216.0
        u2 = (r[2 * h + 1] >> np.uint32(8)).astype(np.float64) / 16777216.0
        rad = np.sqrt(-2.0 * np.log(u1))
        z += [rad * np.cos(2.0 * np.pi * u2), rad * np.sin(2.0 * np.pi * u2)]
    z = np.stack(z, axis=-1)
    return np.take_along_axis(z, (j & 3)[..., None], axis=-1)[..., 0].astype(np.float32)


def normalized_observations_step(norm, raw, clip=5.0):
    """mannequin/gym.py:66-71 on the batch of one time step: `norm` is a
    RunningNormalizeO (mannequin/_normalize.py:34-81)."""
    return np.clip(norm(np.asarray(raw, dtype=np.float64)), -clip, clip)


def synthetic_env_step(env, x, ep_step, act, seed, step_id):
    """The synthetic environment of mq_rollout.cu (not the reference's: gym is
    out of scope).  env: dict(dt, damp, sigma, ctrl, p_end, horizon, b[d,a],
    obs_scale[d], obs_shift[d]).  Returns (x', ep_step', rew, done, raw')."""
    n, d = x.shape
    e = np.arange(n)
    noise = np.stack([philox_normals(seed, step_id, e, 16 + i) for i in range(d)], axis=1)
    f = -env["damp"] * x + act @ env["b"].T
    xn = x + env["dt"] * f + env["sigma"] * noise
    rew = -(np.sum(xn * xn, axis=1) / d) - env["ctrl"] * np.sum(act * act, axis=1) / act.shape[1]
    steps = ep_step + 1
    u = philox4x32(np.uint32(0xFFFFFFFF), np.uint32(step_id), e.astype(np.uint32), np.uint32(1),
                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)[0]
    done = (steps >= env["horizon"]) | ((u >> np.uint32(8)).astype(np.float64) / 16777216.0 < env["p_end"])
    fresh = np.stack([philox_normals(seed, step_id, e, 16 + 64 + i) for i in range(d)], axis=1)
    xn = np.where(done[:, None], fresh, xn)
    raw = env["obs_scale"] * xn + env["obs_shift"]
    return xn.astype(np.float32), np.where(done, 0, steps).astype(np.int32), rew.astype(np.float32), done, \
        raw.astype(np.float32)


def bootstrap_segments(rew, done, v_next, seg_len, gam):
    """benchmarks/gae.py:52-61 at a chunk end (adv beyond the chunk is 0) for
    segments laid end to end: r += gam * v(next), done = True."""
    rew, done = np.array(rew, dtype=np.float32), np.array(done, dtype=bool)
    last = np.arange(len(v_next)) * seg_len + seg_len - 1
    cont = ~done[last]
    rew[last[cont]] += np.float32(gam) * np.asarray(v_next, dtype=np.float32)[cont]
    done[last] = True
    return rew, done
