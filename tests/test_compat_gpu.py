"""The reference's own test-suite (tests/*.py + *.out goldens), run against the drop-in API
of mannequin2_b200 on a B200.  Golden outputs are copied from /root/reference/tests/*.out as
data; printing uses NumPy's legacy 1.13 format, which is what the goldens were written in."""
import io
import sys
from contextlib import redirect_stdout

import numpy as np
import pytest

from conftest import assert_close, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _printopts():
    old = np.get_printoptions()
    np.set_printoptions(precision=3, linewidth=100, suppress=True, legacy="1.13")
    yield
    np.set_printoptions(**old)


def _lines(text):
    return [" ".join(l.split()) for l in text.strip().splitlines()]


def test_adam_py():
    """tests/adam.py vs tests/adam.out (20 vectors; ascent sign, bias correction, per-call lr)."""
    from mannequin2_b200 import Adam
    out = io.StringIO()
    with redirect_stdout(out):
        opt = Adam([10.0, -5.0], lr=2.0, horizon=3)
        for i in range(10):
            opt.apply_gradient(-opt.get_value())
            print(opt.get_value())
        print()
        opt = Adam([0.0, 0.0, 0.0], horizon=1)
        for i in range(10):
            opt.apply_gradient([0.1, 1.0, 10.0] - opt.get_value(), lr=7.0 / (i + 1))
            print(opt.get_value())
    want = """[ 8. -3.]
[ 6.055 -1.155]
[ 4.218  0.335]
[ 2.552  1.247]
[ 1.125  1.509]
[-0.001  1.25 ]
[-0.789  0.703]
[-1.236  0.105]
[-1.376 -0.347]
[-1.276 -0.545]

[ 7.  7.  7.]
[ 2.063  2.129  8.425]
[ 0.958  1.392  9.031]
[ 0.541  1.17   9.353]
[ 0.35   1.084  9.545]
[ 0.25   1.045  9.669]
[ 0.195  1.026  9.753]
[ 0.162  1.015  9.811]
[ 0.142  1.009  9.853]
[ 0.129  1.006  9.884]"""
    assert _lines(out.getvalue()) == _lines(want)
    # and bit-exact fp64 against the real reference's trace (tests/golden/adam.npz)
    g = load_golden("adam")
    opt = Adam(g["adam_values"][0], horizon=10, lr=0.003)
    for grad, lr, want_v in zip(g["adam_grads"], g["adam_lrs"], g["adam_values"][1:]):
        if np.isnan(lr):
            opt.apply_gradient(grad)
        else:
            opt.apply_gradient(grad, lr=lr, epsilon=1e-6)
        assert np.array_equal(np.asarray(opt.get_value()), want_v)
    with pytest.raises(TypeError):
        Adam([1.0], horizon=3).apply_gradient([1.0])          # lr is required, like the reference


def test_basicnet_py():
    """tests/basicnet.py vs tests/basicnet.out: flat layout, batch-MEAN gradient, LReLU / Tanh."""
    import mannequin2_b200.basicnet as net
    m = net.Affine(net.Affine(net.Input(2), 2), 2)
    m.load_params(np.arange(12))

    def check(m, x, g):
        v, backprop = m.evaluate(x)
        return [v, backprop(g).reshape(-1, 6)]

    values = []
    values += check(m, [1.0, 1.0], [1.0, -1.0])
    values += check(m, np.eye(2), [[1.0, -1.0], [-2.0, 2.0]])
    alt = np.mean([m.evaluate([1.0, 0.0])[1]([1.0, -1.0]).reshape(-1, 6),
                   m.evaluate([0.0, 1.0])[1]([-2.0, 2.0]).reshape(-1, 6)], axis=0)
    assert np.allclose(alt, values[-1])
    v = [-18.4, 2.0]
    values += [m(v)]
    values += check(net.LReLU(m), v, [1.0, -1.0])
    values += [m(v)]
    values += check(net.Tanh(m), v, [1.0, -1.0])
    out = io.StringIO()
    with redirect_stdout(out):
        for x in values:
            print(x)
    want = """[ 118.  134.]
[[-1. -1. -1. -1. -1. -1.]
 [ 6. -6.  9. -9.  1. -1.]]
[[  82.   93.]
 [ 110.  125.]]
[[-0.5 -0.5  1.   1.   0.5  0.5]
 [-4.   4.  -5.   5.  -0.5  0.5]]
[-1.2  0.4]
[-0.12  0.4 ]
[[ 117.76  150.88  -12.8   -16.4    -6.4    -8.2 ]
 [   0.8    -8.     -0.74    7.4     0.1    -1.  ]]
[-1.2  0.4]
[-0.834  0.38 ]
[[ 76.532  96.794  -8.319 -10.521  -4.159  -5.261]
 [  2.44   -6.845  -2.257   6.332   0.305  -0.856]]"""
    assert _lines(out.getvalue()) == _lines(want)
    assert m.n_params == 12 and np.array_equal(m.get_params(), np.arange(12, dtype=np.float32))
    with pytest.raises(ValueError):
        m.evaluate([1.0, 2.0, 3.0])


def test_graph_py():
    """tests/graph.py vs graph.out: fan-out DAG, each node evaluated once, user Python layers."""
    from mannequin2_b200.basicnet import Bias, Function, Input, Tanh

    def PrintLayer(inner):
        return Function(inner, f=lambda a: ((print("Forward pass:", a), a)[1], lambda g: (g,)))

    def test(m, inps=[2., 3.]):
        inps = np.array(inps)
        val, bpp = m.evaluate(inps)
        print(val, np.asarray(bpp(np.ones(inps.shape))))

    out = io.StringIO()
    with redirect_stdout(out):
        a = Bias(PrintLayer(Input(2)))
        b = Tanh(a)
        test(a)
        test(b)
        test(a + b)
        test(a * a)
    want = """Forward pass: [ 2.  3.]
[ 2.  3.] [ 1.  1.]
Forward pass: [ 2.  3.]
[ 0.964  0.995] [ 0.071  0.01 ]
Forward pass: [ 2.  3.]
[ 2.964  3.995] [ 1.071  1.01 ]
Forward pass: [ 2.  3.]
[ 4.  9.] [ 4.  6.]"""
    assert _lines(out.getvalue()) == _lines(want)


def test_reshape_py():
    """tests/reshape.py vs reshape.out: Reshape, Multiply with broadcast Params, Clip, batch (1,1)."""
    from mannequin2_b200.basicnet import Function, Input, Params, Reshape

    def PrintLayer(inner):
        return Function(inner, f=lambda a: ((print("Layer output:\n%s\n" % a), a)[1],
                                            lambda g: (print("Gradient:\n%s\n" % g), g)[1:]))

    def Clip(inner, a, b):
        def clip(v):
            def bpp(g):
                g = np.array(g)
                g[np.logical_and(v < a, g < 0.0)] = 0.0
                g[np.logical_and(v > b, g > 0.0)] = 0.0
                return (g,)
            return np.clip(v, a, b), bpp
        return Function(inner, f=clip)

    out = io.StringIO()
    with redirect_stdout(out):
        a = PrintLayer(Input(3, 2))
        a = PrintLayer(Reshape(a, (2, 3)))
        a = PrintLayer(a * Params(3))
        a = PrintLayer(Clip(a, -2.5, 2.5))
        a = PrintLayer(Reshape(a, -1))
        a.load_params([1, 0, -1])
        val, bpp = a.evaluate(np.arange(6).reshape((1, 1, 3, 2)))
        print(np.asarray(bpp(np.ones(val.shape))))
    text = out.getvalue()
    assert "[[[ 0.   0.  -2.   2.5  0.  -2.5]]]" in text          # last Layer output block
    assert _lines(text)[-1] == "[ 0. 5. 7.]"
    grads = text.split("Gradient:")
    assert len(grads) == 6
    assert _lines(grads[3])[:2] == ["[[[[ 1. 1. 1.]", "[ 0. 1. 1.]]]]"]      # after Clip backward
    assert _lines(grads[4])[:2] == ["[[[[ 1. 0. -1.]", "[ 0. 0. -1.]]]]"]    # after Multiply backward


def test_running_mean_and_norm_py():
    """tests/running_mean.py and tests/running_norm.py (1e-10 against np.mean / np.var(ddof=1))."""
    from mannequin2_b200 import RunningMean, RunningNormalize
    m = RunningMean()
    data = np.arange(10) + 10
    for d in [data[:i] for i in range(1, len(data))]:
        assert np.abs(m(d[-1]) - np.mean(d)) < 1e-10
    rs = np.random.RandomState(0)
    m = RunningMean((7,), horizon=2)
    data = rs.rand(4, 7)
    np_data = []
    for i, d in enumerate(data):
        for _ in range(2 ** i):
            np_data.append(d)
        assert np.mean(np.abs(m(d) - np.mean(np_data, axis=0))) < 1e-10
    m = RunningMean((5,))
    data = rs.rand(4, 5)
    m.update(data[0], weight=0.2); m.update(data[0], weight=0.8); m.update(data[1])
    m.update(data[2], weight=1.9); m.update(data[2], weight=0.1); m.update(data[3])
    assert np.mean(np.abs(m.get() - np.mean(data[[0, 1, 2, 2, 3]], axis=0))) < 1e-10

    n = RunningNormalize()
    data = np.arange(10) + 10
    n.update(data[0])
    for d in [data[:i] for i in range(2, len(data))]:
        n.update(d[-1])
        assert np.abs(n.get_mean() - np.mean(d)) < 1e-10
        assert np.abs(n.get_var() - np.var(d, ddof=1)) < 1e-10
    n = RunningNormalize()
    data = np.random.RandomState(123).rand(10, 5)
    for d in [data[:i] for i in range(1, len(data))]:
        n.update(d[-1])
        want = (np.array([4.0, 5.0]) - np.mean(d)) / np.std(d, ddof=1)
        assert np.mean(np.abs(n.apply([4.0, 5.0]) - want)) < 1e-10
    n = RunningNormalize((5,))
    data = rs.rand(10, 5)
    for d in data:
        n.update(d)
    want = (data - np.mean(data, axis=0)) / np.std(data, axis=0, ddof=1)
    assert np.mean(np.abs(n.apply(data) - want)) < 1e-10
    n = RunningNormalize((5,), horizon=2)
    data = rs.rand(4, 5)
    np_data = []
    for i, d in enumerate(data):
        n.update(d)
        for _ in range(2 ** i):
            np_data.append(d)
    np_data = np.array(np_data)
    want = (data - np.mean(np_data, axis=0)) / np.std(np_data, axis=0, ddof=1)
    assert np.mean(np.abs(n.apply(data) - want)) < 1e-10
    # against the real reference's trace (tests/golden/norm.npz), incl. zeros until 2 samples
    g = load_golden("norm")
    n = RunningNormalize(horizon=2)
    pos, res = 0, []
    for s in g["norm_scalar_sizes"]:
        d = g["norm_scalar_in"][pos:pos + s]
        pos += s
        res.append(np.concatenate([n(d), [n.get_mean(), n.get_var()]]))
    assert np.allclose(np.concatenate(res), g["norm_scalar_out"], rtol=1e-12, atol=1e-12)
    n = RunningNormalize(horizon=10)
    got = np.array([n(v) for v in g["norm_single_in"]])
    assert got[0] == 0.0 and np.allclose(got, g["norm_single_out"], rtol=1e-12, atol=1e-12)


def test_initializer_py():
    """tests/initializer.py: normed_columns init gives unit output std; init=10 scales it."""
    from mannequin2_b200.basicnet import Affine, Input, Linear
    rs = np.random.RandomState(1)
    np.random.seed(2)

    def std_of(Model, dims, n=150):
        vals = []
        for _ in range(n):
            a, b = dims()
            v2, _ = Model(Input(a), b).evaluate(rs.randn(a))
            vals.append(np.asarray(v2).reshape(-1))
        return np.std(np.concatenate(vals))

    rand_dims = lambda: rs.randint(16, size=2) + 5
    assert 0.9 < std_of(Linear, lambda: (5, 20)) < 1.1
    assert 0.9 < std_of(Affine, rand_dims) < 1.1
    assert 9.0 < std_of(lambda *p: Linear(*p, init=10.0), rand_dims) < 11.0


def test_predictor_py():
    """tests/predictor1.py + predictor2.py: 64-64 Tanh MLP + Adam(h=10) regression end to end."""
    from mannequin2_b200 import Adam
    from mannequin2_b200.basicnet import Affine, Input, Tanh
    np.random.seed(5)

    def SimplePredictor(in_size, out_size, lr):
        model = Input(in_size)
        for _ in range(2):
            model = Tanh(Affine(model, 64))
        model = Affine(model, out_size, init=0.1)
        opt = Adam(model.get_params(), horizon=10, lr=lr)

        def sgd_step(inps, lbls):
            outs, backprop = model.evaluate(inps)
            opt.apply_gradient(backprop(lbls - outs))
            model.load_params(opt.get_value())
        model.sgd_step = sgd_step
        return model

    pred = SimplePredictor(1, 1, 0.01)
    for _ in range(100):
        x = np.random.randn(128, 1) * 2.0
        pred.sgd_step(x, np.sin(x))
    assert pred([1.0]).shape == (1,)
    x = np.linspace(-5.0, 5.0, 101).reshape(-1, 1)
    y = pred(x)
    assert y.shape == x.shape
    assert np.mean(np.abs(y - np.sin(x))) < 0.2

    func = lambda x, y: [np.exp(x), x * y]
    np.random.seed(5)
    pred = SimplePredictor(2, 2, 0.015)
    for _ in range(500):
        xy = np.random.randn(256).reshape(-1, 2)
        pred.sgd_step(xy, np.array([func(a, b) for a, b in xy]))
    assert pred(np.eye(2)).shape == (2, 2)
    errors = []
    for a in np.linspace(-2.0, 2.0, 21):
        for b in np.linspace(-2.0, 2.0, 21):
            p = pred([a, b])
            assert p.shape == (2,)
            errors.append(np.abs(p - func(a, b)))
    # The reference's "< 0.2" bound is seed dependent (it fails for the reference itself with
    # this seed).  Stronger: the REAL reference, run in the build container with
    # np.random.seed(5) and this exact loop, ends at mean errors [0.21544665, 0.10674201];
    # 500 optimiser steps on the GPU must land on the same numbers.
    assert np.allclose(np.mean(errors, axis=0), [0.21544665, 0.10674201], atol=2e-4)


def test_nets_against_reference_fixtures():
    """evaluate/backprop of the layer API against outputs of the REAL reference (golden/nets.npz)."""
    import mannequin2_b200.basicnet as bn
    g = load_golden("nets")
    m = bn.Input(4, 5)
    for _ in range(2):
        m = bn.LReLU(bn.Affine(m, 16))
    m = bn.Affine(m, 5, init=0.1)
    m.load_params(g["lrelu_theta"])
    y, bpp = m.evaluate(g["lrelu_x"])
    assert_close(y, g["lrelu_y"], rtol=2e-5)
    assert_close(np.asarray(bpp(g["lrelu_dy"])), g["lrelu_grad"], rtol=2e-5, atol_scale=2e-6)
    assert_close(m.prerequisites[0].last_gradient, g["lrelu_dx"], rtol=2e-5, atol_scale=2e-6)
    p = bn.Input(11)
    for _ in range(2):
        p = bn.Tanh(bn.Affine(p, 64))
    p = bn.Affine(p, 3)
    p.load_params(g["tanh_theta"])
    y, bpp = p.evaluate(g["tanh_x"])
    assert_close(y, g["tanh_y"], rtol=2e-5, atol_scale=2e-6)
    assert_close(np.asarray(bpp(g["tanh_dy"])), g["tanh_grad"], rtol=2e-5, atol_scale=2e-6)
    y1, bpp1 = p.evaluate(g["tanh_x"][0])                       # unbatched: no division
    assert_close(y1, g["tanh_y1"], rtol=2e-5, atol_scale=2e-6)
    assert_close(np.asarray(bpp1(g["tanh_dy"][0])), g["tanh_grad1"], rtol=2e-5, atol_scale=2e-6)
    # a net too wide for the fused kernels takes the layered path: same answers as the oracle
    from oracle import mq_oracle as O
    np.random.seed(3)
    w = bn.Input(28, 28)
    for _ in range(2):
        w = bn.LReLU(bn.Affine(w, 128))
    w = bn.Affine(w, 10, init=0.1)
    assert w.n_params == 118282
    rs = np.random.RandomState(4)
    x, dy = rs.rand(128, 28, 28).astype(np.float32), rs.randn(128, 10).astype(np.float32)
    y, bpp = w.evaluate(x)
    spec = O.mlp_spec(784, [128, 128, 10], [O.ACT_LRELU, O.ACT_LRELU, O.ACT_NONE])
    theta = w.get_params()
    outs = O.mlp_forward(theta, spec, x.reshape(128, 784))
    assert_close(y, outs[-1], rtol=2e-5, atol_scale=2e-6)
    assert_close(np.asarray(bpp(dy)), O.mlp_backward(theta, spec, x.reshape(128, 784), outs, dy), rtol=2e-5,
                 atol_scale=2e-6)


def test_gauss_policy_against_reference_fixture():
    """Gauss(mean=net).logprob / .sample against the real reference (golden/gauss.npz)."""
    import mannequin2_b200.basicnet as bn
    from mannequin2_b200.distrib import Gauss
    g = load_golden("gauss")
    p = bn.Input(11)
    for _ in range(2):
        p = bn.Tanh(bn.Affine(p, 64))
    p = bn.Affine(p, 3, init=0.1)
    pol = Gauss(mean=p)
    assert pol.n_params == 5126
    pol.load_params(g["gauss_theta"])
    assert np.array_equal(pol.get_params(), g["gauss_theta"])
    logp, bpp = pol.logprob.evaluate(g["gauss_obs"], sample=g["gauss_act"])
    assert_close(logp, g["gauss_logp"], rtol=1e-5, atol_scale=1e-6)
    assert_close(np.asarray(bpp(g["gauss_g"])), g["gauss_grad"], rtol=3e-5, atol_scale=3e-6)
    # general graph path (no plan) gives the same numbers: evaluate logprob of a single sample
    l1, b1 = pol.logprob.evaluate(g["gauss_obs"][0], sample=g["gauss_act"][0])
    assert abs(float(l1) - g["gauss_logp"][0]) < 1e-4
    smp, bpp = pol.sample.evaluate(g["gauss_obs"][:5])
    mu = pol.mean(g["gauss_obs"][:5])
    noise = np.asarray(smp) - mu
    assert noise.shape == (5, 3) and np.abs(noise).max() > 0
    gs = g["gauss_sample_g"]
    grad = np.asarray(bpp(gs))
    assert_close(grad[-3:], (gs * noise).sum(axis=0) / 5, rtol=1e-4, atol_scale=1e-5)


def test_discrete_policy_against_reference_fixture():
    """Discrete(logits=net).logprob / .sample against the real reference (golden/discrete.npz): the fused plan,
    the node-by-node graph path (single sample) and the sampler."""
    import mannequin2_b200.basicnet as bn
    from mannequin2_b200.distrib import Discrete
    g = load_golden("discrete")
    p = bn.Input(4)
    for _ in range(2):
        p = bn.Tanh(bn.Affine(p, 64))
    p = bn.Affine(p, 3, init=0.1)
    pol = Discrete(logits=p)
    assert pol.n_params == len(g["disc_theta"])
    pol.load_params(g["disc_theta"])
    logp, bpp = pol.logprob.evaluate(g["disc_obs"], sample=g["disc_act"])
    assert_close(logp, g["disc_logp"], rtol=2e-5, atol_scale=1e-6)
    assert_close(np.asarray(bpp(g["disc_g"])), g["disc_grad"], rtol=3e-5, atol_scale=3e-6)
    l1, _ = pol.logprob.evaluate(g["disc_obs"][0], sample=g["disc_act"][0])
    assert abs(float(l1) - g["disc_logp"][0]) < 1e-4
    draws = [pol.sample(g["disc_obs"][0]) for _ in range(20)]
    assert all(0 <= int(a) < 3 for a in draws)


def test_large_batch_ppo_update_against_oracle():
    """The measured path end to end: PPOUpdater.update on 16,384 transitions with 8,192-row minibatches (tcgen05
    gradient kernel, fused reduce / Adam kernel, programmatic dependent launch, value and policy chains on two streams)
    against the oracle's restatement of benchmarks/ppo.py:22-40 + gae.py:34-75 with the same index tables."""
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    from bench import ppo_rollout
    from oracle import mq_oracle as O
    from mannequin2_b200.ppo import PPOUpdater, init_mlp
    rs = np.random.RandomState(41)
    n, mb, epochs = 16384, 8192, 2
    o, a, r, d = ppo_rollout(rs, n, p_done=1.0 / 300, cut=700)
    vi = np.concatenate([rs.permutation(n).astype(np.int32).reshape(n // mb, mb) for _ in range(epochs)])
    pi = np.concatenate([rs.permutation(n).astype(np.int32).reshape(n // mb, mb) for _ in range(epochs)])
    np.random.seed(5)
    pol0 = np.concatenate([init_mlp(11, [64, 64, 3], 0.1), np.zeros(3, np.float32)])
    val0 = init_mlp(11, [64, 64, 1], 0.1)
    upd = PPOUpdater(11, 3, policy_params=pol0.copy(), value_params=val0.copy(), state_dtype="float64")
    pol, val = upd.update(o, a, r, d, vi, pi)
    pol_spec, val_spec = O.policy_spec(11, 3), O.mlp_spec(11, [64, 64, 1], [O.ACT_TANH, O.ACT_TANH, O.ACT_NONE])
    want_pol, want_val, _ = O.ppo_update(pol0.copy(), pol_spec, O.AdamO(pol0, horizon=10), val0.copy(), val_spec,
                                         O.AdamO(val0, horizon=10, lr=0.003), O.RunningNormalizeO(horizon=2),
                                         o, a, r, d.astype(bool), vi, pi)
    steps = len(vi)
    assert np.max(np.abs(pol - pol0)) > 1e-4 and np.max(np.abs(val - val0)) > 1e-4
    assert np.max(np.abs(pol - want_pol)) < 2e-6 * (1 + steps) + 1e-6, np.max(np.abs(pol - want_pol))
    assert np.max(np.abs(val - want_val)) < 2e-5 * (1 + steps), np.max(np.abs(val - want_val))


# Tolerance of a multi-step update against the fp64 oracle.  One Adam step moves a parameter by lr * m / (eps + sqrt(v)),
# a quantity of size ~lr whatever the gradient's scale, so a gradient error e_g on a coordinate of size |g| perturbs that
# coordinate's step by ~lr * e_g / |g|: coordinates whose gradient is at the rounding level move by a sizeable fraction of
# lr in an arbitrary direction in ANY implementation (fp32 FFMA included).  The bound used below is a fraction `frac` of
# a full step per optimiser step: err <= frac * lr * (1 + steps).  frac = 0.7 % is what the fp32 paths meet (the bar of
# test_large_batch_ppo_update_against_oracle: 2e-6 per step at lr 3e-4); the reduced-precision tensor-core modes are
# held to their stated gradient tolerance ratio (bf16x2: 10x).
UPDATE_FRAC = {"exact": 0.007, "fp16x2": 0.007, "bf16x2": 0.07}


@pytest.mark.parametrize("tc_mode", ["exact", "fp16x2", "bf16x2"])
def test_bench_configuration_update_against_oracle(tc_mode):
    """The BENCHMARKED configuration (bench.py default workload, BASELINE configs[2]) against the oracle: 2^20
    transitions, 65,536-row permutation minibatches, DEFAULT fp32 Adam state, one epoch (16 + 16 optimiser steps) through
    PPOUpdater.update_dev -- value forward, GAE, normaliser, old log-probs, packed records, record-fed tcgen05 gradient
    kernel, fused reduce / Adam / operand-image kernel -- vs O.ppo_update (benchmarks/ppo.py:22-40 + gae.py:34-75) on
    the same rollout and index tables."""
    import torch
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    from bench import ppo_rollout
    from oracle import mq_oracle as O
    from mannequin2_b200 import _device as D
    from mannequin2_b200.ppo import PPOUpdater, init_mlp
    rs = np.random.RandomState(3)
    n, mb = 1 << 20, 65536
    o, a, r, d = ppo_rollout(rs, n, p_done=1.0 / 500, cut=1000)
    vi = rs.permutation(n).astype(np.int32).reshape(n // mb, mb)
    pi = rs.permutation(n).astype(np.int32).reshape(n // mb, mb)
    np.random.seed(1)
    pol0 = np.concatenate([init_mlp(11, [64, 64, 3], 0.1), np.zeros(3, np.float32)])
    val0 = init_mlp(11, [64, 64, 1], 0.1)
    upd = PPOUpdater(11, 3, policy_params=pol0.copy(), value_params=val0.copy(), tc_mode=tc_mode)
    assert upd.policy.opt._tdtype == torch.float32                      # the bench's state dtype
    adv, ret = upd.update_dev(*[D.from_host(x) for x in (o, a, r, d, vi, pi)])
    torch.cuda.synchronize()
    assert upd.policy.records is not None and upd.value.records is not None     # the record-fed tensor-core path ran
    pol, val = D.to_host(upd.policy.theta), D.to_host(upd.value.theta)
    pol_spec, val_spec = O.policy_spec(11, 3), O.mlp_spec(11, [64, 64, 1], [O.ACT_TANH, O.ACT_TANH, O.ACT_NONE])
    want_pol, want_val, aux = O.ppo_update(pol0.copy(), pol_spec, O.AdamO(pol0, horizon=10), val0.copy(), val_spec,
                                           O.AdamO(val0, horizon=10, lr=0.003), O.RunningNormalizeO(horizon=2),
                                           o, a, r, d.astype(bool), vi, pi)
    steps = len(vi)
    assert_close(D.to_host(adv), aux["adv"], rtol=1e-4, atol_scale=1e-5)
    e_pol, e_val = float(np.max(np.abs(pol - want_pol))), float(np.max(np.abs(val - want_val)))
    print("bench-config update vs oracle (%s): max|d policy| %.2e, max|d value| %.2e after %d + %d steps"
          % (tc_mode, e_pol, e_val, steps, steps))
    assert np.max(np.abs(pol - pol0)) > 1e-4 and np.max(np.abs(val - val0)) > 1e-4
    frac = UPDATE_FRAC[tc_mode]
    assert e_pol < frac * 3e-4 * (1 + steps), e_pol
    assert e_val < frac * 3e-3 * (1 + steps), e_val


def test_independent_learners_on_concurrent_streams():
    """Independent PPOUpdater instances driven on different CUDA streams share no scratch: each of three learners
    running concurrently (both the single-CTA chain path and the large-batch path) ends bit-identical to the same
    learner run alone."""
    import torch
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    from bench import ppo_rollout
    from mannequin2_b200 import _device as D
    from mannequin2_b200.ppo import PPOUpdater, init_mlp
    for n, mb, steps in ((2048, 64, 40), (16384, 8192, 4)):
        rs = np.random.RandomState(n)
        data, starts = [], []
        for _ in range(3):
            o, a, r, d = ppo_rollout(rs, n)
            vi = rs.randint(n, size=(steps, mb)).astype(np.int32)
            pi = rs.randint(n, size=(steps, mb)).astype(np.int32)
            data.append(tuple(D.from_host(x) for x in (o, a, r, d, vi, pi)))
            np.random.seed(len(data))
            starts.append((np.concatenate([init_mlp(11, [64, 64, 3], 0.1), np.zeros(3, np.float32)]),
                           init_mlp(11, [64, 64, 1], 0.1)))

        def learner(k):
            return PPOUpdater(11, 3, policy_params=starts[k][0].copy(), value_params=starts[k][1].copy())
        alone = []
        for k in range(3):
            u = learner(k)
            for _ in range(2):
                u.update_dev(*data[k])
            torch.cuda.synchronize()
            alone.append((u.policy.theta.cpu().numpy().copy(), u.value.theta.cpu().numpy().copy()))
        upds, streams = [learner(k) for k in range(3)], [torch.cuda.Stream() for _ in range(3)]
        torch.cuda.synchronize()
        for _ in range(2):
            for k in range(3):
                with torch.cuda.stream(streams[k]):
                    upds[k].update_dev(*data[k])
        torch.cuda.synchronize()
        for k in range(3):
            assert np.array_equal(upds[k].policy.theta.cpu().numpy(), alone[k][0]), (n, k)
            assert np.array_equal(upds[k].value.theta.cpu().numpy(), alone[k][1]), (n, k)


def test_gauss_entropy_function():
    """Gauss.entropy (an extension -- the reference's Gauss exports no entropy, distrib.py:70-76):
    H = sum(logstd) + A/2 (1 + ln 2 pi) as a Function of the logstd sub-graph, for a per-row logstd network and for
    the shared Params(A) logstd; the backward pass hands g to every logstd entry."""
    from mannequin2_b200.basicnet import Input, Affine, Tanh
    from mannequin2_b200.distrib import Gauss
    np.random.seed(12)
    rs = np.random.RandomState(13)
    x = Input(6)
    mean = Affine(Tanh(Affine(x, 16)), 4)
    lsn = Affine(Tanh(Affine(x, 8)), 4)
    pol = Gauss(mean=mean, logstd=lsn)
    obs = rs.randn(9, 6).astype(np.float32)
    h, bpp = pol.entropy.evaluate(obs)
    const = 0.5 * 4 * (1.0 + np.log(2.0 * np.pi))
    assert_close(h, np.sum(np.asarray(lsn(obs), dtype=np.float64), axis=-1) + const, rtol=1e-5, atol_scale=1e-6)
    assert pol.entropy.n_params == lsn.n_params
    g = rs.randn(9).astype(np.float32)
    _, b2 = lsn.evaluate(obs)
    assert_close(np.asarray(bpp(g)), np.asarray(b2(np.repeat(g[:, None], 4, axis=1))), rtol=1e-5, atol_scale=1e-6)
    pol2 = Gauss(mean=mean)                                   # shared logstd = the last 4 parameters
    theta = np.array(pol2.get_params())
    theta[-4:] = [0.3, -0.2, 0.1, 0.05]
    pol2.load_params(theta)
    h2, bpp2 = pol2.entropy.evaluate(obs[:1])
    assert abs(float(np.asarray(h2).reshape(-1)[0]) - (0.25 + const)) < 1e-5
    assert np.allclose(np.asarray(bpp2(np.float32(1.5))), 1.5)


def test_batched_actor_rollouts_and_closed_loop():
    """Rollout side (mannequin/gym.py:61-84,115-137 for many environments in lock step):
    (1) collecting through one CUDA graph per chunk gives bit-identical chunks, normaliser statistics and environment
    states as launching the same kernels one by one, and successive chunks differ (fresh noise per replay);
    (2) the first observation of a chunk is the carried-over last observation of the previous one (gae.py:64);
    (3) the closed loop collect -> PPOUpdater.update_rollout learns on the synthetic environment, in the spirit of the
    reference's convergence tests (tests/predictor1.py)."""
    import torch
    from mannequin2_b200 import _device as D
    from mannequin2_b200.ppo import PPOUpdater
    from mannequin2_b200.rollout import BatchedActor, SyntheticEnv
    n_env, t_len = 256, 16
    chunks = {}
    for graph in (False, True):
        np.random.seed(2)
        env = SyntheticEnv(n_env, horizon=10, p_end=0.02, seed=3)
        upd = PPOUpdater(env.n_obs, env.n_act)
        actor = BatchedActor(env, upd.policy, seed=5, use_graph=graph)
        got = []
        for _ in range(3):
            c = actor.collect(t_len)
            torch.cuda.synchronize()
            got.append({k: D.to_host(v).copy() for k, v in c.items()})
        chunks[graph] = (got, D.to_host(actor.normalizer._state).copy(), D.to_host(env.x).copy())
    for a, b in zip(chunks[False][0], chunks[True][0]):
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(chunks[False][1], chunks[True][1]) and np.array_equal(chunks[False][2], chunks[True][2])
    c0, c1, c2 = chunks[True][0]
    assert not np.array_equal(c0["act"], c1["act"]) and not np.array_equal(c1["act"], c2["act"])
    first_rows = np.arange(n_env) * t_len
    assert np.array_equal(c1["obs"][first_rows], c0["last_obs"]) and np.array_equal(c2["obs"][first_rows], c1["last_obs"])
    assert np.max(np.abs(c2["obs"])) <= 5.0 and c0["done"].any() and not c0["done"].all()
    # closed loop
    np.random.seed(0)
    rs = np.random.RandomState(1)
    n_env, t_len, mb, epochs = 512, 64, 8192, 4
    n = n_env * t_len
    env = SyntheticEnv(n_env, horizon=64, seed=3)
    upd = PPOUpdater(env.n_obs, env.n_act)
    actor = BatchedActor(env, upd.policy, seed=5, use_graph=True)
    rewards = []
    for _ in range(12):
        chunk = actor.collect(t_len)
        rewards.append(float(chunk["rew"].mean()))
        vi = D.from_host(np.concatenate([rs.permutation(n).astype(np.int32).reshape(n // mb, mb) for _ in range(epochs)]))
        pi = D.from_host(np.concatenate([rs.permutation(n).astype(np.int32).reshape(n // mb, mb) for _ in range(epochs)]))
        upd.update_rollout(chunk, n_env, t_len, vi, pi)
    assert np.mean(rewards[-3:]) > rewards[0] + 0.04, rewards
    assert np.all(D.to_host(upd.policy.theta)[-3:] < 0.0)          # the policy narrows its noise


def test_mlp_trainer_matches_oracle():
    """mnist/mlp.py:24-43 through MLPTrainer: `sgd_step(host batch, host labels)` (pinned staging + one graph replay per
    step) and `train_block` (a block of steps as one graph; split-K forward, bias gradients inside the dW products,
    parameter-gradient products overlapped on a side stream) both follow the oracle's restatement step for step."""
    from oracle import mq_oracle as O
    from mannequin2_b200.mlp import MLPTrainer
    from mannequin2_b200.ppo import init_mlp
    from mannequin2_b200 import _device as D
    rs = np.random.RandomState(8)
    n, steps, B = 600, 7, 128
    x = rs.rand(n, 784).astype(np.float32)
    y = np.eye(10, dtype=np.float32)[rs.randint(10, size=n)]
    tab = rs.randint(n, size=(steps, B)).astype(np.int32)
    np.random.seed(4)
    theta0 = init_mlp(784, [128, 128, 10], 0.1)
    spec = O.mlp_spec(784, [128, 128, 10], [O.ACT_LRELU, O.ACT_LRELU, O.ACT_NONE])
    opt = O.AdamO(theta0, horizon=10, lr=1e-3)
    want = theta0.copy()
    for s_ in range(steps):
        want = O.mlp_sgd_step(want, spec, opt, x[tab[s_]], y[tab[s_]])
    a = MLPTrainer(params=theta0.copy(), graph_steps=steps)
    for s_ in range(steps):
        a.sgd_step(x[tab[s_]].reshape(B, 28, 28), y[tab[s_]])
    b = MLPTrainer(params=theta0.copy(), graph_steps=steps)
    b.train_block(D.from_host(x), D.from_host(y), tab)
    for got in (a.get_params(), b.get_params()):
        assert np.max(np.abs(got - theta0)) > 1e-3
        assert np.max(np.abs(got - want)) < 2e-6 * (1 + steps) + 1e-6, np.max(np.abs(got - want))
    assert np.max(np.abs(a.get_params() - b.get_params())) < 2e-6 * (1 + steps)


def test_vae_device_step_and_graph_replay():
    """VAETrainer.step_dev / step_graph (the whole mnist/vae.py:57-71 iteration without a host round trip, replayed as
    one CUDA graph, sampling noise from the device Philox generator) follow the host-API `step` exactly: same
    parameters after three iterations, and the noise differs from replay to replay."""
    import torch
    from mannequin2_b200 import _device as D
    from mannequin2_b200.vae import VAETrainer
    rs = np.random.RandomState(21)
    xs = [rs.rand(32, 6, 5).astype(np.float32) for _ in range(3)]
    finals, stats = [], []
    for mode in ("host", "graph"):
        np.random.seed(9)
        tr = VAETrainer(image_shape=(6, 5), hidden=16, latent=3, n_hidden=1, rng=D.DeviceRNG(77))
        for x in xs:
            if mode == "host":
                lp, kl = tr.step(x)
            else:
                lp, kl = (float(v) for v in tr.step_graph(D.from_host(x)))
            stats.append((lp, kl))
        finals.append(np.array(tr.decoder.get_params()))
    assert np.max(np.abs(finals[0] - finals[1])) < 1e-6, np.max(np.abs(finals[0] - finals[1]))
    for a, b in zip(stats[:3], stats[3:]):
        assert abs(a[0] - b[0]) < 1e-4 * (1 + abs(a[0])) and abs(a[1] - b[1]) < 1e-4 * (1 + abs(a[1]))
    rng = D.DeviceRNG(5)
    z1, z2 = D.to_host(rng.randn_dev((4, 3))), D.to_host(rng.randn_dev((4, 3)))
    assert not np.array_equal(z1, z2) and abs(z1.mean()) < 2.0


class _FixedNoise(object):
    """A device generator that replays one recorded N(0,1) draw (the latent sample of mnist/vae.py)."""

    def __init__(self, unit):
        from mannequin2_b200 import _device as D
        self.unit = D.from_host(np.asarray(unit, dtype=np.float32))

    def randn_dev(self, shape):
        assert tuple(shape) == tuple(self.unit.shape)
        return self.unit.clone()


def test_vae_fused_step_against_reference_run():
    """The fused plan of csrc/mq_vae.cu (mq_vae_grad / mq_vae_step: one forward and one backward pass, every product on
    the TMA-fed tcgen05 GEMM) against ONE iteration of mnist/vae.py:57-69 run from the reference's own classes
    (golden/vae_step.npz): log-probs, DKLs, grad1 + grad2, and the parameters after the clipped-momentum update.  The
    fixture's sizes are deliberately awkward for the plan: 12 pixels (head columns padded to 16), 3 latents (padded to
    8), batch 6 (one 128-row tile, one 64-wide K block), two hidden layers per side, log-deviations on both sides of
    Clip's bound.  Operand formats fp16x2 and exact share the fp32 bar (rtol 2e-5 of the largest entry)."""
    from mannequin2_b200 import _device as D
    from mannequin2_b200.vae import VAETrainer
    g = load_golden("vae_step")
    d_img, hidden, latent, n_hidden, batch = (int(v) for v in g["vstep_dims"])
    x = D.from_host(g["vstep_x"].reshape(batch, d_img).astype(np.float32))
    for mode in ("fp16x2", "exact"):
        np.random.seed(1)
        tr = VAETrainer(image_shape=(4, 3), hidden=hidden, latent=latent, n_hidden=n_hidden, rng=_FixedNoise(g["vstep_unit"]),
                        gemm_mode=mode)
        assert tr.fused_supported() and tr.n_params == len(g["vstep_theta"])
        tr.decoder.load_params(g["vstep_theta"])
        grad, lp, kl = tr.grad_fused(x, tr._rng.unit)
        assert_close(D.to_host(lp), g["vstep_logprob"], rtol=2e-5, atol_scale=2e-6, msg=mode)
        assert_close(D.to_host(kl), g["vstep_dkl"], rtol=2e-5, atol_scale=2e-6, msg=mode)
        assert_close(D.to_host(grad), g["vstep_grad"], rtol=2e-5, atol_scale=2e-5, msg=mode)
        tr.momentum = D.from_host(g["vstep_momentum"].astype(np.float32))
        tr.step_fused(x)
        assert np.max(np.abs(D.to_host(tr.momentum) - g["vstep_momentum_new"])) < 5e-5, mode
        assert np.max(np.abs(np.asarray(tr.decoder.get_params()) - g["vstep_theta_new"])) < 2e-7, mode


def test_vae_fused_step_at_baseline_shape():
    """BASELINE configs[4] (784-400-(20+20)-400-(784+784), batch 4096): the fused plan against the fp64 oracle's
    two-pass restatement of vae.py:57-66 (O.vae_grads) AND against the node-by-node device graph with the same latent
    noise; then three iterations of `step_graph` on the fused plan against three of the node-by-node `step_dev`.  The
    log-deviation heads are shifted so that part of the batch sits on Clip's bounds."""
    import torch
    from oracle import mq_oracle as O
    from mannequin2_b200 import _device as D
    from mannequin2_b200.vae import VAETrainer
    rs = np.random.RandomState(5)
    B, D_IMG, H, L = 4096, 784, 400, 20
    x_host = rs.rand(B, 28, 28).astype(np.float32)
    unit = rs.randn(B, L).astype(np.float32)
    trainers = []
    for k in range(2):
        np.random.seed(11)
        tr = VAETrainer(image_shape=(28, 28), hidden=H, latent=L, n_hidden=1, rng=_FixedNoise(unit))
        theta = np.asarray(tr.decoder.get_params(), dtype=np.float32).copy()
        prs = np.random.RandomState(12)
        theta += (prs.randn(len(theta)) * 0.02).astype(np.float32)
        tr.decoder.load_params(theta)
        trainers.append(tr)
    fused, graph = trainers
    x = D.from_host(x_host)
    grad, lp, kl = fused.grad_fused(x.reshape(B, D_IMG), fused._rng.unit)
    o_lp, o_kl, g1, g2 = O.vae_grads(theta, D_IMG, H, L, 1, x_host.reshape(B, D_IMG), unit)
    want = g1.copy()
    want[:len(g2)] += g2
    assert_close(D.to_host(lp), o_lp, rtol=2e-5, atol_scale=2e-6)
    assert_close(D.to_host(kl), o_kl, rtol=2e-5, atol_scale=2e-6)
    assert_close(D.to_host(grad), want, rtol=2e-5, atol_scale=1e-5)
    # the node-by-node graph (two evaluations, vae.py:60-66) with the same noise
    from mannequin2_b200.backprop import gemm_mode
    ones = torch.ones(B, dtype=torch.float32, device=x.device)
    with gemm_mode("fp16x2"):
        _, bp = graph.decoder.logprob.evaluate_dev(x, sample=x)
        n1 = D.to_host(D.as_dev_f32(bp(ones))).astype(np.float64)
        _, bp = graph.dkl.evaluate_dev(x)
        n2 = D.to_host(D.as_dev_f32(bp(-ones))).astype(np.float64)
    n1[:len(n2)] += n2
    assert_close(D.to_host(grad), n1, rtol=2e-5, atol_scale=1e-5)
    # three iterations: CUDA-graph replay of the fused plan against the node-by-node step
    for tr in trainers:
        tr.decoder.load_params(theta)
    for _ in range(3):
        f_stats = [float(v) for v in fused.step_graph(x)]
        g_stats = [float(v) for v in graph.step_dev(x)]
        assert abs(f_stats[0] - g_stats[0]) < 1e-4 * (1 + abs(g_stats[0])) and abs(f_stats[1] - g_stats[1]) < 1e-4 * (1 + abs(g_stats[1]))
    a, b = np.asarray(fused.decoder.get_params(), np.float64), np.asarray(graph.decoder.get_params(), np.float64)
    assert np.max(np.abs(a - b)) < 1e-6, np.max(np.abs(a - b))
    assert np.max(np.abs(a - theta)) > 1e-5


def test_forward_from_packed_records_matches_oracle():
    """mq_fused_forward_rec: the forward-only form of the record-fed tensor-core kernel (value predictions, gae.py:43-45, and
    old log-probabilities, ppo.py:27, over a whole rollout) against the fp64 oracle and against the array-fed
    mq_fused_forward, in the fp16x2 and the three-plane operand formats; 70,001 rows (ragged last tile, several tiles per SM)
    and 9,000 rows (single wave), with and without a row gather."""
    import ctypes
    import torch
    from oracle import mq_oracle as O
    from mannequin2_b200 import _device as D, _lib as L
    lib = L.lib()
    rs = np.random.RandomState(31)
    n_obs, n_act = 11, 3
    for kind in ("policy", "value"):
        if kind == "policy":
            spec = O.policy_spec(n_obs, n_act)
            net = L.make_net(n_obs, [64, 64, n_act], [1, 1, 0], logstd=True)
        else:
            spec = O.mlp_spec(n_obs, [64, 64, 1], [1, 1, 0])
            net = L.make_net(n_obs, [64, 64, 1], [1, 1, 0])
        theta = (rs.randn(spec["n_params"]) * 0.3).astype(np.float32)
        nb = ctypes.byref(net)
        rec_w = int(lib.mq_record_width(nb))
        for rows, gather in ((70001, False), (9000, True)):
            obs = rs.randn(rows, n_obs).astype(np.float32)
            act = rs.randn(rows, n_act).astype(np.float32)
            idx = rs.randint(rows, size=rows).astype(np.int32) if gather else None
            sel = idx if gather else slice(None)
            if kind == "policy":
                want = O.policy_logprob(theta, spec, obs[sel], act[sel])[0]
            else:
                want = O.mlp_forward(theta, spec, obs[sel])[-1].reshape(-1)
            d_theta, d_obs, d_act = D.from_host(theta), D.from_host(obs), D.from_host(act)
            d_idx = D.from_host(idx) if gather else None
            rec = torch.empty(rows * rec_w, dtype=torch.float32, device=D.device())
            D.call("mq_pack_records", nb, D.ptr(d_obs), D.ptr(d_act) if kind == "policy" else None, None, None, rows, D.ptr(rec),
                   D.stream())
            for mode in ("fp16x2", "exact"):
                tc = L.TC_MODES[mode]
                img = torch.zeros(int(lib.mq_tc_image_bytes(nb, tc)), dtype=torch.uint8, device=D.device())
                assert img.numel() > 0
                D.call("mq_tc_image_build", nb, D.ptr(d_theta), tc, D.ptr(img), D.stream())
                got = torch.full((rows,), float("nan"), dtype=torch.float32, device=D.device())
                D.call("mq_fused_forward_rec", nb, D.ptr(d_theta), D.ptr(rec), rec_w, D.ptr(img), D.ptr(d_idx),
                       rows, D.ptr(got) if kind == "value" else None, D.ptr(got) if kind == "policy" else None, tc, D.stream())
                assert_close(D.to_host(got), want, rtol=1e-5, atol_scale=2e-6, msg="%s %s rows=%d" % (kind, mode, rows))
            if not gather:
                ref = torch.empty(rows, dtype=torch.float32, device=D.device())
                D.call("mq_fused_forward", nb, D.ptr(d_theta), D.ptr(d_obs), None, rows, D.ptr(ref) if kind == "value" else None,
                       D.ptr(d_act) if kind == "policy" else None, D.ptr(ref) if kind == "policy" else None, D.stream())
                assert_close(D.to_host(got), D.to_host(ref), rtol=1e-5, atol_scale=2e-6, msg="%s vs array-fed forward" % kind)


def test_parameter_arena_releases_dropped_nets():
    """Params live in 16 MB arena chunks so that a net's parameters are one flat slice; a full chunk must go away with the
    last Params created in it (nets built and dropped in a loop must not accumulate device memory)."""
    import gc
    import weakref
    from mannequin2_b200 import basicnet as bn
    arena = bn._arena()
    big = (arena.CHUNK * 3) // 4
    held = bn.Params(big)                               # forces every later allocation of this size into a new chunk
    before = len(arena.chunks)
    for _ in range(3):
        p = bn.Params(big)
        assert float(p.dev.sum()) == 0.0
        del p
    gc.collect()
    full = [c[0] for c in arena.chunks[before - 1:-1]]
    assert len(full) >= 2 and all(isinstance(r, weakref.ref) for r in full)
    dead = sum(1 for r in full if r() is None)
    assert dead >= len(full) - 1                        # every dropped chunk is gone; the one `held` lives in may remain
    assert float(held.dev.sum()) == 0.0 and held.get_params().shape == (big,)


def test_reinforce_script_loop_and_reward_log():
    """benchmarks/policy.py:29-47 (plain REINFORCE: episode -> Trajectory.discounted -> RunningNormalize -> tanh ->
    logprob backprop -> Adam) written against `mannequin` / `mannequin.gym`, on a small NumPy environment wrapped in
    NormalizedObservations and PrintRewards; the reward log has the format of benchmarks/_env.py:60-63."""
    import mannequin2_b200.compat  # noqa: F401
    from mannequin import Adam, RunningNormalize
    from mannequin.basicnet import Input, Affine, Tanh
    from mannequin.distrib import Gauss
    from mannequin.gym import NormalizedObservations, PrintRewards, episode, log_line

    class PointMass(object):                     # reach the origin: reward -|x|^2, 25 steps per episode
        def __init__(self):
            self.rs, self.t = np.random.RandomState(0), 0

        def reset(self):
            self.x, self.t = self.rs.randn(2), 0
            return self.x * 3.0 + 1.0

        def step(self, action):
            self.x = self.x + 0.3 * np.clip(np.asarray(action, dtype=np.float64).reshape(2), -1, 1)
            self.t += 1
            return self.x * 3.0 + 1.0, -float(np.sum(self.x ** 2)), self.t >= 25, {}

    lines = []
    env = PrintRewards(NormalizedObservations(PointMass(), shape=(2,)), every=250,
                       print=lambda s_, r: lines.append(log_line(s_, r)))
    np.random.seed(1)
    policy = Input(2)
    for _ in range(2):
        policy = Tanh(Affine(policy, 64))
    policy = Gauss(mean=Affine(policy, 2, init=0.1))
    opt = Adam(policy.get_params(), horizon=10)
    normalize = RunningNormalize(horizon=2)
    start = np.array(policy.get_params())
    for _ in range(40):
        traj = episode(env, policy.sample)
        assert len(traj) == 25
        traj = traj.discounted(horizon=500)
        traj = traj.modified(rewards=normalize)
        traj = traj.modified(rewards=np.tanh)
        _, backprop = policy.logprob.evaluate(traj.o, sample=traj.a)
        opt.apply_gradient(backprop(traj.r), lr=0.001)
        policy.load_params(opt.get_value())
    assert np.max(np.abs(np.array(policy.get_params()) - start)) > 1e-3
    assert len(lines) == 4 and all(len(l.split()) == 2 and l.endswith("\n") for l in lines)
    assert lines[0].split()[0] == "250" and "." in lines[0].split()[1] and len(lines[0].split()[1].split(".")[1]) == 2
    assert np.all(np.abs(env.env.get_mean() - 1.0) < 3.0) and np.all(env.env.get_std() > 0.1)


def test_es_population_generations_against_oracle():
    """benchmarks/mutation.py:33-37 in population form through the ESPopulation driver: three generations
    (FFMA evaluation for 100 observations; tensor-core evaluation for 300) follow the oracle -- returns, the
    RunningNormalize(horizon 10) batch update, the mean of noise x normalised return, Adam(lr 0.01)."""
    from oracle import mq_oracle as O
    from mannequin2_b200 import _device as D
    from mannequin2_b200.es import ESPopulation
    from mannequin2_b200.ppo import init_mlp
    for n_obs_rows in (100, 300):
        rs = np.random.RandomState(17 + n_obs_rows)
        np.random.seed(6)
        theta0 = init_mlp(11, [64, 64, 3], 1.0)
        spec = O.mlp_spec(11, [64, 64, 3], [O.ACT_TANH, O.ACT_TANH, O.ACT_NONE])
        es = ESPopulation(11, 3, params=theta0.copy())
        opt, norm = O.AdamO(theta0, horizon=10, lr=0.01), O.RunningNormalizeO(horizon=10)
        want = theta0.copy()
        coef = rs.randn(3).astype(np.float32)
        for _ in range(3):
            noise = (rs.randn(48, len(theta0)) * 0.1).astype(np.float32)
            obs = np.clip(rs.randn(n_obs_rows, 11), -5, 5).astype(np.float32)
            want, rets, _ = O.es_generation(want, spec, opt, norm, noise, obs, coef)
            got_rets = es.generation(D.from_host(noise), D.from_host(obs), D.from_host(coef))
            assert_close(D.to_host(got_rets), rets, rtol=2e-4, atol_scale=2e-5)
        got = es.get_params().astype(np.float32)
        assert np.max(np.abs(got - theta0)) > 1e-3
        assert np.max(np.abs(got - want)) < 2e-4, np.max(np.abs(got - want))


def test_full_size_update_is_bit_reproducible():
    """BASELINE configs[2] at full size (2^20 transitions, 65,536-row minibatches; one epoch to keep it short): the
    update is a pure function of its inputs -- two learners started from the same parameters end with bit-identical
    policy and value parameters (fixed-order reductions everywhere: per-CTA partials, the look-back scan, the
    normaliser), advantages and returns included -- and the result is finite and has moved."""
    import torch
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    from bench import ppo_rollout
    from mannequin2_b200 import _device as D
    from mannequin2_b200.ppo import PPOUpdater, init_mlp
    rs = np.random.RandomState(3)
    n, mb = 1 << 20, 65536
    o, a, r, d = ppo_rollout(rs, n, p_done=1.0 / 500, cut=1000)
    vi = rs.permutation(n).astype(np.int32).reshape(n // mb, mb)
    pi = rs.permutation(n).astype(np.int32).reshape(n // mb, mb)
    dev = [D.from_host(x) for x in (o, a, r, d, vi, pi)]
    np.random.seed(1)
    pol0 = np.concatenate([init_mlp(11, [64, 64, 3], 0.1), np.zeros(3, np.float32)])
    val0 = init_mlp(11, [64, 64, 1], 0.1)
    results = []
    for _ in range(2):
        upd = PPOUpdater(11, 3, policy_params=pol0.copy(), value_params=val0.copy())
        adv, ret = upd.update_dev(*dev)
        torch.cuda.synchronize()
        results.append([D.to_host(t).copy() for t in (upd.policy.theta, upd.value.theta, adv, ret)])
    for x, y in zip(*results):
        assert np.array_equal(x, y)
    pol, val, adv, ret = results[0]
    assert np.all(np.isfinite(pol)) and np.all(np.isfinite(val)) and np.all(np.isfinite(adv))
    assert np.max(np.abs(pol - pol0)) > 1e-4 and np.max(np.abs(val - val0)) > 1e-4
    # episode segmentation at full size: the advantage at an episode end is r - v there, whatever follows
    ends = np.nonzero(d)[0][:2000]
    v_all = ret - adv
    assert np.max(np.abs(adv[ends] - (r[ends] - v_all[ends]))) < 1e-5


def test_vae_heads_against_reference_fixture():
    """DKLUninormal and Clip as device ops inside a two-headed encoder graph against mnist/vae.py:13-34 itself
    (golden/vae.npz, generated by importing the reference module): DKL values, the flat parameter gradient (shared
    trunk, fan-out accumulation, batch mean, Clip's gradient gate at the range boundary) and Clip alone."""
    from mannequin2_b200.basicnet import Input, Affine, Tanh
    from mannequin2_b200.vae import DKLUninormal, Clip
    g = load_golden("vae")
    np.random.seed(42)
    x = Input(5)
    h = Tanh(Affine(x, 8))
    mean = Affine(h, 3, init=0.1)
    logstd = Clip(Affine(h, 3), -6.0, 0.0)
    dkl = DKLUninormal(mean=mean, logstd=logstd)
    assert dkl.n_params == len(g["vae_theta"])
    dkl.load_params(g["vae_theta"])
    val, bpp = dkl.evaluate(g["vae_x"])
    assert_close(val, g["vae_dkl"], rtol=2e-5, atol_scale=2e-6)
    assert_close(np.asarray(bpp(g["vae_g"])), g["vae_grad"], rtol=3e-5, atol_scale=3e-6)
    assert_close(np.asarray(logstd(g["vae_x"])), g["vae_logstd"], rtol=2e-5, atol_scale=2e-6)
    c_in = Input(6)
    clip = Clip(c_in, -2.5, 2.5)
    y, cb = clip.evaluate(g["clip_v"])
    cb(g["clip_g"])
    assert np.array_equal(np.asarray(y), g["clip_y"])
    assert np.array_equal(np.asarray(c_in.last_gradient), g["clip_dx"])


def test_reinforce_steps_against_reference_run():
    """The trajectories benchmarks/policy.py:run() itself sampled (golden/reinforce.npz), pushed through this package's
    API exactly as the script does (policy.py:40-47: Trajectory.discounted -> RunningNormalize -> tanh -> Gauss.logprob
    backprop -> Adam -> load_params): the optimiser values follow the script's, episode by episode."""
    from mannequin2_b200 import Adam, RunningNormalize, Trajectory
    from mannequin2_b200.basicnet import Input, Affine, Tanh
    from mannequin2_b200.distrib import Gauss
    g = load_golden("reinforce")
    np.random.seed(0)
    policy = Input(11)
    for _ in range(2):
        policy = Tanh(Affine(policy, 64))
    policy = Gauss(mean=Affine(policy, 3, init=0.1))
    policy.load_params(g["rf_trace"][0])
    opt = Adam(policy.get_params(), horizon=10)
    normalize = RunningNormalize(horizon=2)
    pos = 0
    for ep, n in enumerate(int(v) for v in g["rf_len"]):
        traj = Trajectory(g["rf_obs"][pos:pos + n], g["rf_act"][pos:pos + n], g["rf_rew"][pos:pos + n])
        pos += n
        traj = traj.discounted(horizon=500)
        traj = traj.modified(rewards=normalize)
        traj = traj.modified(rewards=np.tanh)
        _, backprop = policy.logprob.evaluate(traj.o, sample=traj.a)
        opt.apply_gradient(backprop(traj.r), lr=0.001)
        policy.load_params(opt.get_value())
        err = np.max(np.abs(np.asarray(opt.get_value(), dtype=np.float64) - g["rf_trace"][ep + 1]))
        assert err < 5e-6, (ep, err)


def test_mlp_script_steps_against_reference_run():
    """mnist/mlp.py:21-43 written against `mannequin` (compat alias) with the minibatches mnist/mlp.py:run() itself drew
    (golden/mlp_script.npz): same initial weights from the same global generator, and the optimiser values of the first
    ten steps follow the script's."""
    import os
    import mannequin2_b200.compat  # noqa: F401
    from mannequin import Adam
    from mannequin.basicnet import Input, Affine, LReLU
    g = load_golden("mlp_script")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    try:
        from make_golden import mlp_script_data
    finally:
        sys.path.pop(0)
    data = mlp_script_data()

    def softmax(v):
        v = v.T
        v = np.exp(v - np.amax(v, axis=0))
        v /= np.sum(v, axis=0)
        return v.T

    state = np.random.get_state()
    try:
        np.random.seed(int(g["mlps_seed"][0]))
        model = Input(28, 28)
        for _ in range(2):
            model = LReLU(Affine(model, 128))
        model = Affine(model, 10, init=0.1)
    finally:
        np.random.set_state(state)
    stride, keep = int(g["mlps_stride"][0]), [int(k) for k in g["mlps_keep"]]
    assert np.max(np.abs(np.asarray(model.get_params())[::stride] - g["mlps_trace"][0])) < 1e-6
    opt = Adam(model.get_params(), horizon=10, lr=0.001)
    for step, idx in enumerate(g["mlps_idx"][:10].astype(np.int64)):
        outs, backprop = model.evaluate(data["train_x"][idx])
        grad = data["train_y"][idx] - softmax(outs) - outs * 0.01
        opt.apply_gradient(backprop(grad))
        model.load_params(opt.get_value())
        if step + 1 in keep:
            err = np.max(np.abs(np.asarray(opt.get_value(), dtype=np.float64)[::stride] - g["mlps_trace"][keep.index(step + 1)]))
            assert err < 1e-5, (step, err)


def test_vae_step_graph():
    """mnist/vae.py:39-71 on the device graph (DKLUninormal, Clip, Gauss.sample fan-out, matmul nodes): the gradient of
    mean(logprob) - mean(DKL) agrees with central differences along random directions (sampling noise held fixed), and
    a step of the clipped-momentum rule moves the parameters by lr * clip(0.1 * g)."""
    from mannequin2_b200.vae import VAETrainer
    from mannequin2_b200 import _device as D
    np.random.seed(3)
    rs = np.random.RandomState(4)
    x = rs.rand(8, 6, 5).astype(np.float32)

    class Replay(object):                                   # the same noise for every evaluation
        def randn(self, *shape):
            return np.random.RandomState(77).randn(*shape)
    tr = VAETrainer(image_shape=(6, 5), hidden=16, latent=3, rng=Replay())
    dec, dkl = tr.decoder, tr.dkl
    # keep both logstd heads strictly inside Clip's (-6, 0): its backward rule (vae.py:28-32) is deliberately not
    # the derivative of np.clip outside the range, so central differences only apply inside
    for gauss in (tr.encoder, tr.decoder):
        bias_fn = gauss.logstd.inputs[0]                    # Clip(Bias(Linear(h, W), b))
        lin, b = bias_fn.inputs
        b.dev.fill_(-3.0)
        while len(lin.inputs) == 1:                         # Reshape around the matmul of image-shaped heads
            lin = lin.inputs[0]
        lin.inputs[1].dev.mul_(0.1)
    theta0 = np.asarray(dec.get_params(), dtype=np.float64).copy()
    n_enc = dkl.n_params

    def loss_and_grad(theta):
        dec.load_params(theta.astype(np.float32))
        lp, bpp = dec.logprob.evaluate(x, sample=x)
        g = np.asarray(bpp(np.ones(8, np.float32)), dtype=np.float64).copy()
        kl, bpp2 = dkl.evaluate(x)
        g[:n_enc] += np.asarray(bpp2(-np.ones(8, np.float32)), dtype=np.float64)
        return float(np.mean(lp)) - float(np.mean(kl)), g

    l0, g0 = loss_and_grad(theta0)
    assert np.isfinite(l0) and np.isfinite(g0).all() and np.abs(g0).max() > 0
    for k in range(3):
        d = np.random.RandomState(10 + k).randn(len(theta0))
        d /= np.linalg.norm(d)
        eps = 2e-2
        fd = (loss_and_grad(theta0 + eps * d)[0] - loss_and_grad(theta0 - eps * d)[0]) / (2 * eps)
        an = float(g0 @ d)                                   # batch-mean gradients: d mean(...) / d theta
        assert abs(fd - an) <= 3e-2 * max(1.0, abs(an)), (k, fd, an)
    dec.load_params(theta0.astype(np.float32))
    lp, kl = tr.step(x)
    moved = np.asarray(dec.get_params(), dtype=np.float64) - theta0
    assert_close(moved, 1e-3 * np.clip(0.1 * g0, -1, 1), rtol=1e-3, atol_scale=1e-3)


def test_trajectory_device_gather_and_discounted():
    from mannequin2_b200 import Trajectory, discounted
    g = load_golden("trajectory")
    t = Trajectory(g["traj_o"], g["traj_a"], g["traj_r"])
    o, a, r = t.take_dev(g["traj_idx"])
    assert np.array_equal(o.cpu().numpy().reshape(64, 11), g["traj_take_o"])
    assert np.array_equal(a.cpu().numpy().reshape(64, 3), g["traj_take_a"])
    assert np.array_equal(r.cpu().numpy(), g["traj_take_r"])
    assert np.allclose(t.discounted(horizon=20).r, g["traj_disc_r"], rtol=1e-6)
    n = load_golden("norm")
    assert np.allclose(discounted(n["disc_in"], horizon=500), n["disc_h500"], rtol=1e-11, atol=1e-13)
    with pytest.raises(AssertionError):
        discounted([1.0, 2.0], horizon=0.5)


def test_unmodified_script_loop_via_compat_alias():
    """`import mannequin2_b200.compat` lets reference scripts run unchanged: the inner loop of
    mnist/mlp.py:24-35 written against `mannequin`, on synthetic MNIST-shaped data."""
    import mannequin2_b200.compat  # noqa: F401
    from mannequin import Adam, bar
    from mannequin.basicnet import Input, Affine, LReLU
    np.random.seed(0)
    rs = np.random.RandomState(1)
    proto = rs.rand(10, 28, 28).astype(np.float32)
    labels = rs.randint(10, size=2048)
    x = np.clip(proto[labels] + 0.3 * rs.randn(2048, 28, 28), 0, 1).astype(np.float32)
    y = np.eye(10)[labels]

    def softmax(v):
        v = v.T
        v = np.exp(v - np.amax(v, axis=0))
        v /= np.sum(v, axis=0)
        return v.T

    model = Input(28, 28)
    for _ in range(2):
        model = LReLU(Affine(model, 128))
    model = Affine(model, 10, init=0.1)
    opt = Adam(model.get_params(), horizon=10, lr=0.001)
    for _ in range(60):
        idx = np.random.randint(len(x), size=128)
        outs, backprop = model.evaluate(x[idx])
        grad = y[idx] - softmax(outs) - outs * 0.01
        opt.apply_gradient(backprop(grad))
        model.load_params(opt.get_value())
    acc = np.mean(np.argmax(model(x[:512]), axis=-1) == labels[:512])
    assert acc > 0.9, acc
    assert bar(100.0 * acc).startswith(" ")


@pytest.mark.gpu
def test_adam_get_value_is_a_snapshot():
    """mannequin/_adam.py:47-53 hands out an immutable snapshot: a value handle kept across apply_gradient still reads the
    OLD value (the device state is updated in place, so a held handle is detached onto its own copy first); a handle that
    was dropped costs nothing; a bare Input back-propagated with a host array returns it (basicnet.py:94-99)."""
    import mannequin2_b200.compat  # noqa: F401
    from mannequin import Adam
    from mannequin.basicnet import Input
    rs = np.random.RandomState(3)
    v0 = rs.randn(1000)
    opt = Adam(v0, horizon=10, lr=0.01)
    old = opt.get_value()
    g = rs.randn(1000)
    opt.apply_gradient(g)
    new = opt.get_value()
    assert np.allclose(np.asarray(old), v0) and not np.allclose(np.asarray(new), v0)
    assert np.max(np.abs(np.asarray(new) - np.asarray(old))) > 1e-3
    held = opt.get_value()
    opt.apply_gradient(g)
    assert np.array_equal(np.asarray(held), np.asarray(new))
    inp = Input(3)
    out, backprop = inp.evaluate(np.ones((2, 3), np.float32))
    backprop(np.full((2, 3), 2.0, np.float32))
    assert np.array_equal(inp.last_gradient, np.full((2, 3), 2.0, np.float32))

@pytest.mark.gpu
def test_mlp_fused_block_matches_oracle_and_reference_run():
    """csrc/mq_mlp_fused.cu -- a block of mnist/mlp.py:31-35 steps as ONE persistent launch (batch 128): against the oracle's
    restatement step for step (train_block and sgd_step, fp32 and fp64 optimiser state, the 784-128-128-10 and 784-128-10
    nets, a narrow ragged-slice net), bit-reproducible, and against the optimiser values that
    mnist/mlp.py:run() ITSELF produced on the same minibatches (golden/mlp_script.npz)."""
    import os
    from oracle import mq_oracle as O
    from mannequin2_b200.mlp import MLPTrainer
    from mannequin2_b200.ppo import init_mlp
    from mannequin2_b200 import _device as D
    rs = np.random.RandomState(18)
    n, steps, B = 700, 9, 128
    for n_in, widths in ((784, [128, 128, 10]), (784, [128, 10]), (36, [128, 128, 3])):
        acts = [O.ACT_LRELU] * (len(widths) - 1) + [O.ACT_NONE]
        spec = O.mlp_spec(n_in, widths, acts)
        np.random.seed(4)
        theta0 = init_mlp(n_in, widths, 0.1)
        # A hidden pre-activation within rounding distance of LReLU's kink takes slope 1 in one fp32 summation order and 0.1
        # in another -- both are right, and Adam turns the difference into full-size steps of a whole unit's weights.  The
        # comparison is made on draws where the oracle sees no pre-activation closer than 2e-6 to the kink.
        for attempt in range(20):
            x = rs.rand(n, n_in).astype(np.float32)
            y = np.eye(widths[-1], dtype=np.float32)[rs.randint(widths[-1], size=n)]
            tab = rs.randint(n, size=(steps, B)).astype(np.int32)
            opt = O.AdamO(theta0, horizon=10, lr=1e-3)
            want, margin = theta0.copy(), np.inf
            for s_ in range(steps):
                hidden = O.mlp_forward(want, spec, x[tab[s_]])[:-1]
                margin = min(margin, min(float(np.min(np.where(h < 0, -10.0 * h, h))) for h in hidden))     # |z| (leak 0.1)
                want = O.mlp_sgd_step(want, spec, opt, x[tab[s_]], y[tab[s_]])
            if margin > 2e-6:
                break
        assert margin > 2e-6
        kw = dict(n_in=n_in, widths=tuple(widths), acts=tuple([2] * (len(widths) - 1) + [0]), graph_steps=steps)
        xd, yd = D.from_host(x), D.from_host(y)
        got = {}
        # (the layered path has its own test above; held to the same bar here it depends on the draw: a parameter whose batch
        # gradient is rounding noise gets a full-size Adam step of either sign -- tools/mlp_blocks_check.py)
        for name, state, fused in (("fused32", "float32", True), ("fused64", "float64", True)):
            tr = MLPTrainer(params=theta0.copy(), state_dtype=state, fused=fused, **kw)
            assert tr.fused == fused
            tr.train_block(xd, yd, tab[:5])            # two blocks: the moments survive the launch boundary
            tr.train_block(xd, yd, tab[5:])
            got[name] = tr.get_params()
            err = np.max(np.abs(got[name] - want))
            assert np.max(np.abs(got[name] - theta0)) > 1e-3
            assert err < 2e-6 * (1 + steps) + 1e-6, (n_in, widths, name, err)
            val = D.to_host(tr.opt.state()[0])
            assert np.max(np.abs(val - got[name])) < 1e-6          # optimiser value and model copy agree
        again = MLPTrainer(params=theta0.copy(), **kw)
        again.train_block(xd, yd, tab[:5])
        again.train_block(xd, yd, tab[5:])
        assert np.array_equal(again.get_params(), got["fused32"]), "fused block is not bit-reproducible"
        one = MLPTrainer(params=theta0.copy(), **kw)     # host-side batches, one step per launch
        for s_ in range(steps):
            one.sgd_step(x[tab[s_]], y[tab[s_]])
        assert np.max(np.abs(one.get_params() - want)) < 2e-6 * (1 + steps) + 1e-6
    # the reference script's own run: same initial weights, its minibatches, its optimiser trace
    g = load_golden("mlp_script")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    try:
        from make_golden import mlp_script_data
    finally:
        sys.path.pop(0)
    data = mlp_script_data()
    import mannequin2_b200.compat  # noqa: F401
    from mannequin.basicnet import Input, Affine, LReLU
    state = np.random.get_state()
    try:
        np.random.seed(int(g["mlps_seed"][0]))
        model = Input(28, 28)
        for _ in range(2):
            model = LReLU(Affine(model, 128))
        model = Affine(model, 10, init=0.1)
    finally:
        np.random.set_state(state)
    stride, keep = int(g["mlps_stride"][0]), [int(k) for k in g["mlps_keep"]]
    theta0 = np.asarray(model.get_params(), np.float32)
    tr = MLPTrainer(params=theta0.copy(), state_dtype="float64", graph_steps=32)
    xd = D.from_host(np.ascontiguousarray(data["train_x"], np.float32).reshape(len(data["train_x"]), 784))
    yd = D.from_host(np.ascontiguousarray(data["train_y"], np.float32))
    idx = g["mlps_idx"].astype(np.int32)
    done = 0
    for k in keep:
        if k == 0 or k > len(idx):
            continue
        tr.train_block(xd, yd, idx[done:k])
        done = k
        val = D.to_host(tr.opt.state()[0]).astype(np.float64)
        err = np.max(np.abs(val[::stride] - g["mlps_trace"][keep.index(k)]))
        assert err < 1e-5 if k <= 10 else err < 1e-4, (k, err)
    assert done >= 10

