"""Pin the CPU oracle (oracle/mq_oracle.py) to the reference.

Two kinds of pins:
  * literal known answers from the reference's own golden files
    (/root/reference/tests/adam.out, basicnet.out -- values copied as data);
  * tests/golden/*.npz, produced by running the real reference in the build
    container (tests/golden/make_golden.py).
"""
import numpy as np
import pytest

from conftest import assert_close, load_golden
from oracle import mq_oracle as O


# ---- /root/reference/tests/adam.out:1-10 (Adam([10,-5], lr=2, horizon=3), grad=-value)
ADAM_OUT_1 = [[8., -3.], [6.055, -1.155], [4.218, 0.335], [2.552, 1.247], [1.125, 1.509],
              [-0.001, 1.25], [-0.789, 0.703], [-1.236, 0.105], [-1.376, -0.347], [-1.276, -0.545]]
# ---- adam.out:12-21 (Adam(zeros(3), horizon=1), grad=[.1,1,10]-value, lr=7/(i+1))
ADAM_OUT_2 = [[7., 7., 7.], [2.063, 2.129, 8.425], [0.958, 1.392, 9.031], [0.541, 1.17, 9.353],
              [0.35, 1.084, 9.545], [0.25, 1.045, 9.669], [0.195, 1.026, 9.753],
              [0.162, 1.015, 9.811], [0.142, 1.009, 9.853], [0.129, 1.006, 9.884]]


def test_adam_reference_out_file():
    opt = O.AdamO([10.0, -5.0], horizon=3, lr=2.0)
    for want in ADAM_OUT_1:
        opt.apply_gradient(-opt.get_value())
        assert np.allclose(opt.get_value(), want, atol=6e-4)
    opt = O.AdamO([0.0, 0.0, 0.0], horizon=1)
    for i, want in enumerate(ADAM_OUT_2):
        opt.apply_gradient([0.1, 1.0, 10.0] - opt.get_value(), lr=7.0 / (i + 1))
        assert np.allclose(opt.get_value(), want, atol=6e-4)


@pytest.mark.parametrize("name", ["adam", "adams"])
def test_adam_bit_exact_vs_reference(name):
    g = load_golden("adam")
    opt = O.AdamO(g[name + "_values"][0], horizon=10, lr=0.003, adams=(name == "adams"))
    for grad, lr, want in zip(g[name + "_grads"], g[name + "_lrs"], g[name + "_values"][1:]):
        if np.isnan(lr):
            opt.apply_gradient(grad)
        else:
            opt.apply_gradient(grad, lr=lr, epsilon=1e-6)
        # same fp64 operation order as mannequin/_adam.py + _normalize.py -> identical bits
        assert np.array_equal(opt.get_value(), want)
    opt = O.AdamO([10.0, -5.0], horizon=3, lr=2.0)
    for want in g["adam_kat1"]:
        opt.apply_gradient(-opt.get_value())
        assert np.array_equal(opt.get_value(), want)


def test_basicnet_reference_out_file():
    # /root/reference/tests/basicnet.py:6-22 + basicnet.out:1-7
    spec = O.mlp_spec(2, [2, 2], [O.ACT_NONE, O.ACT_NONE])
    theta = np.arange(12, dtype=np.float32)
    outs = O.mlp_forward(theta, spec, [[1.0, 1.0]])
    assert np.allclose(outs[-1], [[118., 134.]])
    grad = O.mlp_backward(theta, spec, [[1.0, 1.0]], outs, [[1.0, -1.0]])
    assert np.allclose(grad.reshape(-1, 6), [[-1, -1, -1, -1, -1, -1], [6, -6, 9, -9, 1, -1]])
    outs = O.mlp_forward(theta, spec, np.eye(2))
    assert np.allclose(outs[-1], [[82., 93.], [110., 125.]])
    grad = O.mlp_backward(theta, spec, np.eye(2), outs, [[1.0, -1.0], [-2.0, 2.0]])
    assert np.allclose(grad.reshape(-1, 6),
                       [[-0.5, -0.5, 1, 1, 0.5, 0.5], [-4, 4, -5, 5, -0.5, 0.5]])
    # basicnet.out:8-15: LReLU(0.1) / Tanh on top, input [-18.4, 2]
    v = [[-18.4, 2.0]]
    for act, y, g in (
        (O.ACT_LRELU, [-0.12, 0.4], [[117.76, 150.88, -12.8, -16.4, -6.4, -8.2],
                                     [0.8, -8.0, -0.74, 7.4, 0.1, -1.0]]),
        (O.ACT_TANH, [-0.834, 0.38], [[76.532, 96.794, -8.319, -10.521, -4.159, -5.261],
                                      [2.44, -6.845, -2.257, 6.332, 0.305, -0.856]])):
        spec = O.mlp_spec(2, [2, 2], [O.ACT_NONE, act])
        outs = O.mlp_forward(theta, spec, v)
        assert np.allclose(outs[-1], [y], atol=6e-4)
        grad = O.mlp_backward(theta, spec, v, outs, [[1.0, -1.0]])
        assert np.allclose(grad.reshape(-1, 6), g, atol=6e-4)


def test_nets_vs_reference():
    g = load_golden("nets")
    spec = O.mlp_spec(20, [16, 16, 5], [O.ACT_LRELU, O.ACT_LRELU, O.ACT_NONE])
    outs = O.mlp_forward(g["lrelu_theta"], spec, g["lrelu_x"].reshape(7, 20))
    assert_close(outs[-1], g["lrelu_y"])
    grad, dx = O.mlp_backward(g["lrelu_theta"], spec, g["lrelu_x"].reshape(7, 20), outs,
                              g["lrelu_dy"], want_dx=True)
    assert_close(grad, g["lrelu_grad"])
    assert_close(dx.reshape(7, 4, 5), g["lrelu_dx"])

    spec = O.mlp_spec(11, [64, 64, 3], [O.ACT_TANH, O.ACT_TANH, O.ACT_NONE])
    outs = O.mlp_forward(g["tanh_theta"], spec, g["tanh_x"])
    assert_close(outs[-1], g["tanh_y"], rtol=2e-5, atol_scale=2e-6)   # reference is fp32 here
    grad = O.mlp_backward(g["tanh_theta"], spec, g["tanh_x"], outs, g["tanh_dy"])
    assert_close(grad, g["tanh_grad"], rtol=2e-5, atol_scale=2e-6)
    # unbatched: batch of one, no division (basicnet.py:242-249)
    outs = O.mlp_forward(g["tanh_theta"], spec, g["tanh_x"][:1])
    assert_close(outs[-1][0], g["tanh_y1"], rtol=2e-5, atol_scale=2e-6)
    grad = O.mlp_backward(g["tanh_theta"], spec, g["tanh_x"][:1], outs, g["tanh_dy"][:1])
    assert_close(grad, g["tanh_grad1"], rtol=2e-5, atol_scale=2e-6)


def test_running_stats_vs_reference():
    g = load_golden("norm")
    n = O.RunningNormalizeO(horizon=2)
    pos, res = 0, []
    for s in g["norm_scalar_sizes"]:
        d = g["norm_scalar_in"][pos:pos + s]
        pos += s
        res.append(np.concatenate([n(d), [n.get_mean(), n.get_var()]]))
    assert np.allclose(np.concatenate(res), g["norm_scalar_out"], rtol=1e-13, atol=1e-13)

    n = O.RunningNormalizeO(horizon=10)
    got = np.array([n(v) for v in g["norm_single_in"]])
    assert np.allclose(got, g["norm_single_out"], rtol=1e-13, atol=1e-13)
    assert got[0] == 0.0                                   # zeros until 2 samples

    n = O.RunningNormalizeO((11,))
    got = np.array([n(v).reshape(11) for v in g["norm_vec_in"]])
    assert np.allclose(got, g["norm_vec_out"], rtol=1e-12, atol=1e-12)
    assert np.allclose(n.apply(g["norm_vec_in"]), g["norm_vec_batch_apply"], rtol=1e-12)
    assert np.allclose(n.get_mean(), g["norm_vec_mean"], rtol=1e-13)
    assert np.allclose(n.get_var(), g["norm_vec_var"], rtol=1e-13)

    m = O.RunningMeanO((7,), horizon=3)
    for x, w, want in zip(g["rmean_x"], g["rmean_w"], g["rmean_out"]):
        m.update(x, weight=w)
        assert np.array_equal(m.get(), want)

    assert np.array_equal(O.discounted(g["disc_in"], 500), g["disc_h500"])
    assert np.array_equal(O.discounted(g["disc_in"], 3), g["disc_h3"])


def test_running_norm_reference_asserts():
    # /root/reference/tests/running_norm.py:8-16,33-43: == np.mean / np.var(ddof=1)
    n = O.RunningNormalizeO()
    data = np.arange(10) + 10
    n.update(data[0])
    for i in range(2, len(data)):
        n.update(data[i - 1])
        assert abs(n.get_mean() - np.mean(data[:i])) < 1e-10
        assert abs(n.get_var() - np.var(data[:i], ddof=1)) < 1e-10
    rs = np.random.RandomState(0)
    n = O.RunningNormalizeO((5,))
    data = rs.rand(10, 5)
    for d in data:
        n.update(d)
    want = (data - data.mean(axis=0)) / data.std(axis=0, ddof=1)
    assert np.mean(np.abs(n.apply(data) - want)) < 1e-10


def test_trajectory_rules_vs_reference():
    g = load_golden("trajectory")
    o, a, r = (O.standardize_field(g["traj_o"]), O.standardize_field(g["traj_a"]),
               np.asarray(g["traj_r"], np.float32))
    to, ta, tr = O.trajectory_take(o, a, r, g["traj_idx"])
    assert np.array_equal(to, g["traj_take_o"]) and to.dtype == np.float32
    assert np.array_equal(ta, g["traj_take_a"]) and np.array_equal(tr, g["traj_take_r"])
    so, _, sr = O.trajectory_take(o, a, r, slice(5, 33, 3))
    assert np.array_equal(so, g["traj_slice_o"]) and np.array_equal(sr, g["traj_slice_r"])
    assert np.array_equal(O.discounted(r, 20).astype(np.float32), g["traj_disc_r"])
    assert O.standardize_field([3, 1, 2]).dtype == np.int32 == g["traj_int_o"].dtype
    assert np.array_equal(g["traj_int_r"], np.ones(3, np.float32))


def test_gauss_vs_reference_and_fd():
    g = load_golden("gauss")
    spec = O.policy_spec(11, 3)
    assert spec["n_params"] == 5126 == len(g["gauss_theta"])
    logp, outs = O.policy_logprob(g["gauss_theta"], spec, g["gauss_obs"], g["gauss_act"])
    assert_close(logp, g["gauss_logp"], rtol=2e-6, atol_scale=1e-6)
    grad = O.policy_logprob_grad(g["gauss_theta"], spec, g["gauss_obs"], g["gauss_act"],
                                 g["gauss_g"], outs)
    assert_close(grad, g["gauss_grad"], rtol=3e-5, atol_scale=3e-6)
    # closed form vs central finite differences (fp64)
    rs = np.random.RandomState(1)
    mu, ls, s, w = rs.randn(4, 3), rs.randn(3) * 0.3, rs.randn(4, 3), rs.randn(4)
    d_mu, d_ls = O.gauss_logprob_grads(mu, ls, s, w)
    f = lambda m, l: np.sum(O.gauss_logprob(m, l, s) * w)
    for idx in np.ndindex(4, 3):
        e = np.zeros((4, 3)); e[idx] = 1e-6
        assert abs((f(mu + e, ls) - f(mu - e, ls)) / 2e-6 - d_mu[idx]) < 1e-6
    for j in range(3):
        e = np.zeros(3); e[j] = 1e-6
        assert abs((f(mu, ls + e) - f(mu, ls - e)) / 2e-6 - d_ls.sum(axis=0)[j]) < 1e-6
    # sample path: mean + unit*exp(logstd); backward (g, sum_b g*noise)
    theta = g["gauss_theta"]
    mu = O.mlp_forward(theta[:5123], O._mlp_part(spec), g["gauss_obs"][:5])[-1]
    smp, noise = O.gauss_sample(mu, theta[-3:], g["gauss_sample_unit"])
    assert_close(smp, g["gauss_sample_out"], rtol=1e-5)
    gs = g["gauss_sample_g"].astype(np.float64)
    grad = np.zeros(5126, np.float32)
    outs = O.mlp_forward(theta[:5123], O._mlp_part(spec), g["gauss_obs"][:5])
    grad[:5123] = O.mlp_backward(theta[:5123], O._mlp_part(spec), g["gauss_obs"][:5], outs, gs)
    grad[5123:] = (gs * noise).sum(axis=0) / 5
    assert_close(grad, g["gauss_sample_grad"], rtol=3e-5, atol_scale=3e-6)


def test_discrete_vs_reference():
    """Categorical head restatement against mannequin.distrib.Discrete run here (golden/discrete.npz)."""
    g = load_golden("discrete")
    spec = O.mlp_spec(4, [64, 64, 3], [O.ACT_TANH, O.ACT_TANH, O.ACT_NONE])
    assert spec["n_params"] == len(g["disc_theta"])
    outs = O.mlp_forward(g["disc_theta"], spec, g["disc_obs"])
    assert_close(outs[-1], g["disc_logits"], rtol=2e-5, atol_scale=2e-6)
    logp, p = O.discrete_logprob(outs[-1], g["disc_act"])
    assert_close(logp, g["disc_logp"], rtol=2e-5, atol_scale=1e-6)
    assert np.allclose(p.sum(axis=1), 1.0)
    grad = O.discrete_logprob_grad(g["disc_theta"], spec, g["disc_obs"], g["disc_act"], g["disc_g"], outs)
    assert_close(grad, g["disc_grad"], rtol=3e-5, atol_scale=3e-6)


def test_rollout_side_vs_reference():
    """Rollout-side restatements: NormalizedObservations against the real wrapper (mannequin/gym.py:61-84) stepping one
    environment, RunningNormalize fed whole batches (golden/obsnorm.npz, both generated by the reference here), the
    Philox4x32-10 generator against Random123's known-answer vectors, and the chunk-end bootstrap against the literal
    loop of benchmarks/gae.py:52-61."""
    g = load_golden("obsnorm")
    norm = O.RunningNormalizeO(shape=(11,))
    for raw, want in zip(g["on_raw"], g["on_out"]):
        got = O.normalized_observations_step(norm, raw.reshape(1, 11), 5.0).reshape(-1)
        assert np.max(np.abs(got - want)) < 1e-12
    assert np.max(np.abs(norm.get_mean() - g["on_mean"])) < 1e-12
    assert np.max(np.abs(norm.get_var() - g["on_var"])) < 1e-12
    assert np.all(g["on_out"][0] == 0.0) and np.max(np.abs(g["on_out"])) <= 5.0      # zeros before 2 samples; clipped
    norm = O.RunningNormalizeO(shape=(11,))
    for b, want in zip(g["on_batches"], g["on_batch_out"]):
        assert np.max(np.abs(O.normalized_observations_step(norm, b, 5.0) - want)) < 1e-12
    assert np.max(np.abs(norm.get_var() - g["on_batch_var"])) < 1e-12
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert tuple(int(v) for v in O.philox4x32(*ctr, *key)) == want
    z = O.philox_normals(99, 1, 2, np.arange(100000))
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01 and abs(np.mean(z ** 4) - 3.0) < 0.1
    # chunk end: gae.py:52-61 with adv[length] = 0 vs the terminal form on the patched rewards
    rs = np.random.RandomState(3)
    n_seg, seg, gam, lam = 5, 7, 0.99, 0.95
    rew = rs.randn(n_seg * seg).astype(np.float32)
    done = rs.rand(n_seg * seg) < 0.2
    v = rs.randn(n_seg, seg + 1).astype(np.float32)
    want = np.zeros((n_seg, seg))
    for e in range(n_seg):                                  # the reference loop, one chunk per segment
        adv = np.zeros(seg + 1)
        for i in range(seg - 1, -1, -1):
            k = e * seg + i
            if done[k]:
                adv[i] = rew[k] - v[e, i]
            else:
                adv[i] = rew[k] + gam * v[e, i + 1] - v[e, i] + gam * lam * adv[i + 1]
        want[e] = adv[:seg]
    r2, d2 = O.bootstrap_segments(rew, done, v[:, seg], seg, gam)
    flat_v = np.concatenate([v[:, :seg].reshape(-1), [0.0]]).astype(np.float32)
    adv, _ = O.gae_advantages(r2, flat_v, d2, gam, lam)
    assert np.max(np.abs(adv - want.reshape(-1))) < 1e-5


def test_vae_step_vs_reference_run():
    """O.vae_grads / O.vae_step against ONE iteration of mnist/vae.py:57-69 run from the reference's own classes
    (golden/vae_step.npz, tests/golden/make_golden.py:gen_vae_step): the graph of vae.py:39-55 with two hidden layers
    per side, log-deviations on both sides of Clip's upper bound, both evaluations, both backward passes and the
    clipped-momentum update.  The reference computes in float32, the oracle in float64."""
    g = load_golden("vae_step")
    d_img, hidden, latent, n_hidden, batch = (int(v) for v in g["vstep_dims"])
    assert (g["vstep_logstd_d"] == 0.0).any() and (g["vstep_logstd_d"] < 0.0).any() and (g["vstep_logstd_e"] == 0.0).any()
    lp, dkl, g1, g2 = O.vae_grads(g["vstep_theta"], d_img, hidden, latent, n_hidden, g["vstep_x"], g["vstep_unit"])
    assert_close(lp, g["vstep_logprob"], rtol=2e-6, atol_scale=2e-6)
    assert_close(dkl, g["vstep_dkl"], rtol=2e-6, atol_scale=2e-6)
    assert_close(g1, g["vstep_grad1"], rtol=1e-5, atol_scale=2e-6)
    assert_close(g2, g["vstep_grad2"], rtol=1e-5, atol_scale=2e-6)
    th, mom, _, _, grad = O.vae_step(g["vstep_theta"], g["vstep_momentum"], d_img, hidden, latent, n_hidden, g["vstep_x"],
                                     g["vstep_unit"])
    assert_close(grad, g["vstep_grad"], rtol=1e-5, atol_scale=2e-6)
    assert (np.abs(g["vstep_momentum_new"]) == 1.0).any() and (np.abs(g["vstep_momentum_new"]) < 1.0).any()
    assert np.max(np.abs(mom - g["vstep_momentum_new"])) < 2e-5
    assert np.max(np.abs(th - g["vstep_theta_new"])) < 2e-7


def test_vae_heads_vs_reference_module():
    """DKLUninormal and Clip against mnist/vae.py:13-34 ITSELF (the module imported in the build container with
    matplotlib and autograd replaced by stand-ins, golden/vae.npz): forward values and the flat gradient of a small
    encoder graph whose clipped logstd head touches the range boundary; Clip alone on both sides of its range."""
    g = load_golden("vae")
    th, x, gg = g["vae_theta"].astype(np.float64), g["vae_x"].astype(np.float64), g["vae_g"].astype(np.float64)
    w1, b1 = th[:40].reshape(5, 8), th[40:48]
    w2, b2 = th[48:72].reshape(8, 3), th[72:75]
    w3, b3 = th[75:99].reshape(8, 3), th[99:102]
    h = np.tanh(x @ w1 + b1)
    mean, ls_raw = h @ w2 + b2, h @ w3 + b3
    ls = O.clip_forward(ls_raw, -6.0, 0.0)
    assert np.max(np.abs(mean - g["vae_mean"])) < 1e-5 and np.max(np.abs(ls - g["vae_logstd"])) < 1e-5
    assert (ls == 0.0).any() and (ls < 0.0).any()
    assert_close(O.dkl_uninormal(mean, ls), g["vae_dkl"], rtol=2e-5, atol_scale=2e-6)
    d_mean, d_ls = O.dkl_uninormal_grads(mean, ls, gg)
    d_raw = O.clip_backward(ls_raw, d_ls, -6.0, 0.0)
    n = len(x)
    d_h = (d_mean @ w2.T + d_raw @ w3.T) * (1.0 - h * h)
    grad = np.concatenate([(x.T @ d_h).ravel(), d_h.sum(0), (h.T @ d_mean).ravel(), d_mean.sum(0),
                           (h.T @ d_raw).ravel(), d_raw.sum(0)]) / n          # batch-mean rule, basicnet.py:242-249
    assert_close(grad, g["vae_grad"], rtol=3e-5, atol_scale=3e-6)
    assert np.array_equal(O.clip_forward(g["clip_v"], -2.5, 2.5), g["clip_y"])
    assert np.array_equal(O.clip_backward(g["clip_v"], g["clip_g"], -2.5, 2.5).astype(np.float32), g["clip_dx"])


def test_es_update_rule_vs_reference_run():
    """benchmarks/mutation.py:run() itself, 12 generations on a stand-in environment (golden/es.npz): with a population
    of one the oracle's ES update (returns through RunningNormalize(horizon 10) as a batch, Adam(horizon 10, lr 0.01) on
    mean_p noise_p r_p) reproduces the script's optimiser values.  The perturbations are replayed from the global
    NumPy generator exactly as the script drew them."""
    g = load_golden("es")
    state = np.random.get_state()
    try:
        np.random.seed(int(g["es_seed"][0]))
        for shape in ((11, 64), (64, 64), (64, 3)):            # build_policy's normed_columns draws (basicnet.py:265-267)
            np.random.randn(*shape)
        n_gen = len(g["es_returns"])
        diffs = [np.random.randn(5123) * 0.1 for _ in range(n_gen)]
    finally:
        np.random.set_state(state)
    keep = [int(k) for k in g["es_keep"]]
    theta0 = g["es_trace"][0]
    opt, norm = O.AdamO(theta0.astype(np.float32), horizon=10), O.RunningNormalizeO(horizon=10)
    for gen in range(n_gen):
        r = O.es_apply_returns(opt, norm, diffs[gen][None, :], np.array([g["es_returns"][gen]]), lr=0.01)
        assert abs(np.linalg.norm(diffs[gen] * r[0]) - g["es_grad_norms"][gen]) < 1e-9 * (1 + g["es_grad_norms"][gen])
        if gen + 1 in keep:
            want = g["es_trace"][keep.index(gen + 1)]
            assert np.max(np.abs(opt.get_value() - want)) < 1e-10, (gen, np.max(np.abs(opt.get_value() - want)))
    assert g["es_grad_norms"][0] == 0.0 and np.all(g["es_grad_norms"][1:] > 0)     # zeros until two returns were seen


def test_reinforce_loop_vs_reference_run():
    """benchmarks/policy.py:run() itself (plain REINFORCE, policy.py:29-47) for 6 episodes (golden/reinforce.npz): from
    the trajectories the script sampled, the oracle's discounted() -> RunningNormalize(horizon 2) -> tanh ->
    Gauss.logprob gradient (batch mean) -> Adam(horizon 10, lr 0.001) reproduces the script's optimiser values."""
    g = load_golden("reinforce")
    spec = O.policy_spec(11, 3)
    theta = g["rf_trace"][0].astype(np.float32)
    opt, norm = O.AdamO(theta, horizon=10), O.RunningNormalizeO(horizon=2)
    pos = 0
    for ep, n in enumerate(g["rf_len"]):
        o, a, r = g["rf_obs"][pos:pos + n], g["rf_act"][pos:pos + n], g["rf_rew"][pos:pos + n]
        pos += n
        r = O.discounted(r, horizon=500).astype(np.float32)            # Trajectory stores float32 rewards
        r = np.asarray(norm(r)).astype(np.float32)
        r = np.tanh(r).astype(np.float32)
        grad = O.policy_logprob_grad(theta, spec, o, a, r)
        opt.apply_gradient(grad, lr=0.001)
        theta = opt.get_value().astype(np.float32)
        err = np.max(np.abs(opt.get_value() - g["rf_trace"][ep + 1]))
        assert err < 2e-6, (ep, err)
    assert np.max(np.abs(g["rf_trace"][-1] - g["rf_trace"][0])) > 1e-3


def test_mlp_script_vs_reference_run():
    """mnist/mlp.py:run() itself, 30 optimiser steps on synthetic MNIST-shaped data (golden/mlp_script.npz): the oracle's
    column-normalised init drawn from the same global generator, the script's recorded minibatch indices, and
    mlp_sgd_step (784-128-128-10 LReLU net, dy = onehot - softmax - 0.01 out, Adam(horizon 10, lr 1e-3)) reproduce a
    strided subset of the script's optimiser values at steps 0, 1, 2, 3, 6, 10, 15 and 20 to 1e-6.  (Later steps are not
    compared: on 384 memorised samples the trajectory is sensitive to rounding -- the oracle, which evaluates the first
    layer in fp64 where the reference calls an fp32 sgemm, agrees with the script to 3e-6 up to step 20 and is 1e-4 away
    at step 25 and 1e-3 at step 30.)"""
    import os
    import sys
    g = load_golden("mlp_script")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    try:
        from make_golden import mlp_script_data
    finally:
        sys.path.pop(0)
    data = mlp_script_data()
    x, y = data["train_x"].reshape(-1, 784), data["train_y"]
    spec = O.mlp_spec(784, [128, 128, 10], [O.ACT_LRELU, O.ACT_LRELU, O.ACT_NONE])
    state = np.random.get_state()
    try:
        np.random.seed(int(g["mlps_seed"][0]))
        theta = O.mlp_init(spec, np.random, 0.1)               # the script's three normed_columns draws
    finally:
        np.random.set_state(state)
    stride, keep = int(g["mlps_stride"][0]), [int(k) for k in g["mlps_keep"]]
    assert np.max(np.abs(theta[::stride] - g["mlps_trace"][0])) < 1e-7
    opt = O.AdamO(theta, horizon=10, lr=0.001)
    for step, idx in enumerate(g["mlps_idx"].astype(np.int64)):
        theta = O.mlp_sgd_step(theta, spec, opt, x[idx], y[idx])
        if step + 1 in keep:
            err = np.max(np.abs(opt.get_value()[::stride] - g["mlps_trace"][keep.index(step + 1)]))
            assert err < 1e-6, (step, err)
    assert len(g["mlps_bars"]) == 10 and all("[" in str(b) for b in g["mlps_bars"])


def test_vae_heads_fd():
    rs = np.random.RandomState(2)
    m, l, w = rs.randn(5, 3), rs.randn(5, 3) * 0.5, rs.randn(5)
    dm, dl = O.dkl_uninormal_grads(m, l, w)
    f = lambda a, b: np.sum(O.dkl_uninormal(a, b) * w)
    for idx in np.ndindex(5, 3):
        e = np.zeros((5, 3)); e[idx] = 1e-6
        assert abs((f(m + e, l) - f(m - e, l)) / 2e-6 - dm[idx]) < 1e-6
        assert abs((f(m, l + e) - f(m, l - e)) / 2e-6 - dl[idx]) < 1e-6
    # Clip backward: /root/reference/tests/reshape.out "Gradient" blocks 2->3
    v = np.array([[0., 0., -2.], [3., 0., -5.]])
    assert np.array_equal(O.clip_forward(v, -2.5, 2.5), [[0, 0, -2], [2.5, 0, -2.5]])
    assert np.array_equal(O.clip_backward(v, np.ones((2, 3)), -2.5, 2.5), [[1, 1, 1], [0, 1, 1]])


def test_gae_vs_reference_chunks():
    g = load_golden("gae")
    # value predictions are internal to the reference closure; the advantages of the
    # FIRST chunk use the freshly initialised predictor, which make_golden seeds
    # with np.random.seed(79): rebuild it through the oracle's init.
    np_state = np.random.RandomState(79)
    spec = O.mlp_spec(11, [64, 64, 1], [O.ACT_TANH, O.ACT_TANH, O.ACT_NONE])
    theta = O.mlp_init(spec, np_state, head_scale=0.1)
    n = 2048
    obs, rew, done = g["gae_obs"], g["gae_rew"], g["gae_done"]
    value = O.mlp_forward(theta, spec, obs[:n + 1], dtype=np.float32)[-1].reshape(-1)
    adv, _ = O.gae_advantages(rew[:n], value, done[:n], 0.97, 0.9)
    assert_close(adv, g["gae_adv1"], rtol=2e-5, atol_scale=2e-6)
    adv32, _ = O.gae_advantages(rew[:n], value, done[:n], 0.97, 0.9, dtype=np.float32)
    assert_close(adv32, g["gae_adv1"], rtol=2e-5, atol_scale=2e-6)
    # chunk carry-over (gae.py:64): second chunk starts at step 2048
    assert len(obs) == 2 * n + 1
    assert np.array_equal(obs[n], g["gae_o2_first"])
    # segmentation is integer-exact: with gam=lam=1 and integer data the scan is exact
    rs = np.random.RandomState(5)
    r = rs.randint(-3, 4, size=500).astype(np.float64)
    v = rs.randint(-3, 4, size=501).astype(np.float64)
    d = rs.rand(500) < 0.05
    a, _ = O.gae_advantages(r, v, d, 1.0, 1.0)
    want = np.zeros(501)
    for i in range(499, -1, -1):
        want[i] = r[i] - v[i] + (0 if d[i] else v[i + 1] + want[i + 1])
    assert np.array_equal(a, want[:500].astype(np.float32))


def test_ppo_run_vs_reference():
    """Replays benchmarks/ppo.py:run() (two chunks) through the oracle."""
    g = load_golden("ppo")
    n = 2048
    pol_spec = O.policy_spec(11, 3)
    val_spec = O.mlp_spec(11, [64, 64, 1], [O.ACT_TANH, O.ACT_TANH, O.ACT_NONE])
    pol_theta, val_theta = g["ppo_pol_init"], g["ppo_val_init"]
    pol_opt = O.AdamO(pol_theta, horizon=10)
    val_opt = O.AdamO(val_theta, horizon=10, lr=0.003)
    norm = O.RunningNormalizeO(horizon=2)
    ck = {int(s): i for i, s in enumerate(g["ppo_ck_steps"])}
    obs, act, rew, done = g["ppo_obs"], g["ppo_act"], g["ppo_rew"], g["ppo_done"]
    assert len(obs) == 2 * n + 1
    for c in range(2):
        lo = c * n
        # teacher-forced checkpoints: run each chunk with the oracle, compare the
        # parameter trajectories at the recorded steps (tolerance grows with depth
        # because fp32-vs-fp64 rounding feeds back through Adam).
        value = O.mlp_forward(val_theta, val_spec, obs[lo:lo + n + 1])[-1].reshape(-1).astype(np.float32)
        adv, ret = O.gae_advantages(rew[lo:lo + n], value, done[lo:lo + n])
        o = obs[lo:lo + n]
        for k, idx in enumerate(g["ppo_val_idx"][c * 300:(c + 1) * 300]):
            val_theta = O.value_sgd_step(val_theta, val_spec, val_opt, o[idx], ret[idx].reshape(-1, 1))
            s = c * 300 + k
            if s in ck:
                tol = 2e-6 * (1 + k)
                assert np.max(np.abs(val_theta - g["ppo_val_ck"][ck[s]])) < tol + 1e-6, ("val", s)
        adv_n = np.tanh(norm(adv)).astype(np.float32)
        a = act[lo:lo + n]
        logp_old, _ = O.policy_logprob(pol_theta, pol_spec, o, a)
        for k, idx in enumerate(g["ppo_pol_idx"][c * 300:(c + 1) * 300]):
            pol_theta = O.ppo_policy_step(pol_theta, pol_spec, pol_opt, o[idx], a[idx],
                                          adv_n[idx], logp_old[idx])
            s = c * 300 + k
            if s in ck:
                tol = 2e-6 * (1 + k)
                assert np.max(np.abs(pol_theta - g["ppo_pol_ck"][ck[s]])) < tol + 1e-6, ("pol", s)
