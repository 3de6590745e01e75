"""Kernel parity cases: the C-ABI (include/mq_b200.h) against the CPU oracle.

Each case takes a backend `B` with
    B.lib            bound library (mannequin2_b200._lib.bind)
    B.up(a)          host ndarray -> buffer handle
    B.zeros(shape,dtype) -> buffer handle
    B.ptr(h)         address to pass through ctypes (None stays None)
    B.down(h)        buffer handle -> host ndarray (synchronises)
`tests/test_kernels_gpu.py` runs them on a B200 with device buffers (the parity
tests proper); `tools/emu/check_kernels.py` runs the same cases on the
developer-only host emulation.
"""
import ctypes
import os

import numpy as np

from mannequin2_b200 import _lib as L
from oracle import mq_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def ok(B, rc):
    if rc != 0:
        raise RuntimeError("rc=%d: %s" % (rc, B.lib.mq_last_error().decode()))


def close(a, b, rtol=1e-5, atol_scale=2e-6, what=""):
    """allclose(rtol, atol = atol_scale*max|ref|): the fp32 FFMA bar of BASELINE.md 3.6."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    atol = atol_scale * max(1e-30, np.max(np.abs(b)))
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    if not (err <= 0).all():
        i = np.argmax(err)
        raise AssertionError("%s: mismatch at %d: got %r want %r (max|ref|=%g)" % (
            what, i, a.reshape(-1)[i], b.reshape(-1)[i], np.max(np.abs(b))))


def net_of(spec, logstd=False):
    Ls = spec["layers"]
    return L.make_net(spec["n_in"], [l["n_out"] for l in Ls], [l["act"] for l in Ls],
                      leak=[l["leak"] for l in Ls], logstd=logstd)


def head(B, kind, p0=0.0, p1=0.0, dy=None, target=None, adv=None, logp_old=None, tc=0, records=None, rec_w=0,
         image=None):
    h = L.MqHead()
    h.kind, h.p0, h.p1 = kind, p0, p1
    h.dy, h.target, h.adv, h.logp_old = B.ptr(dy), B.ptr(target), B.ptr(adv), B.ptr(logp_old)
    h.tc, h.records, h.rec_w, h.tc_image = tc, B.ptr(records), rec_w, B.ptr(image)
    h._keep = (dy, target, adv, logp_old, records, image)
    return h


def mlp_case(rs, n_in, widths, acts, batch):
    spec = O.mlp_spec(n_in, widths, acts)
    theta = O.mlp_init(spec, rs, head_scale=0.1) + rs.randn(spec["n_params"]).astype(np.float32) * 0.05
    x = rs.randn(batch, n_in).astype(np.float32)
    return spec, theta, x


# ------------------------------------------------------------------------- cases
def case_adam(B):
    """mannequin/_adam.py: fp64 state is bit-exact with the reference's op order."""
    lib = B.lib
    g = golden("adam")
    for name, adams in (("adam", 0), ("adams", 1)):
        vals = g[name + "_values"]
        n = vals.shape[1]
        val, m, v = B.up(vals[0].copy()), B.zeros(n, np.float64), B.zeros(n, np.float64)
        th = B.zeros(n, np.float32)
        tw1 = tw2 = 0.0
        for grad, lr, want in zip(g[name + "_grads"], g[name + "_lrs"], vals[1:]):
            tw1, tw2 = tw1 * 0.9 + 1.0, tw2 * 0.99 + 1.0
            lr, eps = (0.003, 1e-8) if np.isnan(lr) else (float(lr), 1e-6)
            hg = B.up(grad)
            ok(B, lib.mq_adam_step(B.ptr(val), B.ptr(m), B.ptr(v), B.ptr(th), B.ptr(hg), n, L.MQ_F64,
                                   L.MQ_F32, 1.0 / tw1, 1.0 / tw2, lr, eps, adams, B.stream))
            assert np.array_equal(B.down(val), want), "fp64 %s must match the reference bit for bit" % name
            assert np.array_equal(B.down(th), want.astype(np.float32))
    rs = np.random.RandomState(0)
    for n in (5, 37, 1024 + 3, 100003):
        val0 = rs.randn(n)
        grads = rs.randn(4, n).astype(np.float32)
        ref = O.AdamO(val0, horizon=10, lr=0.003)
        val, m, v = B.up(val0.copy()), B.zeros(n, np.float64), B.zeros(n, np.float64)
        f32 = [B.up(val0.astype(np.float32)), B.zeros(n, np.float32), B.zeros(n, np.float32)]
        tw1 = tw2 = 0.0
        for grad in grads:
            ref.apply_gradient(grad)
            tw1, tw2 = tw1 * 0.9 + 1.0, tw2 * 0.99 + 1.0
            hg = B.up(grad)
            ok(B, lib.mq_adam_step(B.ptr(val), B.ptr(m), B.ptr(v), None, B.ptr(hg), n, L.MQ_F64, L.MQ_F32,
                                   1.0 / tw1, 1.0 / tw2, 0.003, 1e-8, 0, B.stream))
            ok(B, lib.mq_adam_step(B.ptr(f32[0]), B.ptr(f32[1]), B.ptr(f32[2]), None, B.ptr(hg), n,
                                   L.MQ_F32, L.MQ_F32, 1.0 / tw1, 1.0 / tw2, 0.003, 1e-8, 0, B.stream))
        assert np.array_equal(B.down(val), ref.get_value())
        close(B.down(f32[0]), ref.get_value(), what="adam fp32 state n=%d" % n)
    # fp64 gradient input (host scripts pass python lists, tests/adam.py)
    val, m, v = B.up(np.array([10.0, -5.0])), B.zeros(2, np.float64), B.zeros(2, np.float64)
    tw1 = tw2 = 0.0
    for want in g["adam_kat1"]:
        tw1, tw2 = tw1 * (1 - 1 / 3.0) + 1.0, tw2 * 0.99 + 1.0
        hg = B.up(-B.down(val))
        ok(B, lib.mq_adam_step(B.ptr(val), B.ptr(m), B.ptr(v), None, B.ptr(hg), 2, L.MQ_F64, L.MQ_F64,
                               1.0 / tw1, 1.0 / tw2, 2.0, 1e-8, 0, B.stream))
        assert np.array_equal(B.down(val), want)


def case_norm(B):
    lib = B.lib
    rs = np.random.RandomState(1)
    for rows, cols, dt in ((2048, 1, np.float32), (5000, 1, np.float64), (9, 11, np.float64),
                           (300, 5, np.float32), (1, 1, np.float32), (70000, 1, np.float32)):
        x = (rs.randn(rows, cols) * 3 + 1).astype(dt)
        code = L.MQ_F64 if dt == np.float64 else L.MQ_F32
        wsb = lib.mq_norm_ws_bytes(rows, cols)
        hx, ws = B.up(x), B.zeros(wsb, np.uint8)
        mean = B.zeros(cols, np.float64)
        ok(B, lib.mq_col_mean(B.ptr(hx), code, rows, cols, 1.0, B.ptr(mean), B.ptr(ws), wsb, B.stream))
        hmean = B.down(mean)
        assert np.allclose(hmean, x.astype(np.float64).mean(axis=0), rtol=1e-13, atol=1e-13)
        mo = rs.randn(cols)
        hmo, cov = B.up(mo), B.zeros(cols, np.float64)
        ok(B, lib.mq_norm_cov(B.ptr(hx), code, rows, cols, B.ptr(hmo), B.ptr(mean), 2.0, B.ptr(cov),
                              B.ptr(ws), wsb, B.stream))
        want = 2.0 * ((x.astype(np.float64) - mo) * (x.astype(np.float64) - hmean)).mean(axis=0)
        assert np.allclose(B.down(cov), want, rtol=1e-12, atol=1e-12)
        var = np.abs(want) + 0.1
        hvar = B.up(var)
        out = B.zeros((rows, cols), np.float32)
        ok(B, lib.mq_norm_apply(B.ptr(hx), code, rows, cols, B.ptr(mean), B.ptr(hvar), B.ptr(out), L.MQ_F32,
                                1, B.stream))
        close(B.down(out), np.tanh(((x - hmean) / np.maximum(1e-8, np.sqrt(var))).astype(np.float32)),
              what="norm apply+tanh")
        out64 = B.zeros((rows, cols), np.float64)
        ok(B, lib.mq_norm_apply(B.ptr(hx), code, rows, cols, B.ptr(mean), B.ptr(hvar), B.ptr(out64),
                                L.MQ_F64, 0, B.stream))
        assert np.allclose(B.down(out64), (x.astype(np.float64) - hmean) / np.sqrt(var), rtol=1e-13, atol=1e-13)
    # one statistics pass + update on the device against RunningNormalize itself (oracle = the reference's code path)
    for cols, dt, horizon, sizes in ((1, np.float32, 2, (1, 4099, 70001, 3)), (1, np.float64, None, (2, 513)),
                                     (11, np.float32, 10, (1, 1, 300)), (5, np.float64, 3, (64, 9))):
        code = L.MQ_F64 if dt == np.float64 else L.MQ_F32
        norm = O.RunningNormalizeO(shape=() if cols == 1 else (cols,), horizon=horizon)
        mean, var, stats = B.zeros(cols, np.float64), B.zeros(cols, np.float64), B.zeros(2 * cols, np.float64)
        tw_m = tw_v = 0.0
        n_seen = 0
        decay = 1.0 if horizon is None else 1.0 - 1.0 / horizon
        for rows in sizes:
            x = (rs.randn(rows, cols) * 3 + 10).astype(dt)
            norm.update(x if cols > 1 else x.reshape(-1))
            wsb = lib.mq_norm_stats_ws_bytes(rows, cols)
            hx, ws = B.up(x), B.zeros(wsb, np.uint8)
            ok(B, lib.mq_norm_stats(B.ptr(hx), code, rows, cols, B.ptr(mean), 1.0, B.ptr(stats), B.ptr(ws), wsb, B.stream))
            tw_m = tw_m * decay + 1.0
            coef_v, scale = 0.0, 1.0
            if n_seen >= 1:
                tw_v = tw_v * decay + 1.0
                coef_v = 1.0 / tw_v
            elif rows >= 2:
                w = (rows - 1.0) / rows
                tw_v = tw_v * decay ** w + w
                coef_v, scale = w / tw_v, 1.0 / w
            ok(B, lib.mq_norm_update_from_stats(B.ptr(mean), B.ptr(var), B.ptr(stats), cols, 1.0 / tw_m, coef_v, scale,
                                                B.stream))
            n_seen += rows
            assert np.allclose(B.down(mean), np.reshape(norm.get_mean(), -1), rtol=1e-12, atol=1e-12), "stats mean"
            if n_seen >= 2:
                assert np.allclose(B.down(var), np.reshape(norm.get_var(), -1), rtol=1e-10, atol=1e-12), "stats var"
    mean0, x = rs.randn(7), rs.randn(7)
    mean, hx, old = B.up(mean0.copy()), B.up(x), B.zeros(7, np.float64)
    ok(B, lib.mq_running_mean_update(B.ptr(mean), B.ptr(hx), L.MQ_F64, 7, 0.3, B.ptr(old), B.stream))
    assert np.array_equal(B.down(mean), mean0 + 0.3 * (x - mean0)) and np.array_equal(B.down(old), mean0)


def case_gather(B):
    """mannequin/_trajectory.py:79-86 t[idx]: bit-exact, NumPy negative-index rule."""
    lib = B.lib
    rs = np.random.RandomState(2)
    for cols, dt in ((11, np.float32), (12, np.float32), (3, np.float32), (1, np.float32), (5, np.uint8),
                     (4, np.int32)):
        src = (rs.randn(40, cols) * 100).astype(dt)
        idx = rs.randint(-40, 40, size=64).astype(np.int32)
        hs, hi, dst = B.up(src), B.up(idx), B.zeros((64, cols), dt)
        ok(B, lib.mq_gather_rows(B.ptr(hs), B.ptr(hi), 64, cols * src.itemsize, 40, B.ptr(dst), B.stream))
        assert np.array_equal(B.down(dst), src[idx])
    g = golden("trajectory")
    o = g["traj_o"].astype(np.float32)
    hs, hi, dst = B.up(o), B.up(g["traj_idx"].astype(np.int32)), B.zeros((64, 11), np.float32)
    ok(B, lib.mq_gather_rows(B.ptr(hs), B.ptr(hi), 64, 44, 40, B.ptr(dst), B.stream))
    assert np.array_equal(B.down(dst), g["traj_take_o"])
    assert lib.mq_gather_rows(B.ptr(hs), B.ptr(hi), 0, 44, 40, B.ptr(dst), B.stream) == 0     # empty index


def case_scan(B, sizes=(1, 2, 700, 2048, 2049, 3088, 4096, 4112, 5000, 70000, 70001)):
    """benchmarks/gae.py:43-64,69 and mannequin/_utils.py:9-20.  Sizes with n % 16 == 0 above 2048 take the hybrid register /
    shared-memory kernel (one CTA, whole CTAs, a ragged last CTA), n % 8 == 0 the register kernel, the rest the staged one."""
    lib = B.lib
    rs = np.random.RandomState(3)
    for n in sizes:
        r = rs.randn(n).astype(np.float32)
        v = rs.randn(n + 1).astype(np.float32)
        d = (rs.rand(n) < 0.02).astype(np.uint8)
        wsb = lib.mq_scan_ws_bytes(n)
        hr, hv, hd, ws = B.up(r), B.up(v), B.up(d), B.zeros(wsb, np.uint8)
        adv, ret = B.zeros(n, np.float32), B.zeros(n, np.float32)
        ok(B, lib.mq_gae(B.ptr(hr), B.ptr(hv), B.ptr(hd), n, 0.99, 0.95, B.ptr(adv), B.ptr(ret), B.ptr(ws),
                         wsb, B.stream))
        wa, wr = O.gae_advantages(r, v, d)
        close(B.down(adv), wa, what="gae adv n=%d" % n)
        close(B.down(ret), wr, what="gae ret n=%d" % n)
        # episode segmentation is integer-exact: integer data with gam=lam=1 has no rounding
        ri = rs.randint(-3, 4, size=n).astype(np.float32)
        vi = rs.randint(-3, 4, size=n + 1).astype(np.float32)
        d2 = (rs.rand(n) < 0.05).astype(np.uint8)
        hr, hv, hd = B.up(ri), B.up(vi), B.up(d2)
        ok(B, lib.mq_gae(B.ptr(hr), B.ptr(hv), B.ptr(hd), n, 1.0, 1.0, B.ptr(adv), B.ptr(ret), B.ptr(ws),
                         wsb, B.stream))
        wa, wr = O.gae_advantages(ri, vi, d2, 1.0, 1.0)
        assert np.array_equal(B.down(adv), wa) and np.array_equal(B.down(ret), wr), "segmentation n=%d" % n
        x = rs.randn(n)
        hx, out = B.up(x), B.zeros(n, np.float64)
        ok(B, lib.mq_discounted(B.ptr(hx), n, 500.0, B.ptr(out), B.ptr(ws), wsb, B.stream))
        assert np.allclose(B.down(out), O.discounted(x, 500), rtol=1e-11, atol=1e-13)
    # all-terminal and no-terminal rollouts
    n = 300
    r, v = rs.randn(n).astype(np.float32), rs.randn(n + 1).astype(np.float32)
    for d in (np.ones(n, np.uint8), np.zeros(n, np.uint8)):
        wsb = lib.mq_scan_ws_bytes(n)
        hr, hv, hd, ws = B.up(r), B.up(v), B.up(d), B.zeros(wsb, np.uint8)
        adv, ret = B.zeros(n, np.float32), B.zeros(n, np.float32)
        ok(B, lib.mq_gae(B.ptr(hr), B.ptr(hv), B.ptr(hd), n, 0.99, 0.95, B.ptr(adv), B.ptr(ret), B.ptr(ws),
                         wsb, B.stream))
        close(B.down(adv), O.gae_advantages(r, v, d)[0], what="gae edge")
    # no episode end at all over 41 CTAs of the hybrid kernel: every look-back walks a full 32-descriptor window of open
    # maps; and a single end in the middle
    n = 4096 * 40 + 16
    r, v = rs.randn(n).astype(np.float32), rs.randn(n + 1).astype(np.float32)
    for d in (np.zeros(n, np.uint8), (np.arange(n) == n // 2).astype(np.uint8)):
        wsb = lib.mq_scan_ws_bytes(n)
        hr, hv, hd, ws = B.up(r), B.up(v), B.up(d), B.zeros(wsb, np.uint8)
        adv, ret = B.zeros(n, np.float32), B.zeros(n, np.float32)
        ok(B, lib.mq_gae(B.ptr(hr), B.ptr(hv), B.ptr(hd), n, 0.99, 0.95, B.ptr(adv), B.ptr(ret), B.ptr(ws),
                         wsb, B.stream))
        wa, wr = O.gae_advantages(r, v, d)
        close(B.down(adv), wa, what="gae open look-back")
        close(B.down(ret), wr, what="gae open look-back (returns)")
    # reference output of benchmarks/gae.py:gae() (tests/golden/gae.npz), first chunk
    g = golden("gae")
    spec = O.mlp_spec(11, [64, 64, 1], [O.ACT_TANH, O.ACT_TANH, O.ACT_NONE])
    theta = O.mlp_init(spec, np.random.RandomState(79), head_scale=0.1)
    n = 2048
    value = O.mlp_forward(theta, spec, g["gae_obs"][:n + 1], dtype=np.float32)[-1].reshape(-1)
    hr, hv, hd = B.up(g["gae_rew"][:n].astype(np.float32)), B.up(value), B.up(g["gae_done"][:n].astype(np.uint8))
    adv, ret = B.zeros(n, np.float32), B.zeros(n, np.float32)
    ok(B, lib.mq_gae(B.ptr(hr), B.ptr(hv), B.ptr(hd), n, 0.97, 0.9, B.ptr(adv), B.ptr(ret), None, 0, B.stream))
    close(B.down(adv), g["gae_adv1"], rtol=2e-5, atol_scale=2e-6, what="gae vs reference chunk")
    assert lib.mq_discounted(B.ptr(hx), 5, 0.5, B.ptr(out), None, 0, B.stream) == L.MQ_ERR_SHAPE


def case_layered_mlp(B):
    lib = B.lib
    rs = np.random.RandomState(4)
    g = golden("nets")
    # reference outputs (tests/golden/nets.npz): LReLU net with image input, Tanh policy trunk
    for name, n_in, widths, acts in (("lrelu", 20, [16, 16, 5], [2, 2, 0]), ("tanh", 11, [64, 64, 3], [1, 1, 0])):
        spec = O.mlp_spec(n_in, widths, acts)
        net = net_of(spec)
        theta, x, dy = g[name + "_theta"], g[name + "_x"].reshape(-1, n_in), g[name + "_dy"]
        batch = len(x)
        ht, hx, hdy = B.up(theta), B.up(x), B.up(dy.copy())
        acts_buf = B.zeros(lib.mq_mlp_acts_floats(net, batch), np.float32)
        ok(B, lib.mq_mlp_forward(net, B.ptr(ht), B.ptr(hx), None, batch, B.ptr(acts_buf), B.stream))
        close(B.down(acts_buf)[-batch * widths[-1]:].reshape(batch, -1), g[name + "_y"].reshape(batch, -1),
              rtol=2e-5, what="layered fwd vs reference " + name)
        wsb = lib.mq_mlp_ws_bytes(net, batch)
        ws, grad, dx = B.zeros(wsb, np.uint8), B.zeros(spec["n_params"], np.float32), B.zeros((batch, n_in), np.float32)
        ok(B, lib.mq_mlp_backward(net, B.ptr(ht), B.ptr(hx), None, B.ptr(acts_buf), B.ptr(hdy), batch,
                                  1.0 / batch, B.ptr(grad), B.ptr(dx), B.ptr(ws), wsb, B.stream))
        close(B.down(grad), g[name + "_grad"], rtol=2e-5, atol_scale=2e-6, what="layered grad vs reference " + name)
        if name == "lrelu":
            close(B.down(dx), g["lrelu_dx"].reshape(batch, n_in), rtol=2e-5, what="layered dx vs reference")
    for n_in, widths, acts, batch in ((70, [33, 10], [2, 1], 65), (784, [128, 128, 10], [2, 2, 0], 128),
                                      (5, [7], [0], 1)):
        spec, theta, x = mlp_case(rs, n_in, widths, acts, batch)
        net = net_of(spec)
        ht, hx = B.up(theta), B.up(x)
        acts_buf = B.zeros(lib.mq_mlp_acts_floats(net, batch), np.float32)
        ok(B, lib.mq_mlp_forward(net, B.ptr(ht), B.ptr(hx), None, batch, B.ptr(acts_buf), B.stream))
        outs = O.mlp_forward(theta, spec, x)
        got, pos = B.down(acts_buf), 0
        for o in outs:
            close(got[pos:pos + o.size].reshape(o.shape), o, what="layered fwd %s" % (widths,))
            pos += o.size
        # the same forward with a workspace (split-K over the 784-long contraction) and gathered rows
        wsb = lib.mq_mlp_ws_bytes(net, batch)
        ws2, acts2 = B.zeros(wsb, np.uint8), B.zeros(lib.mq_mlp_acts_floats(net, batch), np.float32)
        perm = rs.permutation(batch).astype(np.int32)
        hp = B.up(perm)
        ok(B, lib.mq_mlp_forward_ws(net, B.ptr(ht), B.ptr(hx), B.ptr(hp), batch, B.ptr(acts2), B.ptr(ws2), wsb, B.stream))
        got2, pos = B.down(acts2), 0
        for o in outs:
            close(got2[pos:pos + o.size].reshape(o.shape), o[perm], what="layered fwd + workspace %s" % (widths,))
            pos += o.size
        dy = rs.randn(batch, widths[-1]).astype(np.float32)
        want, wdx = O.mlp_backward(theta, spec, x, outs, dy, want_dx=True)
        wsb = lib.mq_mlp_ws_bytes(net, batch)
        ws, grad, dx, hdy = B.zeros(wsb, np.uint8), B.zeros(spec["n_params"], np.float32), \
            B.zeros((batch, n_in), np.float32), B.up(dy.copy())
        ok(B, lib.mq_mlp_backward(net, B.ptr(ht), B.ptr(hx), None, B.ptr(acts_buf), B.ptr(hdy), batch,
                                  1.0 / batch, B.ptr(grad), B.ptr(dx), B.ptr(ws), wsb, B.stream))
        close(B.down(grad), want, what="layered grad %s" % (widths,))
        close(B.down(dx), wdx, what="layered dx %s" % (widths,))
    # gathered rows (Trajectory indexing fused into the load) + split-K over a long batch
    spec, theta, x = mlp_case(rs, 11, [64, 64, 1], [1, 1, 0], 300)
    idx = rs.randint(300, size=1024).astype(np.int32)
    net = net_of(spec)
    ht, hx, hi = B.up(theta), B.up(x), B.up(idx)
    acts_buf = B.zeros(lib.mq_mlp_acts_floats(net, 1024), np.float32)
    ok(B, lib.mq_mlp_forward(net, B.ptr(ht), B.ptr(hx), B.ptr(hi), 1024, B.ptr(acts_buf), B.stream))
    outs = O.mlp_forward(theta, spec, x[idx])
    dy = rs.randn(1024, 1).astype(np.float32)
    wsb = lib.mq_mlp_ws_bytes(net, 1024)
    ws, grad, hdy = B.zeros(wsb, np.uint8), B.zeros(spec["n_params"], np.float32), B.up(dy.copy())
    ok(B, lib.mq_mlp_backward(net, B.ptr(ht), B.ptr(hx), B.ptr(hi), B.ptr(acts_buf), B.ptr(hdy), 1024,
                              1.0 / 1024, B.ptr(grad), None, B.ptr(ws), wsb, B.stream))
    close(B.down(grad), O.mlp_backward(theta, spec, x[idx], outs, dy), what="layered grad gathered/split-K")


def case_fused_heads(B, batches=(64, 200)):
    lib = B.lib
    rs = np.random.RandomState(5)
    cases = [(11, [64, 64, 3], [1, 1, 0], b) for b in batches] + [
        (20, [16, 16, 5], [2, 2, 0], 7), (1, [64, 64, 1], [1, 1, 0], 128), (2, [2, 2], [0, 1], 3),
        (11, [64, 64, 1], [1, 1, 0], 300), (3, [5], [0], 1), (6, [128, 10], [2, 0], 33)]
    for n_in, widths, acts, batch in cases:
        spec, theta, x = mlp_case(rs, n_in, widths, acts, batch)
        net = net_of(spec)
        assert lib.mq_fused_tile_rows(net, 1) > 0
        n_out = widths[-1]
        ht, hx = B.up(theta), B.up(x)
        out = B.zeros((batch, n_out), np.float32)
        ok(B, lib.mq_fused_forward(net, B.ptr(ht), B.ptr(hx), None, batch, B.ptr(out), None, None, B.stream))
        outs = O.mlp_forward(theta, spec, x)
        close(B.down(out), outs[-1], what="fused fwd %s" % (widths,))
        wsb = lib.mq_fused_ws_bytes(net, batch)
        ws, grad = B.zeros(wsb, np.uint8), B.zeros(spec["n_params"], np.float32)
        dy = rs.randn(batch, n_out).astype(np.float32)
        hdy = B.up(dy)
        h = head(B, L.HEAD_DY, dy=hdy)
        out2 = B.zeros((batch, n_out), np.float32)
        ok(B, lib.mq_fused_grad(net, B.ptr(ht), B.ptr(hx), None, batch, ctypes.byref(h), 1.0 / batch,
                                B.ptr(grad), B.ptr(out2), None, B.ptr(ws), wsb, B.stream))
        close(B.down(grad), O.mlp_backward(theta, spec, x, outs, dy), what="fused grad DY %s b=%d" % (widths, batch))
        close(B.down(out2), outs[-1], what="fused grad out")
        tgt = rs.randn(batch, n_out).astype(np.float32)
        h = head(B, L.HEAD_DIFF, target=B.up(tgt))
        ok(B, lib.mq_fused_grad(net, B.ptr(ht), B.ptr(hx), None, batch, ctypes.byref(h), 1.0 / batch,
                                B.ptr(grad), None, None, B.ptr(ws), wsb, B.stream))
        close(B.down(grad), O.mlp_backward(theta, spec, x, outs, tgt - outs[-1]), what="fused grad DIFF")
        if n_out >= 2 and acts[-1] == 0:
            onehot = np.eye(n_out, dtype=np.float32)[rs.randint(n_out, size=batch)]
            h = head(B, L.HEAD_SOFTMAX_CE, p0=0.01, target=B.up(onehot))
            ok(B, lib.mq_fused_grad(net, B.ptr(ht), B.ptr(hx), None, batch, ctypes.byref(h), 1.0 / batch,
                                    B.ptr(grad), None, None, B.ptr(ws), wsb, B.stream))
            close(B.down(grad), O.mlp_backward(theta, spec, x, outs, O.softmax_ce_grad(outs[-1], onehot)),
                  what="fused grad CE")
    # reference fixtures through the fused path
    g = golden("nets")
    spec = O.mlp_spec(11, [64, 64, 3], [1, 1, 0])
    net = net_of(spec)
    ht, hx, hdy = B.up(g["tanh_theta"]), B.up(g["tanh_x"]), B.up(g["tanh_dy"])
    grad, out = B.zeros(spec["n_params"], np.float32), B.zeros((64, 3), np.float32)
    h = head(B, L.HEAD_DY, dy=hdy)
    ok(B, lib.mq_fused_grad(net, B.ptr(ht), B.ptr(hx), None, 64, ctypes.byref(h), 1.0 / 64, B.ptr(grad),
                            B.ptr(out), None, None, 0, B.stream))
    close(B.down(out), g["tanh_y"], rtol=2e-5, what="fused fwd vs reference")
    close(B.down(grad), g["tanh_grad"], rtol=2e-5, atol_scale=2e-6, what="fused grad vs reference")
    # a net too wide for shared memory is refused, not silently mis-handled
    big = L.make_net(784, [128, 128, 10], [2, 2, 0])
    assert lib.mq_fused_tile_rows(big, 1) == 0
    assert lib.mq_fused_grad(big, B.ptr(ht), B.ptr(hx), None, 64, ctypes.byref(h), 1.0, B.ptr(grad), None,
                             None, None, 0, B.stream) == L.MQ_ERR_UNSUPPORTED


def case_fused_gauss_ppo(B, minibatches=(64, 200)):
    lib = B.lib
    rs = np.random.RandomState(6)
    g = golden("gauss")
    spec = O.policy_spec(11, 3)
    net = L.make_net(11, [64, 64, 3], [1, 1, 0], logstd=True)
    theta, obs, act, w = g["gauss_theta"], g["gauss_obs"], g["gauss_act"], g["gauss_g"]
    ht, ho, ha, hw = B.up(theta), B.up(obs), B.up(act), B.up(w)
    logp = B.zeros(64, np.float32)
    ok(B, lib.mq_fused_forward(net, B.ptr(ht), B.ptr(ho), None, 64, None, B.ptr(ha), B.ptr(logp), B.stream))
    close(B.down(logp), g["gauss_logp"], what="fused logp vs reference")
    wsb = lib.mq_fused_ws_bytes(net, 4096)
    ws, grad = B.zeros(wsb, np.uint8), B.zeros(5126, np.float32)
    h = head(B, L.HEAD_GAUSS_LOGP, dy=hw, target=ha)
    ok(B, lib.mq_fused_grad(net, B.ptr(ht), B.ptr(ho), None, 64, ctypes.byref(h), 1.0 / 64, B.ptr(grad), None,
                            B.ptr(logp), B.ptr(ws), wsb, B.stream))
    close(B.down(grad), g["gauss_grad"], rtol=3e-5, atol_scale=3e-6, what="fused gauss grad vs reference")
    n = 300
    o = np.clip(rs.randn(n, 11), -5, 5).astype(np.float32)
    a = rs.randn(n, 3).astype(np.float32)
    adv = np.tanh(rs.randn(n)).astype(np.float32)
    lp_old = (O.policy_logprob(theta, spec, o, a)[0] + rs.randn(n) * 0.2).astype(np.float32)
    ho, ha, hadv, hlp = B.up(o), B.up(a), B.up(adv), B.up(lp_old)
    for mb in minibatches:
        idx = rs.randint(n, size=mb).astype(np.int32)
        hi = B.up(idx)
        h = head(B, L.HEAD_GAUSS_PPO, p0=0.8, p1=1.2, target=ha, adv=hadv, logp_old=hlp)
        ok(B, lib.mq_fused_grad(net, B.ptr(ht), B.ptr(ho), B.ptr(hi), mb, ctypes.byref(h), 1.0 / mb,
                                B.ptr(grad), None, None, B.ptr(ws), wsb, B.stream))
        lp, outs = O.policy_logprob(theta, spec, o[idx], a[idx])
        wgt = O.ppo_clip_weight(lp, lp_old[idx], adv[idx])
        assert (wgt == 0).any() and (wgt != 0).any()          # both branches of the clip rule
        close(B.down(grad), O.policy_logprob_grad(theta, spec, o[idx], a[idx], wgt, outs), rtol=3e-5,
              atol_scale=3e-6, what="fused ppo grad mb=%d" % mb)


def case_fused_train(B, steps=5):
    lib = B.lib
    rs = np.random.RandomState(7)
    spec = O.policy_spec(11, 3)
    net = L.make_net(11, [64, 64, 3], [1, 1, 0], logstd=True)
    theta0 = np.concatenate([O.mlp_init(O._mlp_part(spec), rs, 0.1), np.zeros(3, np.float32)])
    n, mb = 256, 64
    o = np.clip(rs.randn(n, 11), -5, 5).astype(np.float32)
    a = rs.randn(n, 3).astype(np.float32)
    adv = np.tanh(rs.randn(n)).astype(np.float32)
    lp_old = O.policy_logprob(theta0, spec, o, a)[0].astype(np.float32)
    idx = rs.randint(n, size=(steps, mb)).astype(np.int32)
    ho, ha, hadv, hlp, hi = B.up(o), B.up(a), B.up(adv), B.up(lp_old), B.up(idx)
    for dtype, code in ((np.float64, L.MQ_F64), (np.float32, L.MQ_F32)):
        opt = O.AdamO(theta0, horizon=10)
        th_ref = theta0.copy()
        for s in range(steps):
            th_ref = O.ppo_policy_step(th_ref, spec, opt, o[idx[s]], a[idx[s]], adv[idx[s]], lp_old[idx[s]])
        ht = B.up(theta0.copy())
        val, m, v = B.up(theta0.astype(dtype)), B.zeros(5126, dtype), B.zeros(5126, dtype)
        tw = (ctypes.c_double * 2)(0.0, 0.0)
        h = head(B, L.HEAD_GAUSS_PPO, p0=0.8, p1=1.2, target=ha, adv=hadv, logp_old=hlp)
        ok(B, lib.mq_fused_train(net, B.ptr(ht), B.ptr(val), B.ptr(m), B.ptr(v), code, B.ptr(ho),
                                 ctypes.byref(h), B.ptr(hi), steps, mb, 0.9, 0.99, tw, 3e-4, 1e-8, B.stream))
        theta = B.down(ht)
        assert abs(tw[0] - opt.tw1) < 1e-12 and abs(tw[1] - opt.tw2) < 1e-12
        # per-step error bar 2e-6 (fp32 gradient rounding feeding Adam), cf. test_oracle_golden
        assert np.max(np.abs(theta - th_ref)) < 2e-6 * (1 + steps) + 1e-6, np.max(np.abs(theta - th_ref))
        assert np.max(np.abs(theta - theta0)) > 1e-4
        assert np.array_equal(theta, B.down(val).astype(np.float32))
    vspec = O.mlp_spec(11, [64, 64, 1], [1, 1, 0])
    vnet = net_of(vspec)
    vt0 = O.mlp_init(vspec, rs, 0.1)
    ret = rs.randn(n, 1).astype(np.float32)
    opt = O.AdamO(vt0, horizon=10, lr=0.003)
    th_ref = vt0.copy()
    for s in range(steps):
        th_ref = O.value_sgd_step(th_ref, vspec, opt, o[idx[s]], ret[idx[s]])
    ht, hret = B.up(vt0.copy()), B.up(ret)
    val, m, v = B.up(vt0.astype(np.float64)), B.zeros(len(vt0), np.float64), B.zeros(len(vt0), np.float64)
    tw = (ctypes.c_double * 2)(0.0, 0.0)
    h = head(B, L.HEAD_DIFF, target=hret)
    ok(B, lib.mq_fused_train(vnet, B.ptr(ht), B.ptr(val), B.ptr(m), B.ptr(v), L.MQ_F64, B.ptr(ho),
                             ctypes.byref(h), B.ptr(hi), steps, mb, 0.9, 0.99, tw, 0.003, 1e-8, B.stream))
    err = np.max(np.abs(B.down(ht) - th_ref))
    assert err < 2e-5 * (1 + steps), err


def case_ppo_reference_replay(B, n_steps=300):
    """Replays the policy half of benchmarks/ppo.py:run() recorded from the REAL
    reference (tests/golden/ppo.npz, first chunk) through mq_fused_train."""
    lib = B.lib
    g = golden("ppo")
    n = 2048
    pol_spec = O.policy_spec(11, 3)
    val_spec = O.mlp_spec(11, [64, 64, 1], [1, 1, 0])
    obs, act, rew, done = g["ppo_obs"], g["ppo_act"], g["ppo_rew"], g["ppo_done"]
    # value half on the device: forward, GAE, 300 value steps
    vnet, pnet = net_of(val_spec), L.make_net(11, [64, 64, 3], [1, 1, 0], logstd=True)
    hvt = B.up(g["ppo_val_init"].copy())
    ho = B.up(obs[:n + 1].copy())
    value = B.zeros(n + 1, np.float32)
    ok(B, lib.mq_fused_forward(vnet, B.ptr(hvt), B.ptr(ho), None, n + 1, B.ptr(value), None, None, B.stream))
    hr, hd = B.up(rew[:n].astype(np.float32)), B.up(done[:n].astype(np.uint8))
    adv, ret = B.zeros(n, np.float32), B.zeros(n, np.float32)
    ok(B, lib.mq_gae(B.ptr(hr), B.ptr(value), B.ptr(hd), n, 0.99, 0.95, B.ptr(adv), B.ptr(ret), None, 0, B.stream))
    vval, vm, vv = B.up(g["ppo_val_init"].astype(np.float64)), B.zeros(4993, np.float64), B.zeros(4993, np.float64)
    tw = (ctypes.c_double * 2)(0.0, 0.0)
    hvi = B.up(g["ppo_val_idx"][:n_steps].copy())
    h = head(B, L.HEAD_DIFF, target=ret)
    ok(B, lib.mq_fused_train(vnet, B.ptr(hvt), B.ptr(vval), B.ptr(vm), B.ptr(vv), L.MQ_F64, B.ptr(ho),
                             ctypes.byref(h), B.ptr(hvi), n_steps, 64, 0.9, 0.99, tw, 0.003, 1e-8, B.stream))
    ck = {int(s): i for i, s in enumerate(g["ppo_ck_steps"])}
    tol = 2e-6 * (1 + n_steps) + 1e-6
    if n_steps - 1 in ck:
        err = np.max(np.abs(B.down(hvt) - g["ppo_val_ck"][ck[n_steps - 1]]))
        assert err < tol, ("value net vs reference", err)
    # policy half: normalise (host-known stats -> device apply), old log-probs, 300 PPO steps
    adv_h = B.down(adv)
    norm = O.RunningNormalizeO(horizon=2)
    norm.update(adv_h)
    mean, var = B.up(np.array([norm.get_mean()]).reshape(1)), B.up(np.array([norm.get_var()]).reshape(1))
    adv_n = B.zeros(n, np.float32)
    ok(B, lib.mq_norm_apply(B.ptr(adv), L.MQ_F32, n, 1, B.ptr(mean), B.ptr(var), B.ptr(adv_n), L.MQ_F32, 1, B.stream))
    hpt, ha = B.up(g["ppo_pol_init"].copy()), B.up(act[:n].copy())
    lp_old = B.zeros(n, np.float32)
    ok(B, lib.mq_fused_forward(pnet, B.ptr(hpt), B.ptr(ho), None, n, None, B.ptr(ha), B.ptr(lp_old), B.stream))
    pval, pm, pv = B.up(g["ppo_pol_init"].astype(np.float64)), B.zeros(5126, np.float64), B.zeros(5126, np.float64)
    tw = (ctypes.c_double * 2)(0.0, 0.0)
    hpi = B.up(g["ppo_pol_idx"][:n_steps].copy())
    h = head(B, L.HEAD_GAUSS_PPO, p0=0.8, p1=1.2, target=ha, adv=adv_n, logp_old=lp_old)
    ok(B, lib.mq_fused_train(pnet, B.ptr(hpt), B.ptr(pval), B.ptr(pm), B.ptr(pv), L.MQ_F64, B.ptr(ho),
                             ctypes.byref(h), B.ptr(hpi), n_steps, 64, 0.9, 0.99, tw, 3e-4, 1e-8, B.stream))
    if n_steps - 1 in ck:
        err = np.max(np.abs(B.down(hpt) - g["ppo_pol_ck"][ck[n_steps - 1]]))
        assert err < tol, ("policy vs reference", err)
    return B.down(hpt), B.down(hvt)


def case_es(B, members=6, n_obs=150):
    lib = B.lib
    rs = np.random.RandomState(8)
    spec = O.mlp_spec(11, [64, 64, 3], [1, 1, 0])
    net = net_of(spec)
    theta = O.mlp_init(spec, rs)
    noise = (rs.randn(members, spec["n_params"]) * 0.1).astype(np.float32)
    obs = np.clip(rs.randn(n_obs, 11), -5, 5).astype(np.float32)
    coef = np.array([1.0, -0.5, 0.25], np.float32)
    ht, hn, ho, hc = B.up(theta), B.up(noise), B.up(obs), B.up(coef)
    rets = B.zeros(members, np.float32)
    ok(B, lib.mq_es_eval(net, B.ptr(ht), B.ptr(hn), members, B.ptr(ho), n_obs, B.ptr(hc), B.ptr(rets), B.stream))
    want = np.zeros(members)
    for p in range(members):
        th = (theta.astype(np.float64) + noise[p]).astype(np.float32)
        want[p] = np.sum(O.mlp_forward(th, spec, obs)[-1] @ coef.astype(np.float64))
    close(B.down(rets), want, rtol=2e-5, atol_scale=1e-5, what="es returns")
    w = rs.randn(members)
    wsb = lib.mq_es_ws_bytes(members, spec["n_params"])
    hw, ws, out = B.up(w), B.zeros(wsb, np.uint8), B.zeros(spec["n_params"], np.float64)
    ok(B, lib.mq_es_weighted_sum(B.ptr(hn), B.ptr(hw), members, spec["n_params"], 1.0 / members, B.ptr(out),
                                 B.ptr(ws), wsb, B.stream))
    assert np.allclose(B.down(out), (noise.astype(np.float64) * w[:, None]).mean(axis=0), rtol=1e-12, atol=1e-15)


def case_ops(B):
    lib = B.lib
    rs = np.random.RandomState(9)
    a = rs.randn(33, 70).astype(np.float32)
    b = rs.randn(70, 21).astype(np.float32)
    ha, hb, c = B.up(a), B.up(b), B.zeros((33, 21), np.float32)
    ok(B, lib.mq_gemm(B.ptr(ha), 70, 1, B.ptr(hb), 21, 1, B.ptr(c), 33, 21, 70, 1.0, B.stream))
    close(B.down(c), a.astype(np.float64) @ b, what="gemm NN")
    ct = B.zeros((70, 70), np.float32)
    ok(B, lib.mq_gemm(B.ptr(ha), 1, 70, B.ptr(ha), 70, 1, B.ptr(ct), 70, 70, 33, 0.5, B.stream))
    close(B.down(ct), 0.5 * a.T.astype(np.float64) @ a, what="gemm TN")
    cn = B.zeros((33, 33), np.float32)
    ok(B, lib.mq_gemm(B.ptr(ha), 70, 1, B.ptr(ha), 1, 70, B.ptr(cn), 33, 33, 70, 1.0, B.stream))
    close(B.down(cn), a.astype(np.float64) @ a.T, what="gemm NT")
    g = rs.randn(50, 7).astype(np.float32)
    x = rs.randn(50, 7).astype(np.float32)
    hg, hx, out = B.up(g), B.up(x), B.zeros(7, np.float32)
    ok(B, lib.mq_col_sum(B.ptr(hg), B.ptr(hx), 50, 7, 0.5, B.ptr(out), B.stream))
    close(B.down(out), 0.5 * (g * x).sum(axis=0), what="col_sum")
    for with_a in (True, False):                            # tall matrices: the 8-column x 32-row-lane kernel
        gt, xt = rs.randn(1500, 77).astype(np.float32), rs.randn(1500, 77).astype(np.float32)
        hgt, hxt, outt = B.up(gt), B.up(xt), B.zeros(77, np.float32)
        ok(B, lib.mq_col_sum(B.ptr(hgt), B.ptr(hxt) if with_a else None, 1500, 77, 0.25, B.ptr(outt), B.stream))
        close(B.down(outt), 0.25 * (gt.astype(np.float64) * (xt if with_a else 1.0)).sum(axis=0), what="col_sum tall")
    y = B.zeros((50, 7), np.float32)
    ok(B, lib.mq_binary(B.ptr(hx), B.ptr(out), B.ptr(y), 350, 7, 1, B.stream))
    assert np.array_equal(B.down(y), x * B.down(out))
    ok(B, lib.mq_add_bias(B.ptr(hx), B.ptr(out), B.ptr(y), 50, 7, B.stream))
    assert np.array_equal(B.down(y), x + B.down(out))
    for act, leak in ((L.ACT_TANH, 0.0), (L.ACT_LRELU, 0.1), (L.ACT_LRELU, 0.0)):
        ok(B, lib.mq_act_forward(B.ptr(hx), B.ptr(y), 350, act, leak, B.stream))
        want = np.tanh(x) if act == L.ACT_TANH else np.where(x < 0, x * np.float32(leak), x)
        close(B.down(y), want, what="act fwd")
        gy = B.zeros((50, 7), np.float32)
        ok(B, lib.mq_act_backward(B.ptr(y), B.ptr(hg), B.ptr(gy), 350, act, leak, B.stream))
        wantg = g * (1 - np.tanh(x) ** 2) if act == L.ACT_TANH else np.where(x < 0, g * np.float32(leak), g)
        close(B.down(gy), wantg, what="act bwd")
    xv, gv = rs.randn(64, 8).astype(np.float32), rs.randn(64, 8).astype(np.float32)      # n % 4 == 0: 128-bit kernels
    hxv, hgv, yv, ov = B.up(xv), B.up(gv), B.zeros((64, 8), np.float32), B.zeros((64, 8), np.float32)
    ok(B, lib.mq_act_forward(B.ptr(hxv), B.ptr(yv), 512, L.ACT_TANH, 0.0, B.stream))
    close(B.down(yv), np.tanh(xv), what="act fwd x4")
    ok(B, lib.mq_act_backward(B.ptr(yv), B.ptr(hgv), B.ptr(ov), 512, L.ACT_TANH, 0.0, B.stream))
    close(B.down(ov), gv * (1 - np.tanh(xv) ** 2), what="act bwd x4")
    ok(B, lib.mq_clip_forward(B.ptr(hx), B.ptr(y), 350, -0.5, 0.5, B.stream))
    assert np.array_equal(B.down(y), np.clip(x, -0.5, 0.5))
    ok(B, lib.mq_clip_backward(B.ptr(hx), B.ptr(hg), B.ptr(y), 350, -0.5, 0.5, B.stream))
    assert np.array_equal(B.down(y), O.clip_backward(x, g, -0.5, 0.5).astype(np.float32))
    mu, ls, s = rs.randn(9, 3).astype(np.float32), (rs.randn(3) * 0.3).astype(np.float32), rs.randn(9, 3).astype(np.float32)
    hmu, hls, hs, lp = B.up(mu), B.up(ls), B.up(s), B.zeros(9, np.float32)
    ok(B, lib.mq_gauss_logprob(B.ptr(hmu), B.ptr(hls), 3, B.ptr(hs), 9, 3, B.ptr(lp), B.stream))
    close(B.down(lp), O.gauss_logprob(mu, ls, s), what="gauss logp")
    # wide actions with a per-row logstd (the decoder of mnist/vae.py): warp-per-row kernel
    muw, lsw, sw = rs.randn(21, 100).astype(np.float32), (rs.randn(21, 100) * 0.3).astype(np.float32), \
        rs.randn(21, 100).astype(np.float32)
    hmw, hlw, hsw, lpw = B.up(muw), B.up(lsw), B.up(sw), B.zeros(21, np.float32)
    ok(B, lib.mq_gauss_logprob(B.ptr(hmw), B.ptr(hlw), 2100, B.ptr(hsw), 21, 100, B.ptr(lpw), B.stream))
    close(B.down(lpw), O.gauss_logprob(muw, lsw, sw), rtol=2e-5, what="gauss logp wide")
    w = rs.randn(9).astype(np.float32)
    hw, dmu, dls = B.up(w), B.zeros((9, 3), np.float32), B.zeros((9, 3), np.float32)
    ok(B, lib.mq_gauss_logprob_bwd(B.ptr(hmu), B.ptr(hls), 3, B.ptr(hs), B.ptr(hw), 9, 3, B.ptr(dmu), B.ptr(dls), B.stream))
    wm, wl = O.gauss_logprob_grads(mu, ls, s, w)
    close(B.down(dmu), wm, what="gauss dmu")
    close(B.down(dls), wl, what="gauss dls")
    unit = rs.randn(9, 3).astype(np.float32)
    hu, so, nz = B.up(unit), B.zeros((9, 3), np.float32), B.zeros((9, 3), np.float32)
    ok(B, lib.mq_gauss_sample(B.ptr(hmu), B.ptr(hls), 3, B.ptr(hu), 9, 3, B.ptr(so), B.ptr(nz), B.stream))
    ws_, wn_ = O.gauss_sample(mu, ls, unit)
    close(B.down(so), ws_, what="gauss sample")
    close(B.down(nz), wn_, what="gauss noise")
    m, l = rs.randn(5, 3).astype(np.float32), (rs.randn(5, 3) * 0.5).astype(np.float32)
    hm, hl, o5 = B.up(m), B.up(l), B.zeros(5, np.float32)
    ok(B, lib.mq_dkl_forward(B.ptr(hm), B.ptr(hl), 5, 3, B.ptr(o5), B.stream))
    close(B.down(o5), O.dkl_uninormal(m, l), what="dkl")
    w5 = rs.randn(5).astype(np.float32)
    hw5, dm, dl = B.up(w5), B.zeros((5, 3), np.float32), B.zeros((5, 3), np.float32)
    ok(B, lib.mq_dkl_backward(B.ptr(hm), B.ptr(hl), B.ptr(hw5), 5, 3, B.ptr(dm), B.ptr(dl), B.stream))
    wdm, wdl = O.dkl_uninormal_grads(m, l, w5)
    close(B.down(dm), wdm, what="dkl dmean")
    close(B.down(dl), wdl, what="dkl dlogstd")
    lg = rs.randn(6, 10).astype(np.float32)
    oh = np.eye(10, dtype=np.float32)[rs.randint(10, size=6)]
    hlg, hoh, dy = B.up(lg), B.up(oh), B.zeros((6, 10), np.float32)
    ok(B, lib.mq_softmax_ce_grad(B.ptr(hlg), B.ptr(hoh), None, 6, 10, 0.01, B.ptr(dy), B.stream))
    close(B.down(dy), O.softmax_ce_grad(lg, oh), what="softmax ce")
    lpn, lpo, adv = rs.randn(40).astype(np.float32) * 0.3, rs.randn(50).astype(np.float32) * 0.3, rs.randn(50).astype(np.float32)
    idx = rs.randint(50, size=40).astype(np.int32)
    h1, h2, h3, h4, wo = B.up(lpn), B.up(lpo), B.up(adv), B.up(idx), B.zeros(40, np.float32)
    ok(B, lib.mq_ppo_weight(B.ptr(h1), B.ptr(h2), B.ptr(h3), B.ptr(h4), 40, 0.8, 1.2, B.ptr(wo), B.stream))
    close(B.down(wo), O.ppo_clip_weight(lpn, lpo[idx], adv[idx]), what="ppo weight")
    # error contract: bad arguments return codes, never crash
    assert lib.mq_act_forward(B.ptr(hx), B.ptr(y), 350, 7, 0.0, B.stream) == L.MQ_ERR_INVALID
    assert lib.mq_binary(B.ptr(hx), B.ptr(out), B.ptr(y), 350, 8, 0, B.stream) == L.MQ_ERR_SHAPE
    assert b"broadcast" in lib.mq_last_error()


def case_entropy(B, batches=(64, 300, 9000), steps=4):
    """Gaussian entropy (named by the north star, absent from the reference -> pinned analytically): the
    per-op kernel, and the entropy coefficient of the Gaussian heads in the FFMA tile kernel, the tcgen05 kernel
    (9000 rows) and the persistent chain kernel: d/dlogstd gains exactly +ent, nothing else moves."""
    lib = B.lib
    rs = np.random.RandomState(47)
    ls = (rs.randn(37, 5) * 0.3).astype(np.float32)
    gg = rs.randn(37).astype(np.float32)
    h, dls = B.zeros(37, np.float32), B.zeros((37, 5), np.float32)
    hls, hg = B.up(ls), B.up(gg)
    ok(B, lib.mq_gauss_entropy(B.ptr(hls), None, 37, 5, B.ptr(h), None, B.stream))
    ok(B, lib.mq_gauss_entropy(None, B.ptr(hg), 37, 5, None, B.ptr(dls), B.stream))
    close(B.down(h), O.gauss_entropy(ls, 5), what="gauss entropy")
    assert np.array_equal(B.down(dls), np.repeat(gg[:, None], 5, axis=1))
    eps = 1e-3                                              # the closed form against central differences
    fd = (O.gauss_entropy(ls.astype(np.float64) + eps, 5) - O.gauss_entropy(ls.astype(np.float64) - eps, 5)) / (2 * eps)
    assert np.max(np.abs(fd - 5.0)) < 1e-8                  # d/d(all logstd) = A
    spec = O.policy_spec(11, 3)
    net = L.make_net(11, [64, 64, 3], [1, 1, 0], logstd=True)
    theta = np.concatenate([O.mlp_init(O._mlp_part(spec), rs, 0.1), (rs.randn(3) * 0.2).astype(np.float32)])
    n = 12000
    o = np.clip(rs.randn(n, 11), -5, 5).astype(np.float32)
    a = rs.randn(n, 3).astype(np.float32)
    adv = np.tanh(rs.randn(n)).astype(np.float32)
    lp_old = (O.policy_logprob(theta, spec, o, a)[0] + rs.randn(n) * 0.2).astype(np.float32)
    ht, ho, ha, hadv, hlp = B.up(theta), B.up(o), B.up(a), B.up(adv), B.up(lp_old)
    ent = 0.037
    for mb in batches:
        idx = rs.randint(n, size=mb).astype(np.int32)
        hi = B.up(idx)
        wsb = lib.mq_fused_ws_bytes(net, mb)
        ws = B.zeros(wsb, np.uint8)
        grads = []
        for e in (0.0, ent):
            grad = B.zeros(5126, np.float32)
            hd = head(B, L.HEAD_GAUSS_PPO, p0=0.8, p1=1.2, target=ha, adv=hadv, logp_old=hlp)
            hd.ent = e
            ok(B, lib.mq_fused_grad(net, B.ptr(ht), B.ptr(ho), B.ptr(hi), mb, ctypes.byref(hd), 1.0 / mb,
                                    B.ptr(grad), None, None, B.ptr(ws), wsb, B.stream))
            grads.append(B.down(grad))
        lp, outs = O.policy_logprob(theta, spec, o[idx], a[idx])
        want = np.array(O.policy_logprob_grad(theta, spec, o[idx], a[idx], O.ppo_clip_weight(lp, lp_old[idx], adv[idx]),
                                              outs), dtype=np.float64)
        close(grads[0], want, rtol=3e-5, atol_scale=3e-6, what="ppo grad, ent=0, mb=%d" % mb)
        want[5123:] += ent
        close(grads[1], want, rtol=3e-5, atol_scale=3e-6, what="ppo grad with entropy bonus, mb=%d" % mb)
        assert np.array_equal(grads[0][:5123], grads[1][:5123])
        assert np.max(np.abs((grads[1][5123:] - grads[0][5123:]) - ent)) < 1e-6
    # chain kernel: `steps` Adam steps with the bonus against the oracle
    mb = 64
    idx = rs.randint(n, size=(steps, mb)).astype(np.int32)
    opt = O.AdamO(theta, horizon=10)
    th_ref = theta.copy()
    for s_ in range(steps):
        th_ref = O.ppo_policy_step(th_ref, spec, opt, o[idx[s_]], a[idx[s_]], adv[idx[s_]], lp_old[idx[s_]], ent_coef=ent)
    ht2, hi = B.up(theta.copy()), B.up(idx)
    val, m, v = B.up(theta.astype(np.float64)), B.zeros(5126, np.float64), B.zeros(5126, np.float64)
    tw = (ctypes.c_double * 2)(0.0, 0.0)
    hd = head(B, L.HEAD_GAUSS_PPO, p0=0.8, p1=1.2, target=ha, adv=hadv, logp_old=hlp)
    hd.ent = ent
    ok(B, lib.mq_fused_train(net, B.ptr(ht2), B.ptr(val), B.ptr(m), B.ptr(v), L.MQ_F64, B.ptr(ho), ctypes.byref(hd),
                             B.ptr(hi), steps, mb, 0.9, 0.99, tw, 3e-4, 1e-8, B.stream))
    got = B.down(ht2)
    assert np.max(np.abs(got - th_ref)) < 2e-6 * (1 + steps) + 1e-6, np.max(np.abs(got - th_ref))
    assert np.all(got[5123:] > theta[5123:] - 1e-3)


def case_rollout(B, n_env=300):
    """Rollout-side kernels: Philox normals, the observation normaliser against the real
    reference's NormalizedObservations / RunningNormalize (golden/obsnorm.npz) and the oracle, the batched actor /
    synthetic environment step and the segment bootstrap against the oracle."""
    lib = B.lib
    rs = np.random.RandomState(53)
    seed = 0x1234567ABCDEF
    # ---- Philox normals
    n = 1003
    z = B.zeros(n, np.float32)
    ok(B, lib.mq_randn_philox(seed, 7, 9, None, n, B.ptr(z), B.stream))
    want = O.philox_normals(seed, 7, 9, np.arange(n))
    assert np.max(np.abs(B.down(z) - want)) < 1e-5
    ctr = B.up(np.array([5], np.int32))                     # device-resident stream offset
    ok(B, lib.mq_randn_philox(seed, 7, 4, B.ptr(ctr), n, B.ptr(z), B.stream))
    assert np.max(np.abs(B.down(z) - want)) < 1e-5
    # ---- NormalizedObservations: the reference's wrapper, one observation per step (E = 1)
    g = golden("obsnorm")
    st = B.zeros(2 * 11 + 3, np.float64)
    for raw, want in zip(g["on_raw"], g["on_out"]):
        hr, ho = B.up(raw.astype(np.float32).reshape(1, 11)), B.zeros((1, 11), np.float32)
        ok(B, lib.mq_obsnorm_step(B.ptr(hr), 1, 11, 0.0, B.ptr(st), 5.0, B.ptr(ho), B.stream))
        ref = O.normalized_observations_step  # noqa: F841  (oracle form checked below)
        assert np.max(np.abs(B.down(ho).reshape(-1) - want)) < 2e-5 * (1 + np.max(np.abs(want))), "obsnorm E=1"
    sd = B.down(st)
    # the reference saw float64 observations, the kernel their float32 roundings: statistics agree to ~1e-7 relative
    assert np.max(np.abs(sd[:11] - g["on_mean"]) / (1 + np.abs(g["on_mean"]))) < 1e-6
    assert np.max(np.abs(sd[11:22] - g["on_var"]) / g["on_var"]) < 1e-5
    # ---- whole batches (the lock-step form), float32 inputs on both sides: fp64 statistics to 1e-12
    st = B.zeros(2 * 11 + 3, np.float64)
    for b, want in zip(g["on_batches"], g["on_batch_out"]):
        hr, ho = B.up(b), B.zeros(b.shape, np.float32)
        ok(B, lib.mq_obsnorm_step(B.ptr(hr), b.shape[0], 11, 0.0, B.ptr(st), 5.0, B.ptr(ho), B.stream))
        close(B.down(ho), want, rtol=1e-6, atol_scale=1e-6, what="obsnorm batch vs reference")
    sd = B.down(st)
    assert np.max(np.abs(sd[:11] - g["on_batch_mean"])) < 1e-12 * (1 + np.max(np.abs(g["on_batch_mean"])))
    assert np.max(np.abs(sd[11:22] - g["on_batch_var"]) / g["on_batch_var"]) < 1e-12
    # ---- oracle: horizons, widths, batch sizes incl. a first batch of one sample
    for d, horizon, sizes in ((11, None, (1, 1, 5, 4096, 2)), (3, 10.0, (64, 64, 1, 9)), (64, 3.0, (2, 1500))):
        norm = O.RunningNormalizeO(shape=(d,), horizon=horizon)
        st = B.zeros(2 * d + 3, np.float64)
        for m in sizes:
            raw = (rs.randn(m, d) * 3.0 + 10.0).astype(np.float32)
            want = O.normalized_observations_step(norm, raw, 5.0)
            hr, ho = B.up(raw), B.zeros((m, d), np.float32)
            ok(B, lib.mq_obsnorm_step(B.ptr(hr), m, d, horizon or 0.0, B.ptr(st), 5.0, B.ptr(ho), B.stream))
            close(B.down(ho), want, rtol=1e-6, atol_scale=1e-6, what="obsnorm vs oracle d=%d m=%d" % (d, m))
        sd = B.down(st)
        assert np.max(np.abs(sd[:d] - norm.get_mean())) < 1e-11
        assert np.max(np.abs(sd[d:2 * d] - norm.get_var()) / norm.get_var()) < 1e-11
    # ---- actor + synthetic environment
    d, na, t_len = 11, 3, 5
    env = dict(dt=0.1, damp=0.5, sigma=0.05, ctrl=0.1, p_end=0.05, horizon=4,
               b=(rs.randn(d, na) / np.sqrt(na)).astype(np.float32), obs_scale=np.exp(rs.randn(d) * 0.5).astype(np.float32),
               obs_shift=(rs.randn(d) * 2).astype(np.float32))
    hb, hs, hh = B.up(env["b"]), B.up(env["obs_scale"]), B.up(env["obs_shift"])
    ed = L.MqEnv()
    ed.d, ed.a, ed.horizon, ed.dt, ed.damp, ed.sigma, ed.ctrl, ed.p_end = d, na, 4, 0.1, 0.5, 0.05, 0.1, 0.05
    ed.b, ed.obs_scale, ed.obs_shift = B.ptr(hb), B.ptr(hs), B.ptr(hh)
    hx, hep, hraw = B.zeros((n_env, d), np.float32), B.zeros(n_env, np.int32), B.zeros((n_env, d), np.float32)
    ok(B, lib.mq_env_reset(ctypes.byref(ed), n_env, seed, B.ptr(hx), B.ptr(hep), B.ptr(hraw), B.stream))
    e_idx = np.arange(n_env)
    x = np.stack([O.philox_normals(seed, 0xFFFFFFFE, e_idx, i) for i in range(d)], axis=1)
    assert np.max(np.abs(B.down(hx) - x)) < 1e-5
    close(B.down(hraw), env["obs_scale"] * x + env["obs_shift"], what="env reset raw")
    x, ep = B.down(hx).copy(), np.zeros(n_env, np.int32)
    obs_buf, act_buf = B.zeros((n_env * t_len, d), np.float32), B.zeros((n_env * t_len, na), np.float32)
    rew_buf, done_buf = B.zeros(n_env * t_len, np.float32), B.zeros(n_env * t_len, np.uint8)
    base = B.up(np.array([100], np.int32))
    logstd = (rs.randn(na) * 0.3).astype(np.float32)
    hls = B.up(logstd)
    n_done = 0
    for t in range(t_len):
        obs_n = rs.randn(n_env, d).astype(np.float32)
        mean = rs.randn(n_env, na).astype(np.float32)
        ho, hm = B.up(obs_n), B.up(mean)
        ok(B, lib.mq_actor_env_step(ctypes.byref(ed), n_env, t, t_len, seed, t, B.ptr(base), B.ptr(ho), B.ptr(hm), B.ptr(hls),
                                    B.ptr(hx), B.ptr(hep), B.ptr(obs_buf), B.ptr(act_buf), B.ptr(rew_buf), B.ptr(done_buf),
                                    B.ptr(hraw), B.stream))
        sid = 100 + t
        noise = np.stack([O.philox_normals(seed, sid, e_idx, k) for k in range(na)], axis=1)
        act = (mean + np.exp(logstd) * noise).astype(np.float32)                 # distrib.py:50-59
        x, ep, rew, done, raw = O.synthetic_env_step(env, x, ep, act, seed, sid)
        rows = e_idx * t_len + t
        assert np.array_equal(B.down(obs_buf)[rows], obs_n)
        assert np.max(np.abs(B.down(act_buf)[rows] - act)) < 1e-5
        assert np.max(np.abs(B.down(rew_buf)[rows] - rew)) < 1e-4 * (1 + np.max(np.abs(rew)))
        assert np.array_equal(B.down(done_buf)[rows].astype(bool), done)
        assert np.array_equal(B.down(hep), ep)
        assert np.max(np.abs(B.down(hx) - x)) < 2e-5
        assert np.max(np.abs(B.down(hraw) - raw)) < 1e-4
        x = B.down(hx).copy()                                                    # no drift between the two sides
        n_done += int(done.sum())
    assert 0 < n_done < n_env * t_len                                            # both endings and continuations seen
    # ---- segment bootstrap
    n_seg, seg = 37, 6
    rew = rs.randn(n_seg * seg).astype(np.float32)
    done = (rs.rand(n_seg * seg) < 0.3)
    vn = rs.randn(n_seg).astype(np.float32)
    hr, hd, hv = B.up(rew), B.up(done.astype(np.uint8)), B.up(vn)
    ok(B, lib.mq_bootstrap_segments(B.ptr(hr), B.ptr(hd), B.ptr(hv), n_seg, seg, 0.99, B.stream))
    wr, wd = O.bootstrap_segments(rew, done, vn, seg, 0.99)
    assert np.max(np.abs(B.down(hr) - wr)) < 1e-6 and np.array_equal(B.down(hd).astype(bool), wd)


def case_tc_paths(B):
    """tcgen05 / TMEM kernel of mq_tc.cu (separate arrays, exact mode) forced on per call (MQ_TC_FORCE) for every net it
    accepts: the same fp32 bars as the FFMA kernels (three-plane bf16 operands keep 24 mantissa bits)."""
    lib = B.lib
    rs = np.random.RandomState(15)
    F = L.TC_FORCE
    cases = [(11, [64, 64, 3], [1, 1, 0], 64), (11, [64, 64, 3], [1, 1, 0], 1000),
             (11, [64, 64, 1], [1, 1, 0], 40000), (20, [16, 16, 5], [2, 2, 0], 777),
             (6, [32, 10], [2, 0], 333), (17, [64, 48, 6], [1, 2, 1], 2500), (20, [48, 7], [1, 0], 1500)]
    for n_in, widths, acts, batch in cases:
        spec, theta, x = mlp_case(rs, n_in, widths, acts, batch)
        net = net_of(spec)
        n_out = widths[-1]
        ht, hx = B.up(theta), B.up(x)
        out = B.zeros((batch, n_out), np.float32)
        ok(B, lib.mq_fused_forward_tc(net, B.ptr(ht), B.ptr(hx), None, batch, B.ptr(out), None, None, F, B.stream))
        outs = O.mlp_forward(theta, spec, x)
        close(B.down(out), outs[-1], what="tc fwd %s b=%d" % (widths, batch))
        wsb = lib.mq_fused_ws_bytes(net, batch)
        ws, grad = B.zeros(wsb, np.uint8), B.zeros(spec["n_params"], np.float32)
        dy = rs.randn(batch, n_out).astype(np.float32)
        h = head(B, L.HEAD_DY, dy=B.up(dy), tc=F)
        out2 = B.zeros((batch, n_out), np.float32)
        ok(B, lib.mq_fused_grad(net, B.ptr(ht), B.ptr(hx), None, batch, ctypes.byref(h), 1.0 / batch,
                                B.ptr(grad), B.ptr(out2), None, B.ptr(ws), wsb, B.stream))
        close(B.down(grad), O.mlp_backward(theta, spec, x, outs, dy), what="tc grad DY %s b=%d" % (widths, batch))
        close(B.down(out2), outs[-1], what="tc grad out")
        tgt = rs.randn(batch, n_out).astype(np.float32)
        h = head(B, L.HEAD_DIFF, target=B.up(tgt), tc=F)
        ok(B, lib.mq_fused_grad(net, B.ptr(ht), B.ptr(hx), None, batch, ctypes.byref(h), 1.0 / batch,
                                B.ptr(grad), None, None, B.ptr(ws), wsb, B.stream))
        close(B.down(grad), O.mlp_backward(theta, spec, x, outs, tgt - outs[-1]), what="tc grad DIFF %s" % (widths,))
        if n_out >= 2 and acts[-1] == 0:
            onehot = np.eye(n_out, dtype=np.float32)[rs.randint(n_out, size=batch)]
            h = head(B, L.HEAD_SOFTMAX_CE, p0=0.01, target=B.up(onehot), tc=F)
            ok(B, lib.mq_fused_grad(net, B.ptr(ht), B.ptr(hx), None, batch, ctypes.byref(h), 1.0 / batch,
                                    B.ptr(grad), None, None, B.ptr(ws), wsb, B.stream))
            close(B.down(grad), O.mlp_backward(theta, spec, x, outs, O.softmax_ce_grad(outs[-1], onehot)),
                  what="tc grad CE %s" % (widths,))
    # Gaussian policy: reference fixtures, then the PPO head with gathered rows
    g = golden("gauss")
    spec = O.policy_spec(11, 3)
    net = L.make_net(11, [64, 64, 3], [1, 1, 0], logstd=True)
    theta, obs, act, w = g["gauss_theta"], g["gauss_obs"], g["gauss_act"], g["gauss_g"]
    ht, ho, ha, hw = B.up(theta), B.up(obs), B.up(act), B.up(w)
    logp = B.zeros(64, np.float32)
    ok(B, lib.mq_fused_forward_tc(net, B.ptr(ht), B.ptr(ho), None, 64, None, B.ptr(ha), B.ptr(logp), F, B.stream))
    close(B.down(logp), g["gauss_logp"], what="tc logp vs reference")
    wsb = lib.mq_fused_ws_bytes(net, 1 << 16)
    ws, grad = B.zeros(wsb, np.uint8), B.zeros(5126, np.float32)
    h = head(B, L.HEAD_GAUSS_LOGP, dy=hw, target=ha, tc=F)
    logp2 = B.zeros(64, np.float32)
    ok(B, lib.mq_fused_grad(net, B.ptr(ht), B.ptr(ho), None, 64, ctypes.byref(h), 1.0 / 64, B.ptr(grad), None,
                            B.ptr(logp2), B.ptr(ws), wsb, B.stream))
    close(B.down(grad), g["gauss_grad"], rtol=3e-5, atol_scale=3e-6, what="tc gauss grad vs reference")
    close(B.down(logp2), g["gauss_logp"], what="tc grad logp vs reference")
    n = 30000
    o = np.clip(rs.randn(n, 11), -5, 5).astype(np.float32)
    a = rs.randn(n, 3).astype(np.float32)
    adv = np.tanh(rs.randn(n)).astype(np.float32)
    lp_old = (O.policy_logprob(theta, spec, o, a)[0] + rs.randn(n) * 0.2).astype(np.float32)
    ho, ha, hadv, hlp = B.up(o), B.up(a), B.up(adv), B.up(lp_old)
    for mb in (64, 200, 20000):
        idx = rs.randint(n, size=mb).astype(np.int32)
        hi = B.up(idx)
        h = head(B, L.HEAD_GAUSS_PPO, p0=0.8, p1=1.2, target=ha, adv=hadv, logp_old=hlp, tc=F)
        ok(B, lib.mq_fused_grad(net, B.ptr(ht), B.ptr(ho), B.ptr(hi), mb, ctypes.byref(h), 1.0 / mb,
                                B.ptr(grad), None, None, B.ptr(ws), wsb, B.stream))
        lp, outs = O.policy_logprob(theta, spec, o[idx], a[idx])
        wgt = O.ppo_clip_weight(lp, lp_old[idx], adv[idx])
        close(B.down(grad), O.policy_logprob_grad(theta, spec, o[idx], a[idx], wgt, outs), rtol=3e-5,
              atol_scale=3e-6, what="tc ppo grad mb=%d" % mb)
        # the tensor-core and the FFMA kernel agree with each other at the same bar
        h2 = head(B, L.HEAD_GAUSS_PPO, p0=0.8, p1=1.2, target=ha, adv=hadv, logp_old=hlp, tc=L.TC_OFF)
        g2 = B.zeros(5126, np.float32)
        ok(B, lib.mq_fused_grad(net, B.ptr(ht), B.ptr(ho), B.ptr(hi), mb, ctypes.byref(h2), 1.0 / mb,
                                B.ptr(g2), None, None, B.ptr(ws), wsb, B.stream))
        close(B.down(grad), B.down(g2), rtol=3e-5, atol_scale=3e-6, what="tc vs ffma mb=%d" % mb)
    # evolution-strategies population on the tensor pipe (tc_es_kernel, >= 256 observations): several members per CTA,
    # ragged last tile
    case_es(B, members=200, n_obs=300)
    case_es(B, members=9, n_obs=1000)


# stated tolerances of the tensor-core precision modes (include/mq_b200.h): |err| <= rtol |ref| + atol_scale max|ref|
TC_MODE_TOL = {3: (3e-5, 3e-6), 0x82: (3e-5, 3e-6), 2: (1e-4, 3e-5), 1: (5e-2, 2e-2)}      # 0x82: fp16x2 = the exact bar
TC_MODE_LIST = (3, 0x82, 2, 1)


def pack_records(B, net, x, tgt=None, w0=None, w1=None):
    rows = x.shape[0]
    rec_w = B.lib.mq_record_width(net)
    rec = B.zeros((rows, rec_w), np.float32)
    keep = [B.up(x)] + [None if a is None else B.up(a) for a in (tgt, w0, w1)]
    ok(B, B.lib.mq_pack_records(net, B.ptr(keep[0]), B.ptr(keep[1]), B.ptr(keep[2]), B.ptr(keep[3]), rows, B.ptr(rec),
                                B.stream))
    return rec, rec_w


def case_tc_modes(B):
    """Record-fed tensor-core gradient kernel (mq_tc3.cu) in its precision modes -- 3 bf16 planes (exact products), 2 fp16
    planes with the low one scaled (the same bar), 2 bf16 planes, 1 plane -- against the oracle, each at its STATED tolerance (include/mq_b200.h), for every head, ragged
    batches, gathered rows, one and two hidden layers; packed records against the source arrays (bit-exact); operand
    image built by the launch, built ahead (mq_tc_image_build) and refreshed by the optimiser step
    (mq_reduce_step_image) are byte-identical."""
    lib = B.lib
    rs = np.random.RandomState(31)
    F = L.TC_FORCE
    measured = {}
    # ---- records are a bit-exact repacking (mannequin/_trajectory.py:79-86 semantics: float32 rows)
    net = L.make_net(11, [64, 64, 3], [1, 1, 0], logstd=True)
    x = rs.randn(1000, 11).astype(np.float32); t = rs.randn(1000, 3).astype(np.float32)
    a0 = rs.randn(1000).astype(np.float32); a1 = rs.randn(1000).astype(np.float32)
    rec, rec_w = pack_records(B, net, x, t, a0, a1)
    assert rec_w == 16
    r = B.down(rec)
    assert np.array_equal(r[:, :11], x) and np.array_equal(r[:, 11:14], t) and np.array_equal(r[:, 14], a0) \
        and np.array_equal(r[:, 15], a1)
    # ---- plain nets, heads DY / DIFF / SOFTMAX_CE
    cases = [(11, [64, 64, 3], [1, 1, 0], 64), (11, [64, 64, 1], [1, 1, 0], 1000), (11, [64, 64, 1], [1, 1, 0], 40000),
             (13, [16, 16, 5], [2, 2, 0], 777), (6, [32, 10], [2, 0], 333), (12, [48, 2], [1, 0], 1500),
             (20, [64, 64, 4], [1, 1, 0], 3000)]
    for n_in, widths, acts, batch in cases:
        spec, theta, x = mlp_case(rs, n_in, widths, acts, batch)
        net = net_of(spec)
        n_out = widths[-1]
        ht = B.up(theta)
        outs = O.mlp_forward(theta, spec, x)
        wsb = lib.mq_fused_ws_bytes(net, batch)
        ws, grad = B.zeros(wsb, np.uint8), B.zeros(spec["n_params"], np.float32)
        dy = rs.randn(batch, n_out).astype(np.float32)
        tgt = rs.randn(batch, n_out).astype(np.float32)
        onehot = np.eye(n_out, dtype=np.float32)[rs.randint(n_out, size=batch)]
        hx = B.up(x)
        for planes in TC_MODE_LIST:
            rtol, atol = TC_MODE_TOL[planes]
            for kind, mat, dyo in ((L.HEAD_DY, dy, dy), (L.HEAD_DIFF, tgt, tgt - outs[-1]),
                                   (L.HEAD_SOFTMAX_CE, onehot, None)):
                if kind == L.HEAD_SOFTMAX_CE:
                    if not (n_out >= 2 and acts[-1] == 0):
                        continue
                    dyo = O.softmax_ce_grad(outs[-1], onehot)
                rec, rec_w = pack_records(B, net, x, mat)
                hm = B.up(mat)
                h = head(B, kind, p0=0.01 if kind == L.HEAD_SOFTMAX_CE else 0.0, dy=hm if kind == L.HEAD_DY else None,
                         target=None if kind == L.HEAD_DY else hm, tc=F | planes, records=rec, rec_w=rec_w)
                out2 = B.zeros((batch, n_out), np.float32)
                ok(B, lib.mq_fused_grad(net, B.ptr(ht), B.ptr(hx), None, batch, ctypes.byref(h), 1.0 / batch,
                                        B.ptr(grad), B.ptr(out2), None, B.ptr(ws), wsb, B.stream))
                want = O.mlp_backward(theta, spec, x, outs, dyo)
                got = B.down(grad)
                close(got, want, rtol=rtol, atol_scale=atol, what="tc3 P=%d kind=%d %s b=%d" % (planes, kind, widths, batch))
                close(B.down(out2), outs[-1], rtol=rtol, atol_scale=atol, what="tc3 out P=%d" % planes)
                e = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
                measured[planes] = max(measured.get(planes, 0.0), e)
    # ---- Gaussian PPO head with gathered rows (the configs[2] step), categorical PPO head
    spec = O.policy_spec(11, 3)
    net = L.make_net(11, [64, 64, 3], [1, 1, 0], logstd=True)
    theta = golden("gauss")["gauss_theta"]
    n = 70000
    o = np.clip(rs.randn(n, 11), -5, 5).astype(np.float32)
    a = rs.randn(n, 3).astype(np.float32)
    adv = np.tanh(rs.randn(n)).astype(np.float32)
    lp_old = (O.policy_logprob(theta, spec, o, a)[0] + rs.randn(n) * 0.2).astype(np.float32)
    ht, ho, ha, hadv, hlp = B.up(theta), B.up(o), B.up(a), B.up(adv), B.up(lp_old)
    rec, rec_w = pack_records(B, net, o, a, adv, lp_old)
    wsb = lib.mq_fused_ws_bytes(net, 1 << 16)
    ws, grad = B.zeros(wsb, np.uint8), B.zeros(5126, np.float32)
    img_bytes = {pl: lib.mq_tc_image_bytes(net, pl) for pl in TC_MODE_LIST}
    assert 0 < img_bytes[1] < img_bytes[2] < img_bytes[3] <= 96 * 1024 and img_bytes[0x82] == img_bytes[2]
    assert lib.mq_tc_image_bytes(net, 3 | L.TC_FP16) == 0               # the fp16 format exists with two planes only
    for mb in (64, 200, 8192, 20000, 65536):
        idx = rs.randint(n, size=mb).astype(np.int32)
        hi = B.up(idx)
        lp, outs = O.policy_logprob(theta, spec, o[idx], a[idx])
        wgt = O.ppo_clip_weight(lp, lp_old[idx], adv[idx])
        assert (wgt == 0).any() and (wgt != 0).any()                       # both branches of the clip rule
        want = O.policy_logprob_grad(theta, spec, o[idx], a[idx], wgt, outs)
        # reference point: the array-fed exact kernel
        h0 = head(B, L.HEAD_GAUSS_PPO, p0=0.8, p1=1.2, target=ha, adv=hadv, logp_old=hlp, tc=F)
        g0 = B.zeros(5126, np.float32)
        ok(B, lib.mq_fused_grad(net, B.ptr(ht), B.ptr(ho), B.ptr(hi), mb, ctypes.byref(h0), 1.0 / mb, B.ptr(g0), None,
                                None, B.ptr(ws), wsb, B.stream))
        g0 = B.down(g0)
        close(g0, want, rtol=3e-5, atol_scale=3e-6, what="array-fed tc ppo grad mb=%d vs oracle on all rows" % mb)
        for planes in TC_MODE_LIST:
            rtol, atol = TC_MODE_TOL[planes]
            image = B.zeros(img_bytes[planes], np.uint8)
            ok(B, lib.mq_tc_image_build(net, B.ptr(ht), planes, B.ptr(image), B.stream))
            for img in (None, image):
                h = head(B, L.HEAD_GAUSS_PPO, p0=0.8, p1=1.2, target=ha, adv=hadv, logp_old=hlp, tc=F | planes,
                         records=rec, rec_w=rec_w, image=img)
                lpo = B.zeros(mb, np.float32)
                ok(B, lib.mq_fused_grad(net, B.ptr(ht), B.ptr(ho), B.ptr(hi), mb, ctypes.byref(h), 1.0 / mb,
                                        B.ptr(grad), None, B.ptr(lpo), B.ptr(ws), wsb, B.stream))
                got = B.down(grad)
                close(got, want, rtol=rtol, atol_scale=atol, what="tc3 ppo P=%d mb=%d img=%s" % (planes, mb, img is not None))
                close(B.down(lpo), lp, rtol=rtol, atol_scale=max(atol, 3e-6), what="tc3 logp P=%d mb=%d" % (planes, mb))
                if planes == 3:
                    close(got, g0, rtol=3e-5, atol_scale=3e-6, what="tc3 exact vs array-fed kernel mb=%d" % mb)
                e = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
                measured["ppo%d" % planes] = max(measured.get("ppo%d" % planes, 0.0), e)
    # categorical policy (benchmarks/policy.py:15-17): one-hot targets
    spec_d = O.mlp_spec(11, [64, 64, 4], [1, 1, 0])
    net_d = net_of(spec_d)
    theta_d = O.mlp_init(spec_d, rs, head_scale=0.1) + rs.randn(spec_d["n_params"]).astype(np.float32) * 0.05
    acts_d = rs.randint(4, size=n)
    onehot = np.eye(4, dtype=np.float32)[acts_d]
    lpo_d = (O.discrete_logprob(O.mlp_forward(theta_d, spec_d, o)[-1], acts_d)[0] + rs.randn(n) * 0.2).astype(np.float32)
    rec_d, rw_d = pack_records(B, net_d, o, onehot, adv, lpo_d)
    htd, hoh, hlpd = B.up(theta_d), B.up(onehot), B.up(lpo_d)
    gd = B.zeros(spec_d["n_params"], np.float32)
    wsb_d = lib.mq_fused_ws_bytes(net_d, 1 << 16)
    wsd = B.zeros(wsb_d, np.uint8)
    idx = rs.randint(n, size=9000).astype(np.int32)
    hi = B.up(idx)
    outs_d = O.mlp_forward(theta_d, spec_d, o[idx])
    lp_d = O.discrete_logprob(outs_d[-1], acts_d[idx])[0]
    wgt = O.ppo_clip_weight(lp_d, lpo_d[idx], adv[idx])
    want = O.discrete_logprob_grad(theta_d, spec_d, o[idx], acts_d[idx], wgt, outs_d)
    for planes in TC_MODE_LIST:
        rtol, atol = TC_MODE_TOL[planes]
        h = head(B, L.HEAD_DISCRETE_PPO, p0=0.8, p1=1.2, target=hoh, adv=hadv, logp_old=hlpd, tc=F | planes,
                 records=rec_d, rec_w=rw_d)
        ok(B, lib.mq_fused_grad(net_d, B.ptr(htd), B.ptr(ho), B.ptr(hi), 9000, ctypes.byref(h), 1.0 / 9000, B.ptr(gd), None,
                                None, B.ptr(wsd), wsb_d, B.stream))
        close(B.down(gd), want, rtol=rtol, atol_scale=atol, what="tc3 discrete ppo P=%d" % planes)
    # ---- the optimiser step keeps the operand image in step with theta (mq_reduce_step_image)
    for planes in TC_MODE_LIST:
        mb = 8192
        idx = rs.randint(n, size=mb).astype(np.int32)
        hi = B.up(idx)
        th = B.up(theta.copy())
        image = B.zeros(img_bytes[planes], np.uint8)
        ok(B, lib.mq_tc_image_build(net, B.ptr(th), planes, B.ptr(image), B.stream))
        val, m, v = B.up(theta.copy()), B.zeros(5126, np.float32), B.zeros(5126, np.float32)
        h = head(B, L.HEAD_GAUSS_PPO, p0=0.8, p1=1.2, target=ha, adv=hadv, logp_old=hlp, tc=planes, records=rec,
                 rec_w=rec_w, image=image)
        n_parts = ctypes.c_int32(0)
        for step in range(3):
            ok(B, lib.mq_fused_grad_partials(net, B.ptr(th), B.ptr(ho), B.ptr(hi), mb, ctypes.byref(h), None, None,
                                             B.ptr(ws), wsb, ctypes.byref(n_parts), B.stream))
            assert n_parts.value > 1
            ok(B, lib.mq_reduce_step_image(B.ptr(ws), n_parts.value, 5126, 1.0 / mb, B.ptr(val), B.ptr(m), B.ptr(v), B.ptr(th),
                                           None, L.MQ_F32, 1.0 / (step + 1), 1.0 / (step + 1), 3e-3, 1e-8, 0, None, net, planes,
                                           B.ptr(image), B.stream))
        fresh = B.zeros(img_bytes[planes], np.uint8)
        ok(B, lib.mq_tc_image_build(net, B.ptr(th), planes, B.ptr(fresh), B.stream))
        assert np.array_equal(B.down(image), B.down(fresh)), "refreshed operand image differs from a fresh build (P=%d)" % planes
        assert np.max(np.abs(B.down(th) - theta)) > 1e-4
    return measured


def case_reduce_step(B):
    """mq_reduce_step (single GPU): fixed-order sum of partial gradients + Adam in one launch equals
    mq_adam_step on the summed gradient (fp32 state) and the oracle's Adam (fp64 state, bit exact)."""
    lib = B.lib
    rs = np.random.RandomState(21)
    for n, parts in ((5126, 148), (4993, 1), (33, 7), (118282, 40)):
        partial = (rs.randn(parts, n) * 0.1).astype(np.float32)
        scale = 1.0 / 4096
        g64 = partial.astype(np.float64).sum(0) * scale
        val0 = rs.randn(n)
        hp = B.up(partial)
        # fp64 state against the oracle, teacher-forced with the kernel's own fp32 gradient
        val, m, v = B.up(val0.copy()), B.zeros(n, np.float64), B.zeros(n, np.float64)
        th, gout = B.zeros(n, np.float32), B.zeros(n, np.float32)
        ok(B, lib.mq_reduce_step(B.ptr(hp), parts, n, scale, B.ptr(val), B.ptr(m), B.ptr(v), B.ptr(th), B.ptr(gout),
                                 L.MQ_F64, 1.0, 1.0, 1e-3, 1e-8, 0, None, B.stream))
        g = B.down(gout)
        close(g, g64, rtol=1e-5, atol_scale=1e-6, what="reduce_step grad n=%d" % n)
        ref = O.AdamO(val0, horizon=10, lr=1e-3)
        ref.apply_gradient(g)
        assert np.array_equal(B.down(val), ref.get_value()), "reduce_step fp64 Adam n=%d" % n
        assert np.array_equal(B.down(th), ref.get_value().astype(np.float32))
        # fp32 state against mq_adam_step on the same gradient
        v32 = [B.up(val0.astype(np.float32)), B.zeros(n, np.float32), B.zeros(n, np.float32)]
        w32 = [B.up(val0.astype(np.float32)), B.zeros(n, np.float32), B.zeros(n, np.float32)]
        ok(B, lib.mq_reduce_step(B.ptr(hp), parts, n, scale, B.ptr(v32[0]), B.ptr(v32[1]), B.ptr(v32[2]), None, None,
                                 L.MQ_F32, 0.5, 0.25, 3e-4, 1e-8, 0, None, B.stream))
        ok(B, lib.mq_adam_step(B.ptr(w32[0]), B.ptr(w32[1]), B.ptr(w32[2]), None, B.ptr(gout), n, L.MQ_F32, L.MQ_F32,
                               0.5, 0.25, 3e-4, 1e-8, 0, B.stream))
        for a, b in zip(v32, w32):
            assert np.array_equal(B.down(a), B.down(b)), "reduce_step fp32 state n=%d" % n


def case_discrete(B, batches=(64, 300, 9000)):
    """Categorical policy head (mannequin/distrib.py:8-39): per-op kernel, fused heads (FFMA and, for the
    9000-row batch, the tcgen05 kernel) against the reference fixture and the oracle."""
    lib = B.lib
    rs = np.random.RandomState(31)
    g = golden("discrete")
    spec = O.mlp_spec(4, [64, 64, 3], [1, 1, 0])
    net = net_of(spec)
    theta, obs, act, w = g["disc_theta"], g["disc_obs"], g["disc_act"], g["disc_g"]
    # per-op kernel on the reference's logits
    hl, ha, hw = B.up(g["disc_logits"].astype(np.float32)), B.up(act.astype(np.int32)), B.up(w)
    lp, dl = B.zeros(64, np.float32), B.zeros((64, 3), np.float32)
    ok(B, lib.mq_discrete_logprob(B.ptr(hl), B.ptr(ha), B.ptr(hw), 64, 3, B.ptr(lp), B.ptr(dl), B.stream))
    close(B.down(lp), g["disc_logp"], rtol=2e-5, what="discrete logp vs reference")
    wlp, p = O.discrete_logprob(g["disc_logits"], act)
    close(B.down(dl), w[:, None] * (np.eye(3)[act] - p), what="discrete d_logits")
    # fused: forward log-prob and the gradient, against the reference's own backprop
    onehot = B.zeros((64, 3), np.float32)
    ok(B, lib.mq_onehot(B.ptr(ha), 64, 3, B.ptr(onehot), B.stream))
    assert np.array_equal(B.down(onehot), np.eye(3, dtype=np.float32)[act])
    ht, ho = B.up(theta), B.up(obs)
    ok(B, lib.mq_fused_forward(net, B.ptr(ht), B.ptr(ho), None, 64, None, B.ptr(onehot), B.ptr(lp), B.stream))
    close(B.down(lp), g["disc_logp"], rtol=2e-5, what="fused discrete logp vs reference")
    wsb = lib.mq_fused_ws_bytes(net, max(batches))
    ws, grad = B.zeros(wsb, np.uint8), B.zeros(spec["n_params"], np.float32)
    h = head(B, L.HEAD_DISCRETE_LOGP, dy=hw, target=onehot)
    ok(B, lib.mq_fused_grad(net, B.ptr(ht), B.ptr(ho), None, 64, ctypes.byref(h), 1.0 / 64, B.ptr(grad), None,
                            B.ptr(lp), B.ptr(ws), wsb, B.stream))
    close(B.down(grad), g["disc_grad"], rtol=3e-5, atol_scale=3e-6, what="fused discrete grad vs reference")
    # PPO clip weights with gathered rows, several batch sizes
    n = 12000
    o = rs.randn(n, 4).astype(np.float32)
    a = rs.randint(3, size=n).astype(np.int32)
    adv = np.tanh(rs.randn(n)).astype(np.float32)
    lp_all, _ = O.discrete_logprob(O.mlp_forward(theta, spec, o)[-1], a)
    lp_old = (lp_all + rs.randn(n) * 0.2).astype(np.float32)
    ho, hoh, hadv, hlp = B.up(o), B.up(np.eye(3, dtype=np.float32)[a]), B.up(adv), B.up(lp_old)
    for mb in batches:
        idx = rs.randint(n, size=mb).astype(np.int32)
        hi = B.up(idx)
        h = head(B, L.HEAD_DISCRETE_PPO, p0=0.8, p1=1.2, target=hoh, adv=hadv, logp_old=hlp)
        ok(B, lib.mq_fused_grad(net, B.ptr(ht), B.ptr(ho), B.ptr(hi), mb, ctypes.byref(h), 1.0 / mb,
                                B.ptr(grad), None, None, B.ptr(ws), wsb, B.stream))
        outs = O.mlp_forward(theta, spec, o[idx])
        lpn, _ = O.discrete_logprob(outs[-1], a[idx])
        wgt = O.ppo_clip_weight(lpn, lp_old[idx], adv[idx])
        assert (wgt == 0).any() and (wgt != 0).any()
        close(B.down(grad), O.discrete_logprob_grad(theta, spec, o[idx], a[idx], wgt, outs), rtol=3e-5,
              atol_scale=3e-6, what="fused discrete ppo grad mb=%d" % mb)
    # the persistent chain (benchmarks/ppo.py:29-40 for a cartpole-shaped policy): 5 dependent 64-row steps
    steps, mbs = 5, 64
    tab = rs.randint(n, size=(steps, mbs)).astype(np.int32)
    opt = O.AdamO(theta, horizon=10)
    th_ref = theta.copy()
    for s_ in range(steps):
        th_ref = O.ppo_discrete_policy_step(th_ref, spec, opt, o[tab[s_]], a[tab[s_]], adv[tab[s_]], lp_old[tab[s_]])
    np_ = spec["n_params"]
    ht2 = B.up(theta.copy())
    val, m, v = B.up(theta.astype(np.float64)), B.zeros(np_, np.float64), B.zeros(np_, np.float64)
    tw = (ctypes.c_double * 2)(0.0, 0.0)
    h = head(B, L.HEAD_DISCRETE_PPO, p0=0.8, p1=1.2, target=hoh, adv=hadv, logp_old=hlp)
    htab = B.up(tab)
    ok(B, lib.mq_fused_train(net, B.ptr(ht2), B.ptr(val), B.ptr(m), B.ptr(v), L.MQ_F64, B.ptr(ho), ctypes.byref(h),
                             B.ptr(htab), steps, mbs, 0.9, 0.99, tw, 3e-4, 1e-8, B.stream))
    err = np.max(np.abs(B.down(ht2) - th_ref))
    assert err < 2e-6 * (1 + steps) + 1e-6, err
    assert np.max(np.abs(B.down(ht2) - theta)) > 1e-4
    # a categorical head on a net WITH logstd parameters is refused
    h = head(B, L.HEAD_DISCRETE_LOGP, dy=hw, target=onehot)
    gnet = L.make_net(4, [64, 64, 3], [1, 1, 0], logstd=True)
    assert lib.mq_fused_grad(gnet, B.ptr(ht), B.ptr(ho), None, 64, ctypes.byref(h), 1.0, B.ptr(grad), None, None,
                             B.ptr(ws), wsb, B.stream) == L.MQ_ERR_INVALID


# stated tolerances of the GEMM precision modes: |err| <= rtol |ref| + atol_scale max|ref|
GEMM_MODE_TOL = {3: (1e-5, 2e-6), 0x82: (1e-5, 2e-6), 2: (1e-4, 3e-5), 1: (5e-2, 2e-2)}      # 0x82: fp16x2 = the exact bar


def case_gemm_tc(B):
    """TMA-fed tcgen05 GEMM for wide layers (mq_gemm_tma.cu), forced on per call (MQ_TC_FORCE): every operand layout,
    ragged sizes (rows, columns and K not multiples of the tiles), long contractions (split-K, the K = 128 accumulator
    fold), the precision modes (three bf16 planes, two scaled fp16 planes, two bf16 planes, one) at their stated tolerances; then the fused epilogues, row gather and split-K of the
    layered MLP path at the fp32 FFMA parity bar."""
    lib = B.lib
    rs = np.random.RandomState(41)
    F = L.TC_FORCE

    def gemm(ha, a_rs, a_cs, hb, b_rs, b_cs, m, n, k, alpha, tc):
        c = B.zeros((m, n), np.float32)
        wsb = lib.mq_gemm_ws_bytes(m, n, k)
        ws = B.zeros(wsb, np.uint8)
        ok(B, lib.mq_gemm_ex(B.ptr(ha), a_rs, a_cs, B.ptr(hb), b_rs, b_cs, B.ptr(c), m, n, k, alpha, tc, B.ptr(ws), wsb, B.stream))
        return B.down(c)

    for m, n, k in ((128, 128, 32), (256, 64, 64), (300, 200, 100), (130, 70, 2003), (1024, 400, 784), (784, 400, 4096)):
        a = rs.randn(m, k).astype(np.float32)
        b = rs.randn(k, n).astype(np.float32)
        want = 0.5 * (a.astype(np.float64) @ b.astype(np.float64))
        for a_t in (False, True):
            for b_t in (False, True):
                ha = B.up(np.ascontiguousarray(a.T) if a_t else a)
                hb = B.up(np.ascontiguousarray(b.T) if b_t else b)
                a_rs, a_cs = (1, m) if a_t else (k, 1)
                b_rs, b_cs = (1, k) if b_t else (n, 1)
                for planes in (3, 0x82, 2, 1):
                    rtol, atol = GEMM_MODE_TOL[planes]
                    got = gemm(ha, a_rs, a_cs, hb, b_rs, b_cs, m, n, k, 0.5, F | planes)
                    close(got, want, rtol=rtol, atol_scale=atol, what="gemm_tma P=%d %dx%dx%d aT=%d bT=%d" % (planes, m, n, k, a_t, b_t))
    # the tensor-core and the FFMA kernel agree element by element at the same bar
    a, b = rs.randn(512, 300).astype(np.float32), rs.randn(300, 256).astype(np.float32)
    ha, hb = B.up(a), B.up(b)
    c1 = gemm(ha, 300, 1, hb, 256, 1, 512, 256, 300, 1.0, F)
    c2 = gemm(ha, 300, 1, hb, 256, 1, 512, 256, 300, 1.0, L.TC_OFF)
    close(c1, c2, what="gemm tc vs ffma")
    # layered MLP (mnist/mlp.py and mnist/vae.py encoder shapes) at a batch where every layer is a real contraction:
    # bias + activation epilogue, act' epilogue, gathered rows, split-K weight gradients
    # (the third configuration: per-call fp16x2 operands through mq_mlp_forward_tc / mq_mlp_backward_tc, and a batch at which
    # the narrow output layer takes the three skinny-product kernels of mq_gemm.cu)
    for n_in, widths, acts, batch, n_src, tc in ((784, [128, 128, 10], [2, 2, 0], 2048, 3000, None), (784, [400, 40], [1, 0], 4096, 4096, None),
                                                 (784, [128, 128, 10], [2, 2, 0], 8192, 3000, L.TC_MODES["fp16x2"]),
                                                 # MQ_TC_KEEP_X: the forward pass's planes of x feed dW_0 (MN-major), ragged batch
                                                 (784, [128, 128, 10], [2, 2, 0], 8200, 3000, L.TC_MODES["fp16x2"] | L.TC_KEEP_X),
                                                 # (Tanh: a draw with a hidden pre-activation within rounding of LReLU's kink makes
                                                 #  either slope a correct fp32 result and the comparison meaningless)
                                                 (784, [128, 10], [1, 0], 8192, 3000, L.TC_MODES["exact"] | L.TC_KEEP_X)):
        spec, theta, x = mlp_case(rs, n_in, widths, acts, n_src)
        idx = rs.randint(n_src, size=batch).astype(np.int32)
        net = net_of(spec)
        ht, hx, hi = B.up(theta), B.up(x), B.up(idx)
        acts_buf = B.zeros(lib.mq_mlp_acts_floats(net, batch), np.float32)
        wsb = lib.mq_mlp_ws_bytes(net, batch)
        ws = B.zeros(wsb, np.uint8)
        if tc is None:
            ok(B, lib.mq_mlp_forward_ws(net, B.ptr(ht), B.ptr(hx), B.ptr(hi), batch, B.ptr(acts_buf), B.ptr(ws), wsb, B.stream))
        else:
            ok(B, lib.mq_mlp_forward_tc(net, B.ptr(ht), B.ptr(hx), B.ptr(hi), batch, B.ptr(acts_buf), tc, B.ptr(ws), wsb, B.stream))
        outs = O.mlp_forward(theta, spec, x[idx])
        got, pos = B.down(acts_buf), 0
        for o in outs:
            close(got[pos:pos + o.size].reshape(o.shape), o, what="layered tc fwd %s" % (widths,))
            pos += o.size
        dy = rs.randn(batch, widths[-1]).astype(np.float32)
        want, wdx = O.mlp_backward(theta, spec, x[idx], outs, dy, want_dx=True)
        grad, dx, hdy = B.zeros(spec["n_params"], np.float32), B.zeros((batch, n_in), np.float32), B.up(dy.copy())
        if tc is None:
            ok(B, lib.mq_mlp_backward(net, B.ptr(ht), B.ptr(hx), B.ptr(hi), B.ptr(acts_buf), B.ptr(hdy), batch,
                                      1.0 / batch, B.ptr(grad), B.ptr(dx), B.ptr(ws), wsb, B.stream))
        else:
            ok(B, lib.mq_mlp_backward_tc(net, B.ptr(ht), B.ptr(hx), B.ptr(hi), B.ptr(acts_buf), B.ptr(hdy), batch,
                                         1.0 / batch, B.ptr(grad), B.ptr(dx), tc, B.ptr(ws), wsb, B.stream, None, None))
        close(B.down(grad), want, what="layered tc grad %s" % (widths,))
        close(B.down(dx), wdx, what="layered tc dx %s" % (widths,))


ALL_CASES = [case_adam, case_norm, case_gather, case_scan, case_layered_mlp, case_gemm_tc, case_fused_heads,
             case_fused_gauss_ppo, case_discrete, case_entropy, case_rollout, case_tc_paths, case_tc_modes, case_reduce_step, case_fused_train, case_ppo_reference_replay, case_es, case_ops]
