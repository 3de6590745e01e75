#!/usr/bin/env python3
"""Developer-only: run the kernels' LOGIC on the host emulator against the oracle.

    tools/emu/build_emu.sh && python tools/emu/check_kernels.py [filter]

Not part of pytest, not used by the product.  It exists because the build
container has no GPU; it catches indexing / tiling mistakes before B200 time is
spent.  (Races and memory-model bugs are NOT caught: CTAs run sequentially.)
"""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mannequin2_b200 import _lib as L          # noqa: E402  (signature table only)
from oracle import mq_oracle as O               # noqa: E402

emu = L.bind(ctypes.CDLL(os.path.join(ROOT, "tools/emu/_build/libmq_emu.so")))


def ptr(a):
    return None if a is None else a.ctypes.data


def ok(rc):
    if rc != 0:
        raise RuntimeError("rc=%d: %s" % (rc, emu.mq_last_error().decode()))


def close(a, b, rtol=1e-5, atol_scale=2e-6, what=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    atol = atol_scale * max(1e-30, np.max(np.abs(b)))
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    if not (err <= 0).all():
        i = np.argmax(err)
        raise AssertionError("%s: mismatch at %d: got %r want %r (max|b|=%g)" % (
            what, i, a.reshape(-1)[i], b.reshape(-1)[i], np.max(np.abs(b))))


def net_of(spec, logstd=False):
    Ls = spec["layers"]
    return L.make_net(spec["n_in"], [l["n_out"] for l in Ls], [l["act"] for l in Ls],
                      leak=[l["leak"] for l in Ls], logstd=logstd)


CHECKS = []


def check(fn):
    CHECKS.append(fn)
    return fn


@check
def adam():
    rs = np.random.RandomState(0)
    for n in (5, 37, 1024 + 3):
        val0 = rs.randn(n)
        grads = rs.randn(6, n).astype(np.float32)
        ref = O.AdamO(val0, horizon=10, lr=0.003)
        val, m, v = val0.copy(), np.zeros(n), np.zeros(n)
        th = np.zeros(n, np.float32)
        tw1 = tw2 = 0.0
        for g in grads:
            ref.apply_gradient(g)
            tw1 = tw1 * 0.9 + 1.0
            tw2 = tw2 * 0.99 + 1.0
            ok(emu.mq_adam_step(ptr(val), ptr(m), ptr(v), ptr(th), ptr(g), n, L.MQ_F64, L.MQ_F32,
                                1.0 / tw1, 1.0 / tw2, 0.003, 1e-8, 0, None))
            assert np.array_equal(val, ref.get_value()), "fp64 Adam must be bit exact"
            assert np.array_equal(th, val.astype(np.float32))
        # fp32 state
        ref = O.AdamO(val0, horizon=10, lr=0.003)
        val, m, v = val0.astype(np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
        tw1 = tw2 = 0.0
        for g in grads:
            ref.apply_gradient(g)
            tw1 = tw1 * 0.9 + 1.0
            tw2 = tw2 * 0.99 + 1.0
            ok(emu.mq_adam_step(ptr(val), ptr(m), ptr(v), None, ptr(g), n, L.MQ_F32, L.MQ_F32,
                                1.0 / tw1, 1.0 / tw2, 0.003, 1e-8, 0, None))
        close(val, ref.get_value(), rtol=1e-5, what="adam f32")


@check
def norm():
    rs = np.random.RandomState(1)
    for rows, cols, dt in ((2048, 1, np.float32), (5000, 1, np.float64), (9, 11, np.float64), (300, 5, np.float32)):
        x = (rs.randn(rows, cols) * 3 + 1).astype(dt)
        code = L.MQ_F64 if dt == np.float64 else L.MQ_F32
        wsb = emu.mq_norm_ws_bytes(rows, cols)
        ws = np.zeros(wsb, np.uint8)
        mean = np.zeros(cols)
        ok(emu.mq_col_mean(ptr(x), code, rows, cols, ptr(mean), ptr(ws), wsb, None))
        assert np.allclose(mean, x.astype(np.float64).mean(axis=0), rtol=1e-13, atol=1e-13)
        mo = rs.randn(cols)
        cov = np.zeros(cols)
        ok(emu.mq_norm_cov(ptr(x), code, rows, cols, ptr(mo), ptr(mean), 2.0, ptr(cov), ptr(ws), wsb, None))
        want = 2.0 * ((x.astype(np.float64) - mo) * (x.astype(np.float64) - mean)).mean(axis=0)
        assert np.allclose(cov, want, rtol=1e-12, atol=1e-12)
        var = np.abs(cov) + 0.1
        out = np.zeros((rows, cols), np.float32)
        ok(emu.mq_norm_apply(ptr(x), code, rows, cols, ptr(mean), ptr(var), ptr(out), L.MQ_F32, 1, None))
        want = np.tanh(((x - mean) / np.maximum(1e-8, np.sqrt(var))).astype(np.float32))
        close(out, want, what="norm apply tanh")
    mean = rs.randn(7)
    x = rs.randn(7)
    old = np.zeros(7)
    want = mean + 0.3 * (x - mean)
    m0 = mean.copy()
    ok(emu.mq_running_mean_update(ptr(mean), ptr(x), L.MQ_F64, 7, 0.3, ptr(old), None))
    assert np.array_equal(mean, want) and np.array_equal(old, m0)


@check
def gather_scan():
    rs = np.random.RandomState(2)
    src = rs.randn(40, 11).astype(np.float32)
    idx = rs.randint(-40, 40, size=64).astype(np.int32)
    dst = np.zeros((64, 11), np.float32)
    ok(emu.mq_gather_rows(ptr(src), ptr(idx), 64, 44, 40, ptr(dst), None))
    assert np.array_equal(dst, src[idx])
    src4 = rs.randn(40, 12).astype(np.float32)
    dst4 = np.zeros((64, 12), np.float32)
    ok(emu.mq_gather_rows(ptr(src4), ptr(idx), 64, 48, 40, ptr(dst4), None))
    assert np.array_equal(dst4, src4[idx])
    for n in (1, 700, 2048, 5000):
        r = rs.randn(n).astype(np.float32)
        v = rs.randn(n + 1).astype(np.float32)
        d = (rs.rand(n) < 0.02).astype(np.uint8)
        adv, ret = np.zeros(n, np.float32), np.zeros(n, np.float32)
        wsb = emu.mq_scan_ws_bytes(n)
        ws = np.zeros(wsb, np.uint8)
        ok(emu.mq_gae(ptr(r), ptr(v), ptr(d), n, 0.99, 0.95, ptr(adv), ptr(ret), ptr(ws), wsb, None))
        wa, wr = O.gae_advantages(r, v, d)
        close(adv, wa, what="gae adv n=%d" % n)
        close(ret, wr, what="gae ret n=%d" % n)
        # integer data, gam=lam=1: exact
        ri = rs.randint(-3, 4, size=n).astype(np.float32)
        vi = rs.randint(-3, 4, size=n + 1).astype(np.float32)
        ok(emu.mq_gae(ptr(ri), ptr(vi), ptr(d), n, 1.0, 1.0, ptr(adv), ptr(ret), ptr(ws), wsb, None))
        wa, wr = O.gae_advantages(ri, vi, d, 1.0, 1.0)
        assert np.array_equal(adv, wa) and np.array_equal(ret, wr), "segmentation must be exact"
        x = rs.randn(n)
        out = np.zeros(n)
        ok(emu.mq_discounted(ptr(x), n, 500.0, ptr(out), ptr(ws), wsb, None))
        assert np.allclose(out, O.discounted(x, 500), rtol=1e-11, atol=1e-13)


def _mlp_case(rs, n_in, widths, acts, batch, logstd=False):
    spec = O.mlp_spec(n_in, widths, acts)
    theta = O.mlp_init(spec, rs, head_scale=0.1) + rs.randn(spec["n_params"]).astype(np.float32) * 0.05
    if logstd:
        theta = np.concatenate([theta, (rs.randn(widths[-1]) * 0.3).astype(np.float32)])
    x = rs.randn(batch, n_in).astype(np.float32)
    return spec, theta, x


@check
def layered_mlp():
    rs = np.random.RandomState(3)
    for n_in, widths, acts, batch in ((20, [16, 16, 5], [2, 2, 0], 7), (11, [64, 64, 3], [1, 1, 0], 130),
                                      (70, [33, 10], [2, 1], 65)):
        spec, theta, x = _mlp_case(rs, n_in, widths, acts, batch)
        net = net_of(spec)
        acts_buf = np.zeros(emu.mq_mlp_acts_floats(net, batch), np.float32)
        ok(emu.mq_mlp_forward(net, ptr(theta), ptr(x), None, batch, ptr(acts_buf), None))
        outs = O.mlp_forward(theta, spec, x)
        pos = 0
        for o in outs:
            close(acts_buf[pos:pos + o.size].reshape(o.shape), o, what="layered fwd")
            pos += o.size
        dy = rs.randn(batch, widths[-1]).astype(np.float32)
        want, wdx = O.mlp_backward(theta, spec, x, outs, dy, want_dx=True)
        wsb = emu.mq_mlp_ws_bytes(net, batch)
        ws = np.zeros(wsb, np.uint8)
        grad = np.zeros(spec["n_params"], np.float32)
        dx = np.zeros((batch, n_in), np.float32)
        dyc = dy.copy()
        ok(emu.mq_mlp_backward(net, ptr(theta), ptr(x), None, ptr(acts_buf), ptr(dyc), batch, 1.0 / batch,
                               ptr(grad), ptr(dx), ptr(ws), wsb, None))
        close(grad, want, what="layered grad")
        close(dx, wdx, what="layered dx")
    # gathered rows + split-K (batch >= 512)
    spec, theta, x = _mlp_case(rs, 11, [64, 64, 1], [1, 1, 0], 300)
    idx = rs.randint(300, size=1024).astype(np.int32)
    net = net_of(spec)
    acts_buf = np.zeros(emu.mq_mlp_acts_floats(net, 1024), np.float32)
    ok(emu.mq_mlp_forward(net, ptr(theta), ptr(x), ptr(idx), 1024, ptr(acts_buf), None))
    outs = O.mlp_forward(theta, spec, x[idx])
    dy = rs.randn(1024, 1).astype(np.float32)
    want = O.mlp_backward(theta, spec, x[idx], outs, dy)
    wsb = emu.mq_mlp_ws_bytes(net, 1024)
    ws = np.zeros(wsb, np.uint8)
    grad = np.zeros(spec["n_params"], np.float32)
    dyc = dy.copy()
    ok(emu.mq_mlp_backward(net, ptr(theta), ptr(x), ptr(idx), ptr(acts_buf), ptr(dyc), 1024,
                           1.0 / 1024, ptr(grad), None, ptr(ws), wsb, None))
    close(grad, want, what="layered grad gathered/splitK")


def _head(kind, p0=0.0, p1=0.0, dy=None, target=None, adv=None, logp_old=None):
    h = L.MqHead()
    h.kind, h.p0, h.p1 = kind, p0, p1
    h.dy, h.target, h.adv, h.logp_old = ptr(dy), ptr(target), ptr(adv), ptr(logp_old)
    h._keep = (dy, target, adv, logp_old)
    return h


@check
def fused_forward_grad():
    rs = np.random.RandomState(4)
    cases = ((11, [64, 64, 3], [1, 1, 0], 64), (11, [64, 64, 3], [1, 1, 0], 200),
             (20, [16, 16, 5], [2, 2, 0], 7), (1, [64, 64, 1], [1, 1, 0], 128), (2, [2, 2], [0, 1], 3),
             (11, [64, 64, 1], [1, 1, 0], 300))
    for n_in, widths, acts, batch in cases:
        spec, theta, x = _mlp_case(rs, n_in, widths, acts, batch)
        net = net_of(spec)
        n_out = widths[-1]
        out = np.zeros((batch, n_out), np.float32)
        ok(emu.mq_fused_forward(net, ptr(theta), ptr(x), None, batch, ptr(out), None, None, None))
        outs = O.mlp_forward(theta, spec, x)
        close(out, outs[-1], what="fused fwd %s" % (widths,))
        wsb = emu.mq_fused_ws_bytes(net, batch)
        ws = np.zeros(wsb, np.uint8)
        grad = np.zeros(spec["n_params"], np.float32)
        # HEAD_DY
        dy = rs.randn(batch, n_out).astype(np.float32)
        h = _head(L.HEAD_DY, dy=dy)
        out2 = np.zeros((batch, n_out), np.float32)
        ok(emu.mq_fused_grad(net, ptr(theta), ptr(x), None, batch, ctypes.byref(h), 1.0 / batch, ptr(grad),
                             ptr(out2), None, ptr(ws), wsb, None))
        close(grad, O.mlp_backward(theta, spec, x, outs, dy), what="fused grad DY %s b=%d" % (widths, batch))
        close(out2, outs[-1], what="fused grad out")
        # HEAD_DIFF
        tgt = rs.randn(batch, n_out).astype(np.float32)
        h = _head(L.HEAD_DIFF, target=tgt)
        ok(emu.mq_fused_grad(net, ptr(theta), ptr(x), None, batch, ctypes.byref(h), 1.0 / batch, ptr(grad),
                             None, None, ptr(ws), wsb, None))
        close(grad, O.mlp_backward(theta, spec, x, outs, tgt - outs[-1]), what="fused grad DIFF")
        if n_out >= 2 and acts[-1] == 0:
            onehot = np.eye(n_out, dtype=np.float32)[rs.randint(n_out, size=batch)]
            h = _head(L.HEAD_SOFTMAX_CE, p0=0.01, target=onehot)
            ok(emu.mq_fused_grad(net, ptr(theta), ptr(x), None, batch, ctypes.byref(h), 1.0 / batch,
                                 ptr(grad), None, None, ptr(ws), wsb, None))
            close(grad, O.mlp_backward(theta, spec, x, outs, O.softmax_ce_grad(outs[-1], onehot)),
                  what="fused grad CE")


@check
def fused_gauss_ppo():
    rs = np.random.RandomState(5)
    g = dict(np.load(os.path.join(ROOT, "tests/golden/gauss.npz")))
    spec = O.policy_spec(11, 3)
    net = L.make_net(11, [64, 64, 3], [1, 1, 0], logstd=True)
    theta, obs, act, w = g["gauss_theta"], g["gauss_obs"], g["gauss_act"], g["gauss_g"]
    logp = np.zeros(64, np.float32)
    ok(emu.mq_fused_forward(net, ptr(theta), ptr(obs), None, 64, None, ptr(act), ptr(logp), None))
    close(logp, g["gauss_logp"], what="fused logp vs reference")
    wsb = emu.mq_fused_ws_bytes(net, 4096)
    ws = np.zeros(wsb, np.uint8)
    grad = np.zeros(5126, np.float32)
    h = _head(L.HEAD_GAUSS_LOGP, dy=w, target=act)
    ok(emu.mq_fused_grad(net, ptr(theta), ptr(obs), None, 64, ctypes.byref(h), 1.0 / 64, ptr(grad), None,
                         ptr(logp), ptr(ws), wsb, None))
    close(grad, g["gauss_grad"], rtol=3e-5, atol_scale=3e-6, what="fused gauss grad vs reference")
    # PPO head with a gathered minibatch out of a 300-step rollout, multi-tile
    n = 300
    o = np.clip(rs.randn(n, 11), -5, 5).astype(np.float32)
    a = rs.randn(n, 3).astype(np.float32)
    adv = np.tanh(rs.randn(n)).astype(np.float32)
    lp_old = (O.policy_logprob(theta, spec, o, a)[0] + rs.randn(n) * 0.2).astype(np.float32)
    for mb in (64, 200):
        idx = rs.randint(n, size=mb).astype(np.int32)
        h = _head(L.HEAD_GAUSS_PPO, p0=0.8, p1=1.2, target=a, adv=adv, logp_old=lp_old)
        ok(emu.mq_fused_grad(net, ptr(theta), ptr(o), ptr(idx), mb, ctypes.byref(h), 1.0 / mb, ptr(grad),
                             None, None, ptr(ws), wsb, None))
        lp, outs = O.policy_logprob(theta, spec, o[idx], a[idx])
        wgt = O.ppo_clip_weight(lp, lp_old[idx], adv[idx])
        assert (wgt == 0).any() and (wgt != 0).any()
        close(grad, O.policy_logprob_grad(theta, spec, o[idx], a[idx], wgt, outs), rtol=3e-5,
              atol_scale=3e-6, what="fused ppo grad mb=%d" % mb)


@check
def fused_train():
    rs = np.random.RandomState(6)
    spec = O.policy_spec(11, 3)
    net = L.make_net(11, [64, 64, 3], [1, 1, 0], logstd=True)
    theta0 = np.concatenate([O.mlp_init(O._mlp_part(spec), rs, 0.1), np.zeros(3, np.float32)])
    n, steps, mb = 256, 5, 64
    o = np.clip(rs.randn(n, 11), -5, 5).astype(np.float32)
    a = rs.randn(n, 3).astype(np.float32)
    adv = np.tanh(rs.randn(n)).astype(np.float32)
    lp_old = O.policy_logprob(theta0, spec, o, a)[0].astype(np.float32)
    idx = rs.randint(n, size=(steps, mb)).astype(np.int32)
    for dtype, code in ((np.float64, L.MQ_F64), (np.float32, L.MQ_F32)):
        opt = O.AdamO(theta0, horizon=10)
        th_ref = theta0.copy()
        for s in range(steps):
            th_ref = O.ppo_policy_step(th_ref, spec, opt, o[idx[s]], a[idx[s]], adv[idx[s]], lp_old[idx[s]])
        theta = theta0.copy()
        val, m, v = theta0.astype(dtype), np.zeros(5126, dtype), np.zeros(5126, dtype)
        tw = (ctypes.c_double * 2)(0.0, 0.0)
        h = _head(L.HEAD_GAUSS_PPO, p0=0.8, p1=1.2, target=a, adv=adv, logp_old=lp_old)
        ok(emu.mq_fused_train(net, ptr(theta), ptr(val), ptr(m), ptr(v), code, ptr(o), ctypes.byref(h),
                              ptr(idx), steps, mb, 0.9, 0.99, tw, 3e-4, 1e-8, None))
        assert abs(tw[0] - opt.tw1) < 1e-12 and abs(tw[1] - opt.tw2) < 1e-12
        assert np.max(np.abs(theta - th_ref)) < 2e-5, np.max(np.abs(theta - th_ref))
        assert np.max(np.abs(theta - theta0)) > 1e-4
        assert np.array_equal(theta, val.astype(np.float32))
    # value net, DIFF head
    vspec = O.mlp_spec(11, [64, 64, 1], [1, 1, 0])
    vnet = net_of(vspec)
    vt0 = O.mlp_init(vspec, rs, 0.1)
    ret = rs.randn(n, 1).astype(np.float32)
    opt = O.AdamO(vt0, horizon=10, lr=0.003)
    th_ref = vt0.copy()
    for s in range(steps):
        th_ref = O.value_sgd_step(th_ref, vspec, opt, o[idx[s]], ret[idx[s]])
    theta = vt0.copy()
    val, m, v = vt0.astype(np.float64), np.zeros(len(vt0)), np.zeros(len(vt0))
    tw = (ctypes.c_double * 2)(0.0, 0.0)
    h = _head(L.HEAD_DIFF, target=ret)
    ok(emu.mq_fused_train(vnet, ptr(theta), ptr(val), ptr(m), ptr(v), L.MQ_F64, ptr(o), ctypes.byref(h),
                          ptr(idx), steps, mb, 0.9, 0.99, tw, 0.003, 1e-8, None))
    assert np.max(np.abs(theta - th_ref)) < 2e-5, np.max(np.abs(theta - th_ref))


@check
def es():
    rs = np.random.RandomState(7)
    spec = O.mlp_spec(11, [64, 64, 3], [1, 1, 0])
    net = net_of(spec)
    theta = O.mlp_init(spec, rs)
    M, T = 6, 150
    noise = (rs.randn(M, spec["n_params"]) * 0.1).astype(np.float32)
    obs = np.clip(rs.randn(T, 11), -5, 5).astype(np.float32)
    coef = np.array([1.0, -0.5, 0.25], np.float32)
    rets = np.zeros(M, np.float32)
    ok(emu.mq_es_eval(net, ptr(theta), ptr(noise), M, ptr(obs), T, ptr(coef), ptr(rets), None))
    want = np.zeros(M)
    for p in range(M):
        th = (theta.astype(np.float64) + noise[p]).astype(np.float32)
        want[p] = np.sum(O.mlp_forward(th, spec, obs)[-1] @ coef.astype(np.float64))
    close(rets, want, rtol=2e-5, atol_scale=1e-5, what="es returns")
    w = rs.randn(M)
    wsb = emu.mq_es_ws_bytes(M, spec["n_params"])
    ws = np.zeros(wsb, np.uint8)
    out = np.zeros(spec["n_params"])
    ok(emu.mq_es_weighted_sum(ptr(noise), ptr(w), M, spec["n_params"], 1.0 / M, ptr(out), ptr(ws), wsb, None))
    assert np.allclose(out, (noise.astype(np.float64) * w[:, None]).mean(axis=0), rtol=1e-12, atol=1e-15)


@check
def ops():
    rs = np.random.RandomState(8)
    a = rs.randn(33, 70).astype(np.float32)
    b = rs.randn(70, 21).astype(np.float32)
    c = np.zeros((33, 21), np.float32)
    ok(emu.mq_gemm(ptr(a), 70, 1, ptr(b), 21, 1, ptr(c), 33, 21, 70, 1.0, None))
    close(c, a.astype(np.float64) @ b, what="gemm NN")
    ct = np.zeros((70, 70), np.float32)
    ok(emu.mq_gemm(ptr(a), 1, 70, ptr(a), 70, 1, ptr(ct), 70, 70, 33, 0.5, None))
    close(ct, 0.5 * a.T.astype(np.float64) @ a, what="gemm TN")
    g = rs.randn(50, 7).astype(np.float32)
    x = rs.randn(50, 7).astype(np.float32)
    out = np.zeros(7, np.float32)
    ok(emu.mq_col_sum(ptr(g), ptr(x), 50, 7, 0.5, ptr(out), None))
    close(out, 0.5 * (g * x).sum(axis=0), what="col_sum")
    y = np.zeros_like(x)
    ok(emu.mq_binary(ptr(x), ptr(out), ptr(y), 350, 7, 1, None))
    assert np.array_equal(y, x * out)
    mu, ls, s = rs.randn(9, 3).astype(np.float32), (rs.randn(3) * 0.3).astype(np.float32), rs.randn(9, 3).astype(np.float32)
    lp = np.zeros(9, np.float32)
    ok(emu.mq_gauss_logprob(ptr(mu), ptr(ls), 3, ptr(s), 9, 3, ptr(lp), None))
    close(lp, O.gauss_logprob(mu, ls, s), what="gauss logp")
    w = rs.randn(9).astype(np.float32)
    dmu, dls = np.zeros((9, 3), np.float32), np.zeros((9, 3), np.float32)
    ok(emu.mq_gauss_logprob_bwd(ptr(mu), ptr(ls), 3, ptr(s), ptr(w), 9, 3, ptr(dmu), ptr(dls), None))
    wm, wl = O.gauss_logprob_grads(mu, ls, s, w)
    close(dmu, wm, what="gauss dmu")
    close(dls, wl, what="gauss dls")
    m, l = rs.randn(5, 3).astype(np.float32), (rs.randn(5, 3) * 0.5).astype(np.float32)
    o5 = np.zeros(5, np.float32)
    ok(emu.mq_dkl_forward(ptr(m), ptr(l), 5, 3, ptr(o5), None))
    close(o5, O.dkl_uninormal(m, l), what="dkl")
    lg = rs.randn(6, 10).astype(np.float32)
    oh = np.eye(10, dtype=np.float32)[rs.randint(10, size=6)]
    dy = np.zeros((6, 10), np.float32)
    ok(emu.mq_softmax_ce_grad(ptr(lg), ptr(oh), None, 6, 10, 0.01, ptr(dy), None))
    close(dy, O.softmax_ce_grad(lg, oh), what="softmax ce")


if __name__ == "__main__":
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    for fn in CHECKS:
        if flt in fn.__name__:
            t = time.time()
            fn()
            print("ok  %-22s %.1fs" % (fn.__name__, time.time() - t))
