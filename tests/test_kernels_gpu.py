"""Parity tests proper: every kernel, called through the C-ABI with device buffers
on a B200, against the CPU oracle and the reference-generated golden fixtures."""
import numpy as np
import pytest

import kernel_cases as K

pytestmark = pytest.mark.gpu


class CudaBackend(object):
    def __init__(self):
        import torch
        from mannequin2_b200 import _lib
        assert torch.cuda.is_available(), "these tests need a CUDA device"
        self.torch = torch
        self.lib = _lib.lib()
        self.dev = torch.device("cuda", 0)
        self.stream = torch.cuda.current_stream().cuda_stream

    def up(self, a):
        a = np.ascontiguousarray(a)
        if a.dtype == np.bool_:
            a = a.view(np.uint8)
        return self.torch.from_numpy(a.copy()).to(self.dev)

    def zeros(self, shape, dtype):
        return self.up(np.zeros(shape, dtype=dtype))

    def ptr(self, h):
        return None if h is None else h.data_ptr()

    def down(self, h):
        return h.detach().cpu().numpy()


@pytest.fixture(scope="module")
def backend():
    return CudaBackend()


@pytest.mark.parametrize("case", K.ALL_CASES, ids=lambda f: f.__name__)
def test_kernel_case(backend, case):
    case(backend)
    backend.torch.cuda.synchronize()


def test_large_scan_and_adam_properties(backend):
    """BASELINE-size inputs (1M transitions, 118,282 params): size-independent properties."""
    B, lib, L = backend, backend.lib, K.L
    rs = np.random.RandomState(11)
    n = 1 << 20
    r = rs.randn(n).astype(np.float32)
    v = rs.randn(n + 1).astype(np.float32)
    d = (rs.rand(n) < 1.0 / 500).astype(np.uint8)
    d[999::1000] = 1                                       # forced cut every 1000 steps
    wsb = lib.mq_scan_ws_bytes(n)
    hr, hv, hd, ws = B.up(r), B.up(v), B.up(d), B.zeros(wsb, np.uint8)
    adv, ret = B.zeros(n, np.float32), B.zeros(n, np.float32)
    K.ok(B, lib.mq_gae(B.ptr(hr), B.ptr(hv), B.ptr(hd), n, 0.99, 0.95, B.ptr(adv), B.ptr(ret), B.ptr(ws),
                       wsb, B.stream))
    a, rt = B.down(adv), B.down(ret)
    # recurrence residual (checks every element without a sequential pass)
    nxt = np.append(a[1:], np.float32(0.0)).astype(np.float64)
    cont = 1.0 - d.astype(np.float64)
    want = r - v[:-1] + cont * (np.float32(0.99) * v[1:].astype(np.float64) + 0.99 * 0.95 * nxt)
    assert np.max(np.abs(a - want)) < 2e-5
    assert np.array_equal(rt, a + v[:-1])
    term = d.astype(bool)
    assert np.array_equal(a[term], (r - v[:-1])[term])     # episode ends are exact
    # linearity of the scan in the rewards
    hr2 = B.up((2 * r).astype(np.float32))
    hz = B.zeros(n + 1, np.float32)
    adv2 = B.zeros(n, np.float32)
    K.ok(B, lib.mq_gae(B.ptr(hr2), B.ptr(hz), B.ptr(hd), n, 0.99, 0.95, B.ptr(adv2), B.ptr(ret), B.ptr(ws),
                       wsb, B.stream))
    adv1 = B.zeros(n, np.float32)
    K.ok(B, lib.mq_gae(B.ptr(hr), B.ptr(hz), B.ptr(hd), n, 0.99, 0.95, B.ptr(adv1), B.ptr(ret), B.ptr(ws),
                       wsb, B.stream))
    assert np.array_equal(B.down(adv2), 2 * B.down(adv1))
    # Adam at the MNIST MLP size, fp64 state vs oracle, 3 steps: bit exact
    from oracle import mq_oracle as O
    p = 118282
    val0 = rs.randn(p)
    ref = O.AdamO(val0, horizon=10, lr=1e-3)
    val, m, vv, th = B.up(val0.copy()), B.zeros(p, np.float64), B.zeros(p, np.float64), B.zeros(p, np.float32)
    tw1 = tw2 = 0.0
    for _ in range(3):
        g = rs.randn(p).astype(np.float32)
        ref.apply_gradient(g)
        tw1, tw2 = tw1 * 0.9 + 1, tw2 * 0.99 + 1
        hg = B.up(g)
        K.ok(B, lib.mq_adam_step(B.ptr(val), B.ptr(m), B.ptr(vv), B.ptr(th), B.ptr(hg), p, L.MQ_F64, L.MQ_F32,
                                 1 / tw1, 1 / tw2, 1e-3, 1e-8, 0, B.stream))
    assert np.array_equal(B.down(val), ref.get_value())
    assert np.array_equal(B.down(th), ref.get_value().astype(np.float32))


def test_large_batch_fused_grad_is_permutation_and_split_consistent(backend):
    """65,536-row minibatch: gradient of the whole batch equals the mean of the
    gradients of its two halves; a fixed launch is bit-reproducible."""
    import ctypes
    B, lib, L = backend, backend.lib, K.L
    from oracle import mq_oracle as O
    rs = np.random.RandomState(12)
    spec = O.policy_spec(11, 3)
    net = L.make_net(11, [64, 64, 3], [1, 1, 0], logstd=True)
    theta = np.concatenate([O.mlp_init(O._mlp_part(spec), rs, 0.1), (rs.randn(3) * 0.1).astype(np.float32)])
    n = 65536
    o = np.clip(rs.randn(n, 11), -5, 5).astype(np.float32)
    a = rs.randn(n, 3).astype(np.float32)
    adv = np.tanh(rs.randn(n)).astype(np.float32)
    lp_old = (O.policy_logprob(theta, spec, o, a)[0] + rs.randn(n) * 0.2).astype(np.float32)
    ht, ho, ha, hadv, hlp = B.up(theta), B.up(o), B.up(a), B.up(adv), B.up(lp_old)
    wsb = lib.mq_fused_ws_bytes(net, n)
    ws = B.zeros(wsb, np.uint8)

    def grad_of(lo, hi):
        idx = B.up(np.arange(lo, hi, dtype=np.int32))
        g = B.zeros(5126, np.float32)
        h = K.head(B, L.HEAD_GAUSS_PPO, p0=0.8, p1=1.2, target=ha, adv=hadv, logp_old=hlp)
        K.ok(B, lib.mq_fused_grad(net, B.ptr(ht), B.ptr(ho), B.ptr(idx), hi - lo, ctypes.byref(h),
                                  1.0 / (hi - lo), B.ptr(g), None, None, B.ptr(ws), wsb, B.stream))
        return B.down(g)

    full, again = grad_of(0, n), grad_of(0, n)
    assert np.array_equal(full, again)                                  # deterministic reduction order
    halves = 0.5 * (grad_of(0, n // 2).astype(np.float64) + grad_of(n // 2, n))
    K.close(full, halves, rtol=1e-4, atol_scale=1e-5, what="split consistency")
    # the same batch four times over (262,144 rows: 14 tiles per CTA, so the tensor-core kernel folds its TMEM
    # accumulators into memory three times on the way) gives the same mean gradient
    idx4 = B.up(np.tile(np.arange(n, dtype=np.int32), 4))
    g4 = B.zeros(5126, np.float32)
    h = K.head(B, L.HEAD_GAUSS_PPO, p0=0.8, p1=1.2, target=ha, adv=hadv, logp_old=hlp)
    K.ok(B, lib.mq_fused_grad(net, B.ptr(ht), B.ptr(ho), B.ptr(idx4), 4 * n, ctypes.byref(h), 1.0 / (4 * n), B.ptr(g4),
                              None, None, B.ptr(ws), wsb, B.stream))
    K.close(B.down(g4), full, rtol=1e-5, atol_scale=2e-6, what="4x repeated batch")
    # and against the oracle on a 4096-row sample of the same data
    idx = np.arange(4096)
    lp, outs = O.policy_logprob(theta, spec, o[idx], a[idx])
    w = O.ppo_clip_weight(lp, lp_old[idx], adv[idx])
    K.close(grad_of(0, 4096), O.policy_logprob_grad(theta, spec, o[idx], a[idx], w, outs), rtol=3e-5,
            atol_scale=3e-6, what="fused ppo grad 4096")
