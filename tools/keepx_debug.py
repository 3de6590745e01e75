"""Developer aid: layered MLP gradient with / without MQ_TC_KEEP_X per mode and net, error map of the first layer's dW."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import kernel_cases as K
from test_kernels_gpu import CudaBackend
from oracle import mq_oracle as O
B = CudaBackend(); lib = B.lib; L = K.L
for widths, acts in (([128, 10], [2, 0]), ([128, 128, 10], [2, 2, 0])):
    for mode in ("exact", "fp16x2"):
        for keep in (0, L.TC_KEEP_X):
            for batch in (8192, 8200):
                rs = np.random.RandomState(5)
                spec, theta, x = K.mlp_case(rs, 784, widths, acts, 3000)
                idx = rs.randint(3000, size=batch).astype(np.int32)
                net = K.net_of(spec)
                tc = L.TC_MODES[mode] | keep
                ht, hx, hi = B.up(theta), B.up(x), B.up(idx)
                acts_buf = B.zeros(lib.mq_mlp_acts_floats(net, batch), np.float32)
                wsb = lib.mq_mlp_ws_bytes(net, batch)
                ws = B.zeros(wsb, np.uint8)
                K.ok(B, lib.mq_mlp_forward_tc(net, B.ptr(ht), B.ptr(hx), B.ptr(hi), batch, B.ptr(acts_buf), tc, B.ptr(ws), wsb, B.stream))
                outs = O.mlp_forward(theta, spec, x[idx])
                dy = rs.randn(batch, widths[-1]).astype(np.float32)
                want, wdx = O.mlp_backward(theta, spec, x[idx], outs, dy, want_dx=True)
                grad, dx, hdy = B.zeros(spec["n_params"], np.float32), B.zeros((batch, 784), np.float32), B.up(dy.copy())
                K.ok(B, lib.mq_mlp_backward_tc(net, B.ptr(ht), B.ptr(hx), B.ptr(hi), B.ptr(acts_buf), B.ptr(hdy), batch,
                                               1.0 / batch, B.ptr(grad), B.ptr(dx), tc, B.ptr(ws), wsb, B.stream, None, None))
                got = B.down(grad)
                err = np.abs(got - want)
                w0 = err[:785 * widths[0]].reshape(785, widths[0])
                bad = np.argwhere(w0 > 1e-5 * np.max(np.abs(want)) + 1e-7)
                print(widths, mode, "keep" if keep else "pack", batch, "max err %.3g (max|want| %.3g)" % (err.max(), np.abs(want).max()),
                      "bad W0 entries %d" % len(bad), "rows %s cols %s" % (sorted(set(bad[:, 0]))[:12], sorted(set(bad[:, 1]))[:12]) if len(bad) else "")
