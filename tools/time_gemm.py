"""Developer aid: time mq_gemm_ex (pack + TMA-fed tcgen05 GEMM per precision mode, and the FFMA kernel) for the VAE / MNIST
layer shapes.  python tools/time_gemm.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mannequin2_b200 import _device as D          # noqa: E402
from mannequin2_b200 import _lib as L             # noqa: E402

lib = L.lib()
dev = D.device()


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


shapes = [("y = x W   4096x400x784", 4096, 400, 784, False, False), ("dx = g W^T 4096x784x400", 4096, 784, 400, False, True),
          ("dW = x^T g 784x400x4096", 784, 400, 4096, True, False), ("4096^3", 4096, 4096, 4096, False, False),
          ("mlp l1 65536x128x784", 65536, 128, 784, False, False)]
for name, m, n, k, a_t, b_t in shapes:
    a = torch.randn((k, m) if a_t else (m, k), device=dev)
    b = torch.randn((n, k) if b_t else (k, n), device=dev)
    c = torch.empty(m, n, device=dev)
    a_rs, a_cs = (1, m) if a_t else (k, 1)
    b_rs, b_cs = (1, k) if b_t else (n, 1)
    wsb = int(lib.mq_gemm_ws_bytes(m, n, k))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    fl = 2.0 * m * n * k
    line = "%-28s" % name
    for tc, nm in ((3, "exact"), (0x82, "fp16x2"), (2, "bf16x2"), (1, "bf16"), (L.TC_OFF, "ffma")):
        ms = timeit(lambda: D.call("mq_gemm_ex", D.ptr(a), a_rs, a_cs, D.ptr(b), b_rs, b_cs, D.ptr(c), m, n, k, 1.0, tc, D.ptr(ws), wsb,
                                   D.stream()))
        line += "  %s %.1f us = %.0f TF" % (nm, ms * 1e3, fl / ms / 1e9)
    print(line, flush=True)
