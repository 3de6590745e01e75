"""Developer aid: time mq_gemm (tcgen05 vs FFMA) for the VAE / MNIST layer shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mannequin2_b200 import _lib as L, _device as D
lib, dev = L.lib(), D.device()
def timeit(fn, reps=8):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[2:]))
shapes = [("y=xW   4096x400x784", 4096, 400, 784, "nn"), ("dx=gW^T 4096x784x400", 4096, 784, 400, "nt"), ("dW=x^Tg 784x400x4096", 784, 400, 4096, "tn"),
          ("y=xW   8192x128x784", 8192, 128, 784, "nn"), ("square 4096^3", 4096, 4096, 4096, "nn"), ("y=xW 65536x64x64", 65536, 64, 64, "nn")]
for name, m, n, k, lay in shapes:
    a = torch.randn((k, m) if lay[0] == "t" else (m, k), device=dev)
    b = torch.randn((n, k) if lay[1] == "t" else (k, n), device=dev)
    c = torch.empty(m, n, device=dev)
    a_rs, a_cs = (1, m) if lay[0] == "t" else (k, 1)
    b_rs, b_cs = (1, k) if lay[1] == "t" else (n, 1)
    res = []
    for mode in (2, 0):
        lib.mq_debug_set_gemm_tc_mode(mode)
        ms = timeit(lambda: D.call("mq_gemm", D.ptr(a), a_rs, a_cs, D.ptr(b), b_rs, b_cs, D.ptr(c), m, n, k, 1.0, D.stream()))
        res.append((ms, 2.0 * m * n * k / ms / 1e9))
    print("%-24s tcgen05 %8.1f us %7.1f TFLOP/s | ffma %8.1f us %6.1f TFLOP/s" % (name, res[0][0] * 1e3, res[0][1], res[1][0] * 1e3, res[1][1]))
lib.mq_debug_set_gemm_tc_mode(1)
