"""Developer aid: the HBM-bound kernels at 64 Mi elements (GAE scan, normaliser update + apply), timed with CUDA events."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from mannequin2_b200 import _device as D, _lib as L
from mannequin2_b200 import RunningNormalize

lib, dev = L.lib(), D.device()
n = 64 << 20
r, vv = torch.randn(n, device=dev), torch.randn(n + 1, device=dev)
d = (torch.rand(n, device=dev) < 0.002).to(torch.uint8)
adv, ret = torch.empty(n, device=dev), torch.empty(n, device=dev)
wsb = lib.mq_scan_ws_bytes(n)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
norm = RunningNormalize(horizon=2)
o = torch.empty(n, device=dev)


def timeit(fn, reps=6, inner=8):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _i in range(inner):
            fn()
        e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    return float(np.median(ts[1:]))


ms = timeit(lambda: D.call("mq_gae", D.ptr(r), D.ptr(vv), D.ptr(d), n, 0.99, 0.95, D.ptr(adv), D.ptr(ret), D.ptr(ws), wsb,
                           D.stream()))
print("gae 64Mi: %.3f ms  %.0f GB/s (17 B/transition)" % (ms, 17.0 * n / ms / 1e6))
def nrm():
    norm.update_dev(r, n)
    norm.apply_dev(r, n, out=o, fuse_tanh=True)


nrm()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(4):
        nrm()
ms = timeit(g.replay, inner=2) / 4
print("normalise 64Mi (update + apply, graph replay): %.3f ms -> %.0f GB/s (12 B/element: one statistics pass + apply)" % (
    ms, 12.0 * n / ms / 1e6))

ya = torch.empty(n, device=dev)
ms = timeit(lambda: D.call("mq_act_forward", D.ptr(r), D.ptr(ya), n, L.ACT_TANH, 0.0, D.stream()))
print("tanh forward 64Mi: %.4f ms -> %.0f GB/s (8 B/element)" % (ms, 8.0 * n / ms / 1e6))
ms = timeit(lambda: D.call("mq_act_backward", D.ptr(ya), D.ptr(r), D.ptr(o), n, L.ACT_TANH, 0.0, D.stream()))
print("tanh backward 64Mi: %.4f ms -> %.0f GB/s (12 B/element)" % (ms, 12.0 * n / ms / 1e6))
