"""Developer aid: time mq_gemm_ex of a given build of the library (A/B of two builds on the same box).
    python tools/gemm_ab.py path/to/libmq_b200.so"""
import ctypes
import sys

import torch

lib = ctypes.CDLL(sys.argv[1])
lib.mq_gemm_ws_bytes.restype = ctypes.c_int64
lib.mq_gemm_ws_bytes.argtypes = [ctypes.c_int64] * 3
P, I64 = ctypes.c_void_p, ctypes.c_int64
lib.mq_gemm_ex.argtypes = [P, I64, I64, P, I64, I64, P, I64, I64, I64, ctypes.c_float, ctypes.c_int32, P, I64, P]
dev = torch.device("cuda:0")
MODES = {"exact": 3, "fp16x2": 2 | 0x80, "bf16": 1}


def timeit(fn, reps=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for name, m, n, k in (("4096x400x784", 4096, 400, 784), ("4096x784x400", 4096, 784, 400), ("4096^3", 4096, 4096, 4096),
                      ("65536x128x784", 65536, 128, 784)):
    a, b, c = torch.randn(m, k, device=dev), torch.randn(k, n, device=dev), torch.empty(m, n, device=dev)
    wsb = int(lib.mq_gemm_ws_bytes(m, n, k))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    line = "%-16s" % name
    for mode, flag in MODES.items():
        ms = timeit(lambda: lib.mq_gemm_ex(a.data_ptr(), k, 1, b.data_ptr(), n, 1, c.data_ptr(), m, n, k, 1.0, flag | 0x10, ws.data_ptr(), wsb, st))
        line += "  %s %.1f us" % (mode, ms * 1e3)
    print(line)
