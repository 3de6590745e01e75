"""Developer aid: a few launches of one product on the TMA GEMM, for an `ncu --set full` capture.
    python tools/gemm_one.py M N K [exact|fp16x2|bf16x2|bf16]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mannequin2_b200 import _device as D          # noqa: E402
from mannequin2_b200 import _lib as L             # noqa: E402

m, n, k = (int(v) for v in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "fp16x2"
lib, dev = L.lib(), D.device()
a, b, c = torch.randn(m, k, device=dev), torch.randn(k, n, device=dev), torch.empty(m, n, device=dev)
wsb = int(lib.mq_gemm_ws_bytes(m, n, k))
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
for _ in range(4):
    D.call("mq_gemm_ex", D.ptr(a), k, 1, D.ptr(b), n, 1, D.ptr(c), m, n, k, 1.0, L.TC_MODES[mode] | L.TC_FORCE, D.ptr(ws), wsb, D.stream())
torch.cuda.synchronize()
ref = a.double() @ b.double()
print("max rel err", float((c.double() - ref).abs().max() / ref.abs().max()))
