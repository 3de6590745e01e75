#!/usr/bin/env python3
"""Developer-only: run tests/kernel_cases.py on the HOST EMULATION of the kernels.

    tools/emu/build_emu.sh && python tools/emu/check_kernels.py [filter]

Not part of pytest, not used by the product (see README.md).  The parity tests
proper are `pytest -m gpu` (tests/test_kernels_gpu.py), which run the very same
cases on a B200 through the C-ABI with device buffers.
"""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from mannequin2_b200 import _lib as L          # noqa: E402  (signature table only)
import kernel_cases as K                        # noqa: E402


class EmuBackend(object):
    stream = None

    def __init__(self):
        self.lib = L.bind(ctypes.CDLL(os.path.join(ROOT, "tools/emu/_build/libmq_emu.so")))

    def up(self, a):
        a = np.array(a, copy=True, order="C")
        return a.view(np.uint8) if a.dtype == np.bool_ else a

    def zeros(self, shape, dtype):
        return np.zeros(shape, dtype=dtype)

    def ptr(self, h):
        return None if h is None else h.ctypes.data

    def down(self, h):
        return np.array(h, copy=True)


if __name__ == "__main__":
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    B = EmuBackend()
    small = {"case_scan": dict(sizes=(1, 2, 700, 2048, 2049, 5000)),
             "case_ppo_reference_replay": dict(n_steps=20)}
    # tensor-core (tcgen05) and peer-memory cases have no host emulation: they are checked on a B200 only
    gpu_only = {"case_gemm_tc", "case_tc_paths", "case_tc_modes", "case_reduce_step"}
    small["case_discrete"] = dict(batches=(64, 300))
    small["case_entropy"] = dict(batches=(64, 300))
    for fn in K.ALL_CASES:
        if fn.__name__ in gpu_only:
            print("--  %-28s (B200 only)" % fn.__name__)
            continue
        if flt in fn.__name__:
            t = time.time()
            fn(B, **small.get(fn.__name__, {}))
            print("ok  %-28s %.1fs" % (fn.__name__, time.time() - t))
