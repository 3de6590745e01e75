"""Developer aid: time line of the last step of a fused MLP block (csrc/mq_mlp_fused.cu), CTAs 0 and 127, in ns."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mannequin2_b200 import _device as D, _lib as L     # noqa: E402
from mannequin2_b200.mlp import MLPTrainer               # noqa: E402

rs = np.random.RandomState(1)
n = 60000
x = D.from_host(rs.rand(n, 784).astype(np.float32))
y = D.from_host(np.eye(10, dtype=np.float32)[rs.randint(10, size=n)])
tab = D.from_host(rs.randint(n, size=(32, 128)).astype(np.int32))
tr = MLPTrainer(graph_steps=32)
prof = torch.zeros(32, dtype=torch.int64, device="cuda")
for it in range(3):
    tr.train_block(x, y, tab)
torch.cuda.synchronize()
L.lib().mq_debug_set_phase_buffer(prof.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
tr.train_block(x, y, tab)
e1.record()
torch.cuda.synchronize()
L.lib().mq_debug_set_phase_buffer(None)
p = prof.cpu().numpy()
names = {0: "step begin", 1: "x slice landed", 2: "F0 done", 3: "barrier 1 passed", 7: "  MID operands in", 8: "  h2 done", 9: "  logits done",
         10: "  head done", 11: "  dh2 done", 4: "MID done", 5: "barrier 2 passed", 12: "  D operands in", 13: "  dW0 + Adam done", 6: "D done"}
order = [0, 1, 2, 3, 7, 8, 9, 10, 11, 4, 5, 12, 13, 6]
print("block of 32 steps: %.1f us = %.2f us per step (train_block call, events); first step begin -> last step end inside the kernel: %.2f us per step"
      % (e0.elapsed_time(e1) * 1e3, e0.elapsed_time(e1) * 1e3 / 32, (p[15] - p[14]) / 32e3))
for base, who in ((0, "CTA 0"), (16, "CTA 127")):
    print(who)
    prev = p[base]
    for i in order:
        print("  %-20s %7d ns  (+%5d)" % (names[i], p[base + i] - p[0], p[base + i] - prev))
        prev = p[base + i]
