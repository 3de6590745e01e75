"""Developer aid: the layered and the fused MLP paths against the oracle over several block splits."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mq_oracle as O                   # noqa: E402
from mannequin2_b200.mlp import MLPTrainer          # noqa: E402
from mannequin2_b200.ppo import init_mlp            # noqa: E402
from mannequin2_b200 import _device as D            # noqa: E402

rs = np.random.RandomState(18)
n, B = 700, 128
x = rs.rand(n, 784).astype(np.float32)
y = np.eye(10, dtype=np.float32)[rs.randint(10, size=n)]
np.random.seed(4)
theta0 = init_mlp(784, [128, 128, 10], 0.1)
spec = O.mlp_spec(784, [128, 128, 10], [O.ACT_LRELU, O.ACT_LRELU, O.ACT_NONE])
xd, yd = D.from_host(x), D.from_host(y)
for split in ((10,), (10,), (10,), (5, 5), (5, 4), (12,)):
    steps = sum(split)
    tab = rs.randint(n, size=(steps, B)).astype(np.int32)
    opt = O.AdamO(theta0, horizon=10, lr=1e-3)
    want = theta0.copy()
    trace = []
    for s_ in range(steps):
        want = O.mlp_sgd_step(want, spec, opt, x[tab[s_]], y[tab[s_]])
        trace.append(want.copy())
    for fused in (True, False):
        tr = MLPTrainer(params=theta0.copy(), fused=fused, graph_steps=16)
        o = 0
        for k in split:
            tr.train_block(xd, yd, tab[o:o + k])
            o += k
        print(split, "fused" if fused else "layered", float(np.max(np.abs(tr.get_params() - want))), flush=True)
    # step by step on the layered path: where does it leave the oracle?
    tr = MLPTrainer(params=theta0.copy(), fused=False, graph_steps=16)
    errs = []
    for s_ in range(steps):
        tr.train_block(xd, yd, tab[s_:s_ + 1])
        errs.append(float(np.max(np.abs(tr.get_params() - trace[s_]))))
    print("   layered, one step per block:", ["%.1e" % e for e in errs], flush=True)
