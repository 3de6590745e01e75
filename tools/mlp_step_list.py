"""Developer aid: one un-graphed optimiser step of the layered MLP at a given batch (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mannequin2_b200 import _device as D
from mannequin2_b200.mlp import MLPTrainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
rs = np.random.RandomState(0)
x = D.from_host(rs.rand(60000, 784).astype(np.float32)); y = D.from_host(np.eye(10, dtype=np.float32)[rs.randint(10, size=60000)])
tr = MLPTrainer(batch=B, graph_steps=1)
tr.idx.copy_(D.from_host(rs.randint(60000, size=B).astype(np.int32)))
for _ in range(3):
    tr._enqueue_step(x, y, 0)
torch.cuda.synchronize()
