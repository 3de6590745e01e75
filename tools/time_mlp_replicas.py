"""Developer aid: aggregate throughput of R independent batch-128 MLP trainers on one GPU (bench.py MLPWorkload)."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402

w = bench.MLPWorkload(seed=7)
w.setup_gpu()
for r in (1, 4, 8, 16, 32):
    print(json.dumps(w.replica_throughput(r)), flush=True)
