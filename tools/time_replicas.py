"""Aggregate throughput of R independent small-batch PPO learners on one GPU (bench.py PPOWorkload)."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402

w = bench.PPOWorkload(seed=7)
w.setup_gpu()
for r in (1, 2, 4, 8, 16, 32, 64):
    print(json.dumps(w.replica_throughput(r)), flush=True)
