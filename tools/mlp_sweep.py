"""Developer aid: the C1 batch sweep alone.  python tools/mlp_sweep.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
torch.cuda.set_device(0)
wl = bench.MLPWorkload(seed=7)
wl.setup_gpu()
for k, v in wl.sweep(bench.load_peaks()).items():
    print(k, json.dumps({a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()}))
