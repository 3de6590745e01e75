"""Developer aid: repeat bench.py's precision-mode section to catch an intermittent hang (stack dump after a timeout)."""
import faulthandler
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

faulthandler.dump_traceback_later(float(sys.argv[2]) if len(sys.argv) > 2 else 100.0, exit=True, file=sys.stderr)
import torch  # noqa: E402

torch.cuda.set_device(0)
wl = bench.WORKLOADS["ppo_large"](seed=1, tc_mode="fp16x2")
wl.setup_gpu()
wl.step_gpu(0)
torch.cuda.synchronize()
if len(sys.argv) > 3 and sys.argv[3] == "fork":
    bench.cpu_three_figures("ppo_large", 4)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 10):
    out = wl.precision_modes()
    torch.cuda.synchronize()
    sys.stderr.write("iteration %d ok: %s\n" % (it, {k: round(v["ms_per_step"], 2) for k, v in out.items()}))
    sys.stderr.flush()
