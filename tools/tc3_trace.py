"""Developer aid: time line of one CTA of the record-fed tensor-core gradient kernel.  Build the library with
MQ_TC_PROFILE=1 (python -m mannequin2_b200.build --force), then: python tools/tc3_trace.py [planes] [rows]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mannequin2_b200 import _lib as L          # noqa: E402

planes = int(sys.argv[1]) if len(sys.argv) > 1 else 3
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
lib = L.lib()
dev = torch.device("cuda", 0)
rs = np.random.RandomState(3)
net = L.make_net(11, [64, 64, 3], [1, 1, 0], logstd=True)
n = 1 << 18
theta = torch.from_numpy((rs.randn(5126) * 0.1).astype(np.float32)).to(dev)
obs, act, adv, lp = (torch.randn(n, 11, device=dev), torch.randn(n, 3, device=dev), torch.randn(n, device=dev).tanh(),
                     torch.randn(n, device=dev) - 3)
idx = torch.randint(n, (mb,), dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
rec_w = lib.mq_record_width(net)
rec = torch.zeros(n * rec_w, device=dev)
L.check(lib.mq_pack_records(net, obs.data_ptr(), act.data_ptr(), adv.data_ptr(), lp.data_ptr(), n, rec.data_ptr(), stream))
image = torch.zeros(lib.mq_tc_image_bytes(net, planes), dtype=torch.uint8, device=dev)
L.check(lib.mq_tc_image_build(net, theta.data_ptr(), planes, image.data_ptr(), stream))
wsb = lib.mq_fused_ws_bytes(net, mb)
ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
h = L.MqHead()
h.kind, h.p0, h.p1 = L.HEAD_GAUSS_PPO, 0.8, 1.2
h.target, h.adv, h.logp_old = act.data_ptr(), adv.data_ptr(), lp.data_ptr()
h.tc, h.records, h.rec_w, h.tc_image = L.TC_FORCE | planes, rec.data_ptr(), rec_w, image.data_ptr()
prof = torch.zeros(512, dtype=torch.int64, device=dev)
n_parts = ctypes.c_int32(0)
for it in range(3):
    if it == 2:
        lib.mq_debug_set_phase_buffer(prof.data_ptr())
    L.check(lib.mq_fused_grad_partials(net, theta.data_ptr(), obs.data_ptr(), idx.data_ptr(), mb, ctypes.byref(h), None, None,
                                       ws.data_ptr(), wsb, ctypes.byref(n_parts), stream))
    torch.cuda.synchronize()
lib.mq_debug_set_phase_buffer(None)
p = prof.cpu().numpy()
names = {1: "start", 2: "setup done", 3: "pdl wait done", 4: "image landed", 10: "tile: begin", 11: "staged", 20: "F0 done", 21: "F1 done",
         30: "A1 published", 31: "A2 published", 40: "F-last done", 41: "dZ_L published", 52: "dA2 done", 51: "dA1 done",
         62: "dZ1 computed", 61: "dZ0 computed", 72: "dW2 done", 71: "dW1 done", 82: "dZ1 published", 81: "dZ0 published",
         90: "flush begin", 91: "flush end"}
for base, who in ((0, "worker thread 0"), (256, "issuer")):
    print("---- %s (planes %d, %d rows, %d partials)" % (who, planes, mb, n_parts.value))
    t0 = p[1]
    prev = None
    for k in range(127):
        tag, t = int(p[base + 2 * k]), int(p[base + 2 * k + 1])
        if tag == 0:
            break
        if base:
            s, job = (tag % 100) // 10, tag % 10
            nm = ("slot %d job %d: ready seen" if tag < 200 else "slot %d job %d: issued") % (s, job)
        else:
            nm = names.get(tag, str(tag))
        print("%8d  (+%6d)  %s" % (t - t0, 0 if prev is None else t - prev, nm))
        prev = t
