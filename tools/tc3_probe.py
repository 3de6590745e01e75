"""Developer probe: record-fed tensor-core gradient kernel (csrc/mq_tc3.cu) -- error against the oracle and launch time
in each precision mode, next to the array-fed kernel of csrc/mq_tc.cu.  python tools/tc3_probe.py [rows ...]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mannequin2_b200 import _lib as L          # noqa: E402
from oracle import mq_oracle as O              # noqa: E402

lib = L.lib()
dev = torch.device("cuda", 0)
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
rs = np.random.RandomState(3)
spec = O.policy_spec(11, 3)
net = L.make_net(11, [64, 64, 3], [1, 1, 0], logstd=True)
theta = np.concatenate([O.mlp_init(O._mlp_part(spec), rs, 0.1), (rs.randn(3) * 0.1).astype(np.float32)])
n = 1 << 20
o = np.clip(rs.randn(n, 11), -5, 5).astype(np.float32)
a = rs.randn(n, 3).astype(np.float32)
adv = np.tanh(rs.randn(n)).astype(np.float32)
sizes = [int(v) for v in sys.argv[1:]] or [8192, 65536]
check_rows = 1 << 17
lp_old = np.zeros(n, np.float32)
lp_old[:check_rows] = (O.policy_logprob(theta, spec, o[:check_rows], a[:check_rows])[0] + rs.randn(check_rows) * 0.2)
ht, ho, ha, hadv, hlp = map(up, (theta, o, a, adv, lp_old))
stream = torch.cuda.current_stream().cuda_stream
rec_w = lib.mq_record_width(net)
rec = torch.zeros(n * rec_w, device=dev)
L.check(lib.mq_pack_records(net, ho.data_ptr(), ha.data_ptr(), hadv.data_ptr(), hlp.data_ptr(), n, rec.data_ptr(), stream))
wsb = lib.mq_fused_ws_bytes(net, 1 << 16)
ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def head(planes, records, image=None, force=True):
    h = L.MqHead()
    h.kind, h.p0, h.p1 = L.HEAD_GAUSS_PPO, 0.8, 1.2
    h.target, h.adv, h.logp_old = ha.data_ptr(), hadv.data_ptr(), hlp.data_ptr()
    h.tc = (L.TC_FORCE if force else 0) | planes
    if records:
        h.records, h.rec_w = rec.data_ptr(), rec_w
        h.tc_image = None if image is None else image.data_ptr()
    return h


for mb in sizes:
    idx = rs.randint(check_rows, size=mb).astype(np.int32)
    hi = up(idx)
    lp, outs = O.policy_logprob(theta, spec, o[idx], a[idx])
    wgt = O.ppo_clip_weight(lp, lp_old[idx], adv[idx])
    want = O.policy_logprob_grad(theta, spec, o[idx], a[idx], wgt, outs)
    scale = float(np.max(np.abs(want)))
    only = os.environ.get("PROBE_ONLY")
    for name, planes, records in (("array-fed exact", 3, False), ("bf16x3 (exact)", 3, True), ("fp16x2", 0x82, True),
                                  ("bf16x2", 2, True), ("bf16", 1, True), ("fp16x2 static", 0xC2, True), ("bf16 static", 0x41, True)):
        if only and name not in only.split(","):
            continue
        image = None
        if records:
            image = torch.zeros(lib.mq_tc_image_bytes(net, planes), dtype=torch.uint8, device=dev)
            L.check(lib.mq_tc_image_build(net, ht.data_ptr(), planes, image.data_ptr(), stream))
        h = head(planes, records, image)
        g = torch.zeros(5126, device=dev)
        L.check(lib.mq_fused_grad(net, ht.data_ptr(), ho.data_ptr(), hi.data_ptr(), mb, ctypes.byref(h), 1.0 / mb, g.data_ptr(),
                                  None, None, ws.data_ptr(), wsb, stream))
        torch.cuda.synchronize()
        err = float(np.max(np.abs(g.cpu().numpy() - want))) / scale
        n_parts = ctypes.c_int32(0)
        ts = []
        for it in range(12):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(lib.mq_fused_grad_partials(net, ht.data_ptr(), ho.data_ptr(), hi.data_ptr(), mb, ctypes.byref(h), None, None,
                                               ws.data_ptr(), wsb, ctypes.byref(n_parts), stream))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts = sorted(ts[2:])
        reps = 50                                      # event resolution is ~2 us here: also time a back-to-back train
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(reps):
            L.check(lib.mq_fused_grad_partials(net, ht.data_ptr(), ho.data_ptr(), hi.data_ptr(), mb, ctypes.byref(h), None, None,
                                               ws.data_ptr(), wsb, ctypes.byref(n_parts), stream))
        e1.record()
        torch.cuda.synchronize()
        train = e0.elapsed_time(e1) * 1e3 / reps
        print("mb=%6d %-16s max|err|/max|grad| = %.2e   kernel %.1f us (median of 10, L2 flushed), %.2f us (50 back to back), %d partials"
              % (mb, name, err, ts[len(ts) // 2], train, n_parts.value), flush=True)
