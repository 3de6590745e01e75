"""Developer aid: per-phase cycle breakdown of the tcgen05 gradient kernel (CTA 0) on a 65,536-row PPO minibatch."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mannequin2_b200 import _lib as L, _device as D
from mannequin2_b200.ppo import PPOUpdater
from bench import ppo_rollout
rs = np.random.RandomState(1)
N, MB = 1 << 20, int(sys.argv[1]) if len(sys.argv) > 1 else 65536
o, a, r, d = ppo_rollout(rs, N, p_done=1 / 500., cut=1000)
idx = rs.permutation(N).astype(np.int32)[:MB]
np.random.seed(0)
upd = PPOUpdater(11, 3)
od, ad, idxd = D.from_host(o), D.from_host(a), D.from_host(idx)
adv, lp = torch.randn(N, device="cuda").tanh(), torch.randn(N, device="cuda") * 0.1 - 3
head = L.MqHead(); head.kind, head.p0, head.p1 = L.HEAD_GAUSS_PPO, 0.8, 1.2
head.target, head.adv, head.logp_old = D.ptr(ad), D.ptr(adv), D.ptr(lp)
net = upd.policy.net
lib = L.lib()
wsb = lib.mq_fused_ws_bytes(C.byref(net), MB); ws = D.workspace("t", wsb)
lib.mq_debug_set_tc_mode(2)
lib.mq_debug_set_tc_variant(int(os.environ.get("TC_VARIANT", "1")))
def run():
    D.call("mq_fused_grad", C.byref(net), D.ptr(upd.policy.theta), D.ptr(od), D.ptr(idxd), MB,
           C.byref(head), 1.0 / MB, D.ptr(upd.policy.grad), None, None, D.ptr(ws), wsb, D.stream())
for _ in range(3): run()
torch.cuda.synchronize()
prof = torch.zeros(512, dtype=torch.int64, device="cuda")
lib.mq_debug_set_phase_buffer(prof.data_ptr())
run(); torch.cuda.synchronize()
lib.mq_debug_set_phase_buffer(None)
names = {0: "setup", 1: "wait_prev_tile", 2: "stage_x", 4: "prefetch", 5: "wait_fwd0", 6: "wait_fwd1", 7: "wait_fwd2",
         8: "epi0", 9: "epi1", 11: "head", 12: "bwd_gap", 15: "wait_bwd2", 14: "wait_bwd1", 18: "epi_b2", 17: "epi_b1", 19: "wait_last", 20: "final", 21: "setup:alloc", 22: "setup:zero", 23: "setup:ones", 24: "setup:images", 25: "setup:coords"}
p = prof.cpu().numpy()
variant = int(os.environ.get("TC_VARIANT", "1"))      # 1: one-tile kernel (default), 0: two-tile kernel (groups 0 and 1 of CTA 0)
for grp in range(2 if variant == 0 else 1):
    q = p[grp * 32:grp * 32 + 32]
    print("total cycles CTA0 group %d: %d (%.1f us at 1.965 GHz)" % (grp, q.sum(), q.sum() / 1965.0))
    for i in range(32):
        if q[i]: print("  %-16s %8d" % (names.get(i, i), q[i]))
g = p[64:64 + 2 * 148].reshape(148, 2)
t0 = g[:, 0].min()
print("CTA start offsets (us): min %.1f median %.1f max %.1f;  CTA end offsets: min %.1f median %.1f max %.1f;  durations: min %.1f median %.1f max %.1f" % (
    0, np.median(g[:, 0] - t0) / 1e3, (g[:, 0] - t0).max() / 1e3, (g[:, 1] - t0).min() / 1e3, np.median(g[:, 1] - t0) / 1e3, (g[:, 1] - t0).max() / 1e3,
    (g[:, 1] - g[:, 0]).min() / 1e3, np.median(g[:, 1] - g[:, 0]) / 1e3, (g[:, 1] - g[:, 0]).max() / 1e3))
ts = []
for k in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
print("launch (incl. reduce) median %.1f us" % (np.median(ts) * 1e3))
lib.mq_debug_set_tc_mode(1); lib.mq_debug_set_tc_variant(1)
