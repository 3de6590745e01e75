"""Developer aid: time fused_grad on a 65,536-row PPO minibatch for several tile sizes."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mannequin2_b200 import _lib as L, _device as D
from mannequin2_b200.ppo import PPOUpdater
from bench import ppo_rollout
rs = np.random.RandomState(1)
N, MB = 1 << 20, 65536
o, a, r, d = ppo_rollout(rs, N, p_done=1 / 500., cut=1000)
idx = rs.permutation(N).astype(np.int32).reshape(16, MB)
np.random.seed(0)
upd = PPOUpdater(11, 3)
od, ad, idxd = D.from_host(o), D.from_host(a), D.from_host(idx)
adv, lp = torch.randn(N, device="cuda").tanh(), torch.randn(N, device="cuda") * 0.1 - 3
head = L.MqHead(); head.kind, head.p0, head.p1 = L.HEAD_GAUSS_PPO, 0.8, 1.2
head.target, head.adv, head.logp_old = D.ptr(ad), D.ptr(adv), D.ptr(lp)
net = upd.policy.net
lib = L.lib()
wsb = lib.mq_fused_ws_bytes(C.byref(net), MB); ws = D.workspace("t", wsb)
grads = {}
for tb in [int(x) for x in (sys.argv[1:] or ["-1", "-2", "16", "128"])]:
    lib.mq_debug_set_tc_mode(2 if tb < 0 else 0)      # -1: tcgen05 kernel (one tile per SM + issuer warp), -2: two-tiles-in-flight variant
    lib.mq_debug_set_tc_variant(0 if tb == -2 else 1)
    lib.mq_debug_use_chain_grad(0 if tb == 128 else 1)
    lib.mq_debug_set_chain_warps(0 if tb == 128 or tb < 0 else tb)
    lib.mq_debug_set_tile_rows(128 if tb == 128 else 0)
    ts = []
    for k in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        D.call("mq_fused_grad", C.byref(net), D.ptr(upd.policy.theta), D.ptr(od), D.ptr(idxd) + (k % 16) * MB * 4, MB,
               C.byref(head), 1.0 / MB, D.ptr(upd.policy.grad), None, None, D.ptr(ws), wsb, D.stream())
        e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts[3:]))
    grads[tb] = upd.policy.grad.cpu().numpy().copy()
    print("tile_rows=%3d  %.1f us  %.2f TFLOP/s fp32 (%.1f%% of 74.4)" % (tb, ms * 1e3, MB * 28544 / ms / 1e9, MB * 28544 / ms / 1e9 / 74.4 * 100))
lib.mq_debug_set_tile_rows(0); lib.mq_debug_use_chain_grad(1); lib.mq_debug_set_chain_warps(0); lib.mq_debug_set_tc_mode(1); lib.mq_debug_set_tc_variant(1)
k0 = list(grads)[0]
for k in grads: print("max |grad(tb=%d) - grad(tb=%d)| = %.2e" % (k, k0, np.abs(grads[k] - grads[k0]).max()))
