"""Developer aid: per-phase cycle breakdown of the persistent PPO chain kernel (policy chain alone)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mannequin2_b200 import _lib as L, _device as D
from mannequin2_b200.ppo import PPOUpdater
from mannequin2_b200._lib import HEAD_GAUSS_PPO, MqHead
from bench import ppo_rollout
rs = np.random.RandomState(1)
o, a, r, d = ppo_rollout(rs, 2048)
vi = rs.randint(2048, size=(300, 64)).astype(np.int32); pi = rs.randint(2048, size=(300, 64)).astype(np.int32)
np.random.seed(0)
names = {0: "stage+bar1", 2: "head", 5: "adam", 4: "transposes", 6: "bar2", 7: "bar3", 9: "dW0", 12: "dW1", 13: "dX1", 15: "dW2", 16: "dX2", 20: "fwd0", 21: "fwd1", 22: "fwd2"}
for sd in ("float32",):
    upd = PPOUpdater(11, 3, state_dtype=sd)
    dev = [D.from_host(x) for x in (o, a, r, d, vi, pi)]
    upd.update_dev(*dev); torch.cuda.synchronize()
    head = MqHead(); head.kind, head.p0, head.p1 = HEAD_GAUSS_PPO, 0.8, 1.2
    head.target, head.adv, head.logp_old = D.ptr(dev[1]), D.ptr(upd._bufs["adv_n"]), D.ptr(upd._bufs["logp0"])
    prof = torch.zeros(32, dtype=torch.int64, device="cuda")
    L.lib().mq_debug_set_phase_buffer(prof.data_ptr())
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); upd.policy.train(dev[0], head, dev[5], 300, 64, 3e-4); t1.record(); torch.cuda.synchronize()
    L.lib().mq_debug_set_phase_buffer(None)
    p = prof.cpu().numpy()
    print(sd, "chain ms %.3f" % t0.elapsed_time(t1), "total cyc/step %.0f" % (p.sum() / 300))
    print("   ", {names.get(i, i): int(p[i] / 300) for i in range(32) if p[i]})
