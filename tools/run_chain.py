"""Launch the policy chain kernel a few times (target for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mannequin2_b200 import _device as D
from mannequin2_b200.ppo import PPOUpdater
from mannequin2_b200._lib import HEAD_GAUSS_PPO, MqHead
from bench import ppo_rollout
sd = sys.argv[1] if len(sys.argv) > 1 else "float32"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
rs = np.random.RandomState(1)
o, a, r, d = ppo_rollout(rs, 2048)
vi = rs.randint(2048, size=(300, 64)).astype(np.int32); pi = rs.randint(2048, size=(300, 64)).astype(np.int32)
np.random.seed(0)
upd = PPOUpdater(11, 3, state_dtype=sd)
dev = [D.from_host(x) for x in (o, a, r, d, vi, pi)]
upd.update_dev(*dev); torch.cuda.synchronize()
head = MqHead(); head.kind, head.p0, head.p1 = HEAD_GAUSS_PPO, 0.8, 1.2
head.target, head.adv, head.logp_old = D.ptr(dev[1]), D.ptr(upd._bufs["adv_n"]), D.ptr(upd._bufs["logp0"])
for _ in range(3):
    upd.policy.train(dev[0], head, dev[5], steps, 64, 3e-4)
torch.cuda.synchronize()
print("done")
