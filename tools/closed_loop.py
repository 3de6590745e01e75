"""Developer aid: closed PPO loop on the synthetic on-device environment (rollout.py + ppo.py): mean reward per iteration
and the time split between collecting and updating."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mannequin2_b200 import _device as D
from mannequin2_b200.ppo import PPOUpdater
from mannequin2_b200.rollout import BatchedActor, SyntheticEnv

n_env = int(os.environ.get("N_ENV", 512))
t_len = int(os.environ.get("T_LEN", 64))
iters = int(os.environ.get("ITERS", 20))
mb = int(os.environ.get("MB", 8192))
epochs = int(os.environ.get("EPOCHS", 4))
graph = bool(int(os.environ.get("GRAPH", "1")))
np.random.seed(0)
rs = np.random.RandomState(1)
env = SyntheticEnv(n_env, horizon=int(os.environ.get("HORIZON", 64)), seed=3)
upd = PPOUpdater(env.n_obs, env.n_act, pol_lr=float(os.environ.get("LR", 3e-4)))
actor = BatchedActor(env, upd.policy, seed=5, use_graph=graph)
n = n_env * t_len
for it in range(iters):
    torch.cuda.synchronize(); t0 = time.time()
    chunk = actor.collect(t_len)
    torch.cuda.synchronize(); t1 = time.time()
    vi = D.from_host(np.concatenate([rs.permutation(n).astype(np.int32).reshape(n // mb, mb) for _ in range(epochs)]))
    pi = D.from_host(np.concatenate([rs.permutation(n).astype(np.int32).reshape(n // mb, mb) for _ in range(epochs)]))
    torch.cuda.synchronize(); t2 = time.time()
    upd.update_rollout(chunk, n_env, t_len, vi, pi)
    torch.cuda.synchronize(); t3 = time.time()
    print("iter %2d  mean reward %8.4f  logstd %s  collect %.2f ms  update %.2f ms" % (
        it, float(chunk["rew"].mean()), np.round(D.to_host(upd.policy.theta[-3:]), 3), (t1 - t0) * 1e3, (t3 - t2) * 1e3), flush=True)
