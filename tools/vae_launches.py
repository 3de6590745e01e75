"""Developer aid: run eager VAE steps (BASELINE configs[4] shape) so that ncu can list the launches of one step.
    python tools/vae_launches.py [fused|graph]      (default fused: the plan of csrc/mq_vae.cu)"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from mannequin2_b200 import _device as D
from mannequin2_b200.vae import VAETrainer

np.random.seed(0)
tr = VAETrainer(image_shape=(28, 28), hidden=400, latent=20, n_hidden=1, rng=D.DeviceRNG(2))
x = D.from_host(np.random.RandomState(1).rand(4096, 28, 28).astype(np.float32))
fused = not (len(sys.argv) > 1 and sys.argv[1] == "graph")
for _ in range(3):
    (tr.step_fused if fused else tr.step_dev)(x)
torch.cuda.synchronize()
