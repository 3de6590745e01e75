"""Developer aid: the fused VAE plan (csrc/mq_vae.cu) against the fp64 oracle, error per parameter block, for the
reference-run fixture (golden/vae_step.npz) and the BASELINE configs[4] shape; then timings of the fused and the
node-by-node step (CUDA-graph replays)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from mannequin2_b200 import _device as D
from mannequin2_b200.vae import VAETrainer
from oracle import mq_oracle as O


class Fixed(object):
    def __init__(self, unit):
        self.unit = D.from_host(np.asarray(unit, dtype=np.float32))

    def randn_dev(self, shape):
        return self.unit.clone()


def blocks(d_img, hidden, latent, n_hidden):
    enc, dec, off, n = O.vae_layout(d_img, hidden, latent, n_hidden)
    out = []
    for i, l in enumerate(enc["layers"]):
        out += [("enc%d.W" % i, l["w_off"], l["b_off"]), ("enc%d.b" % i, l["b_off"], l["b_off"] + l["n_out"])]
    for nm in ("e_mean", "e_logstd"):
        w0, b0, no = off[nm]
        out += [(nm + ".W", w0, b0), (nm + ".b", b0, b0 + no)]
    for i, l in enumerate(dec["layers"]):
        out += [("dec%d.W" % i, off["dec"] + l["w_off"], off["dec"] + l["b_off"]),
                ("dec%d.b" % i, off["dec"] + l["b_off"], off["dec"] + l["b_off"] + l["n_out"])]
    for nm in ("d_mean", "d_logstd"):
        w0, b0, no = off[nm]
        out += [(nm + ".W", w0, b0), (nm + ".b", b0, b0 + no)]
    return out


def check(tag, tr, theta, x_host, unit, d_img, hidden, latent, n_hidden):
    tr.decoder.load_params(theta)
    x = D.from_host(x_host.reshape(len(x_host), d_img).astype(np.float32))
    grad, lp, kl = tr.grad_fused(x, tr._rng.unit)
    torch.cuda.synchronize()
    grad, lp, kl = D.to_host(grad).astype(np.float64), D.to_host(lp), D.to_host(kl)
    o_lp, o_kl, g1, g2 = O.vae_grads(theta, d_img, hidden, latent, n_hidden, x_host.reshape(len(x_host), d_img), unit)
    want = g1.copy()
    want[:len(g2)] += g2
    print("== %s  logprob err %.3g (max %.3g)  dkl err %.3g (max %.3g)" % (tag, np.max(np.abs(lp - o_lp)), np.max(np.abs(o_lp)),
                                                                         np.max(np.abs(kl - o_kl)), np.max(np.abs(o_kl))))
    for name, a, b in blocks(d_img, hidden, latent, n_hidden):
        e = np.max(np.abs(grad[a:b] - want[a:b]))
        m = np.max(np.abs(want[a:b]))
        print("   %-12s max|g| %10.4g  err %10.3g  rel %8.2g %s" % (name, m, e, e / max(m, 1e-30), "" if e <= 2e-5 * max(m, 1e-30) + 1e-5 * np.max(np.abs(want)) else "  <-- BAD"))
    print("   finite:", np.isfinite(grad).all())


def main():
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "vae_step.npz"))
    d_img, hidden, latent, n_hidden, batch = (int(v) for v in g["vstep_dims"])
    for mode in ("fp16x2", "exact", "bf16x2"):
        tr = VAETrainer(image_shape=(4, 3), hidden=hidden, latent=latent, n_hidden=n_hidden, rng=Fixed(g["vstep_unit"]), gemm_mode=mode)
        check("fixture " + mode, tr, g["vstep_theta"], g["vstep_x"], g["vstep_unit"], d_img, hidden, latent, n_hidden)
    rs = np.random.RandomState(5)
    B, DI, H, L = 4096, 784, 400, 20
    x_host = rs.rand(B, 28, 28).astype(np.float32)
    unit = rs.randn(B, L).astype(np.float32)
    np.random.seed(11)
    tr = VAETrainer(image_shape=(28, 28), hidden=H, latent=L, n_hidden=1, rng=Fixed(unit))
    theta = np.asarray(tr.decoder.get_params(), dtype=np.float32).copy()
    theta += (np.random.RandomState(12).randn(len(theta)) * 0.02).astype(np.float32)
    check("configs[4] fp16x2", tr, theta, x_host, unit, DI, H, L, 1)
    # timings
    x = D.from_host(x_host)
    for fused in (True, False):
        np.random.seed(11)
        t2 = VAETrainer(image_shape=(28, 28), hidden=H, latent=L, n_hidden=1, rng=D.DeviceRNG(3))
        for _ in range(3):
            t2.step_graph(x, fused=fused)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            t2.step_graph(x, fused=fused)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("step_graph fused=%s: %.3f ms per step = %.2f M samples/s" % (fused, ms, B / ms / 1e3))
        if fused:
            t0 = time.time()
            e0.record()
            for _ in range(20):
                t2.step_fused(x)
            e1.record()
            torch.cuda.synchronize()
            print("step_fused eager: %.3f ms per step (host %.3f ms)" % (e0.elapsed_time(e1) / 20, (time.time() - t0) * 50))


if __name__ == "__main__":
    main()
