#!/usr/bin/env python3
"""Benchmark of the B200-native mannequin2 hot path (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ppo|ppo_large|mlp|es]
    python bench.py --impl reference ...        # CPU arm: the oracle port on the host cores

Prints ONE JSON line.  `value` is device-timed throughput with inputs resident in
HBM; `e2e` is the same metric through the public host-array API with the
host<->device copies inside the timed region.  `oracle/` is imported only by the
`cpu_baseline` leg and by `--impl reference`.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L2_FLUSH_BYTES = 256 << 20


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops"]),
                    bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------ synthetic data
def ppo_rollout(rs, n, n_obs=11, n_act=3, p_done=1.0 / 200, cut=None):
    """BASELINE.md 3.4 C2/C3 generators."""
    obs = np.clip(rs.randn(n + 1, n_obs), -5, 5).astype(np.float32)
    act = rs.randn(n, n_act).astype(np.float32)
    rew = rs.randn(n).astype(np.float32)
    done = rs.rand(n) < p_done
    if cut:
        done[cut - 1::cut] = True
    return obs, act, rew, done.astype(np.uint8)


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    samples=len(sm), reasons=sorted(reasons))


# ------------------------------------------------------------------ PPO workload (configs[1])
class PPOWorkload(object):
    """One step = one PPO update on a 2048-step rollout: value forward (2049), GAE
    (gam .99, lam .95), 300 x 64 value-net Adam steps, advantage normalise + tanh, old
    log-probs (2048), 300 x 64 policy Adam steps with the ratio/clip rule -- i.e. one
    iteration of benchmarks/ppo.py:22-40 incl. benchmarks/gae.py:34-75 (script-faithful
    300 with-replacement minibatches; BASELINE.json words it as 10 epochs x 64)."""
    name = "ppo_update_2048x(300+300)x64"
    metric = "ppo_update_transitions_per_s"
    unit = "transitions/s"
    N, STEPS, MB = 2048, 300, 64

    def __init__(self, n_rollouts=8, seed=1):
        rs = np.random.RandomState(seed)
        self.rollouts = []
        for _ in range(n_rollouts):
            o, a, r, d = ppo_rollout(rs, self.N)
            vi = rs.randint(self.N, size=(self.STEPS, self.MB)).astype(np.int32)
            pi = rs.randint(self.N, size=(self.STEPS, self.MB)).astype(np.int32)
            self.rollouts.append((o, a, r, d, vi, pi))
        self.units_per_step = self.N

    # ---- GPU arm
    def setup_gpu(self):
        import torch
        from mannequin2_b200 import _device as D
        from mannequin2_b200.ppo import PPOUpdater
        np.random.seed(0)
        self.torch, self.D = torch, D
        self.upd = PPOUpdater(11, 3)
        self.dev = [tuple(D.from_host(x) for x in r) for r in self.rollouts]
        self.pinned = [tuple(torch.from_numpy(x).pin_memory() for x in r) for r in self.rollouts]
        self.out_pinned = (torch.empty(5126, dtype=torch.float32).pin_memory(),
                           torch.empty(4993, dtype=torch.float32).pin_memory())
        self.h2d = sum(x.nbytes for x in self.rollouts[0])
        self.d2h = (5126 + 4993) * 4

    def step_gpu(self, i):
        self.upd.update_dev(*self.dev[i % len(self.dev)])

    def step_e2e(self, i):
        t = self.torch
        dev = self.D.device()
        args = [x.to(dev, non_blocking=True) for x in self.pinned[i % len(self.pinned)]]
        self.upd.update_dev(*args)
        self.out_pinned[0].copy_(self.upd.policy.theta, non_blocking=True)
        self.out_pinned[1].copy_(self.upd.value.theta, non_blocking=True)
        t.cuda.current_stream().synchronize()

    def launches_per_step(self):
        # value fwd, gae, value train, col_mean+finish, mean update, cov+finish, var update, apply,
        # logp fwd, policy train
        return 13

    def dominant_kernel(self, peaks):
        """fused_train_kernel<double> (policy chain), timed alone on its launch stream."""
        t, D, upd = self.torch, self.D, self.upd
        from mannequin2_b200._lib import HEAD_GAUSS_PPO, MqHead
        obs, act, rew, done, vi, pi = self.dev[0]
        adv_n, logp0 = upd._bufs["adv_n"], upd._bufs["logp0"]
        head = MqHead()
        head.kind, head.p0, head.p1 = HEAD_GAUSS_PPO, 0.8, 1.2
        head.target, head.adv, head.logp_old = D.ptr(act), D.ptr(adv_n), D.ptr(logp0)
        times = []
        for _ in range(5):
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            e0.record()
            upd.policy.train(obs, head, pi, self.STEPS, self.MB, 3e-4)
            e1.record()
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times[1:]))
        p = 5126
        # algorithmic bytes per launch: per step 64 gathered rows x 64 B (obs 44 + act 12 + adv 4 +
        # logp_old 4) + Adam fp32 state 24 B/param (resident on chip: algorithmic, not HBM traffic)
        bytes_per_launch = self.STEPS * (self.MB * 64 + p * 24)
        flops = self.STEPS * self.MB * 28544
        ach = bytes_per_launch / (ms * 1e-3) / 1e9
        return dict(bound="hbm", kernel="chain_train_kernel", achieved=ach, peak=peaks["hbm"],
                    unit="GB/s", frac=ach / peaks["hbm"], traffic=None, peak_source=peaks["source"],
                    ms_per_launch=ms, fp32_gflops=flops / (ms * 1e-3) / 1e9,
                    note="latency-bound: 300 dependent 64-row steps in ONE CTA; working set is on chip")

    # ---- CPU arm (oracle port of the reference)
    def setup_cpu(self):
        from oracle import mq_oracle as O
        self.O = O
        rs = np.random.RandomState(0)
        self.pol_spec = O.policy_spec(11, 3)
        self.val_spec = O.mlp_spec(11, [64, 64, 1], [O.ACT_TANH, O.ACT_TANH, O.ACT_NONE])
        self.pol_theta = np.concatenate([O.mlp_init(O._mlp_part(self.pol_spec), rs, 0.1), np.zeros(3, np.float32)])
        self.val_theta = O.mlp_init(self.val_spec, rs, 0.1)
        self.pol_opt = O.AdamO(self.pol_theta, horizon=10)
        self.val_opt = O.AdamO(self.val_theta, horizon=10, lr=0.003)
        self.norm = O.RunningNormalizeO(horizon=2)

    def step_cpu(self, i):
        o, a, r, d, vi, pi = self.rollouts[i % len(self.rollouts)]
        self.pol_theta, self.val_theta, _ = self.O.ppo_update(
            self.pol_theta, self.pol_spec, self.pol_opt, self.val_theta, self.val_spec, self.val_opt,
            self.norm, o, a, r, d.astype(bool), vi, pi)


WORKLOADS = {"ppo": PPOWorkload}


# ------------------------------------------------------------------ CPU timing helpers
def _cpu_worker(args):
    name, steps, warmup, seed = args
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:
        limiter = None
    wl = WORKLOADS[name](seed=seed + 1)
    wl.setup_cpu()
    for i in range(warmup):
        wl.step_cpu(i)
    t0 = time.perf_counter()
    for i in range(steps):
        wl.step_cpu(warmup + i)
    return time.perf_counter() - t0


def time_cpu(name, steps, warmup, procs):
    """`procs` independent single-thread replicas in parallel (how the reference uses a multi-core
    box: run_benchmarks.sh:98-102, one process per seed); aggregate units/s."""
    import multiprocessing as mp
    wl = WORKLOADS[name]()
    if procs <= 1:
        dt = _cpu_worker((name, steps, warmup, 0))
        return steps * wl.units_per_step / dt, dt
    with mp.get_context("fork").Pool(procs) as pool:
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(name, steps, warmup, s) for s in range(procs)])
        wall = time.perf_counter() - t0
    # wall includes per-process set-up and warm-up: conservative for the CPU (lower bound on its speed)
    dts = None
    return procs * steps * wl.units_per_step / wall, wall


# ------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ppo", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(3, args.warmup)
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        wl = WORKLOADS[args.workload]()
        procs = max(1, min(cores, 32))
        steps = max(1, min(args.steps, 4))
        val, wall = time_cpu(args.workload, steps, 1, procs)
        line = dict(metric=wl.metric, value=val, unit=wl.unit, n_gpus=args.gpus, steps=steps, warmup=1,
                    ms_per_step=1e3 * wall / steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f64", data="synthetic", impl="reference",
                    config=dict(workload=wl.name, note="oracle port of the NumPy reference; %d independent "
                                "single-thread replicas, aggregate" % procs),
                    cpu_baseline=dict(value=val, unit=wl.unit, cores=procs, kind="port",
                                      sample="%d steps x %d replicas" % (steps, procs)),
                    e2e=dict(value=val, unit=wl.unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    else:
        torch.cuda.set_device(0)
    peaks = load_peaks()
    wl = WORKLOADS[args.workload](seed=1 + rank)
    wl.setup_gpu()
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, n, use_events=True):
        """Each step is timed on the device with its own event pair; L2 is flushed
        (256 MB write) between steps, outside the timed intervals."""
        total = 0.0
        for i in range(n):
            flush.fill_(float(i))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step_fn(i)
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
        return total

    for i in range(warmup):
        wl.step_gpu(i)
    barrier()
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    barrier()
    ms = timed(lambda i: wl.step_gpu(warmup + i), args.steps)
    barrier()
    for i in range(2):
        wl.step_e2e(i)
    barrier()
    ms_e2e = timed(lambda i: wl.step_e2e(warmup + i), args.steps)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    units = wl.units_per_step * args.steps * world
    if rank == 0:
        roof = wl.dominant_kernel(peaks)
        cpu = None
        if not args.no_cpu_baseline:
            v1, w1 = time_cpu(args.workload, 2, 1, 1)
            cpu = dict(value=v1, unit=wl.unit, cores=1, kind="port",
                       sample="2 steps of the same workload after 1 warm-up, one process (OpenBLAS default threads)",
                       host_cores=cores)
        line = dict(metric=wl.metric, value=units / (ms * 1e-3), unit=wl.unit, n_gpus=world, steps=args.steps,
                    warmup=warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload=wl.name, l2="flushed between steps (256 MB write)",
                                adam_state="f32 (moments on chip)", parallelism="replicas" if world > 1 else "single"),
                    e2e=dict(value=units / (ms_e2e * 1e-3), unit=wl.unit, h2d_bytes_per_step=wl.h2d,
                             d2h_bytes_per_step=wl.d2h, ms_per_step=ms_e2e / args.steps),
                    gpu_launches=wl.launches_per_step() * args.steps, roofline=roof, cpu_baseline=cpu,
                    clocks=clocks)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
