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


_MEASURED = {}


def load_measured_peaks():
    """fp32 FFMA and tensor-pipe (bf16, TF32) peaks measured on THIS GPU by the library's own microbenchmarks
    ("to be measured by the harness"); cached for the process."""
    if _MEASURED:
        return _MEASURED
    import ctypes as C
    import torch
    from mannequin2_b200 import _lib as L
    lib = L.lib()
    stream = torch.cuda.current_stream().cuda_stream
    scratch = torch.zeros(16, device="cuda")
    fl = C.c_double(0.0)

    def run(fn):
        best = None
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        return fl.value / (best * 1e-3) / 1e12
    try:
        _MEASURED["ffma_tflops"] = run(lambda: L.check(lib.mq_bench_ffma(20000, scratch.data_ptr(), C.byref(fl), stream)))
        _MEASURED["bf16_mma_issue_tflops"] = run(lambda: L.check(lib.mq_bench_tensor(0, 40000, C.byref(fl), stream)))
        _MEASURED["tf32_tflops"] = run(lambda: L.check(lib.mq_bench_tensor(1, 40000, C.byref(fl), stream)))
        _MEASURED["how"] = ("best of 5, CUDA events: mq_bench_ffma (8 independent FFMA chains per thread, 8 CTAs of 256 per SM); "
                            "mq_bench_tensor (back-to-back tcgen05.mma M128 N256 from resident shared-memory operands, one "
                            "CTA per SM: the tensor pipe's issue peak for kind::f16 / kind::tf32)")
    except Exception as exc:                     # pragma: no cover
        _MEASURED["error"] = repr(exc)
    return _MEASURED


def host_info():
    """CPU model and NumPy / BLAS versions of the box the CPU arm runs on."""
    model = None
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    blas = None
    try:
        cfg = np.show_config(mode="dicts")
        b = cfg.get("Build Dependencies", {}).get("blas", {})
        blas = "%s %s" % (b.get("name"), b.get("version"))
    except Exception:
        pass
    return dict(cpu_model=model, logical_cpus=os.cpu_count(), numpy=np.__version__, blas=blas)


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

    def replica_throughput(self, n_replicas, steps=5, warmup=3):
        """Aggregate transitions/s of n_replicas INDEPENDENT learners (own parameters, normaliser and
        Adam state) driven concurrently on their own streams of ONE GPU -- the counterpart of the
        reference arm's one-process-per-core replicas.  One learner's update is a chain of dependent
        single-CTA launches (latency-bound), so replicas fill the other SMs."""
        t = self.torch
        from mannequin2_b200.ppo import PPOUpdater
        upds = [PPOUpdater(11, 3) for _ in range(n_replicas)]
        streams = [t.cuda.Stream() for _ in range(n_replicas)]
        t.cuda.synchronize()

        def one_round(i):
            for r, (u, s) in enumerate(zip(upds, streams)):
                with t.cuda.stream(s):
                    u.update_dev(*self.dev[(i + r) % len(self.dev)])
        for i in range(warmup):
            one_round(i)
        t.cuda.synchronize()
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams:
            s.wait_event(e0)
        for i in range(steps):
            one_round(i)
        for s in streams:
            t.cuda.current_stream().wait_stream(s)
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        return dict(replicas=n_replicas, value=n_replicas * steps * self.N / (ms * 1e-3), unit=self.unit,
                    ms_per_round=ms / steps)

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
            self.norm, o, a, r, d.astype(bool), vi, pi, dtype=np.float32)


class PPOLargeWorkload(object):
    """BASELINE.json configs[2]: large-batch PPO.  ONE rollout of 2^20 synthetic transitions in variable-length episodes
    (done p=1/500, forced cut every 1000 steps); one update = value forward (N+1), GAE, advantage normalise+tanh, old
    log-probs (N), packed records, then 10 epochs x 16 permutation minibatches of 65,536 rows for the value net AND for
    the policy (160 + 160 Adam steps).
    With G GPUs (the default N > 1 form, SURVEY 8d C3 / 8e: STRONG scaling) every rank holds the same rollout and the same
    tables of global minibatches and takes 65,536 / G rows of each (8,192 per GPU at 8 GPUs); the flat gradient is
    exchanged once per optimiser step inside the reduce + Adam kernel; the two forward passes over the rollout are split
    by rows and all-gathered.  `weak=True` is round 1's form (every rank its own 2^20-transition shard, global minibatch
    G x 65,536), reported beside it."""
    name = "ppo_large_1Mx10epochsx65536 (configs[2])"
    metric = "ppo_update_transitions_per_s"
    unit = "transitions/s"
    N, EPOCHS, MB = 1 << 20, 10, 65536

    def __init__(self, n_rollouts=2, seed=1, n=None, epochs=None, weak=False, tc_mode="exact"):
        self.N = n or self.N
        self.EPOCHS = epochs or self.EPOCHS
        self.MB = min(self.MB, self.N)
        self.weak, self.tc_mode = weak, tc_mode
        rs = np.random.RandomState(seed)
        self.rollouts = []
        per = self.N // self.MB
        for _ in range(n_rollouts):
            o, a, r, d = ppo_rollout(rs, self.N, p_done=1.0 / 500, cut=1000)
            tabs = []
            for _t in range(2):
                t = np.concatenate([rs.permutation(self.N).astype(np.int32).reshape(per, self.MB)
                                    for _e in range(self.EPOCHS)])
                tabs.append(t)
            self.rollouts.append((o, a, r, d, tabs[0], tabs[1]))
        self.units_per_step = self.N
        self.steps_per_net = self.EPOCHS * per

    def setup_gpu(self):
        import torch
        import torch.distributed as dist
        from mannequin2_b200 import _device as D
        from mannequin2_b200.ppo import PPOUpdater
        np.random.seed(0)
        self.torch, self.D = torch, D
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.replicated = self.world > 1 and not self.weak
        self.upd = PPOUpdater(11, 3, tc_mode=self.tc_mode)
        self.upd.check_indices = False               # the tables are permutations built above
        self.dev = [tuple(D.from_host(x) for x in r) for r in self.rollouts]
        # end to end: the ROLLOUT (obs, act, rew, done: 64 MB) travels every step from pinned host memory; the minibatch
        # index tables (84 MB of int32 permutations, draws of the training loop, not rollout data) are uploaded once per
        # run and stay resident -- round 1 shipped them with every step
        self.pinned = [tuple(torch.from_numpy(x).pin_memory() for x in r[:4]) for r in self.rollouts[:1]]
        self.out_pinned = (torch.empty(5126, dtype=torch.float32).pin_memory(),
                           torch.empty(4993, dtype=torch.float32).pin_memory())
        self.h2d = sum(x.nbytes for x in self.rollouts[0][:4]) // (self.world if self.replicated else 1)
        self.idx_bytes_once = sum(x.nbytes for x in self.rollouts[0][4:])
        self.d2h = (5126 + 4993) * 4
        self._next = None

    def step_gpu(self, i):
        self.upd.update_dev(*self.dev[i % len(self.dev)], replicated=self.replicated)

    def step_e2e(self, i):
        """The public host-side call PPOUpdater.update: pinned host rollout in, new parameters out (pinned), every step.
        prefetch() starts the upload of step i+1 on a copy stream before step i's kernels are launched, the way a training
        loop feeds rollouts; every step still copies its own inputs H2D and its results D2H inside the timed region."""
        if self._next is None:
            self._next = self.upd.prefetch(*self.pinned[0], replicated=self.replicated)
        cur, self._next = self._next, None
        vi, pi = self.dev[0][4], self.dev[0][5]
        nxt = self.upd.prefetch(*self.pinned[0], replicated=self.replicated)
        self.upd.update(cur, val_idx=vi, pol_idx=pi, replicated=self.replicated, out=self.out_pinned)
        self._next = nxt

    def launches_per_step(self):
        # 2 forwards (record-fed: + pack + image build each), GAE (memset + scan), normaliser (4), 2 x pack, 2 x image build,
        # then per optimiser step the gradient kernel + the reduce / exchange / Adam / image kernel
        return 17 + 2 * self.steps_per_net * 2

    def grad_kernel_time(self, net_obj, head, idx_ptr, mb, reps=12):
        """One gradient launch (record-fed tcgen05 kernel -> per-CTA partials), timed alone with CUDA events on its
        stream, L2 flushed before every launch; median of reps - 2."""
        import ctypes as C
        t, D = self.torch, self.D
        from mannequin2_b200 import _lib as L
        wsb = L.lib().mq_fused_ws_bytes(C.byref(net_obj.net), mb)
        ws = D.workspace("bench", wsb)
        flush = t.empty(L2_FLUSH_BYTES // 4, dtype=t.float32, device="cuda")
        n_parts = C.c_int32(0)
        times = []
        for k in range(reps):
            flush.fill_(float(k))
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            e0.record()
            D.call("mq_fused_grad_partials", C.byref(net_obj.net), D.ptr(net_obj.theta), D.ptr(self.dev[0][0]),
                   idx_ptr + (k % 8) * mb * 4, mb, C.byref(head), None, None, D.ptr(ws), wsb, C.byref(n_parts), D.stream())
            e1.record()
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        del flush
        return float(np.median(times[2:])), n_parts.value

    def policy_head(self, upd):
        from mannequin2_b200 import _lib as L
        D = self.D
        obs, act = self.dev[0][0], self.dev[0][1]
        head = L.MqHead()
        head.kind, head.p0, head.p1 = L.HEAD_GAUSS_PPO, 0.8, 1.2
        head.target, head.adv, head.logp_old = D.ptr(act), D.ptr(upd._bufs["adv_n"]), D.ptr(upd._bufs["logp0"])
        pol = upd.policy
        head.tc = pol.planes
        head.records, head.rec_w, head.tc_image = D.ptr(pol.records), pol.rec_w, D.ptr(pol.image)
        return head

    def dominant_kernel(self, peaks):
        """tc3_grad_kernel on one policy minibatch of this rank (65,536 rows on one GPU), timed alone on its stream."""
        upd = self.upd
        mb = self.MB // (self.world if self.replicated else 1)
        head = self.policy_head(upd)
        ms, n_parts = self.grad_kernel_time(upd.policy, head, self.D.ptr(self.dev[0][5]), mb)
        planes = upd.policy.planes
        bytes_per_launch = mb * 64                         # 64 B gathered per minibatch row
        flops = mb * 28544.0                               # 28,544 FLOP per row (fwd+dW+dX)
        # tensor-core work actually issued: plane pairs i + j < planes per fp32 product (6 / 3 / 1), operands
        # padded to 16 / 64 columns, M = 128 rows per tile of `tile_rows` valid rows, the ones column for the bias sums
        pairs = {3: 6, 2: 3, 1: 1}[planes & 3]
        slots = 148 * (1 if (planes & 3) == 3 else 2)
        waves = -(-mb // (slots * 128))
        tile_rows = min(128, max(16, -(-(-(-mb // (slots * waves))) // 16) * 16))
        tiles = -(-mb // tile_rows)
        ks = tile_rows // 16
        fwd = 2 * 128 * (16 * 64 + 64 * 64 + 64 * 16)                       # Z = A W, three layers, per pair
        dA = 2 * 128 * (16 * 64 + 64 * 64)                                  # dA = dZ W^T, two layers
        dW = ks * 2 * 16 * 64 * (16 + (64 + 8) + (16 + 8))                  # dW^T, db (M = 64 instructions)
        executed = tiles * float(pairs * (fwd + dA + dW) + ks * 2 * 16 * 64 * 16)
        tf = flops / (ms * 1e-3) / 1e12
        measured = load_measured_peaks()
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tj = json.load(f)["tc3_grad_kernel_policy_65536rows"][self.tc_mode]
                traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        except (OSError, KeyError, ValueError):
            pass
        return dict(bound="tensor", kernel="tc3_grad_kernel<planes=%d%s> (policy, %d rows, record-fed tcgen05, %s mode)"
                                           % (planes & 3, ", fp16 scaled" if planes & 0x80 else "", mb, self.tc_mode),
                    achieved=tf, peak=peaks["bf16"], unit="TFLOP/s", frac=tf / peaks["bf16"], traffic=traffic,
                    traffic_note="ncu dram__bytes_read+write of one 65,536-row launch (%s); algorithmic gather 4.2 MB + "
                                 "0.26 MB of indices + the operand image" % traffic_src,
                    peak_source=peaks["source"], ms_per_launch=ms, n_partials=n_parts, rows=mb, mode=self.tc_mode,
                    limiter="a 64-wide MLP is a chain of six dependent (product -> epilogue) phases per 128-row tile with "
                            "K = 16..64: the epilogues (bias, tanh, split into operand planes, shared-memory stores: ~20 "
                            "instructions per element) are instruction-issue bound and the tensor pipe waits for them; "
                            "neither HBM (64 B/row) nor the dense bf16 peak is the bound (DESIGN.md 3.1)",
                    executed_tensor_tflops=executed / (ms * 1e-3) / 1e12,
                    executed_tensor_frac=executed / (ms * 1e-3) / 1e12 / peaks["bf16"],
                    hbm_gbs=bytes_per_launch / (ms * 1e-3) / 1e9, hbm_frac=bytes_per_launch / (ms * 1e-3) / 1e9 / peaks["hbm"],
                    fp32_ffma_peak_tflops=measured.get("ffma_tflops"), fp32_frac_of_ffma_peak=(
                        tf / measured["ffma_tflops"] if measured.get("ffma_tflops") else None),
                    tf32_peak_tflops=measured.get("tf32_tflops"), peaks_measured_here=measured.get("how"))

    def precision_modes(self, steps=3):
        """The looser tensor-core modes next to the exact one (north_star: "a stated looser tolerance for bf16 tensor-core
        paths"): whole-update throughput, gradient-kernel time and the MEASURED error of one 65,536-row policy gradient
        against the exact mode."""
        import ctypes as C
        t, D = self.torch, self.D
        from mannequin2_b200.ppo import PPOUpdater
        from mannequin2_b200 import _lib as L
        out = {}
        grads = {}
        for mode in os.environ.get("BENCH_MODES", "fp16x2,exact,bf16x2,bf16").split(","):
            np.random.seed(0)
            upd = PPOUpdater(11, 3, tc_mode=mode)
            upd.check_indices = False
            for i in range(2):
                upd.update_dev(*self.dev[i % len(self.dev)])
            t.cuda.synchronize()
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                upd.update_dev(*self.dev[i % len(self.dev)])
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1) / steps
            # one gradient on IDENTICAL inputs (the exact updater's parameters and packed records), this mode's arithmetic
            planes = upd.policy.planes
            head = self.policy_head(self.upd)
            image = t.zeros(int(L.lib().mq_tc_image_bytes(C.byref(self.upd.policy.net), planes)), dtype=t.uint8, device="cuda")
            D.call("mq_tc_image_build", C.byref(self.upd.policy.net), D.ptr(self.upd.policy.theta), planes, D.ptr(image), D.stream())
            head.tc, head.tc_image = planes, D.ptr(image)
            kms, _ = self.grad_kernel_time(self.upd.policy, head, D.ptr(self.dev[0][5]), self.MB, reps=8)
            wsb = L.lib().mq_fused_ws_bytes(C.byref(self.upd.policy.net), self.MB)
            ws = D.workspace("bench", wsb)
            g = t.zeros(5126, device="cuda")
            D.call("mq_fused_grad", C.byref(self.upd.policy.net), D.ptr(self.upd.policy.theta), D.ptr(self.dev[0][0]),
                   D.ptr(self.dev[0][5]), self.MB, C.byref(head), 1.0 / self.MB, D.ptr(g), None, None, D.ptr(ws), wsb, D.stream())
            g = g.double().cpu().numpy()
            grads[mode] = g
            out[mode] = dict(value=self.N / (ms * 1e-3), unit=self.unit, ms_per_step=ms, grad_kernel_ms=kms,
                             stated_tolerance={"exact": "rtol 1e-5 (the FFMA bar)", "fp16x2": "rtol 1e-5 (the FFMA bar)",
                                               "bf16x2": 3e-5, "bf16": 1e-1}[mode],
                             note="grad_err_vs_exact = max |g - g_exact| / max |g_exact| of one 65,536-row policy gradient at "
                                  "the same parameters (g_exact: the three-plane bf16 mode)")
            del upd
        for mode in out:
            if "exact" in grads:
                out[mode]["grad_err_vs_exact"] = float(np.max(np.abs(grads[mode] - grads["exact"])) / np.max(np.abs(grads["exact"])))
        return out

    def parity_check(self, rows_per_rank=8192, epochs=1):
        """Outside the timed region: a small update through the SAME dispatch as the benchmark (record-fed tcgen05 kernel
        + in-kernel exchange when sharded) -- every rank ends with bit-identical parameters, and rank 0 compares them with
        the oracle's restatement of benchmarks/ppo.py:22-40 + gae.py:34-75 on the whole problem."""
        import torch.distributed as dist
        t, D = self.torch, self.D
        from mannequin2_b200.ppo import PPOUpdater, init_mlp
        world = self.world
        rank = dist.get_rank() if world > 1 else 0
        mb = rows_per_rank * world
        n = 2 * mb
        rs = np.random.RandomState(77)                  # same problem on every rank
        o, a, r, d = ppo_rollout(rs, n, p_done=1.0 / 300, cut=700)
        vi = np.concatenate([rs.permutation(n).astype(np.int32).reshape(n // mb, mb) for _ in range(epochs)])
        pi = np.concatenate([rs.permutation(n).astype(np.int32).reshape(n // mb, mb) for _ in range(epochs)])
        np.random.seed(5)
        pol0 = np.concatenate([init_mlp(11, [64, 64, 3], 0.1), np.zeros(3, np.float32)])
        val0 = init_mlp(11, [64, 64, 1], 0.1)
        upd = PPOUpdater(11, 3, policy_params=pol0.copy(), value_params=val0.copy(), tc_mode=self.tc_mode)
        upd.update_dev(*[D.from_host(x) for x in (o, a, r, d, vi, pi)], replicated=world > 1)
        t.cuda.synchronize()
        both = t.cat([upd.policy.theta, upd.value.theta])
        identical = True
        if world > 1:
            allp = [t.empty_like(both) for _ in range(world)]
            dist.all_gather(allp, both)
            identical = all(bool(t.equal(allp[0].view(t.int32), q.view(t.int32))) for q in allp)
        res = dict(replicas_identical=identical, exchange=upd.policy.exchange(), rows_per_rank=rows_per_rank,
                   global_minibatch=mb, transitions=n, steps_per_net=len(vi), mode=self.tc_mode,
                   gradient_kernel="tc3_grad_kernel (records)" if upd.policy.records is not None else "FFMA")
        if rank == 0:
            from oracle import mq_oracle as O
            pol_spec = O.policy_spec(11, 3)
            val_spec = O.mlp_spec(11, [64, 64, 1], [O.ACT_TANH, O.ACT_TANH, O.ACT_NONE])
            want_pol, want_val, _ = O.ppo_update(pol0.copy(), pol_spec, O.AdamO(pol0, horizon=10), val0.copy(), val_spec,
                                                 O.AdamO(val0, horizon=10, lr=0.003), O.RunningNormalizeO(horizon=2),
                                                 o, a, r, d.astype(bool), vi, pi)
            e_pol = float(np.max(np.abs(D.to_host(upd.policy.theta) - want_pol)))
            e_val = float(np.max(np.abs(D.to_host(upd.value.theta) - want_val)))
            frac = {"exact": 0.007, "fp16x2": 0.007, "bf16x2": 0.07, "bf16": 1.0}[self.tc_mode]
            steps = len(vi)
            res.update(max_abs_err=max(e_pol, e_val), max_abs_err_policy=e_pol, max_abs_err_value=e_val,
                       tol_policy=frac * 3e-4 * (1 + steps), tol_value=frac * 3e-3 * (1 + steps),
                       within_tolerance=bool(e_pol < frac * 3e-4 * (1 + steps) and e_val < frac * 3e-3 * (1 + steps)),
                       oracle="oracle/mq_oracle.py: ppo_update on the concatenated rows (fp64)")
        del upd
        return res

    # CPU arm: the same per-transition work on a bounded sample
    SAMPLE_N = 65536

    def setup_cpu(self):
        PPOWorkload.setup_cpu(self)
        rs = np.random.RandomState(5)
        n = self.SAMPLE_N
        o, a, r, d = ppo_rollout(rs, n, p_done=1.0 / 500, cut=1000)
        tabs = [np.stack([rs.permutation(n).astype(np.int32) for _e in range(self.EPOCHS)]) for _t in range(2)]
        self.cpu_sample = (o, a, r, d, tabs[0], tabs[1])

    def step_cpu(self, i):
        o, a, r, d, vi, pi = self.cpu_sample
        self.pol_theta, self.val_theta, _ = self.O.ppo_update(
            self.pol_theta, self.pol_spec, self.pol_opt, self.val_theta, self.val_spec, self.val_opt,
            self.norm, o, a, r, d.astype(bool), vi, pi, dtype=np.float32)

    cpu_units_per_step = SAMPLE_N
    cpu_sample_desc = ("65,536 transitions, 10 epochs x 1 minibatch of 65,536 rows per net "
                       "(same 10 passes per transition as the full workload)")


class MLPWorkload(object):
    """BASELINE.json configs[0]: mnist/mlp.py step -- 784-128-128-10 LReLU(0.1) MLP (script shape,
    SURVEY F9), softmax-CE(+0.01 logit L2) gradient, Adam(h=10, lr 1e-3), batch 128 drawn by index
    from 60,000 synthetic MNIST-shaped images.  One bench step = one block of 32 optimiser steps
    (one CUDA-graph replay)."""
    name = "mlp_784-128-128-10_batch128 (configs[0])"
    metric = "mlp_fwd_bwd_adam_samples_per_s"
    unit = "samples/s"
    B, BLOCK = 128, 32

    def __init__(self, seed=1, n_data=60000):
        rs = np.random.RandomState(seed)
        self.x = rs.rand(n_data, 784).astype(np.float32)
        self.y = np.eye(10, dtype=np.float32)[rs.randint(10, size=n_data)]
        self.tables = [rs.randint(n_data, size=(self.BLOCK, self.B)).astype(np.int32) for _ in range(16)]
        self.units_per_step = self.BLOCK * self.B

    def setup_gpu(self):
        import torch
        from mannequin2_b200 import _device as D
        from mannequin2_b200.mlp import MLPTrainer
        np.random.seed(0)
        self.torch, self.D = torch, D
        self.tr = MLPTrainer(batch=self.B, graph_steps=self.BLOCK)
        self.xd, self.yd = D.from_host(self.x), D.from_host(self.y)
        self.tabs_d = [D.from_host(t) for t in self.tables]
        self.h2d = self.BLOCK * self.B * (784 + 10) * 4
        self.d2h = 118282 * 4
        self.pin_x = torch.empty(self.B, 784).pin_memory()
        self.pin_y = torch.empty(self.B, 10).pin_memory()
        self.out_pinned = torch.empty(118282).pin_memory()

    def step_gpu(self, i):
        self.tr.train_block(self.xd, self.yd, self.tabs_d[i % len(self.tabs_d)])

    def step_e2e(self, i):
        # the reference's own call: sgd_step(host batch, host labels), 32 times, then read the params
        tab = self.tables[i % len(self.tables)]
        for s in range(self.BLOCK):
            self.tr.sgd_step(self.x[tab[s]], self.y[tab[s]])
        self.out_pinned.copy_(self.tr.theta, non_blocking=True)
        self.torch.cuda.current_stream().synchronize()

    def launches_per_step(self):
        return 1 if self.tr.fused else self.BLOCK * 11           # fused: ONE persistent launch per block of 32 steps

    def replica_throughput(self, n_replicas, steps=5, warmup=3):
        """Aggregate samples/s of n_replicas INDEPENDENT trainers (own parameters and Adam state) whose graph replays
        run concurrently on their own streams of ONE GPU: a batch-128 step keeps a handful of SMs busy, so independent
        learners (hyper-parameter sweeps, seeds: how the reference uses a multi-core box) fill the rest."""
        t = self.torch
        from mannequin2_b200.mlp import MLPTrainer
        trs = [MLPTrainer(batch=self.B, graph_steps=self.BLOCK) for _ in range(n_replicas)]
        streams = [t.cuda.Stream() for _ in range(n_replicas)]
        t.cuda.synchronize()

        def one_round(i):
            for r, (tr, s) in enumerate(zip(trs, streams)):
                with t.cuda.stream(s):
                    tr.train_block(self.xd, self.yd, self.tabs_d[(i + r) % len(self.tabs_d)])
        for i in range(warmup):
            one_round(i)
        t.cuda.synchronize()
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams:
            s.wait_event(e0)
        for i in range(steps):
            one_round(i)
        for s in streams:
            t.cuda.current_stream().wait_stream(s)
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        return dict(replicas=n_replicas, value=n_replicas * steps * self.units_per_step / (ms * 1e-3), unit=self.unit,
                    ms_per_round=ms / steps)

    def dominant_kernel(self, peaks):
        t = self.torch
        times = []
        for k in range(6):
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            e0.record()
            self.tr.train_block(self.xd, self.yd, self.tabs_d[k])
            e1.record()
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times[2:])) / self.BLOCK
        bytes_per_step = self.B * 3216 + 24 * 118282        # 3.25 MB per step
        ach = bytes_per_step / (ms * 1e-3) / 1e9
        if self.tr.fused:
            kern = ("mlp_block_kernel (csrc/mq_mlp_fused.cu): 32 optimiser steps in ONE persistent launch of 128 CTAs = 16 clusters "
                    "of 8; first layer and its Adam moments resident in shared memory, K-slice partials reduced through "
                    "distributed shared memory, two grid-wide barriers per step; ms_per_launch is per STEP")
            note = ("latency-bound at batch 128 (3.25 MB / 65 MFLOP per step is 0.5 us at roofline): a step is two grid barriers, "
                    "one cluster barrier and a 64 KB reload of W1 per CTA")
        else:
            kern = ("one optimiser step (11 launches: split-K fwd GEMM + finish, 2 fwd GEMMs, head, 5 bwd GEMMs with the bias "
                    "gradients as an extra output row, Adam)")
            note = "latency-bound at batch 128: 3.25 MB / 65 MFLOP per step is 0.5 us at roofline"
        return dict(bound="hbm", kernel=kern, achieved=ach, peak=peaks["hbm"], unit="GB/s", frac=ach / peaks["hbm"], traffic=None,
                    peak_source=peaks["source"], ms_per_launch=ms, note=note)

    def sweep(self, peaks):
        """BASELINE.md 3.4 / SURVEY 8d: the batch sweep B in {128, 1024, 8192, 65536} for the script's net (784-128-128-10,
        mnist/mlp.py:24-27) and BASELINE.json's wording (784-128-10), fwd + bwd + Adam per step, each with its roofline:
        HBM (3,216 B per sample + 24 B per parameter) where the step is memory / latency bound, the measured dense bf16
        peak where the Affine layers are real contractions (from batch ~1k they run on the TMA-fed tcgen05 GEMM)."""
        t = self.torch
        from mannequin2_b200.mlp import MLPTrainer
        from mannequin2_b200 import _lib as L
        rs = np.random.RandomState(11)
        out = {}
        for widths, acts, tag in (((128, 128, 10), (L.ACT_LRELU, L.ACT_LRELU, L.ACT_NONE), "784-128-128-10"),
                                  ((128, 10), (L.ACT_LRELU, L.ACT_NONE), "784-128-10")):
            s_all = sum(a * b for a, b in zip((784,) + widths[:-1], widths))
            s_rest = s_all - 784 * widths[0]
            flop_per_sample = 4 * s_all + 2 * s_rest                     # 2S fwd + 2S dW + 2S' dX
            n_params = s_all + sum(widths)
            for B in (128, 1024, 8192, 65536):
                block = 32 if B <= 1024 else (8 if B <= 8192 else 3)
                np.random.seed(0)
                tr = MLPTrainer(784, widths, acts, batch=B, graph_steps=block)
                tab = self.D.from_host(rs.randint(len(self.x), size=(block, B)).astype(np.int32))
                times = []
                for k in range(5):
                    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
                    e0.record()
                    tr.train_block(self.xd, self.yd, tab)
                    e1.record()
                    e1.synchronize()
                    times.append(e0.elapsed_time(e1))
                ms = float(np.median(times[2:])) / block
                bytes_per_step = B * 3216 + 24 * n_params
                gbs = bytes_per_step / (ms * 1e-3) / 1e9
                tf = B * flop_per_sample / (ms * 1e-3) / 1e12
                bound = "tensor" if tf / peaks["bf16"] > gbs / peaks["hbm"] else "hbm"
                out["%s_B%d" % (tag, B)] = dict(samples_per_s=B / (ms * 1e-3), ms_per_step=ms, fp32_tflops=tf,
                                                frac_of_bf16_peak=tf / peaks["bf16"], issued_bf16_frac=6 * tf / peaks["bf16"],
                                                hbm_gbs=gbs, frac_of_hbm_peak=gbs / peaks["hbm"], bound=bound,
                                                roofline_frac=max(tf / peaks["bf16"], gbs / peaks["hbm"]))
                del tr, tab
                t.cuda.empty_cache()
        return out

    def setup_cpu(self):
        from oracle import mq_oracle as O
        self.O = O
        self.spec = O.mlp_spec(784, [128, 128, 10], [O.ACT_LRELU, O.ACT_LRELU, O.ACT_NONE])
        self.theta = O.mlp_init(self.spec, np.random.RandomState(0), 0.1)
        self.opt = O.AdamO(self.theta, horizon=10, lr=1e-3)

    def step_cpu(self, i):
        tab = self.tables[i % len(self.tables)]
        for s in range(self.BLOCK):
            self.theta = self.O.mlp_sgd_step(self.theta, self.spec, self.opt, self.x[tab[s]], self.y[tab[s]])


class ESWorkload(object):
    """BASELINE.json configs[3]: evolution strategies, population 4096 x 64-64 policy (P=5,123),
    shared batch of 1000 observations, synthetic linear return, RunningNormalize(h=10), Adam(lr .01).
    The population is sharded across GPUs (4096 members per GPU: weak scaling)."""
    name = "es_population4096_T1000 (configs[3])"
    metric = "es_members_per_s"
    unit = "members/s"
    M, T = 4096, 1000

    def __init__(self, seed=1, m=None):
        self.M = m or self.M
        rs = np.random.RandomState(seed)
        self.noise = (rs.randn(self.M, 5123) * 0.1).astype(np.float32)
        self.obs = np.clip(rs.randn(self.T, 11), -5, 5).astype(np.float32)
        self.coef = np.array([1.0, -0.5, 0.25], np.float32)
        self.units_per_step = self.M

    def setup_gpu(self):
        import torch
        from mannequin2_b200 import _device as D
        from mannequin2_b200.es import ESPopulation
        np.random.seed(0)
        self.torch, self.D = torch, D
        self.es = ESPopulation(11, 3)
        self.nd, self.od, self.cd = D.from_host(self.noise), D.from_host(self.obs), D.from_host(self.coef)
        self.pin = torch.from_numpy(self.noise).pin_memory()
        self.out_pinned = torch.empty(self.M).pin_memory()
        self.h2d, self.d2h = self.noise.nbytes, self.M * 4

    def step_gpu(self, i):
        self.es.generation(self.nd, self.od, self.cd)

    def step_e2e(self, i):
        """Noise (84 MB) from pinned host memory every generation; the upload for generation i+1 is issued on a copy
        stream before generation i's kernels, returns are read back every generation."""
        t = self.torch
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream, self._next = t.cuda.Stream(), None

        def upload():
            with t.cuda.stream(self._copy_stream):
                nd_ = self.pin.to(self.D.device(), non_blocking=True)
                ev_ = t.cuda.Event()
                ev_.record(self._copy_stream)
            return nd_, ev_
        if self._next is None:
            self._next = upload()
        nd, ev = self._next
        main = t.cuda.current_stream()
        main.wait_event(ev)
        self._copy_stream.wait_stream(main)
        self._next = upload()
        rets = self.es.generation(nd, self.od, self.cd)
        nd.record_stream(main)
        self.out_pinned.copy_(rets, non_blocking=True)
        main.synchronize()

    def launches_per_step(self):
        return 12

    def dominant_kernel(self, peaks):
        import ctypes as C
        t, D = self.torch, self.D
        value, _, _, theta = self.es.opt.state()
        rets = t.empty(self.M, dtype=t.float32, device=D.device())
        times = []
        for k in range(6):
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            e0.record()
            D.call("mq_es_eval", C.byref(self.es.net), D.ptr(theta), D.ptr(self.nd), self.M, D.ptr(self.od),
                   self.T, D.ptr(self.cd), D.ptr(rets), D.stream())
            e1.record()
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times[2:]))
        bytes_per_launch = self.M * 5123 * 4               # noise read once, 20.5 KB / member
        flops = self.M * 2.0 * 4992 * self.T
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
        tf = flops / (ms * 1e-3) / 1e12
        return dict(bound="tensor", kernel="tc_es_kernel<16,tanh> (4096 members x 1000 observations, tcgen05 bf16x3 planes)",
                    achieved=tf, peak=peaks["bf16"], unit="TFLOP/s", frac=tf / peaks["bf16"], traffic=None,
                    peak_source=peaks["source"], ms_per_launch=ms,
                    limiter="forward-only 64-wide layers: serial MMA phase -> epilogue per layer, one CTA per SM; "
                            "noise is read once (20.5 KB per member = %.0f GB/s)" % (bytes_per_launch / (ms * 1e-3) / 1e9),
                    fp32_tflops=tf, fp32_frac=tf / fp32_peak)

    def setup_cpu(self):
        from oracle import mq_oracle as O
        self.O = O
        self.spec = O.mlp_spec(11, [64, 64, 3], [O.ACT_TANH, O.ACT_TANH, O.ACT_NONE])
        self.theta = O.mlp_init(self.spec, np.random.RandomState(0))
        self.opt = O.AdamO(self.theta, horizon=10, lr=0.01)
        self.norm = O.RunningNormalizeO(horizon=10)

    cpu_units_per_step = 256
    cpu_sample_desc = "256 of the 4096 members (same observations)"

    def step_cpu(self, i):
        self.theta, _, _ = self.O.es_generation(self.theta, self.spec, self.opt, self.norm,
                                                self.noise[:256], self.obs, self.coef)


class VAEWorkload(object):
    """BASELINE.json configs[4]: mnist/vae.py step at the BASELINE shape -- encoder 784-400-(20+20), decoder
    20-400-(784+784), Gaussian reparameterisation, batch 4096 -- as the fused plan of csrc/mq_vae.cu (mq_vae_step: one
    forward and one backward pass for the two evaluations of vae.py:57-66, every product on the TMA-fed tcgen05 GEMM with
    its operands kept as packed fp16x2 planes, clipped-momentum ascent).  `value`: batch resident on the device, the step
    replayed as ONE CUDA graph, sampling noise from the device Philox generator (VAETrainer.step_graph).  `e2e`: the batch
    comes from pinned host memory every step (12.8 MB) and the step's mean log-likelihood / DKL are read back."""
    name = "vae_784-400-20-400-784_batch4096 (configs[4])"
    metric = "vae_step_samples_per_s"
    unit = "samples/s"
    B = 4096

    def __init__(self, seed=1):
        rs = np.random.RandomState(seed)
        self.xs = [rs.rand(self.B, 28, 28).astype(np.float32) for _ in range(4)]
        self.units_per_step = self.B

    def setup_gpu(self):
        import torch
        from mannequin2_b200 import _device as D
        from mannequin2_b200.vae import VAETrainer
        np.random.seed(0)
        self.torch, self.D = torch, D
        self.tr = VAETrainer(image_shape=(28, 28), hidden=400, latent=20, n_hidden=1, rng=D.DeviceRNG(2))
        self.dev = [D.from_host(x) for x in self.xs]
        self.pinned = [torch.from_numpy(x).pin_memory() for x in self.xs]
        self.out_pinned = torch.empty(2).pin_memory()
        self.h2d, self.d2h = self.xs[0].nbytes, 8

    def step_gpu(self, i):
        self.tr.step_graph(self.dev[i % len(self.dev)])

    def _upload(self, i):
        """Batch i from pinned host memory into one of two device buffers, on the copy stream."""
        t = self.torch
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = t.cuda.Stream()
            self._bufs = [t.empty(self.B, 28, 28, device=self.D.device()) for _ in range(2)]
        buf, ev = self._bufs[i % 2], t.cuda.Event()
        with t.cuda.stream(self._copy_stream):
            buf.copy_(self.pinned[i % len(self.pinned)], non_blocking=True)
            ev.record()
        return i, buf, ev

    def step_e2e(self, i):
        # every step's batch is copied from pinned host memory (12.8 MB); the copy of batch i + 1 is issued on a copy stream
        # before batch i's kernels, as a training loop with a prefetching loader does (the PPO arm does the same)
        t = self.torch
        pend = getattr(self, "_pending", None)
        if pend is None or pend[0] != i:
            pend = self._upload(i)
        _, x, ev = pend
        self._pending = self._upload(i + 1)
        t.cuda.current_stream().wait_event(ev)
        lp, kl = self.tr.step_graph(x)
        self.out_pinned[0:1].copy_(lp.reshape(1), non_blocking=True)
        self.out_pinned[1:2].copy_(kl.reshape(1), non_blocking=True)
        t.cuda.current_stream().synchronize()

    def launches_per_step(self):
        # fused plan (csrc/mq_vae.cu, one hidden layer per side): 3 packs, 7 products + 4 split-K dW products and their 4
        # finishes, latent forward / backward, head, momentum step = 22 kernels of this library + the Philox draw
        return 23

    def dominant_kernel(self, peaks):
        """pack + gemm_tma_kernel on the widest layer (4096 x 784 x 400, y = x W), timed alone."""
        t, D = self.torch, self.D
        from mannequin2_b200 import _lib as L
        x, w = t.randn(4096, 784, device=D.device()), t.randn(784, 400, device=D.device())
        y = t.empty(4096, 400, device=D.device())
        wsb = int(L.lib().mq_gemm_ws_bytes(4096, 400, 784))
        ws = t.empty(wsb, dtype=t.uint8, device=D.device())
        times = []
        for k in range(8):
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            e0.record()
            D.call("mq_gemm_ex", D.ptr(x), 784, 1, D.ptr(w), 400, 1, D.ptr(y), 4096, 400, 784, 1.0, L.TC_MODES["fp16x2"], D.ptr(ws), wsb,
                   D.stream())
            e1.record(); e1.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times[2:]))
        tf = 2.0 * 4096 * 784 * 400 / (ms * 1e-3) / 1e12
        return dict(bound="tensor", kernel="pack_planes + gemm_tma_kernel<2, fp16 scaled> (4096x784x400, fp16x2 mode, TMA-fed)",
                    achieved=tf, peak=peaks["bf16"], unit="TFLOP/s", frac=tf / peaks["bf16"], traffic=None,
                    peak_source=peaks["source"], ms_per_launch=ms, issued_bf16_tflops=3 * tf, issued_frac=3 * tf / peaks["bf16"],
                    note="fp32-exact product = 3 pairs of scaled fp16 planes issued as wide tcgen05 instructions from TMA-loaded, "
                         "128-byte-swizzled planes; the time is pack (1 launch) + GEMM of a 2.6 GFLOP product through mq_gemm_ex; inside the fused "
                         "plan the operands arrive packed (DESIGN.md 3.5, 3.8)")


class ClosedLoopWorkload(PPOLargeWorkload):
    """the rollout side and the update path as one closed loop on the device.  One step = 4,096
    synthetic environments advanced 256 steps in lock step by the current policy (observation normaliser, policy
    forward, Gaussian sample, environment: three launches per time step, replayed as one CUDA graph) = 2^20 transitions,
    then the configs[2] update on them (per-segment bootstrap, value forward, GAE, 10 epochs x 16 minibatches of 65,536
    rows for both nets).  Only the minibatch index tables come from the host (the reference draws them with NumPy)."""
    name = "ppo_closed_loop_4096envs_x256steps + configs[2] update"
    metric = "ppo_closed_loop_transitions_per_s"
    N_ENV, T_LEN = 4096, 256
    step_cpu = None                                     # the CPU port covers the update path only

    def __init__(self, n_rollouts=2, seed=1):
        rs = np.random.RandomState(seed)
        self.N = self.N_ENV * self.T_LEN
        per = self.N // self.MB
        self.tables = [tuple(np.concatenate([rs.permutation(self.N).astype(np.int32).reshape(per, self.MB)
                                             for _e in range(self.EPOCHS)]) for _t in range(2))
                       for _ in range(n_rollouts)]
        self.units_per_step = self.N
        self.steps_per_net = self.EPOCHS * per

    def setup_gpu(self):
        import torch
        from mannequin2_b200 import _device as D
        from mannequin2_b200.ppo import PPOUpdater
        from mannequin2_b200.rollout import BatchedActor, SyntheticEnv
        np.random.seed(0)
        self.torch, self.D = torch, D
        self.env = SyntheticEnv(self.N_ENV, horizon=256, seed=3)
        self.world, self.replicated, self.tc_mode = 1, False, "fp16x2"
        self.upd = PPOUpdater(11, 3, tc_mode=self.tc_mode)
        self.upd.check_indices = False
        self.actor = BatchedActor(self.env, self.upd.policy, seed=5, use_graph=True)
        self.dev = [tuple(D.from_host(x) for x in r) for r in self.tables]
        self.pinned = [tuple(torch.from_numpy(x).pin_memory() for x in r) for r in self.tables[:1]]
        self.out_pinned = (torch.empty(5126, dtype=torch.float32).pin_memory(),
                           torch.empty(4993, dtype=torch.float32).pin_memory())
        self.h2d = sum(x.nbytes for x in self.tables[0])
        self.d2h = (5126 + 4993) * 4
        self.mean_rewards = []

    def step_gpu(self, i):
        chunk = self.actor.collect(self.T_LEN)
        self.upd.update_rollout(chunk, self.N_ENV, self.T_LEN, *self.dev[i % len(self.dev)])

    def step_e2e(self, i):
        t = self.torch
        main = t.cuda.current_stream()
        tabs = [x.to(self.D.device(), non_blocking=True) for x in self.pinned[0]]
        chunk = self.actor.collect(self.T_LEN)
        self.upd.update_rollout(chunk, self.N_ENV, self.T_LEN, *tabs)
        self.out_pinned[0].copy_(self.upd.policy.theta, non_blocking=True)
        self.out_pinned[1].copy_(self.upd.value.theta, non_blocking=True)
        main.synchronize()
        self.mean_rewards.append(float(chunk["rew"].mean()))

    def launches_per_step(self):
        return 3 * self.T_LEN + 3 + 13 + 2 * self.steps_per_net * 2

    def dominant_kernel(self, peaks):
        self.step_gpu(0)                                # makes the buffers the parent's timing loop reads
        self.dev = [(self.actor._bufs[self.T_LEN]["obs"], self.actor._bufs[self.T_LEN]["act"], None, None) + d
                    for d in self.dev]
        r = PPOLargeWorkload.dominant_kernel(self, peaks)
        t = self.torch
        times = []
        for _ in range(5):
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            e0.record()
            self.actor.collect(self.T_LEN)
            e1.record()
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        r["collect_ms"] = float(np.median(times))
        r["collect_transitions_per_s"] = self.N / (r["collect_ms"] * 1e-3)
        r["mean_reward_first_last"] = [self.mean_rewards[0], self.mean_rewards[-1]] if self.mean_rewards else None
        return r


WORKLOADS = {"closed_loop": ClosedLoopWorkload, "ppo": PPOWorkload, "ppo_large": PPOLargeWorkload, "mlp": MLPWorkload, "es": ESWorkload, "vae": VAEWorkload}


def kernel_sweep(peaks):
    """HBM-bound kernels at sizes far beyond L2 (126 MB): achieved algorithmic GB/s vs measured peak."""
    import ctypes as C
    import torch
    from mannequin2_b200 import _device as D, _lib as L
    lib, dev, out = L.lib(), D.device(), {}

    def timeit(fn, reps=5, inner=4):
        """Median device time of fn per call, `inner` calls enqueued back to back per event pair: the kernels here take
        0.1-0.4 ms, and a single launch between two events also measures the event records and the launch gap (3-8 us:
        tanh forward read 0.099 ms alone against 0.0925 ms per launch in a train of them).  Arrays of 256 MB - 1 GB: every
        launch streams from HBM, nothing survives in the 126 MB L2 from the previous one."""
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _i in range(inner):
                fn()
            e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1) / inner)
        return float(np.median(ts[1:]))

    n = 64 << 20                                            # 64 Mi params
    val, m, v, g = (torch.zeros(n, device=dev) for _ in range(4))
    g.normal_()
    th = torch.empty(n, device=dev)
    ms = timeit(lambda: D.call("mq_adam_step", D.ptr(val), D.ptr(m), D.ptr(v), D.ptr(th), D.ptr(g), n, L.MQ_F32,
                               L.MQ_F32, 0.1, 0.01, 1e-3, 1e-8, 0, D.stream()))
    out["adam_f32_64Mi"] = dict(bytes_per_elem=32, ms=ms, gbs=32.0 * n / ms / 1e6, frac=32.0 * n / ms / 1e6 / peaks["hbm"])
    del val, m, v, g, th
    n = 32 << 20
    val, m, v = (torch.zeros(n, device=dev, dtype=torch.float64) for _ in range(3))
    g = torch.randn(n, device=dev)
    th = torch.empty(n, device=dev)
    ms = timeit(lambda: D.call("mq_adam_step", D.ptr(val), D.ptr(m), D.ptr(v), D.ptr(th), D.ptr(g), n, L.MQ_F64,
                               L.MQ_F32, 0.1, 0.01, 1e-3, 1e-8, 0, D.stream()))
    out["adam_f64_32Mi"] = dict(bytes_per_elem=56, ms=ms, gbs=56.0 * n / ms / 1e6, frac=56.0 * n / ms / 1e6 / peaks["hbm"])
    del val, m, v, g, th
    n = 64 << 20                                            # 64 Mi transitions
    r, vv = torch.randn(n, device=dev), torch.randn(n + 1, device=dev)
    d = (torch.rand(n, device=dev) < 0.002).to(torch.uint8)
    adv, ret = torch.empty(n, device=dev), torch.empty(n, device=dev)
    wsb = lib.mq_scan_ws_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    ms = timeit(lambda: D.call("mq_gae", D.ptr(r), D.ptr(vv), D.ptr(d), n, 0.99, 0.95, D.ptr(adv), D.ptr(ret),
                               D.ptr(ws), wsb, D.stream()))
    out["gae_64Mi"] = dict(bytes_per_elem=17, ms=ms, gbs=17.0 * n / ms / 1e6, frac=17.0 * n / ms / 1e6 / peaks["hbm"])
    from mannequin2_b200 import RunningNormalize
    norm = RunningNormalize(horizon=2)
    o = torch.empty(n, device=dev)
    def nrm():
        norm.update_dev(r, n)
        norm.apply_dev(r, n, out=o, fuse_tanh=True)
    nrm()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()                        # four short launches per call: replayed as a graph so that the
    with torch.cuda.graph(graph):                         # host does not leave bubbles between them
        for _i in range(4):
            nrm()
    ms = timeit(graph.replay, inner=1) / 4
    del graph
    # 8 B read (one statistics pass + the apply pass) + 4 B write per element
    out["normalize_update_apply_64Mi"] = dict(bytes_per_elem=12, ms=ms, gbs=12.0 * n / ms / 1e6,
                                              frac=12.0 * n / ms / 1e6 / peaks["hbm"])
    # activations (mannequin/backprop.py:53-60): tanh forward 8 B/element, backward 12 B/element
    ya, ga = torch.empty(n, device=dev), torch.empty(n, device=dev)
    ms = timeit(lambda: D.call("mq_act_forward", D.ptr(r), D.ptr(ya), n, L.ACT_TANH, 0.0, D.stream()))
    out["tanh_forward_64Mi"] = dict(bytes_per_elem=8, ms=ms, gbs=8.0 * n / ms / 1e6, frac=8.0 * n / ms / 1e6 / peaks["hbm"])
    ms = timeit(lambda: D.call("mq_act_backward", D.ptr(ya), D.ptr(r), D.ptr(ga), n, L.ACT_TANH, 0.0, D.stream()))
    out["tanh_backward_64Mi"] = dict(bytes_per_elem=12, ms=ms, gbs=12.0 * n / ms / 1e6, frac=12.0 * n / ms / 1e6 / peaks["hbm"])
    del r, vv, d, adv, ret, o, ya, ga
    # wide Affine layers (BASELINE configs[4]: VAE 784-400-...-784 at batch 4096): y = x W, dx = g W^T, dW = x^T g through
    # mq_gemm -- tcgen05 kernel (three bf16 planes per operand, fp32 parity) vs the FFMA kernel it replaces
    bsz, kin, nout = 4096, 784, 400
    x, w, gy = torch.randn(bsz, kin, device=dev), torch.randn(kin, nout, device=dev), torch.randn(bsz, nout, device=dev)
    y, dx, dw = torch.empty(bsz, nout, device=dev), torch.empty(bsz, kin, device=dev), torch.empty(kin, nout, device=dev)
    wsb = max(int(lib.mq_gemm_ws_bytes(bsz, nout, kin)), int(lib.mq_gemm_ws_bytes(bsz, kin, nout)), int(lib.mq_gemm_ws_bytes(kin, nout, bsz)))
    gws = torch.empty(wsb, dtype=torch.uint8, device=dev)

    def three(tc):
        D.call("mq_gemm_ex", D.ptr(x), kin, 1, D.ptr(w), nout, 1, D.ptr(y), bsz, nout, kin, 1.0, tc, D.ptr(gws), wsb, D.stream())    # y = x W
        D.call("mq_gemm_ex", D.ptr(gy), nout, 1, D.ptr(w), 1, nout, D.ptr(dx), bsz, kin, nout, 1.0, tc, D.ptr(gws), wsb, D.stream())  # dx = g W^T
        D.call("mq_gemm_ex", D.ptr(x), 1, kin, D.ptr(gy), nout, 1, D.ptr(dw), kin, nout, bsz, 1.0, tc, D.ptr(gws), wsb, D.stream())   # dW = x^T g
    flops = 3 * 2.0 * bsz * kin * nout
    pairs = {3: 6, 0x82: 3, 2: 3, 1: 1}
    for tc, key in ((3, "affine_4096x784x400_fwd_dx_dw_tcgen05_exact"), (0x82, "affine_4096x784x400_fwd_dx_dw_tcgen05_fp16x2"),
                    (2, "affine_4096x784x400_fwd_dx_dw_tcgen05_bf16x2"),
                    (1, "affine_4096x784x400_fwd_dx_dw_tcgen05_bf16"), (L.TC_OFF, "affine_4096x784x400_fwd_dx_dw_ffma")):
        ms = timeit(lambda: three(tc), reps=8, inner=1)
        out[key] = dict(flop=flops, ms=ms, fp32_tflops=flops / ms / 1e9,
                        frac_of_bf16_peak=flops / ms / 1e9 / peaks["bf16"],
                        issued_bf16_tflops=(pairs[tc] * flops / ms / 1e9) if tc in pairs else None,
                        note="pack of both operands into bf16 planes + TMA-fed tcgen05 GEMM (+ split-K finish for dW), three products")
    # the forward product alone, and a square 4096^3 product, per precision mode (kernel + pack, CUDA events)
    big = 4096
    xa, xb, xc = torch.randn(big, big, device=dev), torch.randn(big, big, device=dev), torch.empty(big, big, device=dev)
    wsb2 = int(lib.mq_gemm_ws_bytes(big, big, big))
    gws2 = torch.empty(wsb2, dtype=torch.uint8, device=dev)
    for tc, name in ((3, "exact"), (0x82, "fp16x2"), (2, "bf16x2"), (1, "bf16")):
        ms = timeit(lambda: D.call("mq_gemm_ex", D.ptr(x), kin, 1, D.ptr(w), nout, 1, D.ptr(y), bsz, nout, kin, 1.0, tc, D.ptr(gws),
                                   wsb, D.stream()), reps=8, inner=1)
        fl = 2.0 * bsz * kin * nout
        out["gemm_4096x784x400_%s" % name] = dict(ms=ms, fp32_tflops=fl / ms / 1e9, frac_of_bf16_peak=fl / ms / 1e9 / peaks["bf16"],
                                                  issued_bf16_tflops=pairs[tc] * fl / ms / 1e9)
        ms = timeit(lambda: D.call("mq_gemm_ex", D.ptr(xa), big, 1, D.ptr(xb), big, 1, D.ptr(xc), big, big, big, 1.0, tc, D.ptr(gws2),
                                   wsb2, D.stream()), reps=5, inner=1)
        fl = 2.0 * big ** 3
        out["gemm_4096cubed_%s" % name] = dict(ms=ms, fp32_tflops=fl / ms / 1e9, frac_of_bf16_peak=fl / ms / 1e9 / peaks["bf16"],
                                               issued_bf16_tflops=pairs[tc] * fl / ms / 1e9,
                                               frac_issued_of_bf16_peak=pairs[tc] * fl / ms / 1e9 / peaks["bf16"])
    del xa, xb, xc, gws2
    return out


# ------------------------------------------------------------------ CPU timing helpers
def _cpu_worker(args):
    name, steps, warmup, seed, threads = args
    limiter = None
    if threads:
        try:
            from threadpoolctl import threadpool_limits
            limiter = threadpool_limits(limits=threads)
        except Exception:
            limiter = None
    wl = WORKLOADS[name](seed=seed + 1, **CPU_KW.get(name, {}))
    wl.setup_cpu()
    for i in range(warmup):
        wl.step_cpu(i)
    t0 = time.perf_counter()
    for i in range(steps):
        wl.step_cpu(warmup + i)
    dt = time.perf_counter() - t0
    del limiter
    return dt


CPU_KW = {"ppo_large": dict(n=65536, n_rollouts=1), "es": dict(m=256), "mlp": dict(n_data=8192)}


def time_cpu(name, steps, warmup, procs, threads=1):
    """`procs` independent replicas in parallel, each limited to `threads` BLAS threads (None: NumPy / OpenBLAS default);
    procs = cores with threads = 1 is how the reference uses a multi-core box (run_benchmarks.sh:98-102, one process per
    seed).  Returns (aggregate units/s, timed seconds of the slowest replica)."""
    import multiprocessing as mp
    wl = WORKLOADS[name](**CPU_KW.get(name, {}))
    units = getattr(wl, "cpu_units_per_step", wl.units_per_step)
    if procs <= 1:
        with mp.get_context("fork").Pool(1) as pool:            # a fresh process: thread limits do not leak into this one
            dt = pool.map(_cpu_worker, [(name, steps, warmup, 0, threads)])[0]
        return steps * units / dt, dt
    with mp.get_context("fork").Pool(procs) as pool:
        dts = pool.map(_cpu_worker, [(name, steps, warmup, s, threads) for s in range(procs)])
    slowest = max(dts)                       # timed region of the slowest replica (set-up/warm-up excluded)
    return procs * steps * units / slowest, slowest


def cpu_three_figures(name, cores):
    """the CPU port timed three ways on this box."""
    v_def, _ = time_cpu(name, 1, 1, 1, threads=None)
    v_one, _ = time_cpu(name, 1, 1, 1, threads=1)
    procs = max(1, min(cores, 32))
    v_rep, _ = time_cpu(name, 1, 1, procs, threads=1)
    return dict(default_threads_1proc=v_def, one_thread_1proc=v_one, replicas=procs, replicas_aggregate=v_rep)


# ------------------------------------------------------------------ main
_T0 = time.time()


def progress(msg):
    """Progress line on stderr (stdout carries only the JSON line)."""
    sys.stderr.write("[bench %7.1fs] %s\n" % (time.time() - _T0, msg))
    sys.stderr.flush()


def measure(wl, steps, warmup, world, rank, flush):
    """W untimed steps, then K steps each timed on the device with its own CUDA-event pair on the
    launching stream; L2 is flushed (256 MB write) between steps, outside the timed intervals.
    Device-resident inputs first (`value`), then the same steps through host buffers (`e2e`).
    Returns (ms_total, ms_e2e_total) as the MAX over ranks."""
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, n):
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
    ms = timed(lambda i: wl.step_gpu(warmup + i), steps)
    barrier()
    wl.step_e2e(0)
    barrier()
    ms_e2e = timed(lambda i: wl.step_e2e(warmup + i), steps)
    barrier()
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), float(t[1])


PARALLELISM = {"ppo_large": "dp%d, STRONG scaling (BASELINE configs[2]): ONE 2^20-transition rollout replicated on every rank, every "
                            "65,536-row global minibatch split %d ways by rows, the flat gradient exchanged over NVLink peer "
                            "memory inside the reduce + Adam kernel (one exchange per optimiser step, no NCCL call on the step "
                            "path); the two forward passes over the rollout are split by rows and all-gathered (NCCL), GAE and "
                            "the normaliser run replicated",
               "ppo": "%d independent replicas (a 64-row minibatch does not shard)",
               "mlp": "%d independent replicas (batch 128 does not shard)",
               "es": "population sharded: %d x 4096 members, all-gather of returns + one all-reduce",
               "vae": "%d independent replicas",
               "closed_loop": "%d independent replicas"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ppo_large", choices=sorted(WORKLOADS))
    ap.add_argument("--no-extras", action="store_true", help="skip the other workloads / kernel sweep")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tc-mode", default="fp16x2", choices=["fp16x2", "exact", "bf16x2", "bf16"],
                    help="operand format of the tensor-core gradient kernel of the headline workload.  fp16x2 (default) and exact "
                         "(three bf16 planes) both meet the fp32 parity bar (rtol 1e-5); bf16x2 / bf16 are the looser modes")
    args = ap.parse_args()
    if os.environ.get("BENCH_WATCHDOG"):                # developer aid: dump every thread's stack after N seconds and exit
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["BENCH_WATCHDOG"]), exit=True, file=sys.stderr)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly ONE line (rank 0's JSON): anything a library prints there (NCCL announces its version on
    # stdout) is sent to stderr instead, and the result is written to the saved descriptor at the end
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(result_fd, (json.dumps(obj) + "\n").encode())
    warmup = max(3, args.warmup)
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        wl = WORKLOADS[args.workload](**CPU_KW.get(args.workload, {}))
        if not hasattr(wl, "step_cpu"):
            emit(dict(impl="reference", unavailable="no CPU port of the %s step in oracle/ yet" % args.workload))
            return 0
        procs = max(1, min(cores, 32))
        steps = max(1, min(args.steps, 3))
        val, wall = time_cpu(args.workload, steps, 1, procs, threads=1)
        sample = getattr(wl, "cpu_sample_desc", "the full workload step")
        line = dict(metric=wl.metric, value=val, unit=wl.unit, n_gpus=args.gpus, steps=steps, warmup=1,
                    ms_per_step=1e3 * wall / steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f32 nets, f64 Adam (the reference's own types)", data="synthetic", impl="reference",
                    config=dict(workload=wl.name, sample=sample,
                                note="CPU restatement (oracle port) of the NumPy reference path -- Tanh nets in float32 as "
                                     "in the reference (basicnet.py:193,210), Adam state float64 (_adam.py:14-18); %d "
                                     "independent single-thread replicas run in parallel, aggregate throughput (how the "
                                     "reference uses a multi-core box: run_benchmarks.sh:98-102)" % procs,
                                host=host_info()),
                    cpu_baseline=dict(value=val, unit=wl.unit, cores=procs, kind="port",
                                      sample="%d steps x %d replicas (OPENBLAS threads = 1 each); each step: %s"
                                             % (steps, procs, sample)),
                    e2e=dict(value=val, unit=wl.unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        emit(line)
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
    large = args.workload == "ppo_large"
    if large:
        wl = WORKLOADS[args.workload](seed=1, tc_mode=args.tc_mode)       # the SAME rollout on every rank (strong scaling)
    else:
        wl = WORKLOADS[args.workload](seed=1 + rank)
    progress("setup %s" % args.workload)
    wl.setup_gpu()
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device="cuda")
    wl.step_gpu(0)
    torch.cuda.synchronize()
    progress("timing %d + %d steps" % (warmup, args.steps))
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    ms, ms_e2e = measure(wl, args.steps, warmup, world, rank, flush)
    clocks = sampler.stop() if rank == 0 else None
    strong = large and world > 1
    units = wl.units_per_step * args.steps * (1 if strong else world)
    progress("timed region done; parity check")
    parity = wl.parity_check() if large else None                           # every rank takes part (outside the timed region)
    progress("dominant kernel")
    roof = wl.dominant_kernel(peaks) if rank == 0 else None
    weak_line = None
    if strong:
        # round 1's form for comparison: every rank its own 2^20-transition shard, global minibatch world x 65,536
        del wl
        torch.cuda.empty_cache()
        ww = WORKLOADS[args.workload](seed=1 + rank, weak=True, tc_mode=args.tc_mode)
        ww.setup_gpu()
        wms, wms_e2e = measure(ww, max(2, args.steps // 2), 3, world, rank, flush)
        k = max(2, args.steps // 2)
        weak_line = dict(scaling="weak", value=ww.units_per_step * k * world / (wms * 1e-3), unit=ww.unit,
                         ms_per_step=wms / k, e2e=ww.units_per_step * k * world / (wms_e2e * 1e-3),
                         note="every rank its own 2^20-transition shard, global minibatch = %d x 65,536 rows" % world)
        wl = ww
    if rank == 0:
        arith = {"exact": "fp32 results (parity rtol 1e-5): minibatch gradients >= 8,192 rows run on tcgen05 as 3 exact bf16 "
                          "planes per operand (6 plane pairs), fp32 accumulation in TMEM; the rest is fp32 FFMA",
                 "fp16x2": "fp32 results (parity rtol 1e-5): minibatch gradients >= 8,192 rows run on tcgen05 as 2 fp16 planes per "
                           "operand, the low plane scaled by 2^11 (x = x0 + 2^-11 x1 carries fp32 to 2^-24 for |x| < 65504; 3 plane "
                           "pairs), fp32 accumulation in TMEM; the rest is fp32 FFMA",
                 "bf16x2": "minibatch gradients >= 8,192 rows on tcgen05 with 2 bf16 planes per operand (3 plane pairs; stated "
                           "tolerance 3e-5 max|grad|), fp32 accumulation; the rest is fp32 FFMA",
                 "bf16": "minibatch gradients >= 8,192 rows on tcgen05 with plain bf16 operands (stated tolerance 1e-1 "
                         "max|grad|), fp32 accumulation; the rest is fp32 FFMA"}[args.tc_mode]
        line = dict(metric=wl.metric, value=units / (ms * 1e-3), unit=wl.unit, n_gpus=world, steps=args.steps,
                    warmup=warmup, ms_per_step=ms / args.steps, higher_is_better=True,
                    scaling="strong" if strong else "weak",
                    vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload=wl.name,
                                l2="flushed (256 MB write) between timed steps; per-step inputs exceed L2",
                                adam_state="f32", tc_mode=args.tc_mode if large else None,
                                arithmetic=arith if large else "fp32",
                                parallelism=PARALLELISM[args.workload] % ((world, world) if
                                args.workload == "ppo_large" else (world,)) if world > 1 else "single GPU"),
                    e2e=dict(value=units / (ms_e2e * 1e-3), unit=wl.unit, h2d_bytes_per_step=wl.h2d,
                             d2h_bytes_per_step=wl.d2h, ms_per_step=ms_e2e / args.steps,
                             api="PPOUpdater.update(prefetch(pinned rollout), val_idx, pol_idx, out=pinned) -- the public "
                                 "host-side call" if large else "public host-side call of the workload",
                             index_tables="uploaded once per run (%d bytes), resident between updates"
                                          % getattr(wl, "idx_bytes_once", 0) if large else None),
                    gpu_launches=wl.launches_per_step() * args.steps, roofline=roof, clocks=clocks)
        if parity is not None:
            line["parity_check"] = parity
        if weak_line is not None:
            line["weak_scaling"] = weak_line
        if not args.no_cpu_baseline and world == 1 and getattr(wl, "step_cpu", None):
            progress("cpu baseline (three figures)")
            three = cpu_three_figures(args.workload, cores)
            line["cpu_baseline"] = dict(
                value=three["replicas_aggregate"], unit=wl.unit, cores=three["replicas"], kind="port", host=host_info(),
                figures=three,
                sample="BASELINE.md 3.3's three figures, each 1 step after 1 warm-up: one process with NumPy / OpenBLAS default "
                       "threads, one process with one BLAS thread, and `cores` single-thread replicas in parallel (value: how "
                       "the reference uses a multi-core box, run_benchmarks.sh:98-102); Tanh nets float32, Adam float64 as in "
                       "the reference; each step: " + getattr(wl, "cpu_sample_desc", "the full workload step"))
        else:
            line["cpu_baseline"] = None
        if large and world == 1 and not args.no_extras:
            progress("precision modes")
            line["modes"] = wl.precision_modes()
        if not args.no_extras and world == 1:
            del wl
            torch.cuda.empty_cache()
            others = {}
            for name in ("closed_loop", "ppo", "mlp", "es", "vae"):
                if name == args.workload:
                    continue
                progress("other workload: %s" % name)
                w2 = WORKLOADS[name](seed=7)
                w2.setup_gpu()
                m1, m2 = measure(w2, 5, 3, 1, 0, flush)
                u = w2.units_per_step * 5
                r2 = w2.dominant_kernel(peaks)
                c1 = time_cpu(name, 1, 1, 1, threads=None)[0] if getattr(w2, "step_cpu", None) else None
                reps = [w2.replica_throughput(r) for r in (8, 16)] if hasattr(w2, "replica_throughput") else None
                extra = {"sweep": w2.sweep(peaks)} if hasattr(w2, "sweep") else {}
                others[name] = dict(extra, workload=w2.name, metric=w2.metric, unit=w2.unit, value=u / (m1 * 1e-3),
                                    e2e=u / (m2 * 1e-3), ms_per_step=m1 / 5, cpu_port_1proc=c1,
                                    independent_replicas=reps, roofline={k: r2[k] for k in r2 if k in ("kernel", "achieved", "frac", "ms_per_launch",
                                                                             "fp32_tflops", "fp32_frac", "note", "collect_ms",
                                                                             "collect_transitions_per_s",
                                                                             "mean_reward_first_last")})
                del w2
                torch.cuda.empty_cache()
            line["other_workloads"] = others
            progress("kernel sweep")
            line["kernels"] = kernel_sweep(peaks)
            progress("done")
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
