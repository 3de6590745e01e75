"""Fused PPO update: one chunk of benchmarks/ppo.py:22-40 + benchmarks/gae.py:34-75.

Everything the reference does per 2048-step chunk in ~1,300 NumPy calls runs as a
handful of launches, with the two 300-step optimiser chains (value predictor,
policy) each inside ONE persistent kernel (csrc/mq_fused.cu: mq_fused_train) and
running concurrently on two CUDA streams:

    value  = V(obs[0..N])                       mq_fused_forward        gae.py:43-45
    adv,ret= GAE(rew, value, done)              mq_gae                  gae.py:48-56,69
    300 x value step on (obs[idx], ret[idx])    mq_fused_train (DIFF)   gae.py:71-73,20-23
    adv_n  = tanh(RunningNormalize_h2(adv))     mq_col_mean/cov/apply   ppo.py:24-25
    logp0  = logprob(obs, act)                  mq_fused_forward        ppo.py:27
    300 x PPO step on minibatch idx             mq_fused_train (PPO)    ppo.py:29-40

Random draws (minibatch index tables) are inputs.  With a process
group (mannequin2_b200.dist) large minibatches are sharded by rows and the flat
gradient is all-reduced once per optimiser step.
"""
import ctypes
import os

import numpy as np
import torch

from . import _device as D
from . import _lib
from . import dist as mqdist
from ._adam import Adam
from ._lib import ACT_NONE, ACT_TANH, HEAD_DIFF, HEAD_DISCRETE_PPO, HEAD_GAUSS_PPO, MQ_F32, TC_MODE_BITS, TC_MODES, MqHead, MqTrain
from ._normalize import RunningNormalize
from .basicnet import normed_columns


def init_mlp(n_in, widths, head_scale):
    """Flat float32 parameters, same draws as Affine() stacks (basicnet.py:265-267,300-311)."""
    parts, k = [], n_in
    for i, n in enumerate(widths):
        w = normed_columns(k, n)
        if i == len(widths) - 1:
            w = w * np.float32(head_scale)
        parts += [w.reshape(-1).astype(np.float32), np.zeros(n, np.float32)]
        k = n
    return np.concatenate(parts)


CHAIN_MAX_MB = 256        # minibatches up to this size run inside the persistent single-CTA chain


class _Net(object):
    """Parameters + Adam state of one net, all device resident."""

    def __init__(self, net, theta_host, horizon, state_dtype, tc_mode="exact"):
        self.net = net
        self.theta = D.from_host(theta_host, np.float32)
        self.opt = Adam(theta_host, horizon=horizon, state_dtype=state_dtype)
        self.code = self.opt._code
        self.grad = D.zeros(net.n_params)
        self._peer = None
        # tensor-core path of the large-minibatch steps (csrc/mq_tc3.cu): precision mode, packed records, operand image
        self.planes = TC_MODES[tc_mode]
        img_bytes = int(_lib.lib().mq_tc_image_bytes(ctypes.byref(net), self.planes))
        self.image = torch.zeros(img_bytes, dtype=torch.uint8, device=D.device()) if img_bytes else None
        self.rec_w = int(_lib.lib().mq_record_width(ctypes.byref(net)))
        self.records = None

    def pack(self, x, tgt, w0, w1, rows):
        """One aligned record per transition (x | target | w0 | w1): the four gathers of benchmarks/ppo.py:30-37 become one."""
        if self.image is None:
            self.records = None
            return
        if getattr(self, "_rec_buf", None) is None or self._rec_buf.numel() < rows * self.rec_w:
            self._rec_buf = torch.empty(rows * self.rec_w, dtype=torch.float32, device=D.device())
        self.records = self._rec_buf[:rows * self.rec_w]
        D.call("mq_pack_records", ctypes.byref(self.net), D.ptr(x), D.ptr(tgt) if tgt is not None else None,
               D.ptr(w0) if w0 is not None else None, D.ptr(w1) if w1 is not None else None, rows, D.ptr(self.records),
               D.stream())

    def forward_records(self, x, sample, rows, out, logp):
        """Forward pass over `rows` rows on the record-fed tensor-core kernel (forward only): packs x | sample, refreshes the
        operand image from the current parameters.  False when the net does not fit that kernel (caller: array-fed forward)."""
        if self.image is None or rows < 8192:
            return False
        self.pack(x, sample, None, None, rows)
        D.call("mq_tc_image_build", ctypes.byref(self.net), D.ptr(self.theta), self.planes, D.ptr(self.image), D.stream())
        try:
            D.call("mq_fused_forward_rec", ctypes.byref(self.net), D.ptr(self.theta), D.ptr(self.records), self.rec_w,
                   D.ptr(self.image), None, rows, D.ptr(out) if out is not None else None,
                   D.ptr(logp) if logp is not None else None, self.planes, D.stream())
        except NotImplementedError:
            return False
        return True

    def train(self, x, head, idx_table, n_steps, mb, lr, eps=1e-8, col0=0, idx_stride=None):
        """n_steps optimiser steps on the mb rows idx_table[s, col0 : col0 + mb] (i32[n_steps, idx_stride]): the whole
        table on one GPU, this rank's columns of the global minibatch when sharded."""
        idx_stride = mb if idx_stride is None else idx_stride
        if mb <= CHAIN_MAX_MB and mqdist.world() == 1:
            assert col0 == 0 and idx_stride == mb
            return self._train_chain(x, head, idx_table, n_steps, mb, lr, eps)
        return self._train_loop(x, head, idx_table, n_steps, mb, lr, eps, col0, idx_stride)

    # -- small minibatches: one persistent CTA runs the whole dependent chain ------------
    def _train_chain(self, x, head, idx_table, n_steps, mb, lr, eps):
        value, m, v, _ = self.opt.state()
        tw = (ctypes.c_double * 2)(*self.opt._tw)
        D.call("mq_fused_train", ctypes.byref(self.net), D.ptr(self.theta), D.ptr(value), D.ptr(m), D.ptr(v),
               self.code, D.ptr(x), ctypes.byref(head), D.ptr(idx_table), n_steps, mb, self.opt._b1,
               self.opt._b2, tw, lr, eps, D.stream())
        self.opt._tw = [tw[0], tw[1]]
        self.opt._out = None

    # -- large minibatches: all SMs per step, one gradient exchange across GPUs ----------------
    def _train_loop(self, x, head, idx_table, n_steps, mb, lr, eps, col0=0, idx_stride=None):
        idx_stride = mb if idx_stride is None else idx_stride
        lib = _lib.lib()
        wsb = lib.mq_fused_ws_bytes(ctypes.byref(self.net), mb)
        ws = D.workspace(("fused", id(self)), wsb)
        scale = 1.0 / float(mb * mqdist.world())            # batch MEAN over the GLOBAL minibatch
        coefs = (ctypes.c_double * (2 * n_steps))()
        for s in range(n_steps):
            coefs[2 * s], coefs[2 * s + 1] = self.opt.next_coefs()
        self.opt._out = None
        value, m, v, _ = self.opt.state()
        if self._peer is None:
            self._peer = mqdist.PeerExchange(self.net.n_params)
        head.tc = (head.tc & ~TC_MODE_BITS) | self.planes
        if self.image is not None and self.records is not None and mb >= 8192:
            # record-fed tensor-core kernel: its operand image is built once here and then follows theta step by step
            head.records, head.rec_w, head.tc_image = D.ptr(self.records), self.rec_w, D.ptr(self.image)
            D.call("mq_tc_image_build", ctypes.byref(self.net), D.ptr(self.theta), self.planes, D.ptr(self.image), D.stream())
        else:
            head.records, head.rec_w, head.tc_image = None, 0, None
            head.tc = head.tc & ~TC_MODE_BITS
        if mqdist.world() == 1 or self._peer.available():
            # per step 2 launches: gradient kernel (per-CTA partials), then ONE kernel that sums them in fixed order,
            # all-reduces over NVLink peer memory when sharded, applies Adam and refreshes the operand image
            # (csrc/mq_step.cu); the loop over steps runs inside the library
            peer = None
            if mqdist.world() > 1:
                peer = ctypes.byref(self._peer.descriptor())
                self._peer.step += n_steps - 1
            D.call("mq_train_steps", ctypes.byref(self.net), D.ptr(self.theta), D.ptr(x), ctypes.byref(head),
                   D.ptr(idx_table) + 4 * col0, idx_stride, n_steps, mb, scale, D.ptr(value), D.ptr(m), D.ptr(v), self.code, coefs, lr, eps, 0, peer, D.ptr(ws), wsb,
                   D.stream())
            return
        # fallback (no peer mapping): gradient, NCCL / gloo all-reduce, Adam
        head.tc_image = None
        base, stride = D.ptr(idx_table) + 4 * col0, idx_stride * 4
        for s in range(n_steps):
            D.call("mq_fused_grad", ctypes.byref(self.net), D.ptr(self.theta), D.ptr(x), base + s * stride, mb,
                   ctypes.byref(head), scale, D.ptr(self.grad), None, None, D.ptr(ws), wsb, D.stream())
            mqdist.allreduce_sum_(self.grad)
            D.call("mq_adam_step", D.ptr(value), D.ptr(m), D.ptr(v), D.ptr(self.theta), D.ptr(self.grad),
                   self.net.n_params, self.code, MQ_F32, coefs[2 * s], coefs[2 * s + 1], lr, eps, 0, D.stream())

    def prepare_pair(self, x, head, idx_table, n_steps, mb, lr, eps, col0, idx_stride):
        """This net's side of a paired training loop (mq_train_steps2); returns (MqTrain, keep-alive objects) or None when
        the net cannot take the record-fed tensor-core path."""
        if self.image is None or self.records is None or mb < 8192:
            return None
        if self._peer is None:
            self._peer = mqdist.PeerExchange(self.net.n_params)
        if mqdist.world() > 1 and not self._peer.available():
            return None
        lib = _lib.lib()
        wsb = lib.mq_fused_ws_bytes(ctypes.byref(self.net), mb)
        ws = D.workspace(("fused", id(self)), wsb)
        coefs = (ctypes.c_double * (2 * n_steps))()
        for s in range(n_steps):
            coefs[2 * s], coefs[2 * s + 1] = self.opt.next_coefs()
        self.opt._out = None
        value, m, v, _ = self.opt.state()
        head.tc = (head.tc & ~TC_MODE_BITS) | self.planes
        head.records, head.rec_w, head.tc_image = D.ptr(self.records), self.rec_w, D.ptr(self.image)
        D.call("mq_tc_image_build", ctypes.byref(self.net), D.ptr(self.theta), self.planes, D.ptr(self.image), D.stream())
        t = MqTrain()
        t.net, t.theta, t.x, t.head = ctypes.addressof(self.net), D.ptr(self.theta), D.ptr(x), ctypes.addressof(head)
        t.idx_table, t.idx_stride = D.ptr(idx_table) + 4 * col0, idx_stride
        t.scale = 1.0 / float(mb * mqdist.world())
        t.value, t.m, t.v = D.ptr(value), D.ptr(m), D.ptr(v)
        t.coefs_host, t.lr, t.eps = ctypes.addressof(coefs), lr, eps
        keep = [coefs, head, ws]
        if mqdist.world() > 1:
            peer = self._peer.descriptor()
            self._peer.step += n_steps - 1
            t.peer = ctypes.addressof(peer)
            keep.append(peer)
        t.ws, t.ws_bytes = D.ptr(ws), wsb
        return t, keep

    def exchange(self):
        """Which gradient exchange the sharded steps use: "ll" | "push" | "pull" (in-kernel, peer memory) or "nccl"."""
        if mqdist.world() == 1:
            return "none"
        if self._peer is None:
            self._peer = mqdist.PeerExchange(self.net.n_params)
        return {2: "ll", 0: "push", 1: "pull"}[self._peer.pull] if self._peer.available() else "nccl"


PAIR_MAX_ROWS = 16384          # per-GPU minibatch rows up to which the value and the policy step share launches


class PPOUpdater(object):
    def __init__(self, n_obs, n_act, *, hid=64, n_hid=2, gam=0.99, lam=0.95, pol_lr=3e-4, val_lr=3e-3,
                 horizon=10, clip=(0.8, 1.2), norm_horizon=2, state_dtype="float32",
                 policy_params=None, value_params=None, discrete=False, ent_coef=0.0, tc_mode="exact"):
        widths = [hid] * n_hid
        acts = [ACT_TANH] * n_hid + [ACT_NONE]
        # discrete=True: categorical policy over n_act actions (benchmarks/policy.py:15-17, distrib.Discrete);
        # `act` is then an int32 vector and travels to the kernels as one-hot rows
        self.discrete = bool(discrete)
        # ent_coef (extension): adds ent_coef * Gauss.entropy to the objective; 0 = the reference's rule
        self.ent_coef = float(ent_coef)
        assert self.ent_coef == 0.0 or not self.discrete
        pnet = _lib.make_net(n_obs, widths + [n_act], acts, logstd=not self.discrete)
        vnet = _lib.make_net(n_obs, widths + [1], acts)
        if policy_params is None:                                   # policy.py:7-27, Gauss logstd zeros
            policy_params = init_mlp(n_obs, widths + [n_act], 0.1)
            if not self.discrete:
                policy_params = np.concatenate([policy_params, np.zeros(n_act, np.float32)])
        if value_params is None:                                    # gae.py:12-18
            value_params = init_mlp(n_obs, widths + [1], 0.1)
        assert len(policy_params) == pnet.n_params and len(value_params) == vnet.n_params
        self.n_obs, self.n_act = n_obs, n_act
        self.gam, self.lam, self.pol_lr, self.val_lr, self.clip = gam, lam, pol_lr, val_lr, clip
        # tc_mode: precision of the tensor-core gradient kernel that serves minibatches >= 8,192 rows: "exact" (three bf16
        # planes, the fp32 bar of the FFMA kernels), "fp16x2" (two fp16 planes, the low one scaled: the same bar at the two-plane cost, operands must stay below 65504),
        # "bf16x2" (|err| <= 3e-5 max|grad|) or "bf16" (<= 2e-2): north_star's
        # "stated looser tolerance for bf16 tensor-core paths".  Smaller minibatches always run the fp32 FFMA kernels.
        assert tc_mode in TC_MODES, tc_mode
        self.tc_mode = tc_mode
        self.policy = _Net(pnet, np.asarray(policy_params, np.float32), horizon, state_dtype, tc_mode)
        self.value = _Net(vnet, np.asarray(value_params, np.float32), horizon, state_dtype, tc_mode)
        self.normalize = RunningNormalize(horizon=norm_horizon)
        # update() validates user-supplied minibatch indices on the host (an out-of-range index would read out of bounds
        # on the device); throughput loops that reuse known-good tables switch it off
        self.check_indices = True
        self._side = torch.cuda.Stream()
        self._ev_gae, self._ev_val = torch.cuda.Event(), torch.cuda.Event()
        self._bufs = {}

    def _buf(self, name, n, dtype=torch.float32):
        t = self._bufs.get(name)
        if t is None or t.numel() != n or t.dtype != dtype:
            t = torch.empty(n, dtype=dtype, device=D.device())
            self._bufs[name] = t
        return t

    def _forward_rows(self, net, obs, n_rows, out, sample=None, logp=None, split=False):
        """Forward pass over n_rows rows; split=True: this rank computes its slice and the slices are all-gathered
        (every rank ends up with the full vector, bit-identical everywhere)."""
        if not split or mqdist.world() == 1:
            if net.forward_records(obs, sample, n_rows, out, logp):
                return
            D.call("mq_fused_forward", ctypes.byref(net.net), D.ptr(net.theta), D.ptr(obs), None, n_rows,
                   D.ptr(out) if out is not None else None, D.ptr(sample) if sample is not None else None,
                   D.ptr(logp) if logp is not None else None, D.stream())
            return
        w, r = mqdist.world(), mqdist.rank()
        chunk = -(-n_rows // w)
        lo = min(r * chunk, n_rows)
        cnt = min(chunk, n_rows - lo)
        dst = out if out is not None else logp
        part = self._buf("fwd_part", chunk)
        full = self._buf("fwd_full", chunk * w)
        if cnt > 0:
            obs_r = obs.reshape(-1)[lo * self.n_obs:(lo + cnt) * self.n_obs]
            smp_r = sample.reshape(-1)[lo * self.n_act:(lo + cnt) * self.n_act] if sample is not None else None
            if not net.forward_records(obs_r, smp_r, cnt, part if out is not None else None, part if logp is not None else None):
                D.call("mq_fused_forward", ctypes.byref(net.net), D.ptr(net.theta), D.ptr(obs_r), None, cnt,
                       D.ptr(part) if out is not None else None, D.ptr(smp_r) if smp_r is not None else None,
                       D.ptr(part) if logp is not None else None, D.stream())
        mqdist.allgather_into(full, part)
        dst[:n_rows].copy_(full[:n_rows])

    def update_dev(self, obs, act, rew, done, val_idx, pol_idx, replicated=False):
        """All arguments are device tensors: obs f32[N+1,D], act f32[N,A], rew f32[N],
        done u8[N], val_idx / pol_idx i32[steps, mb].  Returns device (adv, ret).

        With a process group there are two data-parallel forms:
          replicated=False  every rank passes ITS OWN rollout shard and index tables into it; the global minibatch is
                            world x mb rows (weak scaling: the problem grows with the number of GPUs);
          replicated=True   every rank passes the SAME rollout and the SAME tables of GLOBAL minibatches (BASELINE
                            configs[2]: one 1M-transition rollout, 65,536-row minibatches sharded across the GPUs); rank r
                            takes columns [r mb/world, (r+1) mb/world) of every minibatch, the two forward passes over the
                            rollout are split by rows and all-gathered, GAE and the normaliser run replicated."""
        lib = _lib.lib()
        n = rew.numel()
        assert obs.numel() == (n + 1) * self.n_obs
        if self.discrete:
            assert act.numel() == n and act.dtype == torch.int32
            onehot = self._buf("onehot", n * self.n_act)
            D.call("mq_onehot", D.ptr(act), n, self.n_act, D.ptr(onehot), D.stream())
            act = onehot
        assert act.numel() == n * self.n_act
        world, rank = mqdist.world(), mqdist.rank()
        split = bool(replicated) and world > 1
        v_steps, v_mb, p_steps, p_mb = val_idx.shape[0], val_idx.shape[1], pol_idx.shape[0], pol_idx.shape[1]
        v_col0 = p_col0 = 0
        v_stride, p_stride = v_mb, p_mb
        if split:
            assert v_mb % world == 0 and p_mb % world == 0, "global minibatch must split evenly over the ranks"
            v_mb, p_mb = v_mb // world, p_mb // world
            v_col0, p_col0 = rank * v_mb, rank * p_mb
        pol, val = self.policy, self.value
        value = self._buf("value", n + 1)
        adv, ret = self._buf("adv", n), self._buf("ret", n)
        adv_n, logp0 = self._buf("adv_n", n), self._buf("logp0", n)
        main = torch.cuda.current_stream()
        # value predictions for every state incl. the bootstrap state (gae.py:43-45)
        self._forward_rows(val, obs, n + 1, value, split=split)
        wsb = lib.mq_scan_ws_bytes(n)
        ws = D.workspace("scan", wsb)
        D.call("mq_gae", D.ptr(rew), D.ptr(value), D.ptr(done), n, self.gam, self.lam, D.ptr(adv), D.ptr(ret),
               D.ptr(ws), wsb, D.stream())
        # The value chain and the policy chain are independent: run them on two streams.  Small minibatches: the two
        # 300-step chains are single-CTA kernels side by side.  Large minibatches: each gradient kernel fills the GPU,
        # but the small reduce / Adam kernels and the launch gaps of one chain run under the gradient kernels of the other.
        # MQ_PPO_SERIAL=1 forces one stream.
        # Sharded (world > 1) the two chains share ONE stream: the step kernels spin on peer memory, and a spinning grid of one
        # chain can keep the other chain's gradient kernel (52 K registers per CTA) from becoming resident; with the chains
        # interleaved differently on two ranks neither would progress.  One stream = one fixed order on every rank.
        concurrent = not os.environ.get("MQ_PPO_SERIAL") and world == 1
        vhead = MqHead()
        vhead.kind, vhead.target = HEAD_DIFF, D.ptr(ret)
        head = MqHead()
        head.kind, head.p0, head.p1 = (HEAD_DISCRETE_PPO if self.discrete else HEAD_GAUSS_PPO), self.clip[0], self.clip[1]
        head.ent = self.ent_coef
        head.target, head.adv, head.logp_old = D.ptr(act), D.ptr(adv_n), D.ptr(logp0)
        # Large minibatches on the same schedule (BASELINE configs[2]): the two nets' steps are PAIRED -- per step one
        # gradient launch with each net on half of the SMs and one reduce / exchange / Adam launch for both (csrc/mq_tc3.cu,
        # csrc/mq_step.cu) -- in one stream, hence in one fixed order on every rank.  MQ_PPO_UNPAIRED=1: the two chains apart.
        # Pairing pays while a net has about one tile per SM (a launch is then one tile's chain of dependent phases long
        # whatever the row count: 5.2 against 8.1 ms per update at 16,384 rows per GPU) and whenever the two chains have to
        # share a stream anyway (sharded: 7.0 against 9.4 ms at 32,768 rows per GPU).  On ONE GPU with several tiles per SM the
        # two chains side by side on two streams, each kernel on all SMs, are faster (12.2 against 13.1 ms at 65,536 rows).
        # MQ_PPO_PAIRED=1 / MQ_PPO_UNPAIRED=1 force either form.
        paired = (v_mb >= 8192 and v_mb == p_mb and v_steps == p_steps and val.code == pol.code and
                  (v_mb <= PAIR_MAX_ROWS or world > 1 or os.environ.get("MQ_PPO_PAIRED")) and not os.environ.get("MQ_PPO_UNPAIRED"))
        big = v_mb >= 8192 or p_mb >= 8192
        if big:
            val.pack(obs, ret, None, None, n)
        if paired:
            self.normalize.update_dev(adv, n, sharded=world > 1 and not split)
            self.normalize.apply_dev(adv, n, out=adv_n, fuse_tanh=True)
            self._forward_rows(pol, obs, n, None, sample=act, logp=logp0, split=split)
            pol.pack(obs, act, adv_n, logp0, n)
            ta = val.prepare_pair(obs, vhead, val_idx, v_steps, v_mb, self.val_lr, 1e-8, v_col0, v_stride)
            tb = pol.prepare_pair(obs, head, pol_idx, p_steps, p_mb, self.pol_lr, 1e-8, p_col0, p_stride) if ta else None
            if ta and tb:
                D.call("mq_train_steps2", ctypes.byref(ta[0]), ctypes.byref(tb[0]), v_steps, v_mb, val.code, 0, D.stream())
                return adv, ret
            assert ta is None and tb is None, "only one net of a pair can take the paired path"
            # (neither took it: fall through to the separate chains; the normaliser and logp0 above are already done)
            val.train(obs, vhead, val_idx, v_steps, v_mb, self.val_lr, col0=v_col0, idx_stride=v_stride)
            pol.train(obs, head, pol_idx, p_steps, p_mb, self.pol_lr, col0=p_col0, idx_stride=p_stride)
            return adv, ret
        if concurrent:
            self._ev_gae.record(main)
            with torch.cuda.stream(self._side):              # value predictor chain (gae.py:66-73)
                self._side.wait_event(self._ev_gae)
                val.train(obs, vhead, val_idx, v_steps, v_mb, self.val_lr, col0=v_col0, idx_stride=v_stride)
                self._ev_val.record(self._side)
        else:
            val.train(obs, vhead, val_idx, v_steps, v_mb, self.val_lr, col0=v_col0, idx_stride=v_stride)
        # policy chain on the main stream (ppo.py:24-40)
        self.normalize.update_dev(adv, n, sharded=world > 1 and not split)
        self.normalize.apply_dev(adv, n, out=adv_n, fuse_tanh=True)
        self._forward_rows(pol, obs, n, None, sample=act, logp=logp0, split=split)
        if big:
            pol.pack(obs, act, adv_n, logp0, n)
        pol.train(obs, head, pol_idx, p_steps, p_mb, self.pol_lr, col0=p_col0, idx_stride=p_stride)
        if concurrent:
            main.wait_event(self._ev_val)
        return adv, ret

    def update_rollout(self, chunk, n_seg, seg_len, val_idx, pol_idx):
        """One update from a chunk collected by rollout.BatchedActor: n_seg per-environment segments of seg_len
        transitions laid end to end.  Each segment ends like a chunk of benchmarks/gae.py:43-64 (the advantage beyond
        the chunk is 0, the value of the state after it is bootstrapped): mq_bootstrap_segments turns that into the
        terminal form, then the update is update_dev unchanged."""
        n = n_seg * seg_len
        assert chunk["rew"].numel() == n and chunk["last_obs"].numel() == n_seg * self.n_obs
        v_next = self._buf("v_next", n_seg)
        D.call("mq_fused_forward", ctypes.byref(self.value.net), D.ptr(self.value.theta), D.ptr(chunk["last_obs"]), None,
               n_seg, D.ptr(v_next), None, None, D.stream())
        rew, done = self._buf("rew_seg", n), self._buf("done_seg", n, torch.uint8)
        rew.copy_(chunk["rew"])
        done.copy_(chunk["done"])
        D.call("mq_bootstrap_segments", D.ptr(rew), D.ptr(done), D.ptr(v_next), n_seg, seg_len, self.gam, D.stream())
        return self.update_dev(chunk["obs"], chunk["act"], rew, done, val_idx, pol_idx)

    # ---- public host-side call ------------------------------------------------------------------------------------
    def _to_dev(self, a, dtype):
        """Host array (NumPy, pageable) or host / device torch tensor -> device tensor.  Pinned torch tensors are copied
        asynchronously on the current stream; device tensors pass through."""
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                return a
            return a.to(D.device(), non_blocking=True)
        return D.from_host(np.ascontiguousarray(a, dtype=dtype))

    def _upload_split(self, t, dtype):
        """Replicated data-parallel form: this rank uploads 1/world of the rows of a host tensor and the ranks all-gather the
        pieces over NVLink, instead of every rank pulling the whole rollout through its own PCIe link."""
        w, r = mqdist.world(), mqdist.rank()
        if not isinstance(t, torch.Tensor):
            t = torch.from_numpy(np.ascontiguousarray(t, dtype=dtype))
        if t.is_cuda:
            return t
        rows = t.shape[0]
        rowlen = t.numel() // max(rows, 1)
        chunk = -(-rows // w)
        lo = min(r * chunk, rows)
        cnt = min(chunk, rows - lo)
        part = torch.zeros(chunk * rowlen, dtype=t.dtype, device=D.device())
        if cnt > 0:
            part[:cnt * rowlen].copy_(t.reshape(-1)[lo * rowlen:(lo + cnt) * rowlen], non_blocking=True)
        full = torch.empty(chunk * w * rowlen, dtype=t.dtype, device=D.device())
        mqdist.allgather_into(full, part)
        return full[:rows * rowlen].view(t.shape)

    def prefetch(self, obs, act, rew, done, replicated=False):
        """Start uploading the NEXT rollout (pinned host tensors) on a copy stream while the current update runs; the
        returned handle is passed to update() in place of the four arrays.  replicated=True (every rank holds the same
        rollout): each rank uploads its share of the rows and the shares are all-gathered."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        self._copy_stream.wait_stream(main)
        split = bool(replicated) and mqdist.world() > 1
        with torch.cuda.stream(self._copy_stream):
            dts = (np.float32, np.int32 if self.discrete else np.float32, np.float32, np.uint8)
            if split:
                dev = [self._upload_split(a, dt) for a, dt in zip((obs, act, rew, done), dts)]
            else:
                dev = [self._to_dev(a, dt) for a, dt in zip((obs, act, rew, done), dts)]
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        return ("prefetched", dev, ev)

    def update(self, obs, act=None, rew=None, done=None, val_idx=None, pol_idx=None, *, replicated=False, out=None):
        """Host arrays in, new (policy_params, value_params) host arrays out -- the end-to-end call of one iteration of
        benchmarks/ppo.py:22-40.  obs/act/rew/done: NumPy arrays, pinned torch tensors, or the handle of prefetch();
        val_idx / pol_idx: NumPy arrays or device tensors (index tables that stay resident between updates).
        out = (policy, value) pinned host tensors to receive the parameters (optional)."""
        main = torch.cuda.current_stream()
        if isinstance(obs, tuple) and obs and obs[0] == "prefetched":
            _, dev, ev = obs
            main.wait_event(ev)
            for a in dev:
                a.record_stream(main)
            d_obs, d_act, d_rew, d_done = dev
        else:
            d_obs = self._to_dev(obs, np.float32)
            d_act = self._to_dev(act, np.int32 if self.discrete else np.float32)
            d_rew = self._to_dev(rew, np.float32)
            d_done = self._to_dev(done if isinstance(done, torch.Tensor) else np.asarray(done, dtype=np.uint8), np.uint8)
        d_vi, d_pi = self._to_dev(val_idx, np.int32), self._to_dev(pol_idx, np.int32)
        if self.check_indices:
            n = d_rew.numel()
            for t in (d_vi, d_pi):
                lo, hi = int(t.min()), int(t.max())
                if lo < 0 or hi >= n:
                    raise ValueError("minibatch index out of range: [%d, %d] for %d transitions" % (lo, hi, n))
        self.update_dev(d_obs, d_act, d_rew, d_done, d_vi, d_pi, replicated=replicated)
        if out is not None:
            out[0].copy_(self.policy.theta, non_blocking=True)
            out[1].copy_(self.value.theta, non_blocking=True)
            main.synchronize()
            return out
        return D.to_host(self.policy.theta), D.to_host(self.value.theta)
