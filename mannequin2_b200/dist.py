"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink).

The only exchange on the hot path is ONE all-reduce (sum) of the flat float32
gradient per optimiser step: every rank holds the same
parameters and Adam state, computes the gradient of its rows of the global
minibatch (already scaled by 1/B_global), and applies the identical Adam step
after the reduction -- no broadcast.  The ES population shards by members with
an all-gather of the returns and an all-reduce of the weighted noise sum.
On CPU-only machines the same code runs over `gloo` for the host-logic tests.

On NVLink-connected GPUs the per-step exchange does not go through NCCL at all: `PeerExchange` maps one
small buffer per rank into every peer (torch symmetric memory: plumbing) and csrc/mq_step.cu does the
all-reduce inside the kernel that also sums the per-CTA partial gradients and applies Adam.
"""
import os

import torch
import torch.distributed as dist


def is_initialized():
    return dist.is_available() and dist.is_initialized()


def world():
    return dist.get_world_size() if is_initialized() else 1


def rank():
    return dist.get_rank() if is_initialized() else 0


def allreduce_sum_(t, group=None):
    """In-place sum over ranks; a no-op for a single process."""
    if world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allgather_concat(t, group=None):
    """Concatenate equally sized 1-D shards in rank order."""
    if world() == 1:
        return t
    out = torch.empty(world() * t.numel(), dtype=t.dtype, device=t.device)
    try:
        dist.all_gather_into_tensor(out, t.contiguous().reshape(-1), group=group)      # one call, no concatenation
        return out
    except (RuntimeError, NotImplementedError):                                        # backend without the fused form
        parts = [torch.empty_like(t) for _ in range(world())]
        dist.all_gather(parts, t, group=group)
        return torch.cat(parts)


def allgather_into(out, part, group=None):
    """out[r * len(part) : (r + 1) * len(part)] = rank r's part (equal sizes)."""
    if world() == 1:
        out.copy_(part)
        return out
    try:
        dist.all_gather_into_tensor(out, part.contiguous().reshape(-1), group=group)
    except (RuntimeError, NotImplementedError):                                        # backend without the fused form
        parts = [torch.empty_like(part) for _ in range(world())]
        dist.all_gather(parts, part, group=group)
        out.copy_(torch.cat(parts))
    return out


def shard_range(n_units, r=None, w=None):
    """Contiguous [lo, hi) slice of n_units owned by rank r of w (remainder to low ranks)."""
    r = rank() if r is None else r
    w = world() if w is None else w
    base, rem = divmod(int(n_units), int(w))
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


class PeerExchange(object):
    """Peer-mapped exchange buffer + flag pad for mq_reduce_step (include/mq_b200.h: mq_peer_t).

    One instance per optimised parameter vector.  `available()` is False on a single process, on
    non-CUDA backends (the gloo host-logic tests) or when symmetric memory cannot be set up; the
    callers then fall back to the NCCL all-reduce path (same numbers: both sum in rank order only up to
    NCCL's own reduction order, so the fallback is allclose, not bit-equal, to the fused exchange)."""

    def __init__(self, n_params):
        self.ok = False
        self.step = 0
        # default: the low-latency form (tagged 8-byte words, one one-way trip); MQ_PEER_MODE=push|pull select the
        # flag-based forms (measured against each other in profiles/r1_scaling.md)
        mode = os.environ.get("MQ_PEER_MODE", "ll")
        self.pull = {"pull": 1, "push": 0, "ll": 2}.get(mode, 2)
        if world() == 1 or not torch.cuda.is_available() or os.environ.get("MQ_NO_PEER_EXCHANGE"):
            return
        if dist.get_backend() != "nccl":
            return
        try:
            import torch.distributed._symmetric_memory as symm
            from . import _lib
            w = world()
            self.stride = (int(n_params) + 63) // 64 * 64
            nblk = int(_lib.lib().mq_step_blocks(int(n_params)))
            dev = torch.device("cuda", torch.cuda.current_device())
            # float[2][w][stride] for push / pull; the low-latency form uses 8-byte words: twice the bytes
            self.buf = symm.empty(2 * w * self.stride * 2, dtype=torch.float32, device=dev)
            self.flags = symm.empty(2 * w * nblk, dtype=torch.int32, device=dev)
            self.buf.zero_()
            self.flags.zero_()
            torch.cuda.synchronize()
            hb = symm.rendezvous(self.buf, dist.group.WORLD.group_name)
            hf = symm.rendezvous(self.flags, dist.group.WORLD.group_name)
            self._handles = (hb, hf)
            self.grad_ptrs = torch.tensor([int(q) for q in hb.buffer_ptrs], dtype=torch.int64, device=dev)
            self.flag_ptrs = torch.tensor([int(q) for q in hf.buffer_ptrs], dtype=torch.int64, device=dev)
            dist.barrier()                       # every rank's pads are zeroed and mapped before the first step
            torch.cuda.synchronize()
            self.ok = True
        except Exception as exc:                 # pragma: no cover - depends on the fabric
            import warnings
            warnings.warn("mannequin2_b200: peer-memory exchange unavailable (%r); using NCCL all-reduce" % (exc,))
            self.ok = False

    def available(self):
        return self.ok

    def descriptor(self):
        """mq_peer_t for the NEXT step (step counters advance identically on every rank)."""
        from . import _lib
        self.step += 1
        d = _lib.MqPeer()
        d.rank, d.world = rank(), world()
        d.grad_bufs, d.flag_bufs = self.grad_ptrs.data_ptr(), self.flag_ptrs.data_ptr()
        d.buf_stride, d.step = self.stride, self.step
        d.pull = self.pull
        return d
