"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink).

The only exchange on the hot path is ONE all-reduce (sum) of the flat float32
gradient per optimiser step: every rank holds the same
parameters and Adam state, computes the gradient of its rows of the global
minibatch (already scaled by 1/B_global), and applies the identical Adam step
after the reduction -- no broadcast.  The ES population shards by members with
an all-gather of the returns and an all-reduce of the weighted noise sum.
On CPU-only machines the same code runs over `gloo` for the host-logic tests.
"""
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
    out = [torch.empty_like(t) for _ in range(world())]
    dist.all_gather(out, t, group=group)
    return torch.cat(out)


def shard_range(n_units, r=None, w=None):
    """Contiguous [lo, hi) slice of n_units owned by rank r of w (remainder to low ranks)."""
    r = rank() if r is None else r
    w = world() if w is None else w
    base, rem = divmod(int(n_units), int(w))
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)
