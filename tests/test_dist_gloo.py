"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: rank sharding, the gradient
all-reduce convention (each rank scales by 1/B_global, sum over ranks = global batch mean) and
the all-gather of ES returns.  Gradients here come from the oracle -- the point is the exchange
logic in mannequin2_b200/dist.py, not the kernels."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mannequin2_b200 import dist as mqdist
    from oracle import mq_oracle as O
    assert mqdist.world() == world and mqdist.rank() == rank
    rs = np.random.RandomState(0)
    spec = O.mlp_spec(11, [16, 3], [O.ACT_TANH, O.ACT_NONE])
    theta = O.mlp_init(spec, rs)
    x, dy = rs.randn(64, 11), rs.randn(64, 3)
    lo, hi = mqdist.shard_range(64, rank, world)
    outs = O.mlp_forward(theta, spec, x[lo:hi])
    # local batch-mean gradient times (local rows / global rows) == local SUM / B_global
    g_local = O.mlp_backward(theta, spec, x[lo:hi], outs, dy[lo:hi]).astype(np.float64) * (hi - lo) / 64.0
    g = torch.from_numpy(g_local.copy())
    mqdist.allreduce_sum_(g)
    full = O.mlp_backward(theta, spec, x, O.mlp_forward(theta, spec, x), dy)
    assert np.allclose(g.numpy(), full, rtol=1e-6, atol=1e-8)
    rets = torch.arange(lo, hi, dtype=torch.float32)
    assert torch.equal(mqdist.allgather_concat(rets), torch.arange(64, dtype=torch.float32))
    # identical Adam step on every rank after the reduction -> parameters stay replicated
    opt = O.AdamO(theta, horizon=10, lr=0.01)
    opt.apply_gradient(g.numpy())
    gathered = [torch.zeros(spec["n_params"], dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(opt.get_value().copy()))
    assert all(torch.equal(gathered[0], t) for t in gathered)
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_gradient_exchange(tmp_path):
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def test_single_process_is_a_noop():
    sys.path.insert(0, ROOT)
    from mannequin2_b200 import dist as mqdist
    t = torch.ones(3)
    assert mqdist.world() == 1 and mqdist.rank() == 0
    assert torch.equal(mqdist.allreduce_sum_(t), torch.ones(3))
    assert mqdist.allgather_concat(t) is t
