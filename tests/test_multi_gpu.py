"""2-GPU NCCL parity: a row-sharded PPO update (one gradient all-reduce per step) must give the
same parameters as one GPU working on the concatenated rows.  Needs >= 2 CUDA devices."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _data(seed, n):
    sys.path.insert(0, ROOT)
    from bench import ppo_rollout
    rs = np.random.RandomState(seed)
    o, a, r, d = ppo_rollout(rs, n, p_done=1.0 / 100, cut=500)
    vi = np.stack([rs.permutation(n).astype(np.int32) for _ in range(3)]).reshape(6, n // 2)
    pi = np.stack([rs.permutation(n).astype(np.int32) for _ in range(3)]).reshape(6, n // 2)
    return o, a, r, d, vi, pi


def _worker(rank, world, port, out_dir, peer):
    sys.path.insert(0, ROOT)
    if not peer:
        os.environ["MQ_NO_PEER_EXCHANGE"] = "1"           # NCCL all-reduce + separate Adam launch
    else:
        os.environ["MQ_PEER_MODE"] = {True: "ll"}.get(peer, peer)      # ll (default): tagged 8-byte words; push / pull: flags
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    from mannequin2_b200 import _device as D
    from mannequin2_b200.ppo import PPOUpdater
    np.random.seed(0)
    upd = PPOUpdater(11, 3)
    shard = _data(10 + rank, 4096)
    upd.update_dev(*[D.from_host(x) for x in shard])
    torch.cuda.synchronize()
    assert (upd.policy._peer is not None and upd.policy._peer.available()) == bool(peer), "exchange path"
    np.save(os.path.join(out_dir, "pol%d.npy" % rank), D.to_host(upd.policy.theta))
    np.save(os.path.join(out_dir, "val%d.npy" % rank), D.to_host(upd.value.theta))
    dist.destroy_process_group()


@pytest.mark.parametrize("peer", [True, "pull", "push", False],
                         ids=["peer_memory_ll", "peer_memory_pull", "peer_memory_push", "nccl_allreduce"])
def test_sharded_update_matches_single_gpu(tmp_path, peer):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 2000 + {True: 7, "pull": 13, "push": 19, False: 0}[peer]
    mp.spawn(_worker, args=(2, port, str(tmp_path), peer), nprocs=2, join=True)
    pol = [np.load(os.path.join(tmp_path, "pol%d.npy" % r)) for r in range(2)]
    val = [np.load(os.path.join(tmp_path, "val%d.npy" % r)) for r in range(2)]
    assert np.array_equal(pol[0], pol[1]) and np.array_equal(val[0], val[1])     # replicas stay identical
    # single-GPU reference run of the same global problem is not identical in structure (GAE and the
    # normaliser see each shard separately), so compare against the oracle run on the two shards
    from oracle import mq_oracle as O
    rs = np.random.RandomState(0)
    from mannequin2_b200.ppo import init_mlp
    np.random.seed(0)
    pol_theta = np.concatenate([init_mlp(11, [64, 64, 3], 0.1), np.zeros(3, np.float32)])
    val_theta = init_mlp(11, [64, 64, 1], 0.1)
    pol_spec = O.policy_spec(11, 3)
    val_spec = O.mlp_spec(11, [64, 64, 1], [1, 1, 0])
    shards = [_data(10 + r, 4096) for r in range(2)]
    n = 4096
    vals = [O.mlp_forward(val_theta, val_spec, s[0])[-1].reshape(-1).astype(np.float32) for s in shards]
    gae = [O.gae_advantages(s[2], v, s[3].astype(bool)) for s, v in zip(shards, vals)]
    vopt = O.AdamO(val_theta, horizon=10, lr=0.003)
    for k in range(6):
        o = np.concatenate([s[0][:n][s[4][k]] for s in shards])
        t = np.concatenate([g[1][s[4][k]] for s, g in zip(shards, gae)]).reshape(-1, 1)
        val_theta = O.value_sgd_step(val_theta, val_spec, vopt, o, t)
    assert np.max(np.abs(val[0] - val_theta)) < 2e-5
    norm = O.RunningNormalizeO(horizon=2)
    adv_all = np.concatenate([g[0] for g in gae])
    norm.update(adv_all)
    adv_n = [np.tanh(norm.apply(g[0])).astype(np.float32) for g in gae]
    lp0 = [O.policy_logprob(pol_theta, pol_spec, s[0][:n], s[1])[0] for s in shards]
    popt = O.AdamO(pol_theta, horizon=10)
    for k in range(6):
        o = np.concatenate([s[0][:n][s[5][k]] for s in shards])
        a = np.concatenate([s[1][s[5][k]] for s in shards])
        ad = np.concatenate([x[s[5][k]] for s, x in zip(shards, adv_n)])
        l0 = np.concatenate([x[s[5][k]] for s, x in zip(shards, lp0)])
        pol_theta = O.ppo_policy_step(pol_theta, pol_spec, popt, o, a, ad, l0)
    assert np.max(np.abs(pol[0] - pol_theta)) < 2e-5


def _es_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    from mannequin2_b200 import _device as D
    from mannequin2_b200.es import ESPopulation
    noise, obs, coef, theta0 = _es_data()
    es = ESPopulation(11, 3, params=theta0.copy())
    m = noise.shape[1] // world
    for g in range(noise.shape[0]):                          # this rank's members of every generation
        es.generation(D.from_host(noise[g, rank * m:(rank + 1) * m]), D.from_host(obs[g]), D.from_host(coef))
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, "es%d.npy" % rank), es.get_params().astype(np.float32))
    dist.destroy_process_group()


def _es_data():
    from mannequin2_b200.ppo import init_mlp
    rs = np.random.RandomState(23)
    np.random.seed(6)
    theta0 = init_mlp(11, [64, 64, 3], 1.0)
    noise = (rs.randn(3, 48, len(theta0)) * 0.1).astype(np.float32)
    obs = np.clip(rs.randn(3, 120, 11), -5, 5).astype(np.float32)
    return noise, obs, rs.randn(3).astype(np.float32), theta0


def test_sharded_es_population_matches_oracle(tmp_path):
    """the ES population sharded over two GPUs (24 members each, no communication during evaluation, returns
    all-gathered, one all-reduce of the weighted noise sum) takes the same three Adam steps as the oracle on all 48."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_es_worker, args=(2, 29500 + os.getpid() % 2000 + 29, str(tmp_path)), nprocs=2, join=True)
    got = [np.load(os.path.join(tmp_path, "es%d.npy" % r)) for r in range(2)]
    assert np.array_equal(got[0], got[1])                     # replicas stay identical
    from oracle import mq_oracle as O
    noise, obs, coef, theta0 = _es_data()
    spec = O.mlp_spec(11, [64, 64, 3], [O.ACT_TANH, O.ACT_TANH, O.ACT_NONE])
    opt, norm = O.AdamO(theta0, horizon=10, lr=0.01), O.RunningNormalizeO(horizon=10)
    want = theta0.copy()
    for g in range(noise.shape[0]):
        want, _, _ = O.es_generation(want, spec, opt, norm, noise[g], obs[g], coef)
    assert np.max(np.abs(got[0] - theta0)) > 1e-3
    assert np.max(np.abs(got[0] - want)) < 2e-4, np.max(np.abs(got[0] - want))


def _strong_data(n, mb, epochs):
    sys.path.insert(0, ROOT)
    from bench import ppo_rollout
    rs = np.random.RandomState(91)
    o, a, r, d = ppo_rollout(rs, n, p_done=1.0 / 300, cut=700)
    vi = np.concatenate([rs.permutation(n).astype(np.int32).reshape(n // mb, mb) for _ in range(epochs)])
    pi = np.concatenate([rs.permutation(n).astype(np.int32).reshape(n // mb, mb) for _ in range(epochs)])
    return o, a, r, d, vi, pi


def _strong_worker(rank, world, port, out_dir, tc_mode):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    from mannequin2_b200 import _device as D
    from mannequin2_b200.ppo import PPOUpdater, init_mlp
    np.random.seed(5)
    pol0 = np.concatenate([init_mlp(11, [64, 64, 3], 0.1), np.zeros(3, np.float32)])
    val0 = init_mlp(11, [64, 64, 1], 0.1)
    upd = PPOUpdater(11, 3, policy_params=pol0, value_params=val0, tc_mode=tc_mode)
    data = _strong_data(32768, 16384, 2)                    # the SAME rollout and global minibatch tables on every rank
    upd.update_dev(*[D.from_host(x) for x in data], replicated=True)
    torch.cuda.synchronize()
    assert upd.policy.records is not None, "8,192 rows per rank must take the record-fed tensor-core kernel"
    assert upd.policy.exchange() == "ll", upd.policy.exchange()
    np.save(os.path.join(out_dir, "pol%d.npy" % rank), D.to_host(upd.policy.theta))
    np.save(os.path.join(out_dir, "val%d.npy" % rank), D.to_host(upd.value.theta))
    dist.destroy_process_group()


@pytest.mark.parametrize("tc_mode", ["exact", "fp16x2", "bf16x2"])
def test_strong_scaling_form_matches_oracle(tmp_path, tc_mode):
    """BASELINE configs[2] as SURVEY 8e words it: ONE rollout, every global minibatch split by rows over the GPUs (8,192
    rows per rank here, so the record-fed tcgen05 gradient kernel and the in-kernel low-latency exchange run TOGETHER),
    forward passes split and all-gathered.  Replicas end bit-identical and equal the oracle's update of the whole problem
    (benchmarks/ppo.py:22-40 + gae.py:34-75) -- the sharded run IS the single-GPU problem."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 2000 + (41 if tc_mode == "exact" else 43)
    mp.spawn(_strong_worker, args=(2, port, str(tmp_path), tc_mode), nprocs=2, join=True)
    pol = [np.load(os.path.join(tmp_path, "pol%d.npy" % r)) for r in range(2)]
    val = [np.load(os.path.join(tmp_path, "val%d.npy" % r)) for r in range(2)]
    assert np.array_equal(pol[0], pol[1]) and np.array_equal(val[0], val[1])     # replicas stay identical
    from oracle import mq_oracle as O
    from mannequin2_b200.ppo import init_mlp
    np.random.seed(5)
    pol0 = np.concatenate([init_mlp(11, [64, 64, 3], 0.1), np.zeros(3, np.float32)])
    val0 = init_mlp(11, [64, 64, 1], 0.1)
    o, a, r, d, vi, pi = _strong_data(32768, 16384, 2)
    pol_spec, val_spec = O.policy_spec(11, 3), O.mlp_spec(11, [64, 64, 1], [1, 1, 0])
    want_pol, want_val, _ = O.ppo_update(pol0.copy(), pol_spec, O.AdamO(pol0, horizon=10), val0.copy(), val_spec,
                                         O.AdamO(val0, horizon=10, lr=0.003), O.RunningNormalizeO(horizon=2),
                                         o, a, r, d.astype(bool), vi, pi)
    steps = len(vi)
    frac = {"exact": 0.007, "fp16x2": 0.007, "bf16x2": 0.07}[tc_mode]         # tests/test_compat_gpu.py: UPDATE_FRAC
    assert np.max(np.abs(pol[0] - pol0)) > 1e-4
    assert np.max(np.abs(pol[0] - want_pol)) < frac * 3e-4 * (1 + steps), np.max(np.abs(pol[0] - want_pol))
    assert np.max(np.abs(val[0] - want_val)) < frac * 3e-3 * (1 + steps), np.max(np.abs(val[0] - want_val))
