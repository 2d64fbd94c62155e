"""world_size-2 gloo tests (CPU) of the host-side logic of the data-parallel path: NCCL unique-id
bootstrap through torch.distributed, env sharding, per-env sampler streams, gradient bucket table,
and the identity "mean over ranks of per-rank means == reference DDP result" on the oracle.
Pattern follows the reference's drl/utils/test_distributed.py:8-30 (mp.spawn + gloo)."""
import os

import numpy as np
import pytest
import torch as tc
import torch.distributed as dist
import torch.multiprocessing as mp

WORLD = 2


def _worker(rank, world, port, fn_name, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        globals()[fn_name](rank, world, q)
    finally:
        dist.destroy_process_group()


def _run(fn_name, port):
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    mp.spawn(_worker, args=(WORLD, port, fn_name, q), nprocs=WORLD, join=True)
    out = []
    while not q.empty():
        out.append(q.get())
    return out


def _bootstrap(rank, world, q):
    from pytorch_drl_b200.comm import bootstrap_unique_id
    ident = bootstrap_unique_id(rank, world)
    q.put((rank, ident))


def test_unique_id_bootstrap_over_gloo():
    out = dict(_run('_bootstrap', 29731))
    assert len(out) == WORLD and len(out[0]) == 128
    assert out[0] == out[1] and any(b != 0 for b in out[0])


def _sharding(rank, world, q):
    from pytorch_drl_b200.comm import shard_envs, Communicator
    from pytorch_drl_b200.algos import IndependentSampler
    envs = shard_envs(8, rank, world)
    perms = []
    for e in envs:      # env e draws the stream its own reference rank would have (launch.py:33-45)
        s = IndependentSampler(16, 8, np.random.RandomState(10000 * e + 0))
        s.resample()
        perms.append(np.concatenate(list(iter(s))).tolist())
    # the metric mean of per-rank values (global_mean, metrics.py:8-26) through gloo
    v = tc.tensor([float(rank + 1), 10.0 * (rank + 1)])
    dist.all_reduce(v)
    q.put((rank, list(envs), perms, (v / world).tolist(), Communicator.grad_buckets([0, 1, 2, 3, 4, 5, 6, 7, 8])))


def test_env_sharding_and_sampler_streams():
    out = {r: (envs, perms, mean, buckets) for r, envs, perms, mean, buckets in _run('_sharding', 29732)}
    assert out[0][0] == [0, 1, 2, 3] and out[1][0] == [4, 5, 6, 7]
    for r in range(WORLD):
        for e, p in zip(out[r][0], out[r][1]):
            ref = np.random.RandomState(10000 * e)
            ref.permutation(16)
            assert p == ref.permutation(16).tolist()
        assert out[r][2] == [1.5, 15.0]
        assert out[r][3] == [0, 6, 8]
    with pytest.raises(ValueError):
        from pytorch_drl_b200.comm import shard_envs
        shard_envs(7, 0, 2)


def _oracle_ddp(rank, world, q):
    """Each rank computes its env's oracle gradients; gloo-averaging them must equal the oracle's
    multi-env step (the semantics the GPU path implements with one SUM all-reduce + 1/world)."""
    from oracle import ppo_oracle as po
    A, rk, T, B = 6, ['extrinsic'], 8, 4
    params = po.init_params(A, rk, seed=11)
    credit = {'extrinsic': dict(gamma=0.99, lambda_=0.95, use_dones=True)}
    mbs = []
    for r in range(world):
        raw = po.make_synthetic_env_rollout(r, 0, T, A, rk)
        ro = po.collect_from_raw(params, raw, T, credit, True)
        mbs.append(po.slice_env(ro, np.arange(B)))
    hp = po.Hyper({'extrinsic': 1.0})
    _, g_mine = po.minibatch_losses_and_grads(params, [mbs[rank]], hp)
    key = po.POLICY_KEYS[0]
    t = tc.from_numpy(g_mine[key].copy())
    dist.all_reduce(t)
    t /= world
    _, g_all = po.minibatch_losses_and_grads(params, mbs, hp)
    q.put((rank, float((t - tc.from_numpy(g_all[key])).abs().max()), float(tc.from_numpy(g_all[key]).abs().max())))


def test_mean_over_ranks_equals_multi_env_step():
    for rank, err, scale in _run('_oracle_ddp', 29733):
        assert err <= 1e-6 * max(scale, 1e-3) + 1e-9
