"""Two-GPU checks of the data-parallel learner step (skipped when fewer than 2 GPUs are visible).

The gradient exchange attached to the network handle (`drl_net_set_comm`: fc + heads range all-reduced on a side
stream while the convolution gradients are still being computed) must return exactly the gradients of the
plain "backward, then all-reduce" order, on both engines, and they must equal the sum of the per-rank gradients.
"""
import os
import socket

import numpy as np
import pytest
import torch as tc

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from oracle import ppo_oracle as po
    from pytorch_drl_b200 import ops
    from pytorch_drl_b200.algos import RolloutStorage
    from pytorch_drl_b200.comm import Communicator
    from pytorch_drl_b200.engine import LearnerEngine
    tc.cuda.set_device(rank)
    dev = tc.device('cuda', rank)
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    comm = Communicator(rank, world, rank, p2p_elems=1 << 21)   # NVLink peer-memory all-reduce
    comm_nccl = Communicator(rank, world, rank, p2p_elems=0)  # plain NCCL
    A, rk, n_envs, T, B = 6, ['extrinsic'], 2, 16, 8
    params = po.init_params(A, rk, seed=5)
    credit = {'extrinsic': dict(gamma=0.99, lambda_=0.95, use_dones=True)}
    raws = [po.make_synthetic_env_rollout(rank * n_envs + e, 0, T, A, rk) for e in range(n_envs)]
    rollouts = [po.collect_from_raw(params, raw, T, credit, True) for raw in raws]
    st = RolloutStorage(n_envs, T, rk, dev)
    d = st.as_dict()
    for e in range(n_envs):
        st.load_env(e, raws[e]['observations'], raws[e]['actions'], raws[e]['rewards'], raws[e]['dones'])
        d['logprobs'][e, :T] = tc.from_numpy(rollouts[e]['logprobs']).to(dev)
        for name in ('advantages', 'vpreds', 'vtargs'):
            d[name]['extrinsic'][e, :T] = tc.from_numpy(rollouts[e][name]['extrinsic']).to(dev)
    rows = tc.as_tensor(np.concatenate([np.arange(B) * 2 % T + e * st.Tp for e in range(n_envs)]).astype(np.int64), device=dev)
    hyper = ops.make_hyper(A, [1.0], 0.2, 0.01, 1.0, False, True)
    batch = ops.make_batch(d['actions'], d['logprobs'], [d['advantages']['extrinsic']], [d['vpreds']['extrinsic']],
                           [d['vtargs']['extrinsic']])
    result = {}
    for precision in ('fp32', 'bf16'):
        eng = LearnerEngine(A, ['value_extrinsic'], n_envs * B, precision)
        eng.load_named({k: tc.from_numpy(v) for k, v in params.items()})
        eng.ppo_step(d['observations'], rows, batch, hyper)
        local = eng.grads.clone()
        plain = comm.allreduce_(local.clone(), Communicator.grad_buckets(eng.offsets))     # backward, then all-reduce
        nccl = comm_nccl.allreduce_(local.clone(), Communicator.grad_buckets(eng.offsets))
        small = comm.mean_(tc.arange(5, device=dev, dtype=tc.float32) + rank)              # 5-float metric vector
        comm.attach(eng)
        for _ in range(3):                                                                  # repeated: buffers are reused
            eng.ppo_step(d['observations'], rows, batch, hyper)
        tc.cuda.synchronize()
        fused = eng.grads.clone()
        # the network's vector-Jacobian product (drl_net_backward, the autograd path of Agent.forward) with the communicator
        # still ATTACHED must stay a purely local computation: no collective is started on the caller's scratch buffer
        dl = tc.full((n_envs * B, A), 0.01 * (rank + 1), device=dev)
        dv = tc.full((n_envs * B, 1), -0.02 * (rank + 1), device=dev)
        vjp_attached = eng.backward(d['observations'], rows, dl, dv).clone()
        check_ok = __import__('pytorch_drl_b200')._lib.lib().drl_net_set_comm(eng._h, None)
        assert check_ok == 0
        vjp_detached = eng.backward(d['observations'], rows, dl, dv).clone()
        tc.cuda.synchronize()
        result[precision] = dict(local=local.cpu().numpy(), plain=plain.cpu().numpy(), fused=fused.cpu().numpy(),
                                 nccl=nccl.cpu().numpy(), small=small.cpu().numpy(), p2p=np.array([int(comm.p2p)]),
                                 vjp_a=vjp_attached.cpu().numpy(), vjp_d=vjp_detached.cpu().numpy())
        del eng
    np.savez(os.path.join(out_dir, f'rank{rank}.npz'), **{f'{p}_{k}': v for p, r in result.items() for k, v in r.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(tc.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_overlapped_gradient_exchange_matches_plain_allreduce(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = [np.load(tmp_path / f'rank{i}.npz') for i in range(2)]
    for precision in ('fp32', 'bf16'):
        total = r[0][f'{precision}_local'] + r[1][f'{precision}_local']
        for i in range(2):
            np.testing.assert_array_equal(r[i][f'{precision}_plain'], total)       # 2-rank fp32 sum is order-free
            np.testing.assert_array_equal(r[i][f'{precision}_nccl'], total)        # NCCL and peer-memory paths agree
            np.testing.assert_allclose(r[i][f'{precision}_small'], np.arange(5) + 0.5, rtol=0, atol=1e-6)
            assert int(r[i][f'{precision}_p2p'][0]) == 1, 'peer-memory all-reduce was not enabled'
            # the fused result comes from a re-run of the step: its red.global.add partial sums arrive in a different
            # order, so compare as a whole (a missing or doubled range would show up as O(1))
            f, p = r[i][f'{precision}_fused'].astype(np.float64), r[i][f'{precision}_plain'].astype(np.float64)
            assert np.linalg.norm(f - p) <= 1e-5 * np.linalg.norm(p), precision
            va, vd = r[i][f'{precision}_vjp_a'].astype(np.float64), r[i][f'{precision}_vjp_d'].astype(np.float64)
            assert np.linalg.norm(va - vd) <= 1e-5 * np.linalg.norm(vd), (precision, 'drl_net_backward exchanged gradients')
        # ... and the two ranks' (different) cotangents gave different local results: nothing was summed across ranks
        assert np.linalg.norm(r[0][f'{precision}_vjp_a'] - r[1][f'{precision}_vjp_a']) > 1e-3 * np.linalg.norm(r[0][f'{precision}_vjp_a'])
        assert np.abs(total).max() > 0
