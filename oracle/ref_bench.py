"""`bench.py --impl reference`: time the UNMODIFIED reference's learner path on the host CPUs.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/ppo_oracle.py for who may import oracle/).

The reference is pure Python on torch-CPU.  `__graft_entry__.build()` copies its package (`/root/reference/drl`) into the
git-ignored `oracle/_ref/` when the reference tree is mounted, so that it travels to the GPU box with the snapshot; this module
imports it from there (or from /root/reference) under the gym / moviepy stub and runs it in the reference's own deployment
shape (script.py:108-116): W gloo ranks, one synthetic env each, DDP(Agent(NatureCNN)) + torch.optim.Adam, and per step

    annotate (abstract.py:371-411) -> credit_assignment (:413-445) -> PPO.optimize (ppo.py:242-281)

on a T = 128, B = 32 rollout per rank (opt_epochs = 3 incl. the per-epoch metrics passes) — the same work `PPO.optimize_host`
of this repository does for its envs.  Update samples per step = opt_epochs * W * T.
"""
import contextlib
import io
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch as tc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_COPY = os.path.join(ROOT, 'oracle', '_ref')


def reference_root():
    for p in (REF_COPY, os.environ.get('DRL_REFERENCE_ROOT', '/root/reference')):
        if os.path.isdir(os.path.join(p, 'drl', 'algos')):
            return p
    return None


def vendor_reference() -> bool:
    """Build step (this container only): copy the reference's python package into oracle/_ref (git-ignored, not
    gpurun-ignored).  Returns True when a copy exists afterwards."""
    import shutil
    src = os.path.join(os.environ.get('DRL_REFERENCE_ROOT', '/root/reference'), 'drl')
    if os.path.isdir(src):
        dst = os.path.join(REF_COPY, 'drl')
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns('__pycache__', 'test_*.py'))
    return os.path.isdir(os.path.join(REF_COPY, 'drl', 'algos'))


def _rank_main(rank, world, port, threads, T, B, A, epochs, warmup, steps, out_path):
    sys.path.insert(0, ROOT)
    from oracle import gym_stub
    gym_stub.install()
    sys.path.insert(0, reference_root())
    sys.dont_write_bytecode = True
    import gym
    from drl.envs.wrappers.stateless.abstract import RewardSpec
    from drl.agents.integration.agent import Agent
    from drl.agents.preprocessing import ToChannelMajor
    from drl.agents.architectures import NatureCNN, Linear
    from drl.agents.heads import CategoricalPolicyHead, SimpleValueHead
    from drl.algos.common.credit_assignment import GAE
    from drl.algos.ppo import PPO
    from torch.nn.parallel import DistributedDataParallel as DDP
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    tc.distributed.init_process_group('gloo', rank=rank, world_size=world)
    tc.set_num_threads(threads)

    class SynthEnv(gym.core.Env):
        def __init__(self):
            self.observation_space = gym.spaces.Box(0., 1., (84, 84, 4), np.float32)
            self.action_space = gym.spaces.Discrete(A)

        @property
        def reward_spec(self):
            return RewardSpec(keys=['extrinsic_raw', 'extrinsic'])

        def reset(self):
            return np.zeros((84, 84, 4), np.float32)

        def step(self, a):
            return self.reset(), {'extrinsic_raw': 0.0, 'extrinsic': 0.0}, False, {}

    ortho = lambda g: (lambda w: tc.nn.init.orthogonal_(w, gain=g))          # noqa: E731
    tc.manual_seed(10000 * rank)
    np.random.seed(10000 * rank)
    agent = Agent([ToChannelMajor()], NatureCNN(4, ortho(2 ** 0.5), tc.nn.init.zeros_),
                  {'policy': CategoricalPolicyHead(512, A, Linear, {}, ortho(0.01), tc.nn.init.zeros_),
                   'value_extrinsic': SimpleValueHead(512, Linear, {}, ortho(0.01), tc.nn.init.zeros_)})
    net = DDP(agent)
    opt = tc.optim.Adam(net.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-5)
    tmp = tempfile.mkdtemp()
    algo = PPO(rank=rank, world_size=world, rollout_len=T, extra_steps=0,
               credit_assignment_ops={'extrinsic': GAE(gamma=0.99, use_dones=True, lambda_=0.95)}, stats_window_len=100,
               non_learning_steps=0, max_steps=10 ** 9, checkpoint_frequency=10 ** 9, checkpoint_dir=tmp, log_dir=tmp,
               media_dir=tmp, global_step=1, env=SynthEnv(), policy_net=net, policy_optimizer=opt, policy_scheduler=None,
               value_net=None, value_optimizer=None, value_scheduler=None, standardize_adv=True, opt_epochs=epochs,
               learner_batch_size=B, clip_param_init=0.2, clip_param_final=0.0, ent_coef_init=0.01, ent_coef_final=0.01,
               vf_loss_cls='MSELoss', vf_loss_coef=1.0, vf_loss_clipping=False, vf_simple_weighting=True, use_pcgrad=False,
               reward_weights=None)
    if rank == 0:
        algo._writer.add_scalar = lambda *a, **k: None
    rs = np.random.RandomState(1234 + rank)
    obs = tc.from_numpy(rs.randint(0, 256, size=(T + 1, 84, 84, 4), dtype=np.uint8)).float() * float(np.float32(1.0 / 255.0))
    actions = tc.from_numpy(rs.randint(0, A, size=T + 1).astype(np.int64))
    rewards = tc.from_numpy(rs.randint(-1, 2, size=T).astype(np.float32))
    dones = tc.from_numpy((rs.uniform(size=T) < 0.01).astype(np.float32))

    def one():
        rollout = dict(observations=obs, actions=actions, rewards={'extrinsic': rewards}, dones=dones)
        rollout = algo.credit_assignment(algo.annotate(rollout, no_grad=True))
        algo.optimize(rollout)

    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(warmup):
            one()
        tc.distributed.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        tc.distributed.barrier()
        el = time.perf_counter() - t0
    if rank == 0:
        json.dump({'seconds': el, 'steps': steps, 'world': world, 'threads_per_rank': threads,
                   'samples_per_step': epochs * world * T, 'torch': tc.__version__}, open(out_path, 'w'))
    tc.distributed.destroy_process_group()


def run(warmup: int, steps: int, T: int = 128, B: int = 32, A: int = 6, epochs: int = 3, port: int = 29871):
    """Returns dict(rate=update samples/s, ...) or None when no copy of the reference is available."""
    if reference_root() is None:
        return None
    import torch.multiprocessing as mp
    cores = os.cpu_count() or 1
    world = 8 if cores >= 8 else max(1, cores)
    threads = max(1, cores // world)
    out_path = os.path.join(tempfile.mkdtemp(), 'ref_bench.json')
    env_pp = os.environ.get('PYTHONPATH', '')
    os.environ['PYTHONPATH'] = ROOT + (os.pathsep + env_pp if env_pp else '')
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU arm must still use all the host threads it can
    os.environ['OMP_NUM_THREADS'] = str(threads)
    mp.spawn(_rank_main, args=(world, port, threads, T, B, A, epochs, warmup, steps, out_path), nprocs=world, join=True)
    r = json.load(open(out_path))
    r['rate'] = r['samples_per_step'] * r['steps'] / r['seconds']
    r['cores'] = world * threads
    return r
