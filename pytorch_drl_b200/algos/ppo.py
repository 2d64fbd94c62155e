"""Drop-in for drl/algos/ppo.py (`PPO`) and the learner half of drl/algos/abstract.py
(`ActorCriticAlgo.annotate` / `credit_assignment`), running on one B200 per process.

The reference runs one env per rank.  Here one process drives `envs_per_rank` envs on
one GPU; every env keeps the per-rank semantics of the reference (own advantage standardisation,
own minibatch permutation stream), and a learner step consumes `learner_batch_size` samples of
each env, so that losses, gradients and updates equal those of a reference job with
world_size = n_gpus * envs_per_rank (mean over ranks of per-rank means, equal shards).
"""
from typing import Any, Dict, List, Mapping, Optional, Sequence

import numpy as np
import torch as tc

from .. import ops
from ..utils.nested import MinibatchView
from .credit_assignment import GAE
from .samplers import IndependentSampler
from .schedules import LinearSchedule

METRIC_NAMES = ops.LOSS_NAMES


def extract_reward_name(prediction_key: str) -> str:
    """'value_extrinsic' -> 'extrinsic' (drl/algos/abstract.py, helper of annotate)."""
    for prefix in ('action_value_', 'value_'):
        if prediction_key.startswith(prefix):
            return prediction_key[len(prefix):]
    raise ValueError(prediction_key)


class RolloutStorage:
    """Device-resident rollout of `n_envs` trajectories: every array is [n_envs, Tp, ...] with
    Tp = T+1 rounded up to a multiple of 4, so that ONE flat row number `env*Tp + t` addresses the
    uint8 frame and all per-step scalars of a sample, and rows stay 16-byte aligned for the
    vectorised GAE loads.  This is the storage side of drl/algos/common/rollout.py:105-224 for
    the learner (frames stay uint8: 28 224 B/sample instead of the reference's 112 896 B)."""

    def __init__(self, n_envs: int, rollout_len: int, reward_keys: Sequence[str], device, extra_steps: int = 0):
        self.n_envs, self.T, self.extra_steps = n_envs, rollout_len, extra_steps
        self.Tp = (rollout_len + extra_steps + 1 + 3) // 4 * 4        # T + n frames, n = extra_steps + 1 (rollout.py:105-224)
        self.reward_keys = list(reward_keys)
        f = lambda dt, *tail: tc.zeros((n_envs, self.Tp) + tail, dtype=dt, device=device)  # noqa: E731
        self.data: Dict[str, Any] = dict(
            observations=f(tc.uint8, 84, 84, 4), actions=f(tc.int64), dones=f(tc.float32),
            logprobs=f(tc.float32), entropies=f(tc.float32),
            rewards={k: f(tc.float32) for k in reward_keys},
            vpreds={k: f(tc.float32) for k in reward_keys},
            advantages={k: f(tc.float32) for k in reward_keys},
            vtargs={k: f(tc.float32) for k in reward_keys})

    def as_dict(self) -> Dict[str, Any]:
        return self.data

    def load_env(self, e: int, observations, actions, rewards: Mapping[str, Any], dones) -> None:
        """Fill env `e` from host/device arrays: T+n frames/actions, T+n-1 rewards/dones (n = extra_steps + 1)."""
        d, T = self.data, self.T + self.extra_steps
        dev = d['actions'].device
        d['observations'][e, :T + 1].copy_(tc.as_tensor(observations).to(dev))
        d['actions'][e, :T + 1].copy_(tc.as_tensor(actions).to(dev))
        d['dones'][e, :T].copy_(tc.as_tensor(dones).to(dev))
        for k in self.reward_keys:
            d['rewards'][k][e, :T].copy_(tc.as_tensor(rewards[k]).to(dev))

    def rows(self, t_count: int) -> tc.Tensor:
        """Flat row numbers of steps [0, t_count) of every env, env-major."""
        dev = self.data['actions'].device
        e = tc.arange(self.n_envs, device=dev, dtype=tc.int64).unsqueeze(1) * self.Tp
        return (e + tc.arange(t_count, device=dev, dtype=tc.int64).unsqueeze(0)).reshape(-1)


class PPO:
    """Proximal Policy Optimization (clip variant) with the reference's constructor
    (drl/algos/ppo.py:33-67) and learner methods.  Extra keyword arguments (all optional):
      envs_per_rank  number of envs this process/GPU owns (reference: 1),
      comm           a `pytorch_drl_b200.comm.Communicator` replacing the DDP wrapper,
      sampler_seed   if given, env e draws its permutations from RandomState(10000*e_global+seed)
                     — the stream its own rank would have had (launch.py:33-45); else np.random.
    """

    def __init__(self, rank: int, world_size: int, rollout_len: int, extra_steps: int,
                 credit_assignment_ops: Mapping[str, GAE], stats_window_len: int = 100,
                 non_learning_steps: int = 0, max_steps: int = 10 ** 7, checkpoint_frequency: int = 10 ** 9,
                 checkpoint_dir: str = '', log_dir: str = '', media_dir: str = '', global_step: int = 0,
                 env=None, policy_net=None, policy_optimizer=None, policy_scheduler=None, value_net=None,
                 value_optimizer=None, value_scheduler=None, standardize_adv: bool = True,
                 opt_epochs: int = 3, learner_batch_size: int = 64, clip_param_init: float = 0.2,
                 clip_param_final: float = 0.0, ent_coef_init: float = 0.01, ent_coef_final: float = 0.01,
                 vf_loss_cls: str = 'MSELoss', vf_loss_coef: float = 1.0, vf_loss_clipping: bool = False,
                 vf_simple_weighting: bool = True, use_pcgrad: bool = False,
                 reward_weights: Optional[Mapping[str, float]] = None, envs_per_rank: int = 1, comm=None,
                 sampler_seed: Optional[int] = None) -> None:
        if value_net is not None:
            raise NotImplementedError('separate value network (use_shared_architecture: False) is not fused')
        if use_pcgrad:
            raise NotImplementedError('PCGrad is off in both shipped configs and not on the fused path')
        if extra_steps < 0:
            raise ValueError('extra_steps must be >= 0')
        assert rollout_len % learner_batch_size == 0                    # samplers.py:51
        self._rank, self._world_size = rank, world_size
        self._rollout_len, self._extra_steps = rollout_len, extra_steps
        self._credit_assignment_ops = credit_assignment_ops
        self._global_step = global_step
        self._env = env
        self._non_learning_steps, self._max_steps = non_learning_steps, max_steps
        self._checkpoint_frequency, self._checkpoint_dir = checkpoint_frequency, checkpoint_dir
        self._log_dir, self._media_dir = log_dir, media_dir
        self._reward_weights = reward_weights
        self._policy_net_wrapped = policy_net                           # what gets checkpointed (a DDPShim keeps the `module.` keys)
        self._policy_net = getattr(policy_net, 'module', policy_net)    # accept a DDP-style wrapper
        self._policy_optimizer, self._policy_scheduler = policy_optimizer, policy_scheduler
        self._standardize_adv = standardize_adv
        self._opt_epochs, self._learner_batch_size = opt_epochs, learner_batch_size
        self._clip_param = LinearSchedule(clip_param_init, clip_param_final, max_steps)
        self._ent_coef = LinearSchedule(ent_coef_init, ent_coef_final, max_steps)
        self._vf_loss_cls, self._vf_loss_coef = vf_loss_cls, vf_loss_coef
        self._vf_loss_clipping, self._vf_simple_weighting = vf_loss_clipping, vf_simple_weighting
        self._envs_per_rank, self._comm = envs_per_rank, comm
        self._engine = self._policy_net._require_engine()
        if policy_optimizer is not None:
            # the fused Adam's flat state becomes the torch optimizer's `state` (views), so policy_optimizer.state_dict() is
            # complete and a state loaded before construction (launch.py:377-383) or later is adopted (ppo.py:295-307)
            from ..utils.checkpoint import bind_flat_adam
            bind_flat_adam(self._engine, policy_optimizer, list(self._policy_net.parameters()))
        if comm is not None and hasattr(comm, 'attach') and getattr(comm, 'overlap', False):
            comm.attach(self._engine)                # opt-in: gradient exchange overlapped with the backward pass
        self._samplers = []
        for e in range(envs_per_rank):
            rng = None
            if sampler_seed is not None:
                rng = np.random.RandomState(10000 * (rank * envs_per_rank + e) + sampler_seed)
            self._samplers.append(IndependentSampler(rollout_len, learner_batch_size, rng))
        self.last_epoch_metrics: List[Dict[str, float]] = []
        self.on_epoch_metrics = None       # optional callback(epoch, metrics) (tensorboard hook)

    # ---- reward bookkeeping (abstract.py:108-134) ---------------------------------------------------
    def _get_reward_keys(self, omit_raw: bool = True) -> List[str]:
        if self._env is not None and getattr(self._env, 'reward_spec', None) is not None:
            keys = list(self._env.reward_spec.keys)
        else:
            keys = [extract_reward_name(k) for k in self._policy_net.value_keys]
        if omit_raw:
            keys = [k for k in keys if not k.endswith('_raw')]
            assert len(keys) > 0
        return keys

    def _get_reward_weights(self) -> Dict[str, float]:
        rw = {'extrinsic': 1.0}
        if self._reward_weights is not None:
            rw.update(self._reward_weights)
        assert set(rw.keys()) == set(self._get_reward_keys(omit_raw=True))
        return rw

    def _stream_order(self) -> List[str]:
        return [extract_reward_name(k) for k in self._policy_net.value_keys]

    def _hyper(self):
        rw = self._get_reward_weights()
        order = self._stream_order()
        return ops.make_hyper(self._engine.A, [rw[k] for k in order],
                              self._clip_param.value(self._global_step),
                              self._ent_coef.value(self._global_step), self._vf_loss_coef,
                              self._vf_loss_clipping, self._vf_simple_weighting, self._vf_loss_cls)

    def _batch(self, rollout):
        order = self._stream_order()
        return ops.make_batch(rollout['actions'], rollout['logprobs'],
                              [rollout['advantages'][k] for k in order],
                              [rollout['vpreds'][k] for k in order],
                              [rollout['vtargs'][k] for k in order])

    # ---- annotate / credit assignment ---------------------------------------------------------------
    @tc.no_grad()
    def annotate(self, rollout: Dict[str, Any], no_grad: bool = True, envs: Optional[range] = None) -> Dict[str, Any]:
        """abstract.py:371-411 for all envs at once (or the env slice `envs`): one no-grad forward over the T+1 frames of
        every env; fills logprobs/entropies/vpreds in place and returns the rollout."""
        obs = rollout['observations']
        Tp = obs.shape[1]
        e0, e1 = (0, obs.shape[0]) if envs is None else (envs.start, envs.stop)
        n = e1 - e0
        T1 = self._rollout_len + self._extra_steps + 1                 # T + n frames (abstract.py:383-397)
        e = tc.arange(e0, e1, device=obs.device, dtype=tc.int64).unsqueeze(1) * Tp
        rows = (e + tc.arange(T1, device=obs.device, dtype=tc.int64).unsqueeze(0)).reshape(-1)
        _, values, logp, ent = self._engine.forward(obs, rows, actions=rollout['actions'])
        rollout['logprobs'][e0:e1, :T1] = logp.view(n, T1)
        rollout['entropies'][e0:e1, :T1] = ent.view(n, T1)
        for i, k in enumerate(self._stream_order()):
            rollout['vpreds'][k][e0:e1, :T1] = values[:, i].reshape(n, T1)
        return rollout

    @tc.no_grad()
    def credit_assignment(self, rollout: Dict[str, Any]) -> Dict[str, Any]:
        """abstract.py:413-445: per reward stream GAE -> vtargs = A + V -> optional per-env
        standardisation, one fused launch per stream."""
        T, n = self._rollout_len, self._extra_steps + 1
        for k in self._get_reward_keys(omit_raw=True):
            op = self._credit_assignment_ops[k]
            if n == 1 and hasattr(op, 'call_fused'):
                op.call_fused(T, rollout['rewards'][k], rollout['vpreds'][k], rollout['dones'],
                              self._standardize_adv, rollout['advantages'][k], rollout['vtargs'][k])
                continue
            # n-step estimators / extra steps: estimator over T + n - 1 steps, then the reference's three lines
            # (abstract.py:427-439): vtargs = A + V[:T], optional standardisation, both kept for the first T steps
            adv = op(rollout_len=T, extra_steps=n - 1, rewards=rollout['rewards'][k][:, :T + n - 1],
                     vpreds=rollout['vpreds'][k][:, :T + n], dones=rollout['dones'][:, :T + n - 1])
            rollout['vtargs'][k][:, :T] = adv + rollout['vpreds'][k][:, :T]
            rollout['advantages'][k][:, :T] = ops.standardize(adv.contiguous()) if self._standardize_adv else adv
        return rollout

    # ---- losses -------------------------------------------------------------------------------------
    def _all_rows(self, rollout) -> tc.Tensor:
        obs = rollout['observations']
        n, Tp = obs.shape[0], obs.shape[1]
        e = tc.arange(n, device=obs.device, dtype=tc.int64).unsqueeze(1) * Tp
        return (e + tc.arange(self._rollout_len, device=obs.device, dtype=tc.int64).unsqueeze(0)).reshape(-1)

    def compute_losses_and_metrics(self, minibatch: Dict[str, Any], no_grad: bool) -> Dict[str, tc.Tensor]:
        """ppo.py:172-207.  `minibatch` is a MinibatchView (rollout + row indices) or a whole
        rollout dict.  With no_grad=False the backward is fused in and the gradient lands in the
        engine's flat gradient buffer (there is no autograd tape to return)."""
        rows = minibatch.row_idx if isinstance(minibatch, MinibatchView) else self._all_rows(minibatch)
        hyper, batch = self._hyper(), self._batch(minibatch)
        eng = self._engine
        if not no_grad:
            if rows.numel() > eng.max_batch:
                raise ValueError('minibatch larger than the engine max_batch')
            out = eng.ppo_step(minibatch['observations'], rows, batch, hyper, grads=True).clone()
        else:
            out = tc.zeros(5, dtype=tc.float32, device=eng.device)
            total = rows.numel()
            for s in range(0, total, eng.max_batch):
                part = rows[s:s + eng.max_batch]
                out += eng.ppo_step(minibatch['observations'], part, batch, hyper, grads=False) * (part.numel() / total)
        return {n: out[i] for i, n in enumerate(METRIC_NAMES)}

    # ---- optimize -----------------------------------------------------------------------------------
    def _optimizer_hparams(self):
        g = self._policy_optimizer.param_groups[0] if self._policy_optimizer is not None else {}
        return (float(g.get('lr', 1e-3)), tuple(g.get('betas', (0.9, 0.999))), float(g.get('eps', 1e-8)))

    def _minibatch_rows(self, rollout, blocks_per_env: List[np.ndarray]) -> tc.Tensor:
        """Flat row numbers of ONE minibatch given every env's index block (tests / benchmarks; optimize uses _epoch_rows)."""
        Tp = rollout['observations'].shape[1]
        idx = np.stack([b.astype(np.int64) + e * Tp for e, b in enumerate(blocks_per_env)]).reshape(-1)
        return tc.as_tensor(idx, device=self._engine.device)

    def learner_step(self, minibatch, hyper=None, batch=None) -> tc.Tensor:
        """One learner step on a minibatch (ppo.py:250-257): forward, fused losses, backward,
        gradient all-reduce across GPUs (the DDP mean, launch.py:310), Adam.  Returns the device
        tensor losses[5] (policy_loss, value_loss, composite_loss, meanent, clipfrac).  `hyper` / `batch` are the C-ABI
        argument blocks of this rollout (built once per optimize call; rebuilt here when omitted)."""
        eng = self._engine
        rows = minibatch.row_idx
        if rows.numel() > eng.max_batch:
            raise ValueError('minibatch larger than the engine max_batch')
        losses = eng.ppo_step(minibatch['observations'], rows, batch if batch is not None else self._batch(minibatch),
                              hyper if hyper is not None else self._hyper(), grads=True)
        world = self._world_size if self._comm is not None else 1
        if self._comm is not None and getattr(eng, 'comm_attached', None) is not self._comm:
            self._comm.allreduce_grads(eng)          # not attached: exchange after the backward pass
        lr, betas, eps = self._optimizer_hparams()
        eng.adam_step(lr, betas, eps, grad_scale=1.0 / world)
        return losses

    def _epoch_rows(self, rollout) -> tc.Tensor:
        """Resample every env's permutation (samplers.py:60-63) and return the epoch's minibatches as ONE device tensor
        [T/B steps, n_envs * B] of flat row numbers: a single host-to-device copy per epoch instead of one per step."""
        for s in self._samplers:
            s.resample()
        Tp = rollout['observations'].shape[1]
        B, steps = self._learner_batch_size, self._rollout_len // self._learner_batch_size
        blocks = []
        for e, s in enumerate(self._samplers):
            it = iter(s)
            blocks.append(np.stack([next(it) for _ in range(steps)]).astype(np.int64) + e * Tp)      # [steps, B]
        idx = np.ascontiguousarray(np.stack(blocks, axis=1).reshape(steps, -1))                       # [steps, n_envs * B]
        # pinned staging (two buffers: an epoch's host sync at its metrics read separates the reuses) + a copy KERNEL: the copy
        # engines stay free for rollout uploads, and nothing here synchronises the host
        from .. import _lib
        ring = getattr(self, '_idx_ring', None)
        if ring is None or ring[0].numel() != idx.size:
            ring = [tc.empty(idx.size, dtype=tc.int64).pin_memory() for _ in range(2)]
            self._idx_ring, self._idx_turn = ring, 0
            self._idx_events = [None, None]
        slot = self._idx_turn % 2
        self._idx_turn += 1
        if self._idx_events[slot] is not None:
            self._idx_events[slot].synchronize()      # the kernel that read this staging buffer two epochs ago has run
        host = ring[slot]
        host.numpy()[:] = idx.reshape(-1)
        dev = tc.empty(idx.shape, dtype=tc.int64, device=self._engine.device)
        _lib.check(_lib.lib().drl_copy_by_kernel(_lib.ptr(dev), host.data_ptr(), idx.size * 8, _lib.current_stream()), 'copy_by_kernel')
        ev = tc.cuda.Event()
        ev.record()
        self._idx_events[slot] = ev
        return dev

    def optimize(self, rollout: Dict[str, Any]) -> None:
        """ppo.py:242-281: opt_epochs x T/B learner steps (+ the per-epoch full-rollout metrics pass
        and the scheduler step).  One learner step = B samples of each env of every GPU."""
        self.last_epoch_metrics = []
        hyper, batch = self._hyper(), self._batch(rollout)        # global_step (the schedules' clock) is fixed during optimize
        for opt_epoch in range(self._opt_epochs):
            epoch_rows = self._epoch_rows(rollout)
            for step in range(epoch_rows.shape[0]):
                minibatch = MinibatchView(rollout, epoch_rows[step])
                self._update_trainable_wrappers(minibatch)
                if self._global_step <= self._non_learning_steps:
                    continue
                self.learner_step(minibatch, hyper, batch)
            metrics = self.compute_losses_and_metrics(rollout, no_grad=True)
            vec = tc.stack([metrics[n] for n in METRIC_NAMES])
            if self._comm is not None:
                vec = self._comm.mean_(vec)
            host = vec.tolist()                       # the step's device->host read
            m = dict(zip(METRIC_NAMES, host))
            self.last_epoch_metrics.append(m)
            if self.on_epoch_metrics is not None and self._rank == 0:
                self.on_epoch_metrics(opt_epoch, m)
        if self._policy_scheduler:
            self._policy_scheduler.step()

    def upload_rollout(self, host: Mapping[str, Any], storage: 'RolloutStorage', chunk_envs: int = 128):
        """Start the host-to-device copy of a rollout the actors left in (pinned) HOST memory: scalars first, then the uint8 frames
        in env chunks on a copy stream.  Returns the per-chunk events `annotate_uploaded` waits on.  The copy stream first waits
        for the work already queued on the current stream (earlier readers of `storage`)."""
        d = storage.as_dict()
        dev = self._engine.device
        main = tc.cuda.current_stream(dev)
        if getattr(self, '_copy_stream', None) is None:
            self._copy_stream = tc.cuda.Stream(device=dev)
        cs = self._copy_stream
        cs.wait_stream(main)
        events = []
        with tc.cuda.stream(cs):
            for k in ('actions', 'dones'):
                d[k].copy_(host[k], non_blocking=True)
            for k in storage.reward_keys:
                d['rewards'][k].copy_(host['rewards'][k], non_blocking=True)
            for e0 in range(0, storage.n_envs, chunk_envs):
                e1 = min(storage.n_envs, e0 + chunk_envs)
                # copies of <= 32 MB: the host-to-device copy engine serves one copy at a time, so anything else that needs it (the
                # per-epoch index upload of an optimize() running concurrently) waits for at most one of these, not for 0.5 GB
                for c0 in range(e0, e1, 8):
                    c1 = min(e1, c0 + 8)
                    d['observations'][c0:c1].copy_(host['observations'][c0:c1], non_blocking=True)
                ev = tc.cuda.Event()
                ev.record(cs)
                events.append((e0, e1, ev))
        return events

    def annotate_uploaded(self, storage: 'RolloutStorage', events) -> Dict[str, Any]:
        """annotate (abstract.py:383-411) chunk by chunk as the uploads land, then credit assignment."""
        d = storage.as_dict()
        main = tc.cuda.current_stream(self._engine.device)
        for e0, e1, ev in events:
            main.wait_event(ev)
            self.annotate(d, envs=range(e0, e1))
        return self.credit_assignment(d)

    def optimize_host(self, host: Mapping[str, Any], storage: 'RolloutStorage', chunk_envs: int = 128) -> Dict[str, Any]:
        """collect() -> optimize() for a rollout in pinned host memory: the uint8 frames are uploaded ONCE per rollout (not once
        per minibatch: a rollout is consumed opt_epochs times); each env chunk's annotate forward runs as soon as it has landed,
        overlapping the rest of the upload; then credit assignment and the epochs.  `host` holds observations
        [n_envs, >= T+1, 84, 84, 4] uint8, actions, dones and rewards{key}, laid out like `storage`.  Returns the device rollout.
        (The learner steps cannot start earlier: every step draws B samples of EVERY env, and the advantages need the whole
        trajectory's value predictions.)"""
        rollout = self.annotate_uploaded(storage, self.upload_rollout(host, storage, chunk_envs))
        self.optimize(rollout)
        return rollout

    # ---- the rest of Algo's surface (drl/algos/abstract.py:173-221, ppo.py:283-307) ---------------------------------------
    @property
    def checkpointables(self) -> Dict[str, Any]:
        """What `persist` saves (ppo.py:295-307): the policy network (with the `module.` prefix if it was handed over wrapped),
        its optimizer and scheduler, and the env wrappers' checkpointables."""
        out = {'policy_net': self._policy_net_wrapped, 'policy_optimizer': self._policy_optimizer,
               'policy_scheduler': self._policy_scheduler, 'value_net': None, 'value_optimizer': None, 'value_scheduler': None}
        w = self._env
        while w is not None:                         # every wrapper of the chain (a reference wrapper already merges its inner ones)
            for k, v in (getattr(w, 'checkpointables', None) or {}).items():
                out.setdefault(k, v)
            w = getattr(w, 'env', None)
        return out

    def collect(self):
        """abstract.py:173-186.  Env stepping is the actor's side (out of scope): a caller supplies
        `rollout_source()` -> (RolloutStorage dict with frames / actions / rewards / dones filled, metadata); annotate and
        credit assignment run here, on the device."""
        src = getattr(self, 'rollout_source', None)
        mgr = getattr(self, 'rollout_mgr', None)
        if src is not None:
            rollout, metadata = src()
        elif mgr is not None:                      # a BatchedRolloutManager (algos/rollout.py), like self._rollout_mgr of the reference
            rollout = mgr.generate()
            metadata = rollout.pop('metadata')
        else:
            raise NotImplementedError('set algo.rollout_mgr = BatchedRolloutManager(envs, policy_net, ...) or algo.rollout_source = '
                                      'callable returning (rollout, metadata)')
        rollout = self.credit_assignment(self.annotate(rollout))
        self._global_step += self._world_size * self._envs_per_rank * self._rollout_len        # abstract.py:185
        return rollout, metadata

    def persist(self, metadata: Optional[Mapping[str, Any]] = None) -> None:
        """ppo.py:283-307 without the tensorboard writer: rank 0 saves every checkpointable at multiples of
        checkpoint_frequency in the reference's file format."""
        if self._rank == 0 and self._checkpoint_dir and self._global_step % self._checkpoint_frequency == 0:
            from ..utils.checkpoint import save_checkpoints
            save_checkpoints(self._checkpoint_dir, self.checkpointables, self._global_step)

    def training_loop(self) -> None:
        """abstract.py:211-221."""
        while self._global_step < self._max_steps:
            rollout, metadata = self.collect()
            self.optimize(rollout)
            self.persist(metadata)

    def _update_trainable_wrappers(self, minibatch) -> None:
        """drl/algos/common/wrapper_ops.py:9-25: every wrapper of the chain (linked through `.env`) that has a `learn`."""
        w = self._env
        while w is not None:
            if hasattr(w, 'learn'):
                w.learn(minibatch)
            w = getattr(w, 'env', None)
