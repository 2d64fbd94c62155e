"""Drop-in for drl/envs/wrappers/stateful/normalize_reward.py on the GPU, for all environments of a rank at once.

The reference wraps ONE environment per process: `NormalizeRewardWrapper.step` normalises the step's reward with
the synchronised `ReturnAcc` and feeds the raw reward into the un-synchronised one (:171-194); `learn()` replaces
the synchronised moments by the mean over all processes and copies them back (:144-156, 196-210).  Here one object
owns the accumulators of the rank's N environments and works on [N, T] blocks of a rollout; the mean over
environments is a local mean followed by the communicator's all-reduce (same value as the reference's
`global_mean` over world_size = total environments)."""
from typing import Dict, Optional

import torch as tc

from .. import _lib
from .._lib import check, current_stream, ptr


class BatchedRewardNormalizer:
    def __init__(self, n_envs: int, gamma: float, use_dones: bool, device=None, comm=None,
                 clip_low: float = -10.0, clip_high: float = 10.0, env=None):
        self.env = env                      # next trainable wrapper of the chain (wrapper_ops.py:21-25)
        if not tc.cuda.is_available():
            raise RuntimeError('BatchedRewardNormalizer needs a CUDA device; there is no CPU fallback')
        dev = tc.device('cuda', tc.cuda.current_device()) if device is None else tc.device(device)
        self._L = _lib.lib()
        self.n_envs, self.gamma, self.use_dones = int(n_envs), float(gamma), bool(use_dones)
        self.clip_low, self.clip_high = float(clip_low), float(clip_high)
        self.trace_length = int(5 / (1 - gamma))                  # normalize_reward.py:27
        self._comm = comm
        W = 2 * self.trace_length
        self._win_r = tc.zeros((W, n_envs), dtype=tc.float32, device=dev)
        self._win_d = tc.zeros((W, n_envs), dtype=tc.float32, device=dev)
        self._count = tc.zeros(n_envs, dtype=tc.int32, device=dev)
        # un-synchronised accumulators, one per environment
        self.steps = tc.zeros(n_envs, dtype=tc.float32, device=dev)
        self.moment1 = tc.zeros(n_envs, dtype=tc.float32, device=dev)
        self.moment2 = tc.zeros(n_envs, dtype=tc.float32, device=dev)
        # synchronised normaliser: (steps, moment1, moment2)
        self.synced = tc.zeros(3, dtype=tc.float32, device=dev)

    def normalize(self, rewards: tc.Tensor) -> tc.Tensor:
        """ReturnAcc.forward of the synchronised accumulator on any block of raw rewards (shift=False)."""
        x = rewards.to(self.synced.device, tc.float32).contiguous()
        y = tc.empty_like(x)
        check(self._L.drl_reward_norm_apply_f32(ptr(x), x.numel(), ptr(self.synced), 1e-8, self.clip_low, self.clip_high,
                                                ptr(y), current_stream()), 'reward_norm_apply')
        return y

    def step_block(self, rewards: tc.Tensor, dones: tc.Tensor) -> tc.Tensor:
        """[N, T] raw rewards and dones of the next T steps of every environment -> [N, T] normalised rewards;
        the raw rewards are fed to the un-synchronised accumulators (NormalizeRewardWrapper.step, :186-188)."""
        if rewards.shape != dones.shape or rewards.dim() != 2 or rewards.shape[0] != self.n_envs:
            raise ValueError('rewards and dones must both be [n_envs, T]')
        r = rewards.to(self.synced.device, tc.float32).contiguous()
        d = dones.to(self.synced.device, tc.float32).contiguous()
        out = self.normalize(r)
        T = r.shape[1]
        check(self._L.drl_return_acc_update_f32(ptr(r), T, ptr(d), T, self.n_envs, T, self.gamma, int(self.use_dones),
                                                self.trace_length, ptr(self._win_r), ptr(self._win_d), ptr(self._count),
                                                ptr(self.steps), ptr(self.moment1), ptr(self.moment2), current_stream()),
              'return_acc_update')
        return out

    def learn(self, minibatch=None, **kwargs) -> None:
        """_sync_normalizers_global then _sync_normalizers_local (:196-210)."""
        local = tc.stack([self.steps.sum(), self.moment1.sum(), self.moment2.sum()])
        n = float(self.n_envs)
        if self._comm is not None:
            self._comm.allreduce_(local)
            n *= self._comm.world
        self.synced.copy_(local / n)
        self.steps.fill_(self.synced[0])
        self.moment1.fill_(self.synced[1])
        self.moment2.fill_(self.synced[2])

    @property
    def checkpointables(self):
        """The reference checkpoints the synchronised ReturnAcc's Normalizer buffers (:158-162): an object with
        state_dict() / load_state_dict() under the same keys."""
        from ..utils.checkpoint import DictCheckpointable

        def get():
            s = self.synced.detach().cpu()
            return {'_normalizer._steps': s[0].clone(), '_normalizer._moment1': s[1].clone(), '_normalizer._moment2': s[2].clone()}
        return {'reward_normalizer': DictCheckpointable(get, self.load_checkpointables)}

    def load_checkpointables(self, sd: Dict[str, tc.Tensor]) -> None:
        vals = [float(sd['_normalizer._steps']), float(sd['_normalizer._moment1']), float(sd['_normalizer._moment2'])]
        self.synced.copy_(tc.tensor(vals, dtype=tc.float32))
        self.steps.fill_(vals[0]); self.moment1.fill_(vals[1]); self.moment2.fill_(vals[2])
