"""Learner half of drl/envs/wrappers/stateful/intrinsic/random_network_distillation.py on the GPU.

`RandomNetworkDistillation.learn(mb)` is what `update_trainable_wrappers` calls once per PPO
minibatch (drl/algos/common/wrapper_ops.py:21-25; wrapper `learn`, :209-235): random row dropping,
last-channel + normalise + clip (fused into the first conv), teacher/student forward, predictor MSE,
student backward, gradient all-reduce, Adam, normaliser sync.  `intrinsic_reward(frames)` is the
batched form of the per-step reward of `step` (:186-201).  Env stepping itself stays out of scope.
"""
import ctypes
from typing import Dict, Mapping, Optional

import numpy as np
import torch as tc

from .. import _lib, ops
from .._lib import check, current_stream, ptr
from .normalize import Normalizer

TEACHER_KEYS = ['_network.0._network.1.weight', '_network.0._network.1.bias',
                '_network.0._network.3.weight', '_network.0._network.3.bias',
                '_network.0._network.5.weight', '_network.0._network.5.bias',
                '_network.1.weight', '_network.1.bias']
STUDENT_KEYS = TEACHER_KEYS[:6] + ['_network.1.weight', '_network.1.bias', '_network.3.weight',
                                   '_network.3.bias', '_network.5.weight', '_network.5.bias']
_TRUNK_SHAPES = [(16, 1, 8, 8), (16,), (32, 16, 4, 4), (32,), (32, 32, 4, 4), (32,)]
TEACHER_SHAPES = _TRUNK_SHAPES + [(512, 1152), (512,)]
STUDENT_SHAPES = _TRUNK_SHAPES + [(256, 1152), (256,), (256, 256), (256,), (512, 256), (512,)]


def _views(flat: tc.Tensor, keys, shapes) -> Dict[str, tc.Tensor]:
    out, o = {}, 0
    for k, s in zip(keys, shapes):
        n = int(np.prod(s))
        out[k] = flat[o:o + n].view(s)
        o += n
    assert o == flat.numel()
    return out


class RandomNetworkDistillation:
    def __init__(self, rnd_optimizer_args: Mapping[str, float], world_size: int, widening: int = 1,
                 non_learning_steps: int = 0, max_batch: int = 1024, comm=None, device=None, seed: int = 0,
                 precision: str = 'bf16', env=None):
        self.env = env                      # next trainable wrapper of the chain (wrapper_ops.py:21-25), e.g. a BatchedRewardNormalizer
        if not tc.cuda.is_available():
            raise RuntimeError('RND needs a CUDA device; there is no CPU fallback')
        self.device = tc.device('cuda', tc.cuda.current_device()) if device is None else tc.device(device)
        self._L = _lib.lib()
        self._world_size, self._comm = world_size, comm
        self._non_learning_steps = non_learning_steps
        self._opt = dict(lr=1e-4, betas=(0.9, 0.999), eps=1e-8)
        self._opt.update(rnd_optimizer_args)
        self._data_shape = [84, 84, 1]
        self._synced_normalizer = Normalizer(self._data_shape, -5, 5, self.device)
        self._unsynced_normalizer = Normalizer(self._data_shape, -5, 5, self.device)
        nt, ns = self._L.drl_rnd_param_count(0, widening), self._L.drl_rnd_param_count(1, widening)
        if nt < 0:
            raise NotImplementedError('only widening=1 is built (atari_rnd/config.yaml:55)')
        f = lambda n: tc.zeros(n, dtype=tc.float32, device=self.device)  # noqa: E731
        self.teacher_params, self.student_params = f(nt), f(ns)
        self.student_grads, self.exp_avg, self.exp_avg_sq = f(ns), f(ns), f(ns)
        self.adam_steps = 0
        self.loss = f(1)
        self._init_weights(seed)
        h = ctypes.c_void_p()
        # 'bf16' = tcgen05 engine (csrc/umma_rnd.cu), 'fp32' = CUDA-core strict-parity engine
        self.precision_name = precision
        prec = {'fp32': _lib.PRECISION_FP32, 'bf16': _lib.PRECISION_BF16}[precision]
        check(self._L.drl_rnd_create(ctypes.byref(h), self.device.index or 0, max_batch, widening, prec), 'rnd_create')
        self._h = h
        self.max_batch = max_batch
        self._bind()

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h:
            self._L.drl_rnd_destroy(h)

    def _bind(self):
        check(self._L.drl_rnd_set_params(self._h, ptr(self.teacher_params), ptr(self.student_params), current_stream()))

    def _init_weights(self, seed: int) -> None:
        """orthogonal_(gain sqrt 2) weights, zero biases (random_network_distillation.py:38-42,58-62)."""
        g = tc.Generator().manual_seed(seed)
        for flat, keys, shapes in ((self.teacher_params, TEACHER_KEYS, TEACHER_SHAPES),
                                   (self.student_params, STUDENT_KEYS, STUDENT_SHAPES)):
            for k, v in _views(flat, keys, shapes).items():
                if v.dim() > 1:
                    w = tc.empty(v.shape)
                    tc.nn.init.orthogonal_(w, gain=2 ** 0.5, generator=g)
                    v.copy_(w)

    # ---- state_dict-compatible access ---------------------------------------------------------------
    def teacher_views(self):
        return _views(self.teacher_params, TEACHER_KEYS, TEACHER_SHAPES)

    def student_views(self, of: Optional[tc.Tensor] = None):
        return _views(self.student_params if of is None else of, STUDENT_KEYS, STUDENT_SHAPES)

    def load(self, teacher: Mapping[str, tc.Tensor], student: Mapping[str, tc.Tensor]) -> None:
        for k, v in self.teacher_views().items():
            v.copy_(tc.as_tensor(teacher[k]).to(self.device))
        for k, v in self.student_views().items():
            v.copy_(tc.as_tensor(student[k]).to(self.device))
        self._bind()                      # the tcgen05 engine keeps packed operand copies of the parameters

    @property
    def checkpointables(self):
        """random_network_distillation.py:175-184: objects with state_dict() / load_state_dict() under the reference's key
        names (the networks are DDP-wrapped there: `module.` prefix; the optimizer is a torch.optim.Adam state_dict)."""
        from ..utils.checkpoint import FlatAdamCheckpointable, ViewsCheckpointable

        def set_steps(n):
            self.adam_steps = n
        return {'rnd_observation_normalizer': self._synced_normalizer,
                'rnd_teacher_net': ViewsCheckpointable(self.teacher_views, 'module.', self._bind),
                'rnd_student_net': ViewsCheckpointable(self.student_views, 'module.', self._bind),
                'rnd_optimizer': FlatAdamCheckpointable(STUDENT_SHAPES, self.exp_avg, self.exp_avg_sq, lambda: self.adam_steps,
                                                        set_steps, lambda: self._opt)}

    # ---- actor side, batched ----------------------------------------------------------------------------
    def intrinsic_reward(self, frames_u8: tc.Tensor, row_idx: Optional[tc.Tensor] = None) -> tc.Tensor:
        """mean_d (teacher - student)^2 of the normalised last channel of each frame (:193-196)."""
        nb = int(row_idx.numel() if row_idx is not None else frames_u8.shape[0])
        out = tc.empty(nb, dtype=tc.float32, device=self.device)
        nz = self._synced_normalizer
        for s in range(0, nb, self.max_batch):          # the handle's workspace holds max_batch frames per call
            m = min(self.max_batch, nb - s)
            ridx = row_idx[s:s + m] if row_idx is not None else tc.arange(s, s + m, device=self.device, dtype=tc.int64)
            check(self._L.drl_rnd_reward(self._h, ptr(frames_u8), ptr(ridx), m, ptr(nz.moment1), ptr(nz.moment2),
                                         ptr(out[s:s + m]), current_stream()), 'rnd_reward')
        return out

    def observe(self, last_channel_frames: tc.Tensor) -> None:
        """`self._unsynced_normalizer.update(obs_)` for a stack of frames [n,84,84,1] in order (:192)."""
        self._unsynced_normalizer.update(last_channel_frames)

    # ---- learner side ---------------------------------------------------------------------------------
    def _random_drop(self, rows: tc.Tensor) -> tc.Tensor:
        keepfrac = min(self._world_size / 32, 1)                      # :203-207
        keepnum = int(keepfrac * rows.numel())
        idxs = np.random.permutation(rows.numel())[:keepnum]
        return rows[tc.as_tensor(idxs, device=rows.device)]

    def learn(self, mb, apply_dropping: bool = True, **kwargs) -> None:
        rows = mb.row_idx if hasattr(mb, 'row_idx') else None
        obs = mb['observations']
        if self._synced_normalizer.steps >= self._non_learning_steps:
            if rows is None:
                rows = tc.arange(obs.shape[0], device=self.device, dtype=tc.int64)
            if apply_dropping:
                rows = self._random_drop(rows)
            if rows.numel() > 0:
                self.learn_rows(obs, rows)
        self._sync_normalizers_global()
        self._sync_normalizers_local()

    def learn_rows(self, obs_u8: tc.Tensor, rows: tc.Tensor) -> tc.Tensor:
        """Predictor MSE step on the given frame rows; returns the device loss scalar."""
        nz = self._synced_normalizer
        check(self._L.drl_rnd_learn(self._h, ptr(obs_u8), ptr(rows), rows.numel(), ptr(nz.moment1), ptr(nz.moment2),
                                    ptr(self.student_grads), ptr(self.loss), current_stream()), 'rnd_learn')
        world = 1
        if self._comm is not None:
            self._comm.allreduce_(self.student_grads)
            world = self._comm.world
        self.adam_steps += 1
        b1, b2 = self._opt['betas']
        check(self._L.drl_rnd_adam_step(self._h, ptr(self.student_params), ptr(self.student_grads), ptr(self.exp_avg),
                                        ptr(self.exp_avg_sq), float(self._opt['lr']), float(b1), float(b2), float(self._opt['eps']),
                                        self.adam_steps, 1.0 / world, current_stream()), 'rnd_adam_step')
        return self.loss

    def _sync_normalizers_global(self) -> None:
        """global_mean of steps / moment1 / moment2 (:159-168) as ONE coalesced all-reduce."""
        u, s = self._unsynced_normalizer, self._synced_normalizer
        if self._comm is not None and self._comm.world > 1:
            flat = tc.cat([tc.tensor([u.steps], device=self.device, dtype=tc.float32), u.moment1.reshape(-1),
                           u.moment2.reshape(-1)])
            flat = self._comm.mean_(flat)
            s.steps = float(flat[0].item())
            n = u.moment1.numel()
            s.moment1 = flat[1:1 + n].view_as(u.moment1).clone()
            s.moment2 = flat[1 + n:].view_as(u.moment2).clone()
        else:
            s.steps = u.steps
            s.moment1 = u.moment1.clone()
            s.moment2 = u.moment2.clone()

    def _sync_normalizers_local(self) -> None:
        """The reference assigns the SAME tensor objects (:170-173), so later in-place updates of the
        unsynced moments also move the synced ones; reproduced by aliasing the tensors."""
        u, s = self._unsynced_normalizer, self._synced_normalizer
        u.steps, u.moment1, u.moment2 = s.steps, s.moment1, s.moment2
