"""Drop-in for the advantage estimators of drl/algos/common/credit_assignment.py."""
import abc
import importlib

import torch as tc

from .. import _lib, ops
from .._lib import check, current_stream, ptr


def get_credit_assignment_op(cls_name, cls_args):
    """Name lookup like credit_assignment.py:28-42 (ValueError when the class is unknown)."""
    module = importlib.import_module('pytorch_drl_b200.algos.credit_assignment')
    if not hasattr(module, cls_name):
        raise ValueError(f'Unsupported credit assignment op: {cls_name}')
    return getattr(module, cls_name)(**cls_args)


class AdvantageEstimator(metaclass=abc.ABCMeta):
    def __init__(self, gamma: float, use_dones: bool):
        self._gamma = gamma
        self._use_dones = use_dones

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)


class GAE(AdvantageEstimator):
    """Generalized Advantage Estimation with the reference's constructor and call signature
    (credit_assignment.py:172-214).  Inputs are CUDA tensors of shape [T+n-1] / [T+n] (one
    trajectory, as in the reference) or [N, .] (N envs at once, one warp each)."""

    def __init__(self, gamma: float, use_dones: bool, lambda_: float):
        super().__init__(gamma, use_dones)
        self._lambda = lambda_

    def call(self, rollout_len: int, extra_steps: int, rewards: tc.Tensor, vpreds: tc.Tensor,
             dones: tc.Tensor) -> tc.Tensor:
        single = rewards.dim() == 1
        r = rewards.unsqueeze(0) if single else rewards
        v = vpreds.unsqueeze(0) if single else vpreds
        d = dones.unsqueeze(0) if single else dones
        if single:
            r, v, d = r.contiguous(), v.contiguous(), d.contiguous()
        # extra_steps = n - 1 > 0: the scan runs over T + n - 1 steps, the first T are returned (:204-214)
        adv, _ = ops.gae(r, v, d, self._gamma, self._lambda, self._use_dones, False, rollout_len + extra_steps)
        adv = adv[:, :rollout_len]
        return adv[0] if single else adv

    def call_fused(self, rollout_len, rewards, vpreds, dones, standardize, adv_out=None, vtarg_out=None):
        """GAE + vtarg + per-env standardisation in one launch (abstract.py:427-439)."""
        return ops.gae(rewards, vpreds, dones, self._gamma, self._lambda, self._use_dones, standardize,
                       rollout_len, adv_out, vtarg_out)


def _rows(t: tc.Tensor, single: bool) -> tc.Tensor:
    t = t.unsqueeze(0) if single else t
    return t.to(tc.float32).contiguous()


class NStepAdvantageEstimator(AdvantageEstimator):
    """N-step advantages with the reference's constructor and call signature (credit_assignment.py:215-249).
    Inputs: CUDA tensors [T+n-1] / [T+n] (one trajectory) or [N, .] (N envs per launch), n = extra_steps + 1."""

    def call(self, rollout_len: int, extra_steps: int, rewards: tc.Tensor, vpreds: tc.Tensor,
             dones: tc.Tensor) -> tc.Tensor:
        single = rewards.dim() == 1
        n = extra_steps + 1
        r, v, d = _rows(rewards, single), _rows(vpreds, single), _rows(dones, single)
        d = d[:, :r.shape[1]].contiguous() if d.shape[1] != r.shape[1] else d        # the reference accepts longer dones
        if r.shape[1] < rollout_len + n - 1 or v.shape[1] < rollout_len + n:
            raise ValueError('rewards need T+n-1 and vpreds T+n entries')
        out = tc.empty((r.shape[0], rollout_len), dtype=tc.float32, device=r.device)
        check(_lib.lib().drl_nstep_advantages_f32(ptr(r), ptr(v), ptr(d), r.shape[0], rollout_len,
                                                     extra_steps, r.shape[1], v.shape[1], rollout_len, self._gamma,
                                                     int(self._use_dones), ptr(out), current_stream()),
                  'nstep_advantages')
        return out[0] if single else out


class BellmanOperator(AdvantageEstimator):
    pass


class SimpleDiscreteBellmanOptimalityOperator(BellmanOperator):
    """n-step Bellman optimality backups, optionally double-Q (credit_assignment.py:252-305).  qpreds / tgt_qpreds
    are [T+n, A] or [N, T+n, A]."""

    def __init__(self, gamma: float, use_dones: bool, double_q: bool):
        super().__init__(gamma, use_dones)
        self._double_q = double_q

    def call(self, rollout_len: int, extra_steps: int, rewards: tc.Tensor, qpreds: tc.Tensor,
             tgt_qpreds: tc.Tensor, dones: tc.Tensor) -> tc.Tensor:
        single = rewards.dim() == 1
        n = extra_steps + 1
        r, d = _rows(rewards, single), _rows(dones, single)
        d = d[:, :r.shape[1]].contiguous() if d.shape[1] != r.shape[1] else d
        q, qt = _rows(qpreds, single), _rows(tgt_qpreds, single)
        if q.shape != qt.shape or q.dim() != 3:
            raise ValueError('qpreds and tgt_qpreds must have the same [.., T+n, A] shape')
        if r.shape[1] < rollout_len + n - 1 or q.shape[1] < rollout_len + n:
            raise ValueError('rewards need T+n-1 and q tables T+n time steps')
        out = tc.empty((r.shape[0], rollout_len), dtype=tc.float32, device=r.device)
        check(_lib.lib().drl_bellman_backups_f32(ptr(r), ptr(q), ptr(qt), ptr(d), r.shape[0], rollout_len,
                                                    extra_steps, q.shape[2], r.shape[1], q.shape[1], rollout_len,
                                                    self._gamma, int(self._use_dones), int(self._double_q), ptr(out),
                                                    current_stream()), 'bellman_backups')
        return out[0] if single else out
