"""Drop-in for the advantage estimators of drl/algos/common/credit_assignment.py."""
import abc
import importlib

import torch as tc

from .. import ops


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
        if extra_steps != 0:
            # GAE ignores the extra steps' targets only through A_{T+n-1}=0; both shipped configs
            # use extra_steps: 0 (models_dir/atari_ppo/config.yaml:10).
            raise NotImplementedError('GAE with extra_steps > 0 is not built yet')
        single = rewards.dim() == 1
        r = rewards.unsqueeze(0) if single else rewards
        v = vpreds.unsqueeze(0) if single else vpreds
        d = dones.unsqueeze(0) if single else dones
        if single:
            r, v, d = r.contiguous(), v.contiguous(), d.contiguous()
        adv, _ = ops.gae(r, v, d, self._gamma, self._lambda, self._use_dones, False, rollout_len)
        return adv[0] if single else adv

    def call_fused(self, rollout_len, rewards, vpreds, dones, standardize, adv_out=None, vtarg_out=None):
        """GAE + vtarg + per-env standardisation in one launch (abstract.py:427-439)."""
        return ops.gae(rewards, vpreds, dones, self._gamma, self._lambda, self._use_dones, standardize,
                       rollout_len, adv_out, vtarg_out)
