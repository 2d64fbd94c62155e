"""Functional wrappers over the C-ABI kernels (tensors in, tensors out; current CUDA stream).

Each function names the reference interface it stands in for.  Inputs must be CUDA tensors:
there is no CPU path (`_lib.lib()` raises if the CUDA library is missing).
"""
import ctypes
from typing import Dict, Optional, Sequence

import torch as tc

from . import _lib
from ._lib import check, current_stream, ptr


def _cuda(t: tc.Tensor, dtype, name: str) -> tc.Tensor:
    if not isinstance(t, tc.Tensor) or not t.is_cuda:
        raise ValueError(f'{name} must be a CUDA tensor (drl_b200 has no CPU fallback)')
    if t.dtype != dtype:
        raise TypeError(f'{name} must have dtype {dtype}, got {t.dtype}')
    return t


def gae(rewards: tc.Tensor, vpreds: tc.Tensor, dones: tc.Tensor, gamma: float, lambda_: float,
        use_dones: bool, standardize: bool, rollout_len: Optional[int] = None,
        adv_out: Optional[tc.Tensor] = None, vtarg_out: Optional[tc.Tensor] = None):
    """GAE.call + vtarg + optional per-env standardize (credit_assignment.py:195-214,
    abstract.py:427-439).  rewards/dones [N, >=T], vpreds [N, >=T+1] (row strides kept);
    returns (advantages [N, T], vtargs [N, T])."""
    r, v, d = _cuda(rewards, tc.float32, 'rewards'), _cuda(vpreds, tc.float32, 'vpreds'), _cuda(dones, tc.float32, 'dones')
    if r.dim() != 2 or v.dim() != 2 or d.dim() != 2:
        raise ValueError('rewards/vpreds/dones must be [n_envs, time]')
    T = rollout_len if rollout_len is not None else r.shape[1]
    n = r.shape[0]
    if v.shape[0] != n or d.shape[0] != n or v.shape[1] < T + 1 or r.shape[1] < T or d.shape[1] < T:
        raise ValueError('inconsistent shapes for GAE')
    ld = lambda t: t.stride(0) if t.shape[0] > 1 else t.shape[1]     # size-1 dims carry arbitrary strides
    if r.stride(1) != 1 or v.stride(1) != 1 or d.stride(1) != 1 or ld(d) != ld(r):
        raise ValueError('GAE inputs must be row-contiguous with equal reward/done row strides')
    if adv_out is not None:
        adv, vt = _cuda(adv_out, tc.float32, 'adv_out'), _cuda(vtarg_out, tc.float32, 'vtarg_out')
        if adv.shape[0] != n or adv.shape[1] < T or ld(adv) != ld(vt) or adv.stride(1) != 1:
            raise ValueError('adv_out/vtarg_out must be [n_envs, >=T] with equal strides')
        ld_o = ld(adv)
    else:
        ld_o = (T + 3) // 4 * 4
        adv = tc.empty((n, ld_o), dtype=tc.float32, device=r.device)
        vt = tc.empty((n, ld_o), dtype=tc.float32, device=r.device)
    check(_lib.lib().drl_gae_f32(ptr(r), ptr(v), ptr(d), n, T, ld(r), ld(v), ld_o,
                                 gamma, lambda_, int(use_dones), int(standardize), ptr(adv), ptr(vt),
                                 current_stream()), 'gae')
    return adv[:, :T], vt[:, :T]


def standardize(x: tc.Tensor) -> tc.Tensor:
    """drl/utils/stats.py:4-17 on the last dimension of a 1-D or 2-D CUDA tensor."""
    x = _cuda(x, tc.float32, 'tensor')
    x2 = x.reshape(1, -1) if x.dim() == 1 else x
    x2 = x2.contiguous()
    y = tc.empty_like(x2)
    check(_lib.lib().drl_standardize_f32(ptr(x2), x2.shape[0], x2.shape[1], x2.shape[1], ptr(y),
                                         current_stream()), 'standardize')
    return y.reshape(x.shape)


def make_hyper(num_actions: int, reward_weights: Sequence[float], clip_param: float, ent_coef: float,
               vf_coef: float, vf_loss_clipping: bool, vf_simple_weighting: bool,
               vf_loss_cls: str = 'MSELoss') -> _lib.PpoHyper:
    kinds = {'MSELoss': 0, 'SmoothL1Loss': 1}
    if vf_loss_cls not in kinds:
        raise ValueError(f'unsupported vf_loss_cls {vf_loss_cls}')      # losses.py:15-17 getattr failure
    h = _lib.PpoHyper()
    h.num_actions, h.num_streams = num_actions, len(reward_weights)
    for i, w in enumerate(reward_weights):
        h.reward_weights[i] = float(w)
    h.clip_param, h.ent_coef, h.vf_coef = float(clip_param), float(ent_coef), float(vf_coef)
    h.vf_loss_clipping, h.vf_simple_weighting = int(vf_loss_clipping), int(vf_simple_weighting)
    h.vf_loss_kind = kinds[vf_loss_cls]
    return h


def make_batch(actions: tc.Tensor, logp_old: tc.Tensor, adv: Sequence[tc.Tensor],
               v_old: Sequence[Optional[tc.Tensor]], vtarg: Sequence[tc.Tensor]) -> _lib.PpoBatch:
    b = _lib.PpoBatch()
    b.actions = ptr(_cuda(actions, tc.int64, 'actions'))
    b.logp_old = ptr(_cuda(logp_old, tc.float32, 'logprobs'))
    for k in range(len(adv)):
        b.adv[k] = ptr(_cuda(adv[k], tc.float32, 'advantages'))
        b.vtarg[k] = ptr(_cuda(vtarg[k], tc.float32, 'vtargs'))
        b.v_old[k] = ptr(v_old[k]) if v_old[k] is not None else None
    # the struct carries raw device pointers: keep the tensors alive as long as the struct (a caller that passes temporaries
    # would otherwise hand the kernels freed memory)
    b._keepalive = (actions, logp_old, tuple(adv), tuple(v_old), tuple(vtarg))
    return b


LOSS_NAMES = ('policy_loss', 'value_loss', 'composite_loss', 'meanent', 'clipfrac')


def ppo_loss(logits: tc.Tensor, values: tc.Tensor, batch: _lib.PpoBatch, hyper: _lib.PpoHyper,
             row_idx: Optional[tc.Tensor] = None, want_grad: bool = True):
    """Fused losses + backward on given logits [nb, A] / values [nb, K] (ppo.py:310-417, 195-207).
    Returns (losses[5] tensor, logp, entropy, dlogits|None, dvalues|None)."""
    logits = _cuda(logits, tc.float32, 'logits').contiguous()
    values = _cuda(values, tc.float32, 'values').contiguous()
    nb = logits.shape[0]
    dev = logits.device
    losses = tc.empty(5, dtype=tc.float32, device=dev)
    logp = tc.empty(nb, dtype=tc.float32, device=dev)
    ent = tc.empty(nb, dtype=tc.float32, device=dev)
    dl = tc.empty_like(logits) if want_grad else None
    dv = tc.empty_like(values) if want_grad else None
    check(_lib.lib().drl_ppo_loss_fwd_bwd(ptr(logits), ptr(values), ptr(row_idx), nb,
                                          ctypes.byref(batch), ctypes.byref(hyper), ptr(losses),
                                          ptr(logp), ptr(ent), ptr(dl), ptr(dv), current_stream()),
          'ppo_loss')
    return losses, logp, ent, dl, dv


def adam_step(params: tc.Tensor, grads: tc.Tensor, exp_avg: tc.Tensor, exp_avg_sq: tc.Tensor,
              step: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, grad_scale: float = 1.0):
    """torch.optim.Adam.step on flat fp32 CUDA buffers (optimization.py:29-35)."""
    for t, n in ((params, 'params'), (grads, 'grads'), (exp_avg, 'exp_avg'), (exp_avg_sq, 'exp_avg_sq')):
        _cuda(t, tc.float32, n)
        if not t.is_contiguous():
            raise ValueError(f'{n} must be contiguous')
    check(_lib.lib().drl_adam_step(ptr(params), ptr(grads), ptr(exp_avg), ptr(exp_avg_sq),
                                   params.numel(), lr, betas[0], betas[1], eps, step, grad_scale,
                                   current_stream()), 'adam')


def normalizer_apply(x: tc.Tensor, m1: tc.Tensor, m2: tc.Tensor, clip_low=-5.0, clip_high=5.0,
                     eps: float = 1e-8) -> tc.Tensor:
    """Normalizer.forward (normalize.py:129-162); x [n, *shape], moments [*shape]."""
    x = _cuda(x, tc.float32, 'item').contiguous()
    d = m1.numel()
    y = tc.empty_like(x)
    lo = clip_low if clip_low is not None else -float('inf')
    hi = clip_high if clip_high is not None else float('inf')
    check(_lib.lib().drl_normalizer_apply_f32(ptr(x), ptr(m1), ptr(m2), x.numel() // d, d, eps, lo, hi,
                                              ptr(y), current_stream()), 'normalizer_apply')
    return y


def normalizer_update(m1: tc.Tensor, m2: tc.Tensor, items: tc.Tensor, steps_before: float) -> float:
    """Normalizer.update applied to items[0], items[1], ... in order (normalize.py:105-127)."""
    items = _cuda(items, tc.float32, 'items').contiguous()
    d = m1.numel()
    count = items.numel() // d
    check(_lib.lib().drl_normalizer_update_f32(ptr(m1), ptr(m2), ptr(items), count, d, float(steps_before),
                                               current_stream()), 'normalizer_update')
    return steps_before + count
