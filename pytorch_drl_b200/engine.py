"""`LearnerEngine`: the Python owner of one C-ABI NatureCNN agent handle (`drl_net_t`) and of the
flat parameter / gradient / Adam-state buffers it works on.

The buffers use the reference's own parameter order and layouts (Agent.parameters()), so the
torch views handed out by `named_views()` load and save reference state_dicts unchanged
(drl/utils/checkpointing.py:64; key names from drl/agents/integration/agent.py:34-36).
"""
import ctypes
from typing import Dict, List, Optional, Sequence

import torch as tc

from . import _lib, ops
from ._lib import check, current_stream, ptr

TRUNK_NAMES = ['_architecture._network.0.weight', '_architecture._network.0.bias',
               '_architecture._network.2.weight', '_architecture._network.2.bias',
               '_architecture._network.4.weight', '_architecture._network.4.bias',
               '_architecture._network.7.weight', '_architecture._network.7.bias']
TRUNK_SHAPES = [(32, 4, 8, 8), (32,), (64, 32, 4, 4), (64,), (64, 64, 3, 3), (64,), (512, 3136), (512,)]


def param_names(value_keys: Sequence[str]) -> List[str]:
    names = TRUNK_NAMES + ['_predictors.policy._policy_head._network.0.weight',
                           '_predictors.policy._policy_head._network.0.bias']
    for k in value_keys:
        names += [f'_predictors.{k}._value_head._network.0.weight',
                  f'_predictors.{k}._value_head._network.0.bias']
    return names


class LearnerEngine:
    def __init__(self, num_actions: int, value_keys: Sequence[str], max_batch: int,
                 precision: str = 'bf16', device: Optional[tc.device] = None):
        if not tc.cuda.is_available():
            raise RuntimeError('drl_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        self.device = tc.device('cuda', tc.cuda.current_device()) if device is None else tc.device(device)
        L = _lib.lib()
        _lib.require_device(self.device.index or 0)
        self.A, self.value_keys, self.K = int(num_actions), list(value_keys), len(value_keys)
        self.max_batch = int(max_batch)
        self.precision = {'fp32': _lib.PRECISION_FP32, 'bf16': _lib.PRECISION_BF16}[precision]
        self.precision_name = precision
        n = L.drl_net_param_count(self.A, self.K)
        if n < 0:
            raise ValueError('unsupported number of actions / value heads')
        self.nparams = int(n)
        offs = (ctypes.c_int64 * (2 * (5 + self.K) + 1))()
        check(L.drl_net_param_offsets(self.A, self.K, offs, len(offs)), 'param_offsets')
        self.offsets = list(offs)
        self.params = tc.zeros(self.nparams, dtype=tc.float32, device=self.device)
        self.grads = tc.zeros(self.nparams, dtype=tc.float32, device=self.device)
        self.exp_avg = tc.zeros(self.nparams, dtype=tc.float32, device=self.device)
        self.exp_avg_sq = tc.zeros(self.nparams, dtype=tc.float32, device=self.device)
        self.adam_steps = 0
        self.losses = tc.zeros(5, dtype=tc.float32, device=self.device)
        h = ctypes.c_void_p()
        check(L.drl_net_create(ctypes.byref(h), self.device.index or 0, self.max_batch, self.A, self.K,
                               self.precision), 'net_create')
        self._h = h
        self._L = L
        self.names = param_names(self.value_keys)
        self.shapes = TRUNK_SHAPES + [(self.A, 512), (self.A,)] + [(1, 512), (1,)] * self.K
        self.sync_params()

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h:
            self._L.drl_net_destroy(h)

    # ---- parameters ------------------------------------------------------------------------------
    def named_views(self, of: Optional[tc.Tensor] = None) -> Dict[str, tc.Tensor]:
        """name -> view of the flat buffer `of` (default: parameters) with the reference's shape."""
        buf = self.params if of is None else of
        return {n: buf[self.offsets[i]:self.offsets[i + 1]].view(s)
                for i, (n, s) in enumerate(zip(self.names, self.shapes))}

    def load_named(self, tensors: Dict[str, tc.Tensor]) -> None:
        views = self.named_views()
        for n, v in views.items():
            v.copy_(tc.as_tensor(tensors[n]).to(self.device, tc.float32).reshape(v.shape))
        self.sync_params()

    def sync_params(self) -> None:
        """Tell the handle the flat buffer changed (refreshes packed bf16 operand copies)."""
        check(self._L.drl_net_set_params(self._h, ptr(self.params), current_stream()), 'set_params')

    # ---- compute ---------------------------------------------------------------------------------
    def forward(self, obs_u8: tc.Tensor, row_idx: Optional[tc.Tensor] = None, nb: Optional[int] = None,
                actions: Optional[tc.Tensor] = None):
        """No-grad forward: returns (logits [nb,A], values [nb,K], logp|None, entropy)."""
        ops._cuda(obs_u8, tc.uint8, 'observations')
        nb = int(nb if nb is not None else (row_idx.numel() if row_idx is not None else obs_u8.shape[0]))
        out = []
        for s in range(0, nb, self.max_batch):
            m = min(self.max_batch, nb - s)
            logits = tc.empty((m, self.A), dtype=tc.float32, device=self.device)
            values = tc.empty((m, self.K), dtype=tc.float32, device=self.device)
            logp = tc.empty(m, dtype=tc.float32, device=self.device) if actions is not None else None
            ent = tc.empty(m, dtype=tc.float32, device=self.device)
            if row_idx is not None:
                ridx, o = row_idx[s:s + m], obs_u8
            else:
                ridx = tc.arange(s, s + m, device=self.device, dtype=tc.int64)
                o = obs_u8
            check(self._L.drl_net_forward(self._h, ptr(o), ptr(ridx), m, ptr(logits), ptr(values),
                                          ptr(actions), ptr(logp), ptr(ent), current_stream()), 'net_forward')
            out.append((logits, values, logp, ent))
        if len(out) == 1:
            return out[0]
        cat = lambda i: None if out[0][i] is None else tc.cat([o[i] for o in out])  # noqa: E731
        return cat(0), cat(1), cat(2), cat(3)

    def features(self, obs_u8: tc.Tensor, row_idx: Optional[tc.Tensor] = None, nb: Optional[int] = None) -> tc.Tensor:
        """No-grad trunk forward: the [nb, 512] NatureCNN features (for heads outside the fused ones, e.g. the DQN heads)."""
        ops._cuda(obs_u8, tc.uint8, 'observations')
        nb = int(nb if nb is not None else (row_idx.numel() if row_idx is not None else obs_u8.shape[0]))
        feat = tc.empty((nb, 512), dtype=tc.float32, device=self.device)
        for s in range(0, nb, self.max_batch):
            m = min(self.max_batch, nb - s)
            ridx = row_idx[s:s + m] if row_idx is not None else tc.arange(s, s + m, device=self.device, dtype=tc.int64)
            check(self._L.drl_net_features(self._h, ptr(obs_u8), ptr(ridx), m, ptr(feat[s:s + m]), current_stream()),
                  'net_features')
        return feat

    def ppo_step(self, obs_u8: tc.Tensor, row_idx: tc.Tensor, batch: _lib.PpoBatch,
                 hyper: _lib.PpoHyper, grads: bool = True) -> tc.Tensor:
        """forward + fused losses (+ backward into self.grads). Returns the device losses[5]."""
        nb = row_idx.numel()
        if grads:
            check(self._L.drl_net_ppo_step(self._h, ptr(obs_u8), ptr(row_idx), nb, ctypes.byref(batch),
                                           ctypes.byref(hyper), ptr(self.grads), ptr(self.losses),
                                           current_stream()), 'ppo_step')
        else:
            check(self._L.drl_net_ppo_metrics(self._h, ptr(obs_u8), ptr(row_idx), nb, ctypes.byref(batch),
                                              ctypes.byref(hyper), ptr(self.losses), current_stream()),
                  'ppo_metrics')
        return self.losses

    def backward(self, obs_u8: tc.Tensor, row_idx: Optional[tc.Tensor], dlogits: tc.Tensor, dvalues: tc.Tensor,
                 out: Optional[tc.Tensor] = None) -> tc.Tensor:
        """Vector-Jacobian product of the network for a custom loss: d(sum(dlogits*logits) + sum(dvalues*values))/dparams
        as a flat tensor (drl_net_backward; the activations are recomputed, minibatches larger than max_batch are
        processed in chunks and summed)."""
        ops._cuda(obs_u8, tc.uint8, 'observations')
        nb = int(row_idx.numel() if row_idx is not None else obs_u8.shape[0])
        dl = ops._cuda(dlogits.contiguous(), tc.float32, 'dlogits')
        dv = ops._cuda(dvalues.contiguous(), tc.float32, 'dvalues')
        if dl.shape != (nb, self.A) or dv.shape != (nb, self.K):
            raise ValueError('dlogits must be [nb, A] and dvalues [nb, K]')
        total = out if out is not None else tc.zeros(self.nparams, dtype=tc.float32, device=self.device)
        part = tc.empty(self.nparams, dtype=tc.float32, device=self.device)
        for s in range(0, nb, self.max_batch):
            m = min(self.max_batch, nb - s)
            ridx = row_idx[s:s + m] if row_idx is not None else tc.arange(s, s + m, device=self.device, dtype=tc.int64)
            check(self._L.drl_net_backward(self._h, ptr(obs_u8), ptr(ridx), m, ptr(dl[s:s + m]), ptr(dv[s:s + m]), ptr(part),
                                           current_stream()), 'net_backward')
            total += part
        return total

    def adam_step(self, lr: float, betas=(0.9, 0.999), eps: float = 1e-5, grad_scale: float = 1.0) -> None:
        from .utils import checkpoint
        checkpoint.check_flat_adam_binding(self)       # a torch optimizer bound to this state may have been re-loaded
        self.adam_steps += 1
        if getattr(self, '_bound_opt', None) is not None:
            self._bound_opt[2].fill_(float(self.adam_steps))
        check(self._L.drl_net_adam_step(self._h, ptr(self.params), ptr(self.grads), ptr(self.exp_avg),
                                        ptr(self.exp_avg_sq), lr, betas[0], betas[1], eps, self.adam_steps,
                                        grad_scale, current_stream()), 'adam_step')

    def activation(self, which: int, nb: int) -> tc.Tensor:
        """Copy of an intermediate activation of the last step (parity tests only)."""
        p, n, b = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int()
        check(self._L.drl_net_activation(self._h, which, ctypes.byref(p), ctypes.byref(n), ctypes.byref(b)))
        per = n.value // self.max_batch
        dtype = tc.float32 if b.value == 4 else tc.bfloat16
        out = tc.empty(nb * per, dtype=dtype, device=self.device)
        check(self._L.drl_memcpy_d2d(ptr(out), p.value, out.numel() * b.value, current_stream()))
        return out
