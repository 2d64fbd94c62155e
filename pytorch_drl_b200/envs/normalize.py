"""Drop-in for drl/envs/wrappers/stateful/normalize.py (`Normalizer`) on the GPU."""
from typing import List, Optional

import torch as tc

from .. import ops


class Normalizer:
    """Running first/second non-central moments of a data stream (normalize.py:9-162).
    `steps`, `moment1`, `moment2` are public like the reference's properties; the moments are fp32
    CUDA tensors updated in place by one kernel (items applied in order)."""

    def __init__(self, data_shape: List[int], clip_low: Optional[float] = None,
                 clip_high: Optional[float] = None, device=None):
        if not tc.cuda.is_available():
            raise RuntimeError('Normalizer needs a CUDA device; there is no CPU fallback')
        self._data_shape = list(data_shape)
        self._clip_low, self._clip_high = clip_low, clip_high
        dev = tc.device('cuda', tc.cuda.current_device()) if device is None else tc.device(device)
        self.steps = 0.0
        self.moment1 = tc.zeros(data_shape, dtype=tc.float32, device=dev)
        self.moment2 = tc.zeros(data_shape, dtype=tc.float32, device=dev)

    def update(self, item: tc.Tensor) -> None:
        """One item of shape data_shape, or a stack [n, *data_shape] applied in order."""
        self.steps = ops.normalizer_update(self.moment1, self.moment2, item.to(self.moment1.device, tc.float32),
                                           self.steps)

    def __call__(self, item: tc.Tensor, shift: bool = True, scale: bool = True, eps: float = 1e-8) -> tc.Tensor:
        if not (shift and scale):
            raise NotImplementedError('only the shift+scale form is fused')
        return ops.normalizer_apply(item, self.moment1, self.moment2, self._clip_low, self._clip_high, eps)

    forward = __call__

    def state_dict(self):
        return {'_steps': tc.tensor(self.steps), '_moment1': self.moment1.clone(), '_moment2': self.moment2.clone()}

    def load_state_dict(self, sd):
        self.steps = float(sd['_steps'])
        self.moment1.copy_(sd['_moment1'])
        self.moment2.copy_(sd['_moment2'])
