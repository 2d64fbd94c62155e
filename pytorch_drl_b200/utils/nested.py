"""Drop-in for drl/utils/nested.py.

`slice_nested_tensor(rollout, indices)` of the reference materialises a gathered copy of every
tensor (3.6 MB of frames per 32-sample minibatch, nested.py:6-12).  Here an index array applied to
a device rollout returns a `MinibatchView`: the rollout plus the row indices, which the kernels
consume directly (the gather is folded into conv1's loads and the loss kernel).
"""
from typing import Mapping, Union

import numpy as np
import torch as tc


class MinibatchView(dict):
    """A rollout dict plus `row_idx` (int64 CUDA tensor of flat row numbers)."""

    def __init__(self, rollout: Mapping, row_idx: tc.Tensor):
        super().__init__(rollout)
        self.row_idx = row_idx

    def materialize(self):
        """Gathered copy, with the reference's semantics (for tests / interop)."""
        def g(v):
            if isinstance(v, tc.Tensor):
                return v.reshape((-1,) + tuple(v.shape[2:]))[self.row_idx]
            return {k: g(x) for k, x in v.items()}
        return {k: g(v) for k, v in self.items()}


def slice_nested_tensor(nt, indices):
    if isinstance(indices, tc.Tensor) and indices.dtype == tc.int64 and indices.is_cuda \
            and isinstance(nt, Mapping) and 'observations' in nt:
        return MinibatchView(nt, indices)
    if isinstance(nt, tc.Tensor):
        if isinstance(indices, np.ndarray):
            indices = tc.as_tensor(indices, device=nt.device)
        return nt[indices]
    return {k: slice_nested_tensor(nt[k], indices) for k in nt}


def clone_nested_tensor(nt):
    if isinstance(nt, tc.Tensor):
        return nt.clone()
    return {k: clone_nested_tensor(nt[k]) for k in nt}
