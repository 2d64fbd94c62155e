"""Prioritized-replay sum-tree on the GPU (Schaul et al. 2015, proportional variant).

The reference lists the paper as supported (README.md:15) but ships no replay buffer, sum-tree or
DQN learner, so this API is the builder's; its arithmetic contract is
`oracle/sumtree.c` and results are bit-exact against it ("parity unpinned" w.r.t. the reference).
"""
from typing import Optional, Tuple

import torch as tc

from .. import _lib
from .._lib import check, current_stream, ptr


class SumTree:
    """fp32 implicit-heap sum-tree with `capacity` leaves (power of two) in one CUDA tensor."""

    def __init__(self, capacity: int, device=None):
        if capacity < 1 or capacity & (capacity - 1):
            raise ValueError('capacity must be a power of two')
        if not tc.cuda.is_available():
            raise RuntimeError('SumTree needs a CUDA device; there is no CPU fallback')
        self.capacity = capacity
        self.device = tc.device('cuda', tc.cuda.current_device()) if device is None else tc.device(device)
        self.tree = tc.zeros(2 * capacity, dtype=tc.float32, device=self.device)
        self._L = _lib.lib()

    @property
    def total(self) -> tc.Tensor:
        return self.tree[1]

    @property
    def leaves(self) -> tc.Tensor:
        return self.tree[self.capacity:]

    def set_leaves(self, priorities: tc.Tensor) -> None:
        self.tree[self.capacity:self.capacity + priorities.numel()] = priorities.to(self.device, tc.float32)
        check(self._L.drl_sumtree_rebuild(ptr(self.tree), self.capacity, current_stream()), 'sumtree_rebuild')

    def update(self, idx: tc.Tensor, priorities: tc.Tensor) -> None:
        """leaf[idx[j]] = priorities[j] (highest j wins on duplicates) + ancestor recompute."""
        if idx.dtype != tc.int64 or priorities.dtype != tc.float32 or not idx.is_cuda or not priorities.is_cuda:
            raise TypeError('idx must be int64 CUDA, priorities float32 CUDA')
        if idx.numel() != priorities.numel():
            raise ValueError('idx and priorities differ in length')
        check(self._L.drl_sumtree_update(ptr(self.tree), self.capacity, ptr(idx.contiguous()),
                                         ptr(priorities.contiguous()), idx.numel(), current_stream()),
              'sumtree_update')

    def sample(self, u: tc.Tensor) -> Tuple[tc.Tensor, tc.Tensor]:
        """Prefix-sum descent for masses u in [0, total): returns (leaf indices, priorities)."""
        if u.dtype != tc.float32 or not u.is_cuda:
            raise TypeError('u must be float32 CUDA')
        idx = tc.empty(u.numel(), dtype=tc.int64, device=self.device)
        pr = tc.empty(u.numel(), dtype=tc.float32, device=self.device)
        check(self._L.drl_sumtree_sample(ptr(self.tree), self.capacity, ptr(u.contiguous()), u.numel(),
                                         ptr(idx), ptr(pr), current_stream()), 'sumtree_sample')
        return idx, pr

    def stratified_u(self, uniform01: tc.Tensor) -> tc.Tensor:
        out = tc.empty_like(uniform01)
        check(self._L.drl_sumtree_stratified_u(ptr(self.tree), ptr(uniform01.contiguous()), uniform01.numel(),
                                               ptr(out), current_stream()), 'sumtree_stratified_u')
        return out


class PrioritizedReplayIndex:
    """Index side of proportional prioritized replay: priorities p_i^alpha in a SumTree, stratified
    sampling of a batch, importance weights (N * P(i))^-beta / max_w, priority updates."""

    def __init__(self, capacity: int, alpha: float = 0.6, eps: float = 1e-6, device=None,
                 generator: Optional[tc.Generator] = None):
        self.tree = SumTree(capacity, device)
        self.alpha, self.eps = alpha, eps
        self.size = 0
        self._gen = generator

    def add(self, idx: tc.Tensor, td_error: Optional[tc.Tensor] = None, max_priority: float = 1.0) -> None:
        pr = tc.full((idx.numel(),), max_priority, dtype=tc.float32, device=self.tree.device) \
            if td_error is None else (td_error.abs() + self.eps).pow(self.alpha).float()
        self.tree.update(idx, pr)
        self.size = min(self.tree.capacity, max(self.size, int(idx.max().item()) + 1))

    def sample(self, batch_size: int, beta: float = 0.4, uniform01: Optional[tc.Tensor] = None):
        if uniform01 is None:
            uniform01 = tc.rand(batch_size, device=self.tree.device, generator=self._gen)
        u = self.tree.stratified_u(uniform01.float())
        idx, pr = self.tree.sample(u)
        prob = pr / self.tree.total
        w = (self.size * prob).pow(-beta)
        return idx, pr, w / w.max()

    def update_priorities(self, idx: tc.Tensor, td_error: tc.Tensor) -> None:
        self.tree.update(idx, (td_error.abs() + self.eps).pow(self.alpha).float())
