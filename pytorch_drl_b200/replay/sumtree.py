"""Prioritized replay on the GPU (Schaul et al. 2015, proportional variant): sum-tree, index, transition ring.

The reference lists the paper as supported (README.md:15) but ships no replay buffer, sum-tree or
DQN learner, so this API is the builder's; its arithmetic contract is
`oracle/sumtree.c` and results are bit-exact against it ("parity unpinned" w.r.t. the reference).
"""
from typing import Dict, Optional, Tuple

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

    def update(self, idx: tc.Tensor, priorities: tc.Tensor, validate: bool = False) -> None:
        """leaf[idx[j]] = priorities[j] (highest j wins on duplicates) + ancestor recompute.  Indices outside [0, capacity)
        are ignored by the kernel; `validate=True` raises instead (one host synchronisation)."""
        if idx.dtype != tc.int64 or priorities.dtype != tc.float32 or not idx.is_cuda or not priorities.is_cuda:
            raise TypeError('idx must be int64 CUDA, priorities float32 CUDA')
        if idx.numel() != priorities.numel():
            raise ValueError('idx and priorities differ in length')
        if validate and idx.numel() and bool(((idx < 0) | (idx >= self.capacity)).any().item()):
            raise ValueError('leaf index out of range')
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

    def per_sample(self, uniform01: tc.Tensor, size: float, beta: float) -> Tuple[tc.Tensor, tc.Tensor, tc.Tensor]:
        """One fused launch: stratified masses, descent, importance weights (size * p / total)^-beta / max (batch <= 1024)."""
        if uniform01.dtype != tc.float32 or not uniform01.is_cuda:
            raise TypeError('uniform01 must be float32 CUDA')
        n = uniform01.numel()
        idx = tc.empty(n, dtype=tc.int64, device=self.device)
        pr = tc.empty(n, dtype=tc.float32, device=self.device)
        w = tc.empty(n, dtype=tc.float32, device=self.device)
        check(self._L.drl_per_sample(ptr(self.tree), self.capacity, ptr(uniform01.contiguous()), n, float(size), float(beta),
                                     ptr(idx), ptr(pr), ptr(w), current_stream()), 'per_sample')
        return idx, pr, w


class PrioritizedReplayIndex:
    """Index side of proportional prioritized replay: priorities p_i^alpha in a SumTree, stratified
    sampling of a batch with importance weights (N * P(i))^-beta / max_w in one kernel, priority updates."""

    def __init__(self, capacity: int, alpha: float = 0.6, eps: float = 1e-6, device=None,
                 generator: Optional[tc.Generator] = None):
        self.tree = SumTree(capacity, device)
        self.alpha, self.eps = alpha, eps
        self.size = 0
        self._gen = generator

    def add(self, idx: tc.Tensor, td_error: Optional[tc.Tensor] = None, max_priority: float = 1.0,
            count: Optional[int] = None) -> None:
        """New transitions enter with the maximal priority (Algorithm 1 line 6) unless TD errors are given.  `count` = number of
        valid transitions afterwards when the caller tracks it (avoids the host read of idx.max())."""
        pr = tc.full((idx.numel(),), max_priority, dtype=tc.float32, device=self.tree.device) \
            if td_error is None else (td_error.abs() + self.eps).pow(self.alpha).float()
        self.tree.update(idx, pr)
        if count is None:
            count = int(idx.max().item()) + 1
        self.size = min(self.tree.capacity, max(self.size, count))

    def sample(self, batch_size: int, beta: float = 0.4, uniform01: Optional[tc.Tensor] = None):
        if uniform01 is None:
            uniform01 = tc.rand(batch_size, device=self.tree.device, generator=self._gen)
        return self.tree.per_sample(uniform01.float(), max(self.size, 1), beta)

    def update_priorities(self, idx: tc.Tensor, td_error: tc.Tensor) -> None:
        self.tree.update(idx, (td_error.abs() + self.eps).pow(self.alpha).float())


class PrioritizedReplayBuffer:
    """Device-resident transition ring + prioritized index (BASELINE config 4: 1 M transitions, batch 512).

    Transition i = (observation stack frames[i] uint8 [84, 84, 4], action, reward, done); its successor stack is frames[i + 1]
    (consecutive writes of one environment stream; at 28 224 B per stack a 2^20-entry ring is 29.6 GB of the B200's 180 GB).
    The NEWEST transition has no successor yet (frames[cursor] still holds the oldest stack, or nothing): it keeps priority 0
    until the next `add` delivers its successor, so it can never be sampled with a wrong s'.
    `sample` returns the batch's indices, importance weights and scalars, the frame ring itself (the fused NatureCNN trunk
    gathers rows by index: no copy is needed to run `QAgent` on the batch) and, on request, gathered copies of the
    observation / successor stacks (`drl_ring_gather`)."""

    def __init__(self, capacity: int, alpha: float = 0.6, eps: float = 1e-6, device=None, generator=None,
                 frame_shape=(84, 84, 4)):
        self.index = PrioritizedReplayIndex(capacity, alpha, eps, device, generator)
        dev = self.index.tree.device
        self.capacity, self.frame_shape = capacity, tuple(frame_shape)
        self.frames = tc.zeros((capacity,) + self.frame_shape, dtype=tc.uint8, device=dev)
        self.actions = tc.zeros(capacity, dtype=tc.int64, device=dev)
        self.rewards = tc.zeros(capacity, dtype=tc.float32, device=dev)
        self.dones = tc.zeros(capacity, dtype=tc.float32, device=dev)
        self.cursor, self.count = 0, 0
        self._open = None              # slot of the newest transition (priority 0 until its successor is written)
        self._L = _lib.lib()

    def add(self, frames: tc.Tensor, actions: tc.Tensor, rewards: tc.Tensor, dones: tc.Tensor, max_priority: float = 1.0) -> tc.Tensor:
        """Append n transitions at the ring cursor (wrapping); they get the maximal priority (Schaul et al. 2015, Algorithm 1
        line 6) except the newest one, which waits for its successor.  Returns their slots."""
        n = frames.shape[0]
        if n > self.capacity:
            raise ValueError('more transitions than the ring holds')
        dev = self.frames.device
        slots = (tc.arange(n, device=dev, dtype=tc.int64) + self.cursor) % self.capacity
        first = min(n, self.capacity - self.cursor)
        for dst, src in ((self.frames, frames), (self.actions, actions), (self.rewards, rewards), (self.dones, dones)):
            src = src.to(dev, dst.dtype)
            dst[self.cursor:self.cursor + first] = src[:first]
            if first < n:
                dst[:n - first] = src[first:]
        self.cursor = (self.cursor + n) % self.capacity
        self.count = min(self.capacity, self.count + n)
        # max priority for the new transitions and for the one that was waiting for its successor; 0 for the newest one
        upd = slots if self._open is None else tc.cat([tc.tensor([self._open], device=dev, dtype=tc.int64), slots])
        pr = tc.full((upd.numel(),), max_priority, dtype=tc.float32, device=dev)
        pr[-1] = 0.0
        self.index.tree.update(upd, pr)
        self.index.size = self.count
        self._open = int((self.cursor - 1) % self.capacity)
        return slots

    def gather_frames(self, idx: tc.Tensor, shift: int = 0) -> tc.Tensor:
        row = 1
        for d in self.frame_shape:
            row *= d
        out = tc.empty((idx.numel(),) + self.frame_shape, dtype=tc.uint8, device=self.frames.device)
        check(self._L.drl_ring_gather(ptr(self.frames), self.capacity, row, ptr(idx.contiguous()), idx.numel(), shift, ptr(out),
                                      current_stream()), 'ring_gather')
        return out

    def sample(self, batch_size: int, beta: float = 0.4, uniform01: Optional[tc.Tensor] = None,
               gather: bool = False) -> Dict[str, tc.Tensor]:
        idx, pr, w = self.index.sample(batch_size, beta, uniform01)
        nxt = (idx + 1) % self.capacity
        out = {'idx': idx, 'next_idx': nxt, 'priorities': pr, 'weights': w, 'frames': self.frames,
               'actions': self.actions[idx], 'rewards': self.rewards[idx], 'dones': self.dones[idx]}
        if gather:
            out['observations'] = self.gather_frames(idx, 0)
            out['next_observations'] = self.gather_frames(idx, 1)
        return out

    def update_priorities(self, idx: tc.Tensor, td_error: tc.Tensor) -> None:
        self.index.update_priorities(idx, td_error)
