"""Drop-in for drl/algos/common/samplers.py (host-side index generation; numpy, like the
reference: the permutation stream IS the parity-critical contract)."""
import abc
from typing import Iterator, Optional

import numpy as np


class Sampler(metaclass=abc.ABCMeta):
    def __init__(self, rollout_len: int) -> None:
        self._rollout_len = rollout_len

    @abc.abstractmethod
    def __iter__(self) -> Iterator:
        pass

    @abc.abstractmethod
    def resample(self) -> None:
        pass


class IndicesIterator:
    def __init__(self, indices: np.ndarray, stepsize: int) -> None:
        assert len(indices.shape) == 1
        assert indices.shape[0] % stepsize == 0
        self._indices, self._stepsize, self._offset = indices, stepsize, 0

    def __iter__(self):
        return self

    def __next__(self) -> np.ndarray:
        nxt = self._indices[self._offset:self._offset + self._stepsize]
        if nxt.size == 0:
            raise StopIteration
        self._offset += self._stepsize
        return nxt


class IndependentSampler(Sampler):
    """`np.random.permutation(T)` chopped into T/B blocks (samplers.py:49-63).  `rng` defaults to
    the global numpy stream exactly like the reference; pass a RandomState to give every env of a
    GPU the stream its own rank would have had (seed 10000*rank+seed, launch.py:33-45)."""

    def __init__(self, rollout_len: int, batch_size: int, rng: Optional[np.random.RandomState] = None):
        assert rollout_len % batch_size == 0
        super().__init__(rollout_len)
        self._batch_size = batch_size
        self._rng = rng if rng is not None else np.random
        self._iterator = None
        self.resample()

    def __iter__(self):
        return self._iterator

    def resample(self):
        self._iterator = IndicesIterator(self._rng.permutation(self._rollout_len), self._batch_size)
