"""In-memory stand-ins for `gym` and `moviepy`, which the reference imports but which are
absent from this image.

TEST INFRASTRUCTURE ONLY.  Nothing under `oracle/` is on the product path; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline leg may import it.

The surface below is exactly what the reference touches off the Atari path:
`gym.core.Env` (isinstance checks at drl/envs/wrappers/stateless/abstract.py:98,230),
`gym.spaces.Box/Discrete` (drl/algos/common/rollout.py:136,140; drl/envs/testing/lockstep.py:34)
and an empty `moviepy.editor` (drl/algos/abstract.py:12, only used by `video_loop`).
"""
import sys
import types

import numpy as np


def install():
    """Insert the stub modules into sys.modules (idempotent)."""
    if 'gym' in sys.modules and getattr(sys.modules['gym'], '_drl_b200_stub', False):
        return
    gym = types.ModuleType('gym')
    gym._drl_b200_stub = True
    core = types.ModuleType('gym.core')
    spaces = types.ModuleType('gym.spaces')

    class Env:
        metadata = {}
        reward_range = (-float('inf'), float('inf'))
        spec = None
        observation_space = None
        action_space = None

        @property
        def unwrapped(self):
            return self

        def reset(self):
            raise NotImplementedError

        def step(self, action):
            raise NotImplementedError

        def seed(self, seed=None):
            return [seed]

        def close(self):
            pass

    class Space:
        def __init__(self, shape=None, dtype=None):
            self.shape = None if shape is None else tuple(shape)
            self.dtype = None if dtype is None else np.dtype(dtype)

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                shape = np.asarray(low).shape
            super().__init__(shape, dtype)
            self.low = np.full(self.shape, low, dtype=self.dtype) if np.isscalar(low) else np.asarray(low)
            self.high = np.full(self.shape, high, dtype=self.dtype) if np.isscalar(high) else np.asarray(high)

    class Discrete(Space):
        def __init__(self, n):
            super().__init__((), np.int64)
            self.n = int(n)

        def __contains__(self, x):
            return 0 <= int(x) < self.n

        def sample(self):
            return int(np.random.randint(self.n))

    core.Env = Env
    gym.Env = Env
    gym.core = core
    spaces.Space, spaces.Box, spaces.Discrete = Space, Box, Discrete
    gym.spaces = spaces

    def _make(*a, **k):
        raise RuntimeError('gym.make is unavailable in the stub (no ROMs / emulator here)')

    gym.make = _make
    moviepy = types.ModuleType('moviepy')
    editor = types.ModuleType('moviepy.editor')
    moviepy.editor = editor
    sys.modules.update({
        'gym': gym, 'gym.core': core, 'gym.spaces': spaces,
        'moviepy': moviepy, 'moviepy.editor': editor,
    })
