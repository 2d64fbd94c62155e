"""Import the UNMODIFIED reference (`/root/reference/drl`) in this container.

TEST INFRASTRUCTURE ONLY (golden-vector generation and oracle pinning).  `/root/reference` does
not exist on the GPU box, so nothing that runs there may call `load()`; `available()` lets tests
skip cleanly.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get('DRL_REFERENCE_ROOT', '/root/reference')


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'drl'))


def load():
    """Returns the imported `drl` package of the reference."""
    if not available():
        raise RuntimeError(f'reference not found under {REFERENCE_ROOT}')
    from oracle import gym_stub
    gym_stub.install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.dont_write_bytecode = True   # /root/reference is read-only
    import drl  # noqa: F401
    import drl.algos  # noqa: F401
    import drl.agents  # noqa: F401
    return sys.modules['drl']
