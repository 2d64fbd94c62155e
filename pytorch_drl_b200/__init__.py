"""pytorch_drl_b200 — B200-native learner hot path of lucaslingle/pytorch_drl.

The PPO/RND minibatch update over uint8 84x84x4 rollouts (NatureCNN forward/backward, GAE(lambda),
fused PPO losses, flat Adam, NCCL gradient exchange) and a prioritized-replay sum-tree, written as
CUDA kernels for sm_100a behind the C-ABI of `include/drl_b200.h`, and exposed through classes
with the reference's own Python signatures (`algos.PPO`, `algos.GAE`, `agents.Agent`,
`agents.NatureCNN`, ...).  There is no CPU fallback: importing works anywhere, running needs the
built library and a B200.
"""
__version__ = '0.1.0'

from . import _lib  # noqa: F401


def library_path() -> str:
    return _lib.LIB_PATH


def register(drl_module=None):
    """Expose the fused classes under new `cls_name`s in the reference's lookup namespaces
    (script.py:55-58 algos, launch.py:174-176 architectures, credit_assignment.py:40-42), so that a
    YAML config selects them by name."""
    from . import algos, agents
    if drl_module is None:
        import drl as drl_module  # the reference package, if importable
    import importlib
    importlib.import_module('drl.algos').FusedPPO = algos.PPO
    importlib.import_module('drl.algos.common.credit_assignment').FusedGAE = algos.GAE
    importlib.import_module('drl.agents.architectures').FusedNatureCNN = agents.NatureCNN
    return drl_module
