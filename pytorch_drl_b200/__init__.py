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
    """Expose the fused classes under new `cls_name`s in the reference's lookup namespaces, so that a YAML config selects
    them by name:
      algos            getattr(drl.algos, cls_name)                              script.py:55-58         FusedPPO
      credit ops       getattr(drl.algos.common.credit_assignment, cls_name)      credit_assignment.py:40-42  FusedGAE, FusedNStepAdvantageEstimator
      architectures    getattr(drl.agents.architectures, cls_name)                launch.py:174-176       FusedNatureCNN, FusedLinear
      heads            getattr(drl.agents.heads, cls_name)                        launch.py:232-233       FusedCategoricalPolicyHead, FusedSimpleValueHead
      preprocessing    getattr(drl.agents.preprocessing, cls_name)                launch.py:140-142       FusedToChannelMajor
    The reference's own names stay untouched."""
    from . import algos, agents
    from .algos import credit_assignment as ca
    import importlib
    if drl_module is None:
        drl_module = importlib.import_module('drl')  # the reference package, if importable
    table = {
        'drl.algos': {'FusedPPO': algos.PPO},
        'drl.algos.common.credit_assignment': {'FusedGAE': algos.GAE, 'FusedNStepAdvantageEstimator': ca.NStepAdvantageEstimator},
        'drl.agents.architectures': {'FusedNatureCNN': agents.NatureCNN, 'FusedLinear': agents.Linear},
        'drl.agents.heads': {'FusedCategoricalPolicyHead': agents.CategoricalPolicyHead, 'FusedSimpleValueHead': agents.SimpleValueHead},
        'drl.agents.preprocessing': {'FusedToChannelMajor': agents.ToChannelMajor},
    }
    for mod, names in table.items():
        m = importlib.import_module(mod)
        for name, cls in names.items():
            setattr(m, name, cls)
    return drl_module
