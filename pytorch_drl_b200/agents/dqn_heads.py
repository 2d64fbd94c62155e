"""Drop-ins for the DQN head blocks the reference ships: `MLP`, `DuelingArchitecture`,
`SimpleDiscreteActionValueHead`, `EpsilonGreedyCategoricalPolicyHead` — same constructor signatures and state_dict key
names, forward passes on the CUDA kernels of csrc/dueling.cu.  The reference has no DQN learner, so these
heads are forward-only here: acting and target-Q evaluation; there is no CPU fallback."""
from typing import Any, Mapping, Optional, Type

import torch as tc

from .. import _lib
from .._lib import check, current_stream, ptr


def _apply_init(network: tc.nn.Module, w_init, b_init) -> None:
    # HeadEligibleArchitecture._init_weights (drl/agents/architectures/abstract.py:28-35)
    for m in network:
        if hasattr(m, 'weight') and w_init is not None:
            w_init(m.weight)
        if hasattr(m, 'bias') and b_init is not None:
            b_init(m.bias)


class MLP(tc.nn.Module):
    """drl/agents/architectures/stateless/mlp.py:8-45: `num_layers` x (Linear, ReLU) — the last Linear is followed by
    a ReLU as well, exactly like the reference's list comprehension."""

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int, num_layers: int, w_init=None, b_init=None):
        super().__init__()
        if not num_layers > 1:
            raise ValueError('MLP requires num_layers > 1.')
        self._input_dim, self._output_dim = input_dim, output_dim
        self._network = tc.nn.Sequential(*[
            tc.nn.Linear(in_features=input_dim if l // 2 == 0 else hidden_dim,
                         out_features=output_dim if l // 2 == num_layers - 1 else hidden_dim) if l % 2 == 0 else tc.nn.ReLU()
            for l in range(2 * num_layers)])
        _apply_init(self._network, w_init, b_init)


class DuelingArchitecture(tc.nn.Module):
    """drl/agents/architectures/stateless/dueling.py:9-57 (Wang et al. 2016)."""

    def __init__(self, input_dim: int, output_dim: int, widening: int, w_init=None, b_init=None):
        super().__init__()
        self._input_dim, self._output_dim, self._widening = input_dim, output_dim, widening
        self._proj = tc.nn.Sequential(tc.nn.Linear(input_dim, 50 * widening), tc.nn.ReLU())
        self._advantage_stream = MLP(50 * widening, 25 * widening, output_dim, 2, w_init, b_init)
        self._value_stream = MLP(50 * widening, 25 * widening, 1, 2, w_init, b_init)
        _apply_init(self._proj, w_init, b_init)

    @property
    def input_shape(self):
        return [self._input_dim]

    @property
    def output_dim(self):
        return self._output_dim

    @tc.no_grad()
    def forward(self, features: tc.Tensor, **kwargs) -> tc.Tensor:
        if not features.is_cuda:
            raise RuntimeError('DuelingArchitecture runs on the CUDA kernels only (no CPU fallback): move the module and '
                               'the features to a B200')
        f = features.detach().to(tc.float32).contiguous()
        if f.dim() != 2 or f.shape[1] != self._input_dim:
            raise ValueError(f'features must be [batch_size, {self._input_dim}]')
        ws = [self._proj[0].weight, self._proj[0].bias,
              self._advantage_stream._network[0].weight, self._advantage_stream._network[0].bias,
              self._advantage_stream._network[2].weight, self._advantage_stream._network[2].bias,
              self._value_stream._network[0].weight, self._value_stream._network[0].bias,
              self._value_stream._network[2].weight, self._value_stream._network[2].bias]
        if any(w.device != f.device for w in ws):
            raise RuntimeError('module parameters and features live on different devices')
        ws = [w.detach().to(tc.float32).contiguous() for w in ws]
        q = tc.empty((f.shape[0], self._output_dim), dtype=tc.float32, device=f.device)
        check(_lib.lib().drl_dueling_forward_f32(ptr(f), f.shape[0], self._input_dim, self._output_dim, self._widening,
                                                 *[ptr(w) for w in ws], ptr(q), current_stream()), 'dueling_forward')
        return q


class SimpleDiscreteActionValueHead(tc.nn.Module):
    """drl/agents/heads/action_value_heads.py:85-140."""

    def __init__(self, num_features: int, num_actions: int, head_architecture_cls: Type = DuelingArchitecture,
                 head_architecture_cls_args: Optional[Mapping[str, Any]] = None, w_init=None, b_init=None, **kwargs):
        super().__init__()
        if head_architecture_cls is not DuelingArchitecture:
            raise NotImplementedError('only the DuelingArchitecture action-value head is fused')
        self._num_features, self._num_actions = num_features, num_actions
        self._action_value_head = head_architecture_cls(input_dim=num_features, output_dim=num_actions, w_init=w_init,
                                                        b_init=b_init, **(head_architecture_cls_args or {}))

    @property
    def num_features(self):
        return self._num_features

    @property
    def num_actions(self):
        return self._num_actions

    def forward(self, features: tc.Tensor, **kwargs) -> tc.Tensor:
        return self._action_value_head(features)


class EpsilonGreedyCategoricalPolicyHead(tc.nn.Module):
    """drl/agents/heads/policy_heads.py:166-226: Categorical over (1 - eps) * one_hot(argmax Q) + eps * uniform."""

    def __init__(self, action_value_head: SimpleDiscreteActionValueHead, **kwargs):
        super().__init__()
        self._num_features, self._num_actions = action_value_head.num_features, action_value_head.num_actions
        self._action_value_head = action_value_head
        self._epsilon = None

    @property
    def epsilon(self):
        return self._epsilon

    @epsilon.setter
    def epsilon(self, value):
        assert 0. <= value <= 1.
        self._epsilon = value

    @tc.no_grad()
    def forward(self, features: tc.Tensor, **kwargs) -> tc.distributions.Categorical:
        if not isinstance(self._action_value_head, SimpleDiscreteActionValueHead):
            raise NotImplementedError('Action-value head class not supported.')
        q = self._action_value_head(features)
        probs = tc.empty_like(q)
        check(_lib.lib().drl_epsilon_greedy_probs_f32(ptr(q), q.shape[0], q.shape[1], float(self._epsilon), ptr(probs), None,
                                                      current_stream()), 'epsilon_greedy_probs')
        return tc.distributions.Categorical(probs=probs)


class QAgent(tc.nn.Module):
    """The reference's `Agent` (drl/agents/integration/agent.py:14-75) for the DQN predictor set that
    `get_predictors` builds (drl/utils/launch.py:268-282): a `DiscreteActionValueHead` under its own key plus the
    epsilon-greedy 'policy' derived from it.  Forward only (acting, target-network evaluation): the NatureCNN trunk
    runs on the fused engine (`drl_net_features`), the heads on csrc/dueling.cu.  state_dict keys are the
    reference's (`_architecture._network.*`, `_predictors.<key>._action_value_head.*`)."""

    def __init__(self, preprocessing, architecture, predictors: Mapping[str, tc.nn.Module], detach_input: bool = True):
        super().__init__()
        from . import NatureCNN
        if not isinstance(architecture, NatureCNN):
            raise NotImplementedError('only NatureCNN is fused')
        heads = {k: v for k, v in predictors.items() if isinstance(v, SimpleDiscreteActionValueHead)}
        if len(heads) != 1:
            raise ValueError('QAgent needs exactly one SimpleDiscreteActionValueHead predictor')
        self._preprocessing = tc.nn.Sequential(*preprocessing)
        self._architecture = architecture
        self._predictors = tc.nn.ModuleDict(dict(predictors))
        if 'policy' not in self._predictors:               # launch.py:278-282
            self._predictors['policy'] = EpsilonGreedyCategoricalPolicyHead(action_value_head=next(iter(heads.values())))
        self._q_key = next(iter(heads.keys()))
        self._detach_input = detach_input
        self.engine = None

    def fuse(self, max_batch: int, precision: str = 'bf16', device=None) -> 'QAgent':
        from ..engine import TRUNK_NAMES, LearnerEngine
        eng = LearnerEngine(self._predictors[self._q_key].num_actions, ['value_extrinsic'], max_batch, precision, device)
        named = dict(self.named_parameters())
        views = eng.named_views()
        for n in TRUNK_NAMES:                               # the engine's own policy / value heads stay zero and unused
            views[n].copy_(named[n].detach().to(eng.device, tc.float32))
            named[n].data = views[n]
        eng.sync_params()
        self._predictors.to(eng.device)
        self.engine = eng
        return self

    def sync(self) -> None:
        """Call after changing trunk parameters in place (e.g. a target-network copy): refreshes the packed operands."""
        self.engine.sync_params()

    @tc.no_grad()
    def forward(self, observations: tc.Tensor, predict, **kwargs):
        if self.engine is None:
            raise RuntimeError('call fuse() first: there is no unfused / CPU path')
        from . import Agent
        obs = Agent.to_uint8_frames(observations.detach() if self._detach_input else observations)
        feats = self.engine.features(obs.to(self.engine.device).contiguous())
        return {key: self._predictors[key](feats) for key in predict}
