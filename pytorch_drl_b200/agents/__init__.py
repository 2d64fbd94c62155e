"""Drop-ins for the reference's agent stack on the learner hot path:
`Agent`, `ToChannelMajor`, `NatureCNN`, `Linear`, `CategoricalPolicyHead`, `SimpleValueHead`,
with the reference's constructor signatures and state_dict key names."""
from typing import Any, Callable, List, Mapping, Optional, Type

import torch as tc

from ..engine import LearnerEngine
from ..utils.checkpoint import DDPShim  # noqa: F401
from .dqn_heads import (MLP, DuelingArchitecture, EpsilonGreedyCategoricalPolicyHead, QAgent,  # noqa: F401
                        SimpleDiscreteActionValueHead)


class ToChannelMajor(tc.nn.Module):
    """drl/agents/preprocessing/vision.py:4-9.  Kept for configuration compatibility: the fused
    path consumes NHWC directly (the permute to NCHW never materialises)."""

    def forward(self, x):
        return x.permute(0, 3, 1, 2)


def _init_weights(module: tc.nn.Module, w_init, b_init) -> None:
    # drl/agents/architectures/abstract.py:28-35
    for m in module.modules():
        if hasattr(m, 'weight') and isinstance(m.weight, tc.Tensor) and w_init is not None:
            w_init(m.weight)
        if hasattr(m, 'bias') and isinstance(m.bias, tc.Tensor) and b_init is not None:
            b_init(m.bias)


class NatureCNN(tc.nn.Module):
    """drl/agents/architectures/stateless/nature_cnn.py:8-50 (same constructor, `_network` layout
    and therefore the same state_dict keys)."""

    def __init__(self, img_channels: int, w_init: Optional[Callable] = None, b_init: Optional[Callable] = None):
        super().__init__()
        if img_channels != 4:
            raise ValueError('the fused NatureCNN is built for 4-frame stacks (img_channels=4)')
        self._img_channels = img_channels
        self._num_features = 512
        self._network = tc.nn.Sequential(
            tc.nn.Conv2d(img_channels, 32, kernel_size=(8, 8), stride=(4, 4)), tc.nn.ReLU(),
            tc.nn.Conv2d(32, 64, kernel_size=(4, 4), stride=(2, 2)), tc.nn.ReLU(),
            tc.nn.Conv2d(64, 64, kernel_size=(3, 3), stride=(1, 1)), tc.nn.ReLU(),
            tc.nn.Flatten(), tc.nn.Linear(3136, self._num_features), tc.nn.ReLU())
        _init_weights(self._network, w_init, b_init)

    @property
    def input_shape(self) -> List[int]:
        return [self._img_channels, 84, 84]

    @property
    def output_dim(self) -> int:
        return self._num_features

    def forward(self, x, **kwargs):
        raise RuntimeError('NatureCNN runs fused inside pytorch_drl_b200.agents.Agent; '
                           'call the Agent (there is no module-by-module or CPU path)')


class Linear(tc.nn.Module):
    """drl/agents/architectures/stateless/linear.py:8-31."""

    def __init__(self, input_dim: int, output_dim: int, w_init=None, b_init=None):
        super().__init__()
        self._input_dim, self._output_dim = input_dim, output_dim
        self._network = tc.nn.Sequential(tc.nn.Linear(input_dim, output_dim))
        _init_weights(self._network, w_init, b_init)


class CategoricalPolicyHead(tc.nn.Module):
    """drl/agents/heads/policy_heads.py:60-109."""

    def __init__(self, num_features: int, num_actions: int, head_architecture_cls: Type = Linear,
                 head_architecture_cls_args: Optional[Mapping[str, Any]] = None, w_init=None, b_init=None,
                 **kwargs):
        super().__init__()
        if head_architecture_cls is not Linear:
            raise NotImplementedError('only the Linear head architecture is fused')
        self._num_features, self._num_actions = num_features, num_actions
        self._policy_head = head_architecture_cls(input_dim=num_features, output_dim=num_actions,
                                                  w_init=w_init, b_init=b_init,
                                                  **(head_architecture_cls_args or {}))

    @property
    def num_actions(self):
        return self._num_actions


class SimpleValueHead(tc.nn.Module):
    """drl/agents/heads/value_heads.py:20-65."""

    def __init__(self, num_features: int, head_architecture_cls: Type = Linear,
                 head_architecture_cls_args: Optional[Mapping[str, Any]] = None, w_init=None, b_init=None,
                 **kwargs):
        super().__init__()
        if head_architecture_cls is not Linear:
            raise NotImplementedError('only the Linear head architecture is fused')
        self._value_head = head_architecture_cls(input_dim=num_features, output_dim=1, w_init=w_init,
                                                 b_init=b_init, **(head_architecture_cls_args or {}))


class _FusedForward(tc.autograd.Function):
    """logits, values = f(parameters; frames) with the backward pass delegated to the fused kernels."""

    @staticmethod
    def forward(ctx, anchor, agent, obs):
        eng = agent.engine
        eng.sync_params()                      # parameters may have been changed through torch since the last call
        logits, values, _, _ = eng.forward(obs)
        ctx.agent, ctx.obs = agent, obs
        return logits, values

    @staticmethod
    def backward(ctx, dlogits, dvalues):
        eng = ctx.agent.engine
        nb = ctx.obs.shape[0]
        dl = dlogits if dlogits is not None else tc.zeros((nb, eng.A), device=eng.device)
        dv = dvalues if dvalues is not None else tc.zeros((nb, eng.K), device=eng.device)
        params = dict(ctx.agent.named_parameters())
        if any(p.grad is None for p in params.values()):        # optimizer.zero_grad(set_to_none=True): start from zero
            eng.grads.zero_()
            gviews = eng.named_views(eng.grads)
            for k, p in params.items():
                p.grad = gviews[k]
        eng.backward(ctx.obs, None, dl.contiguous(), dv.contiguous(), out=eng.grads)    # accumulate, like autograd
        return None, None, None


class Agent(tc.nn.Module):
    """drl/agents/integration/agent.py:10-75 with the whole forward fused on the GPU.

    Construction mirrors the reference (`preprocessing`, `architecture`, `predictors`,
    `detach_input`).  `fuse()` moves the parameters into one flat CUDA buffer owned by a
    `LearnerEngine`; every `nn.Parameter` then aliases a slice of that buffer, so `state_dict()`,
    `load_state_dict()` and torch optimisers keep working and stay checkpoint-compatible with the
    reference (same key names, same layouts)."""

    def __init__(self, preprocessing: List[tc.nn.Module], architecture: NatureCNN,
                 predictors: Mapping[str, tc.nn.Module], detach_input: bool = True):
        super().__init__()
        self._preprocessing = tc.nn.Sequential(*preprocessing)
        self._architecture = architecture
        self._predictors = tc.nn.ModuleDict(predictors)
        self._detach_input = detach_input
        self._check_predict_keys()
        self.engine: Optional[LearnerEngine] = None

    def _check_predict_keys(self):
        for key in self.keys:
            if key == 'policy' or key.startswith('value_'):
                continue
            if key.startswith('action_value_'):
                raise NotImplementedError('action-value heads are not on the fused path')
            raise ValueError(f'Prediction key {key} not supported.')        # agent.py:47
        if 'policy' not in self.keys or list(self.keys)[0] != 'policy':
            raise ValueError("the first predictor must be 'policy'")

    @property
    def keys(self):
        return self._predictors.keys()

    @property
    def value_keys(self) -> List[str]:
        return [k for k in self.keys if k.startswith('value_')]

    def fuse(self, max_batch: int, precision: str = 'bf16', device=None) -> 'Agent':
        if not isinstance(self._architecture, NatureCNN):
            raise NotImplementedError('only NatureCNN is fused')
        A = self._predictors['policy'].num_actions
        eng = LearnerEngine(A, self.value_keys, max_batch, precision, device)
        named = dict(self.named_parameters())
        if list(named.keys()) != eng.names:
            raise RuntimeError('parameter naming diverged from the reference layout')
        eng.load_named({k: v.detach() for k, v in named.items()})
        views = eng.named_views()
        gviews = eng.named_views(eng.grads)
        for k, p in named.items():
            p.data = views[k]
            p.grad = gviews[k]
        self.engine = eng
        return self

    def sync(self) -> None:
        """Call after parameters were modified through torch (optimizer.step, load_state_dict)."""
        self._require_engine().sync_params()

    def load_state_dict(self, state_dict, strict: bool = True):
        from ..utils.checkpoint import strip_module_prefix
        state_dict = strip_module_prefix(state_dict)       # checkpoints of the reference's DDP-wrapped Agent (launch.py:310)
        if self.engine is None:
            return super().load_state_dict(state_dict, strict)
        with tc.no_grad():
            for k, p in self.named_parameters():
                if k in state_dict:
                    p.data.copy_(tc.as_tensor(state_dict[k]).to(p.device, p.dtype))
                elif strict:
                    raise RuntimeError(f'missing key {k}')
        self.sync()
        return tc.nn.modules.module._IncompatibleKeys([], [])

    def _autograd_anchor(self) -> tc.Tensor:
        """A leaf that ties the fused forward into the autograd graph (the parameter gradients themselves are
        written by the backward kernels straight into the `.grad` views)."""
        if getattr(self, '_anchor', None) is None or self._anchor.device != self.engine.device:
            self._anchor = tc.zeros(1, device=self.engine.device, requires_grad=True)
        return self._anchor

    def _require_engine(self) -> LearnerEngine:
        if self.engine is None:
            raise RuntimeError('Agent.fuse(max_batch, precision) must be called first '
                               '(the forward runs only as fused CUDA kernels)')
        return self.engine

    @staticmethod
    def to_uint8_frames(observations: tc.Tensor) -> tc.Tensor:
        """Frames as the env emits them.  float32 inputs are the reference's pre-scaled frames
        x*float32(1/255) (scale_observations.py:35) and map back to the original bytes exactly."""
        if observations.dtype == tc.uint8:
            return observations
        return (observations * 255.0).round_().clamp_(0, 255).to(tc.uint8)

    def forward(self, observations: tc.Tensor, predict: List[str], **kwargs):
        """agent.py:53-75.  Under `torch.enable_grad()` the outputs are differentiable with respect to the
        parameters: `loss.backward()` runs the fused backward kernels (drl_net_backward) and ACCUMULATES into every
        parameter's `.grad` (which aliases the engine's flat gradient buffer), so a torch optimiser or
        `engine.adam_step` can follow — custom losses work like with the reference.  Observations get no gradient
        (the reference detaches them, agent.py:66-67)."""
        eng = self._require_engine()
        obs = self.to_uint8_frames(observations.detach() if self._detach_input else observations)
        obs = obs.to(eng.device).contiguous()
        if tc.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            logits, values = _FusedForward.apply(self._autograd_anchor(), self, obs)
        else:
            logits, values, _, _ = eng.forward(obs)
        out = {}
        for key in predict:
            if key == 'policy':
                out[key] = tc.distributions.Categorical(logits=logits)
            else:
                out[key] = values[:, eng.value_keys.index(key)]
        return out
