"""Checkpoint compatibility with drl/utils/checkpointing.py: one `{kind}_{steps}.pth` file per checkpointable holding its
`state_dict()`, the last five kept (checkpointing.py:14-43, 71-82), all kinds required at the same step on load (:85-117).

The fused path keeps parameters and optimizer state in FLAT device buffers; the classes below present them under the
reference's own key names and formats so that a run can move between the reference (CPU, DDP-wrapped modules: keys carry a
`module.` prefix) and this path in either direction:

  * `ViewsCheckpointable`   a named set of tensor views as a module-like state_dict (RND teacher / student networks),
  * `FlatAdamCheckpointable` flat exp_avg / exp_avg_sq / step as a `torch.optim.Adam.state_dict()` (RND optimizer),
  * `DictCheckpointable`    get / set callables (reward normaliser),
  * `DDPShim`               a `module.`-prefixing wrapper for the fused Agent (what DistributedDataParallel does to the keys),
  * `bind_flat_adam`        exposes a LearnerEngine's flat Adam state as the `state` of the caller's `torch.optim.Adam`, so
                            `policy_optimizer.state_dict()` / `.load_state_dict()` (ppo.py:295-307, launch.py:377-383) are complete,
  * `save_checkpoints` / `maybe_load_checkpoints`  the reference's file protocol, for callers without the reference installed.
"""
import os
import re
from typing import Callable, Dict, List, Mapping, Optional, Sequence

import torch as tc


def strip_module_prefix(sd: Mapping[str, tc.Tensor]) -> Dict[str, tc.Tensor]:
    """DistributedDataParallel prefixes every key with `module.` (launch.py:310); accept both forms."""
    return {(k[len('module.'):] if k.startswith('module.') else k): v for k, v in sd.items()}


class ViewsCheckpointable:
    def __init__(self, views: Callable[[], Mapping[str, tc.Tensor]], prefix: str = 'module.',
                 on_load: Optional[Callable[[], None]] = None):
        self._views, self._prefix, self._on_load = views, prefix, on_load

    def state_dict(self) -> Dict[str, tc.Tensor]:
        return {self._prefix + k: v.detach().clone() for k, v in self._views().items()}

    def load_state_dict(self, sd: Mapping[str, tc.Tensor], strict: bool = True) -> None:
        sd = strip_module_prefix(sd)
        views = self._views()
        missing = [k for k in views if k not in sd]
        unexpected = [k for k in sd if k not in views]
        if strict and (missing or unexpected):
            raise RuntimeError(f'state_dict mismatch: missing {missing}, unexpected {unexpected}')
        with tc.no_grad():
            for k, v in views.items():
                if k in sd:
                    v.copy_(tc.as_tensor(sd[k]).to(v.device, v.dtype).reshape(v.shape))
        if self._on_load is not None:
            self._on_load()


class DictCheckpointable:
    def __init__(self, get: Callable[[], Dict[str, tc.Tensor]], set_: Callable[[Mapping[str, tc.Tensor]], None]):
        self._get, self._set = get, set_

    def state_dict(self):
        return self._get()

    def load_state_dict(self, sd):
        self._set(sd)


_ADAM_GROUP_DEFAULTS = dict(weight_decay=0, amsgrad=False, maximize=False, foreach=None, capturable=False,
                            differentiable=False, fused=None)


class FlatAdamCheckpointable:
    """torch.optim.Adam.state_dict() view of flat Adam buffers whose parameters are `shapes` in order."""

    def __init__(self, shapes: Sequence[Sequence[int]], exp_avg: tc.Tensor, exp_avg_sq: tc.Tensor,
                 get_steps: Callable[[], int], set_steps: Callable[[int], None], hyper: Callable[[], Dict]):
        self._shapes = [tuple(s) for s in shapes]
        self._m, self._v = exp_avg, exp_avg_sq
        self._get_steps, self._set_steps, self._hyper = get_steps, set_steps, hyper

    def _slices(self, flat):
        out, o = [], 0
        for s in self._shapes:
            n = 1
            for d in s:
                n *= d
            out.append(flat[o:o + n].view(s))
            o += n
        return out

    def state_dict(self):
        steps = self._get_steps()
        state = {}
        if steps > 0:                      # torch creates the state lazily at the first step
            for i, (m, v) in enumerate(zip(self._slices(self._m), self._slices(self._v))):
                state[i] = {'step': tc.tensor(float(steps)), 'exp_avg': m.detach().clone(), 'exp_avg_sq': v.detach().clone()}
        h = self._hyper()
        group = dict(lr=h['lr'], betas=tuple(h['betas']), eps=h['eps'], **_ADAM_GROUP_DEFAULTS)
        group['params'] = list(range(len(self._shapes)))
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, sd):
        st = sd['state']
        if len(st) == 0:
            self._m.zero_(); self._v.zero_(); self._set_steps(0)
            return
        if sorted(st.keys()) != list(range(len(self._shapes))):
            raise RuntimeError('optimizer state does not match the parameter list')
        with tc.no_grad():
            for i, (m, v) in enumerate(zip(self._slices(self._m), self._slices(self._v))):
                m.copy_(tc.as_tensor(st[i]['exp_avg']).to(m.device, m.dtype).reshape(m.shape))
                v.copy_(tc.as_tensor(st[i]['exp_avg_sq']).to(v.device, v.dtype).reshape(v.shape))
        self._set_steps(int(float(st[0]['step'])))     # int in torch 1.10 (the reference's pin), tensor since 1.12


class DDPShim(tc.nn.Module):
    """`DDPShim(agent).state_dict()` carries the `module.` prefix of the reference's DistributedDataParallel-wrapped
    networks (launch.py:310), and loads both prefixed and plain checkpoints.  The gradient exchange itself is the
    Communicator's job (comm.py), not this wrapper's."""

    def __init__(self, module: tc.nn.Module):
        super().__init__()
        self.module = module

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def load_state_dict(self, state_dict, strict: bool = True):
        return self.module.load_state_dict(strip_module_prefix(state_dict), strict)


def bind_flat_adam(engine, optimizer: tc.optim.Optimizer, params: List[tc.nn.Parameter]) -> None:
    """Make `optimizer.state` the engine's flat Adam state: every parameter's `exp_avg` / `exp_avg_sq` entry is a VIEW of
    the flat buffers the fused Adam kernel updates, `step` one shared scalar tensor the engine keeps current.  State the
    optimizer already holds (a checkpoint loaded before the algorithm was built, launch.py:377-383, or
    `optimizer.load_state_dict` at any later time — detected by `engine.adam_step` through the broken aliasing) is
    adopted first."""
    if not isinstance(optimizer, tc.optim.Adam) or len(params) != len(engine.names):
        return
    held = [optimizer.state.get(p) for p in params]
    if all(h is not None and 'exp_avg' in h for h in held):
        m, v = engine.named_views(engine.exp_avg), engine.named_views(engine.exp_avg_sq)
        with tc.no_grad():
            for name, h in zip(engine.names, held):
                if h['exp_avg'].data_ptr() != m[name].data_ptr():
                    m[name].copy_(h['exp_avg'].to(m[name].device).reshape(m[name].shape))
                    v[name].copy_(h['exp_avg_sq'].to(v[name].device).reshape(v[name].shape))
        engine.adam_steps = int(float(held[0]['step']))
    step_t = tc.tensor(float(engine.adam_steps))
    m, v = engine.named_views(engine.exp_avg), engine.named_views(engine.exp_avg_sq)
    for name, p in zip(engine.names, params):
        optimizer.state[p] = {'step': step_t, 'exp_avg': m[name], 'exp_avg_sq': v[name]}
    engine._bound_opt = (optimizer, params, step_t)


def check_flat_adam_binding(engine) -> None:
    """Called by the engine before every fused Adam step: re-adopt the optimizer's state if somebody replaced it."""
    bound = getattr(engine, '_bound_opt', None)
    if bound is None:
        return
    optimizer, params, step_t = bound
    st = optimizer.state.get(params[0])
    if st is None or st.get('step') is not step_t or st['exp_avg'].data_ptr() != engine.exp_avg.data_ptr():
        bind_flat_adam(engine, optimizer, params)


# ---- the reference's file protocol (drl/utils/checkpointing.py) --------------------------------------------------------
def _parse(filename: str):
    m = re.match(r'(\w+)_([0-9]+).([a-z]+)', filename)
    return (m.group(1), int(m.group(2))) if m else (None, None)


def save_checkpoints(checkpoint_dir: str, checkpointables: Mapping[str, object], steps: int, keep: int = 5) -> None:
    os.makedirs(checkpoint_dir, exist_ok=True)
    for kind, obj in checkpointables.items():
        if obj is None:
            continue
        tc.save(obj.state_dict(), os.path.join(checkpoint_dir, f'{kind}_{steps}.pth'))
        mine = sorted(s for k, s in map(_parse, os.listdir(checkpoint_dir)) if k == kind)
        for s in mine[:-keep]:
            os.remove(os.path.join(checkpoint_dir, f'{kind}_{s}.pth'))


def maybe_load_checkpoints(checkpoint_dir: str, checkpointables: Mapping[str, object], map_location='cpu',
                           steps: Optional[int] = None) -> int:
    os.makedirs(checkpoint_dir, exist_ok=True)
    found = []
    for kind, obj in checkpointables.items():
        if obj is None:
            continue
        have = sorted(s for k, s in map(_parse, os.listdir(checkpoint_dir)) if k is not None and k.startswith(kind))
        s = (have[-1] if have else None) if steps is None else steps
        path = os.path.join(checkpoint_dir, f'{kind}_{s}.pth')
        if s is None or not os.path.exists(path):
            found.append(0)                                  # "Running from scratch." (checkpointing.py:58-62)
            continue
        obj.load_state_dict(tc.load(path, map_location=map_location))
        found.append(s)
    if len(set(found)) != 1:
        raise RuntimeError('Checkpoint steps not aligned.')  # checkpointing.py:114-116
    return found[0]
