"""Double / dueling DQN learner step around the prioritized replay buffer (BASELINE config "Double/Dueling DQN with prioritized
replay, 1M-transition sum-tree, batch 512 sampling and priority update").

The reference ships the BLOCKS — the NatureCNN trunk, `SimpleDiscreteActionValueHead(DuelingArchitecture)`
(drl/agents/heads/action_value_heads.py:85-140, architectures/stateless/dueling.py:9-57), the epsilon-greedy policy derived from it
(policy_heads.py:166-226), the double-Q Bellman optimality operator (algos/common/credit_assignment.py:252-305) and the agent
builder for them (utils/launch.py:268-282) — but no DQN algorithm and no replay buffer.  This module is the learner a
user of those blocks would write, on the CUDA path only:

    sample 512 transitions by priority (drl_per_sample)            -> idx, importance weights
    Q_online(s'), Q_target(s')   trunk on the frame ring by row index (drl_net_features), dueling head (drl_dueling_forward_f32)
    y = r + (1 - d) gamma Q_target(s', argmax_a Q_online(s', a)),  loss = mean w crit(Q_online(s, a) - y) and its gradient
        through the head (drl_dueling_loss_backward_f32) and the trunk (drl_net_backward_features)
    Adam on trunk + head (drl_net_adam_step / drl_adam_step), priorities |td| back into the sum-tree (drl_sumtree_update),
    target network <- online network every `target_update_interval` steps.

Parity: the network blocks and the Bellman operator are pinned to the unmodified reference (tests/golden/dqn.npz is written by
the reference's own modules + torch autograd); the ALGORITHM around them has no reference counterpart ("parity unpinned", like
the sum-tree).  There is no CPU fallback."""
from typing import Dict, Optional

import torch as tc

from .. import _lib
from .._lib import check, current_stream, ptr
from ..agents.dqn_heads import QAgent
from ..replay import PrioritizedReplayBuffer

# parameters() order of the reference's DuelingArchitecture = the flat layout drl_dueling_loss_backward_f32 expects
HEAD_KEYS = ['_proj.0.weight', '_proj.0.bias',
             '_advantage_stream._network.0.weight', '_advantage_stream._network.0.bias',
             '_advantage_stream._network.2.weight', '_advantage_stream._network.2.bias',
             '_value_stream._network.0.weight', '_value_stream._network.0.bias',
             '_value_stream._network.2.weight', '_value_stream._network.2.bias']


class FlatHead:
    """The dueling head's parameters as ONE flat fp32 buffer; the module's parameters become views of it (state_dict keys and
    values stay the reference's), so the head kernels and the flat Adam see what `load_state_dict` wrote."""

    def __init__(self, agent: QAgent):
        head = agent._predictors[agent._q_key]._action_value_head
        named = dict(head.named_parameters())
        params = [named[k] for k in HEAD_KEYS]
        self.device = params[0].device
        self.shapes = [tuple(p.shape) for p in params]
        self.numel = sum(p.numel() for p in params)
        self.flat = tc.empty(self.numel, dtype=tc.float32, device=self.device)
        o = 0
        for p in params:
            v = self.flat[o:o + p.numel()].view(p.shape)
            v.copy_(p.detach().to(tc.float32))
            p.data = v
            o += p.numel()
        self.in_dim, self.num_actions, self.widening = head._input_dim, head._output_dim, head._widening

    def views(self, of: Optional[tc.Tensor] = None) -> Dict[str, tc.Tensor]:
        buf = self.flat if of is None else of
        out, o = {}, 0
        for k, s in zip(HEAD_KEYS, self.shapes):
            n = 1
            for d in s:
                n *= d
            out[k] = buf[o:o + n].view(s)
            o += n
        return out


class DQN:
    """One learner for a fused online `QAgent`, a fused target `QAgent` and a `PrioritizedReplayBuffer`.

    Args mirror the reference's YAML vocabulary where it has one (gamma / use_dones / double_q of
    SimpleDiscreteBellmanOptimalityOperator, the torch.optim.Adam arguments, the criterion names of
    drl/algos/common/losses.py:6-18)."""

    def __init__(self, online: QAgent, target: QAgent, replay: PrioritizedReplayBuffer, batch_size: int = 512, gamma: float = 0.99,
                 use_dones: bool = True, double_q: bool = True, loss: str = 'SmoothL1Loss', lr: float = 1e-4,
                 betas=(0.9, 0.999), eps: float = 1e-8, target_update_interval: int = 1000, beta: float = 0.4):
        if online.engine is None or target.engine is None:
            raise RuntimeError('fuse() both agents first: there is no unfused / CPU path')
        if loss not in ('SmoothL1Loss', 'MSELoss'):
            raise ValueError('loss must be SmoothL1Loss or MSELoss (drl/algos/common/losses.py:6-18)')
        self.online, self.target, self.replay = online, target, replay
        self.batch_size, self.gamma, self.use_dones, self.double_q = batch_size, gamma, use_dones, double_q
        self.smooth_l1 = loss == 'SmoothL1Loss'
        self.lr, self.betas, self.eps = lr, betas, eps
        self.target_update_interval, self.beta = target_update_interval, beta
        self.head, self.target_head = FlatHead(online), FlatHead(target)
        if online.engine.max_batch < batch_size or target.engine.max_batch < batch_size:
            raise ValueError('the engines were fused for a smaller batch')
        dev = self.head.device
        self.device = dev
        self._L = _lib.lib()
        n = self.head.numel
        self.head_grads = tc.zeros(n, dtype=tc.float32, device=dev)
        self.head_exp_avg = tc.zeros(n, dtype=tc.float32, device=dev)
        self.head_exp_avg_sq = tc.zeros(n, dtype=tc.float32, device=dev)
        A = self.head.num_actions
        self.q = tc.empty((batch_size, A), dtype=tc.float32, device=dev)
        self.td = tc.empty(batch_size, dtype=tc.float32, device=dev)
        self.y = tc.empty(batch_size, dtype=tc.float32, device=dev)
        self.loss = tc.zeros(1, dtype=tc.float32, device=dev)
        self.dfeat = tc.empty((batch_size, self.head.in_dim), dtype=tc.float32, device=dev)
        self._ws = tc.empty(batch_size * 50 * self.head.widening, dtype=tc.float32, device=dev)
        self.steps = 0
        self.sync_target()

    # ---- pieces (public: the parity tests drive them one by one) -----------------------------------------------------
    def sync_target(self) -> None:
        """Target network <- online network (trunk flat buffer + head flat buffer), then refresh the packed operands."""
        self.target.engine.params.copy_(self.online.engine.params)
        self.target.sync()
        self.target_head.flat.copy_(self.head.flat)

    def q_values(self, agent: QAgent, head: FlatHead, frames: tc.Tensor, rows: tc.Tensor) -> tc.Tensor:
        feat = agent.engine.features(frames, rows)
        return self._head_forward(head, feat)

    def _head_forward(self, head: FlatHead, feat: tc.Tensor) -> tc.Tensor:
        v = head.views()
        q = tc.empty((feat.shape[0], head.num_actions), dtype=tc.float32, device=self.device)
        check(self._L.drl_dueling_forward_f32(ptr(feat), feat.shape[0], head.in_dim, head.num_actions, head.widening,
                                              *[ptr(v[k]) for k in HEAD_KEYS], ptr(q), current_stream()), 'dueling_forward')
        return q

    def loss_and_grads(self, frames: tc.Tensor, rows: tc.Tensor, next_rows: tc.Tensor, actions: tc.Tensor, rewards: tc.Tensor,
                       dones: tc.Tensor, weights: Optional[tc.Tensor]) -> tc.Tensor:
        """Forward + TD loss + backward for one batch: fills `self.loss` (device scalar), `self.td`, `self.y`, `self.q`,
        `self.head_grads` and the online engine's flat gradient buffer.  Returns the device loss."""
        nb = rows.numel()
        eng = self.online.engine
        q_next_tgt = self.q_values(self.target, self.target_head, frames, next_rows)
        q_next_on = self.q_values(self.online, self.head, frames, next_rows) if self.double_q else None
        feat = eng.features(frames, rows)
        self.loss.zero_()
        self.head_grads.zero_()
        w = weights.contiguous() if weights is not None else None
        check(self._L.drl_dueling_loss_backward_f32(
            ptr(feat), nb, self.head.in_dim, self.head.num_actions, self.head.widening, ptr(self.head.flat),
            ptr(actions.contiguous()), ptr(rewards.contiguous()), ptr(dones.contiguous()),
            ptr(q_next_on) if q_next_on is not None else None, ptr(q_next_tgt), ptr(w) if w is not None else None,
            float(self.gamma), int(self.use_dones), int(self.double_q), int(self.smooth_l1), ptr(self.q), ptr(self.td), ptr(self.y),
            ptr(self.loss), ptr(self.dfeat), ptr(self.head_grads), ptr(self._ws), current_stream()), 'dueling_loss_backward')
        # the online engine's last pass was features(frames, rows) just above (one chunk: nb <= max_batch): its activations are live
        check(self._L.drl_net_backward_features(eng._h, ptr(frames), ptr(rows), nb, ptr(self.dfeat), ptr(eng.grads), 1,
                                                current_stream()), 'net_backward_features')
        return self.loss

    def apply_gradients(self) -> None:
        eng = self.online.engine
        eng.adam_step(self.lr, self.betas, self.eps)
        check(self._L.drl_adam_step(ptr(self.head.flat), ptr(self.head_grads), ptr(self.head_exp_avg), ptr(self.head_exp_avg_sq),
                                    self.head.numel, float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                    eng.adam_steps, 1.0, current_stream()), 'adam_step')

    # ---- the step ------------------------------------------------------------------------------------------------------
    def learner_step(self, uniform01: Optional[tc.Tensor] = None) -> tc.Tensor:
        """Sample by priority, update the online network, write the new priorities back.  No host synchronisation:
        returns the device loss scalar."""
        b = self.replay.sample(self.batch_size, self.beta, uniform01)
        loss = self.loss_and_grads(b['frames'], b['idx'], b['next_idx'], b['actions'], b['rewards'], b['dones'], b['weights'])
        self.apply_gradients()
        self.replay.update_priorities(b['idx'], self.td[:b['idx'].numel()])
        self.steps += 1
        if self.steps % self.target_update_interval == 0:
            self.sync_target()
        return loss
