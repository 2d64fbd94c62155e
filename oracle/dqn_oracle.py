"""TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product).

CPU restatement (torch float32 + autograd) of the double / dueling DQN learner step that pytorch_drl_b200/algos/dqn.py runs on the
CUDA path.  The BLOCKS follow the reference:
    NatureCNN trunk                            drl/agents/architectures/stateless/nature_cnn.py:27-36   (ppo_oracle.trunk_forward)
    SimpleDiscreteActionValueHead / Dueling    drl/agents/heads/action_value_heads.py:124-140, architectures/stateless/dueling.py:50-57
      with MLP's trailing ReLU                 drl/agents/architectures/stateless/mlp.py:32-41
    Bellman optimality backup, double-Q        drl/algos/common/credit_assignment.py:283-303 (n = 1)
    criterion                                  drl/algos/common/losses.py:6-18 (MSELoss / SmoothL1Loss, reduction 'none')
The reference has NO DQN algorithm: the loss `mean_i w_i crit(Q(s_i, a_i) - y_i)` (Schaul et al. 2015, Algorithm 1) is this
repository's; tests/golden/dqn.npz (oracle/make_golden.py::gen_dqn) is written by the reference's own modules + torch autograd on
the same loss, and pins this restatement (tests/test_oracle_golden.py)."""
from typing import Dict, Sequence

import numpy as np
import torch as tc
import torch.nn.functional as F

from oracle import ppo_oracle as po
from oracle.dueling_oracle import PARAM_KEYS as HEAD_KEYS

F32 = np.float32
HEAD_PREFIX = '_predictors.action_value_extrinsic._action_value_head.'


def head_shapes(num_actions: int, in_dim: int = 512):
    return [(50, in_dim), (50,), (25, 50), (25,), (num_actions, 25), (num_actions,), (25, 50), (25,), (1, 25), (1,)]


def init_params(num_actions: int, seed: int) -> Dict[str, np.ndarray]:
    """Seeded trunk (ppo_oracle.init_params) + dueling head parameters; biases are positive enough that the trailing ReLUs of
    both streams are active for a good share of the samples (otherwise every gradient would be zero)."""
    out = {k: v for k, v in po.init_params(num_actions, ['extrinsic'], seed).items() if k.startswith('_architecture')}
    rs = np.random.RandomState(seed + 1)
    for k, shp in zip(HEAD_KEYS, head_shapes(num_actions)):
        if k.endswith('weight'):
            out[HEAD_PREFIX + k] = (rs.standard_normal(shp) * np.sqrt(2.0 / shp[1])).astype(F32)
        else:
            out[HEAD_PREFIX + k] = (0.1 + 0.1 * rs.standard_normal(shp)).astype(F32)
    return out


def dueling_forward_t(p: Dict[str, tc.Tensor], feat: tc.Tensor) -> tc.Tensor:
    h = [p[HEAD_PREFIX + k] for k in HEAD_KEYS]
    proj = F.relu(F.linear(feat, h[0], h[1]))                                                   # dueling.py:51
    adv = F.relu(F.linear(F.relu(F.linear(proj, h[2], h[3])), h[4], h[5]))                      # mlp.py:32-41: trailing ReLU
    adv = adv - adv.mean(dim=-1, keepdim=True)                                                  # dueling.py:53
    val = F.relu(F.linear(F.relu(F.linear(proj, h[6], h[7])), h[8], h[9]))
    return adv + val                                                                            # dueling.py:55


def q_values(params: Dict[str, np.ndarray], obs_u8: np.ndarray) -> np.ndarray:
    p = {k: tc.from_numpy(v) for k, v in params.items()}
    with tc.no_grad():
        return dueling_forward_t(p, po.trunk_forward(p, po.scale_obs(tc.from_numpy(obs_u8)))).numpy()


def bellman_targets(rewards, dones, q_next_online, q_next_target, gamma: float, use_dones: bool, double_q: bool) -> np.ndarray:
    """credit_assignment.py:283-303 with rollout_len = 1, extra_steps = 0, one call per transition."""
    sel = q_next_online if double_q else q_next_target
    a = np.argmax(sel, axis=-1)                                        # first maximum, like torch.argmax
    boot = q_next_target[np.arange(len(a)), a].astype(F32)
    nonterminal = (F32(1) - dones.astype(F32)) if use_dones else np.ones_like(boot)
    return (rewards.astype(F32) + nonterminal * F32(gamma) * boot).astype(F32)


def loss_and_grads(params: Dict[str, np.ndarray], target_params: Dict[str, np.ndarray], obs_u8: np.ndarray, next_obs_u8: np.ndarray,
                   actions: np.ndarray, rewards: np.ndarray, dones: np.ndarray, weights: np.ndarray, gamma: float, use_dones: bool,
                   double_q: bool, loss: str, y_override=None):
    """Returns (dict of scalars / vectors, dict of gradients keyed like `params`).  `y_override`: use these targets instead of
    the oracle's own (bf16 tests: the argmax over near-tied Q(s') may differ; the targets are then checked separately)."""
    y = bellman_targets(rewards, dones, q_values(params, next_obs_u8), q_values(target_params, next_obs_u8), gamma, use_dones, double_q) \
        if y_override is None else np.asarray(y_override, F32)
    p = {k: tc.from_numpy(v.copy()).requires_grad_(True) for k, v in params.items()}
    q = dueling_forward_t(p, po.trunk_forward(p, po.scale_obs(tc.from_numpy(obs_u8))))
    q_taken = q.gather(1, tc.from_numpy(actions.astype(np.int64)).unsqueeze(1)).squeeze(1)
    crit = {'MSELoss': tc.nn.MSELoss, 'SmoothL1Loss': tc.nn.SmoothL1Loss}[loss](reduction='none')       # losses.py:6-18
    per = crit(q_taken, tc.from_numpy(y))
    total = (tc.from_numpy(weights.astype(F32)) * per).mean()
    total.backward()
    grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros(v.shape, F32)) for k, v in p.items()}
    return {'loss': float(total.detach()), 'q': q.detach().numpy(), 'y': y, 'td': (q_taken.detach().numpy() - y).astype(F32)}, grads


def make_case(seed: int, nb: int, num_actions: int):
    """Seeded synthetic batch of transitions (frames i.i.d. uint8 like bench.py; clipped-reward range; ~10 % terminals)."""
    rs = np.random.RandomState(seed)
    obs = rs.randint(0, 256, size=(nb, 84, 84, 4), dtype=np.uint8)
    nxt = rs.randint(0, 256, size=(nb, 84, 84, 4), dtype=np.uint8)
    actions = rs.randint(0, num_actions, size=nb).astype(np.int64)
    rewards = rs.randint(-1, 2, size=nb).astype(F32)
    dones = (rs.uniform(size=nb) < 0.1).astype(F32)
    weights = rs.uniform(0.2, 1.0, size=nb).astype(F32)
    return obs, nxt, actions, rewards, dones, weights


def summarize_grads(grads: Dict[str, np.ndarray], keys: Sequence[str]) -> Dict[str, np.ndarray]:
    """Small tensors whole; larger ones (conv2, conv3, fc weights) as every 53rd element + its norm (keeps the fixture small)."""
    out = {}
    for k in keys:
        g = grads[k]
        out['g_' + k] = g if g.size <= 10000 else g.reshape(-1)[::53].copy()
        out['n_' + k] = np.array(np.linalg.norm(g.astype(np.float64)))
    return out


def head_loss_and_grads(params: Dict[str, np.ndarray], feat: np.ndarray, y: np.ndarray, actions: np.ndarray, weights: np.ndarray,
                        loss: str):
    """The head's share of the step for GIVEN trunk features and targets: (scalars, head gradients, d loss / d features).  Used to
    check the fp32 head kernels tightly when the features come from the bf16 trunk."""
    p = {k: tc.from_numpy(v.copy()).requires_grad_(k.startswith(HEAD_PREFIX)) for k, v in params.items()}
    f = tc.from_numpy(np.asarray(feat, F32).copy()).requires_grad_(True)
    q = dueling_forward_t(p, f)
    q_taken = q.gather(1, tc.from_numpy(actions.astype(np.int64)).unsqueeze(1)).squeeze(1)
    crit = {'MSELoss': tc.nn.MSELoss, 'SmoothL1Loss': tc.nn.SmoothL1Loss}[loss](reduction='none')
    total = (tc.from_numpy(weights.astype(F32)) * crit(q_taken, tc.from_numpy(np.asarray(y, F32)))).mean()
    total.backward()
    grads = {k: v.grad.numpy() for k, v in p.items() if k.startswith(HEAD_PREFIX)}
    return {'loss': float(total.detach()), 'q': q.detach().numpy()}, grads, f.grad.numpy()


def trunk_vjp(params: Dict[str, np.ndarray], obs_u8: np.ndarray, dfeat: np.ndarray) -> Dict[str, np.ndarray]:
    """Trunk gradients for a given d loss / d features (what drl_net_backward_features computes)."""
    p = {k: tc.from_numpy(v.copy()).requires_grad_(True) for k, v in params.items() if k.startswith('_architecture')}
    feat = po.trunk_forward(p, po.scale_obs(tc.from_numpy(obs_u8)))
    feat.backward(tc.from_numpy(np.asarray(dfeat, F32)))
    return {k: v.grad.numpy() for k, v in p.items()}
