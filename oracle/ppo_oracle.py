"""CPU oracle for the PPO learner hot path: a plain numpy / torch-fp32 restatement of the
reference algorithm, function by function.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline
leg may import this module; the product path (`pytorch_drl_b200`) never does and fails loudly
when its CUDA library is missing.

Pinning: `oracle/make_golden.py` runs the UNMODIFIED reference (imported from /root/reference
under the gym stub, incl. a 2-rank gloo DDP run of `PPO.optimize`) on seeded inputs and commits
the outputs under `tests/golden/`; `tests/test_oracle_golden.py` checks every function below
against those vectors and against the reference's own known-answer tests.

All `file:line` citations are relative to /root/reference/.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch as tc
import torch.nn.functional as F

F32 = np.float32
SCALE_1_255 = np.float32(1.0 / 255.0)   # drl/envs/wrappers/stateless/scale_observations.py:35 (0x3b808081)

# state_dict names of the reference Agent (drl/agents/integration/agent.py:34-36,
# nature_cnn.py:27-36, policy_heads.py:95, value_heads.py:45).
TRUNK_KEYS = [
    '_architecture._network.0.weight', '_architecture._network.0.bias',
    '_architecture._network.2.weight', '_architecture._network.2.bias',
    '_architecture._network.4.weight', '_architecture._network.4.bias',
    '_architecture._network.7.weight', '_architecture._network.7.bias',
]
POLICY_KEYS = ['_predictors.policy._policy_head._network.0.weight',
               '_predictors.policy._policy_head._network.0.bias']


def value_keys(reward_key: str) -> List[str]:
    p = f'_predictors.value_{reward_key}._value_head._network.0.'
    return [p + 'weight', p + 'bias']


def param_keys(reward_keys: Sequence[str]) -> List[str]:
    """Parameter order == `Agent.parameters()` order of the reference (module registration order:
    architecture, then predictors in dict order: policy first, then value_<k>)."""
    keys = TRUNK_KEYS + POLICY_KEYS
    for k in reward_keys:
        keys += value_keys(k)
    return keys


def param_shapes(num_actions: int, reward_keys: Sequence[str]) -> Dict[str, Tuple[int, ...]]:
    shapes = {
        TRUNK_KEYS[0]: (32, 4, 8, 8), TRUNK_KEYS[1]: (32,),
        TRUNK_KEYS[2]: (64, 32, 4, 4), TRUNK_KEYS[3]: (64,),
        TRUNK_KEYS[4]: (64, 64, 3, 3), TRUNK_KEYS[5]: (64,),
        TRUNK_KEYS[6]: (512, 3136), TRUNK_KEYS[7]: (512,),
        POLICY_KEYS[0]: (num_actions, 512), POLICY_KEYS[1]: (num_actions,),
    }
    for k in reward_keys:
        w, b = value_keys(k)
        shapes[w] = (1, 512)
        shapes[b] = (1,)
    return shapes


def init_params(num_actions: int, reward_keys: Sequence[str], seed: int,
                head_gain: float = 0.01) -> Dict[str, np.ndarray]:
    """Seeded, platform-stable synthetic initialisation (numpy RandomState only, no LAPACK):
    Gaussian with std gain/sqrt(fan_in) (trunk gain sqrt(2), heads `head_gain`), small random
    biases so that bias paths are exercised.  Parity does not depend on orthogonality."""
    rs = np.random.RandomState(seed)
    out = {}
    for k, shp in param_shapes(num_actions, reward_keys).items():
        if k.endswith('weight'):
            fan_in = int(np.prod(shp[1:]))
            gain = np.sqrt(2.0) if k.startswith('_architecture') else head_gain
            out[k] = (rs.standard_normal(shp) * gain / np.sqrt(fan_in)).astype(F32)
        else:
            out[k] = (rs.standard_normal(shp) * 0.01).astype(F32)
    return out


# --------------------------------------------------------------------------------------------
# credit assignment
# --------------------------------------------------------------------------------------------
def gae(rollout_len: int, extra_steps: int, rewards: np.ndarray, vpreds: np.ndarray,
        dones: np.ndarray, gamma: float, lambda_: float, use_dones: bool) -> np.ndarray:
    """GAE(lambda) reverse scan, restating drl/algos/common/credit_assignment.py:195-214 with the
    same float32 operation order (every product/sum rounded to fp32, python-scalar coefficients
    folded the way torch folds wrapped numbers)."""
    T, n = rollout_len, extra_steps + 1
    r = np.asarray(rewards, F32)
    v = np.asarray(vpreds, F32)
    d = np.asarray(dones, F32)
    adv = np.zeros(T + n, F32)
    g32 = F32(gamma)
    l32 = F32(lambda_)
    gl32 = F32(gamma * lambda_)          # `1. * gamma * lambda` is python-float math when !use_dones
    for t in reversed(range(0, T + n - 1)):
        if use_dones:
            nt = F32(1.0) - d[t]
            c_v = F32(nt * g32)
            c_a = F32(F32(nt * g32) * l32)
        else:
            c_v = g32
            c_a = gl32
        delta = F32(F32(-v[t] + r[t]) + F32(c_v * v[t + 1]))
        adv[t] = F32(delta + F32(c_a * adv[t + 1]))
    return adv[0:T].copy()


def standardize(x: np.ndarray) -> np.ndarray:
    """(x - mean) / std with the UNBIASED std and no epsilon: drl/utils/stats.py:15-16."""
    t = tc.from_numpy(np.asarray(x, F32))
    return ((t - tc.mean(t)) / tc.std(t)).numpy()


def credit_assignment(rollout_len: int, extra_steps: int, rewards: np.ndarray, vpreds: np.ndarray,
                      dones: np.ndarray, gamma: float, lambda_: float, use_dones: bool,
                      standardize_adv: bool) -> Tuple[np.ndarray, np.ndarray]:
    """One reward stream of drl/algos/abstract.py:427-439: adv = GAE; vtarg = adv + V[:T];
    adv optionally standardised AFTER vtarg is formed."""
    adv = gae(rollout_len, extra_steps, rewards, vpreds, dones, gamma, lambda_, use_dones)
    vtarg = (adv + np.asarray(vpreds, F32)[:rollout_len]).astype(F32)
    if standardize_adv:
        adv = standardize(adv)
    return adv, vtarg


# --------------------------------------------------------------------------------------------
# sampler
# --------------------------------------------------------------------------------------------
def minibatch_indices(rollout_len: int, batch_size: int, rs=np.random) -> List[np.ndarray]:
    """drl/algos/common/samplers.py:49-63: one `np.random.permutation(T)` chopped into T/B blocks
    (asserts T % B == 0, samplers.py:51)."""
    assert rollout_len % batch_size == 0
    perm = rs.permutation(rollout_len)
    return [perm[i:i + batch_size] for i in range(0, rollout_len, batch_size)]


# --------------------------------------------------------------------------------------------
# network (torch fp32, functional, on the reference's parameter tensors/layouts)
# --------------------------------------------------------------------------------------------
def scale_obs(obs_u8: tc.Tensor) -> tc.Tensor:
    """uint8 NHWC frames -> the float32 NHWC tensor the reference's rollout stores:
    float32(x) * float32(1/255) (scale_observations.py:35; atari/constants.py:16)."""
    return obs_u8.to(tc.float32) * tc.tensor(float(SCALE_1_255), dtype=tc.float32)


def trunk_forward(p: Dict[str, tc.Tensor], obs_f32_nhwc: tc.Tensor, keep: Optional[dict] = None,
                  masks: Optional[dict] = None) -> tc.Tensor:
    """ToChannelMajor (vision.py:8-9) + NatureCNN (nature_cnn.py:27-36).  `masks` (tests only): 0/1 tensors act1 / act2 / act3
    (NCHW) / feat that REPLACE the ReLU decisions (`z * mask` instead of `relu(z)`) — the gradient of the network with another
    forward's ReLU pattern, used to separate the tensor-core engine's arithmetic error from its ReLU-mask flips."""
    relu = (lambda z, name: F.relu(z)) if masks is None else (lambda z, name: z * masks[name])
    x = obs_f32_nhwc.permute(0, 3, 1, 2)
    a1 = relu(F.conv2d(x, p[TRUNK_KEYS[0]], p[TRUNK_KEYS[1]], stride=4), 'act1')
    a2 = relu(F.conv2d(a1, p[TRUNK_KEYS[2]], p[TRUNK_KEYS[3]], stride=2), 'act2')
    a3 = relu(F.conv2d(a2, p[TRUNK_KEYS[4]], p[TRUNK_KEYS[5]], stride=1), 'act3')
    feat = relu(F.linear(a3.flatten(1), p[TRUNK_KEYS[6]], p[TRUNK_KEYS[7]]), 'feat')
    if keep is not None:
        keep.update(act1=a1, act2=a2, act3=a3, feat=feat)
    return feat


def heads_forward(p: Dict[str, tc.Tensor], feat: tc.Tensor, reward_keys: Sequence[str]):
    """CategoricalPolicyHead logits (policy_heads.py:107) and SimpleValueHead (value_heads.py:64)."""
    logits = F.linear(feat, p[POLICY_KEYS[0]], p[POLICY_KEYS[1]])
    values = {}
    for k in reward_keys:
        w, b = value_keys(k)
        values[k] = F.linear(feat, p[w], p[b]).squeeze(-1)
    return logits, values


def categorical_logprob_entropy(logits: tc.Tensor, actions: tc.Tensor):
    """torch.distributions.Categorical(logits=l): normalised = l - logsumexp(l); log_prob gathers;
    entropy = -sum(p * clamp(normalised, min=finfo.min)) (abstract.py:401-402)."""
    norm = logits - logits.logsumexp(dim=-1, keepdim=True)
    logp = norm.gather(-1, actions.long().unsqueeze(-1)).squeeze(-1)
    probs = F.softmax(norm, dim=-1)
    ent = -(tc.clamp(norm, min=tc.finfo(norm.dtype).min) * probs).sum(-1)
    return logp, ent


# --------------------------------------------------------------------------------------------
# losses (drl/algos/ppo.py:310-417, 195-200)
# --------------------------------------------------------------------------------------------
def ppo_losses(logp_new, ent_new, v_new: Dict[str, tc.Tensor], logp_old, adv: Dict[str, tc.Tensor],
               v_old: Dict[str, tc.Tensor], vtarg: Dict[str, tc.Tensor], reward_weights: Dict[str, float],
               clip_param: float, ent_coef: float, vf_coef: float, vf_clipping: bool,
               vf_simple_weighting: bool, vf_loss: str = 'MSELoss') -> Dict[str, tc.Tensor]:
    meanent = tc.mean(ent_new)                                    # ppo.py:323
    bonus = ent_coef * meanent                                    # ppo.py:324
    ratio = tc.exp(logp_new - logp_old)                           # ppo.py:353
    clipped = tc.clip(ratio, 1. - clip_param, 1. + clip_param)    # ppo.py:354-355
    acc = tc.zeros_like(logp_old)
    for k in reward_weights:                                      # ppo.py:358-359
        acc = acc + reward_weights[k] * adv[k]
    surr1, surr2 = acc * ratio, acc * clipped
    surrogate = tc.mean(tc.min(surr1, surr2))                     # ppo.py:362
    clipfrac = tc.mean(tc.greater(surr1, surr2).float())          # ppo.py:363
    crit = {'MSELoss': lambda i, t: (i - t) ** 2,
            'SmoothL1Loss': lambda i, t: F.smooth_l1_loss(i, t, reduction='none')}[vf_loss]
    vf = tc.tensor(0.0)
    for k in reward_weights:                                      # ppo.py:402-416
        if vf_clipping:
            vclip = v_old[k] + tc.clip(v_new[k] - v_old[k], -clip_param, clip_param)
            lk = tc.mean(tc.max(crit(v_new[k], vtarg[k]), crit(vclip, vtarg[k])))
        else:
            lk = tc.mean(crit(v_new[k], vtarg[k]))
        vf = vf + (lk if vf_simple_weighting else (reward_weights[k] ** 2) * lk)
    policy_loss = -(surrogate + bonus)                            # ppo.py:195-197
    composite = policy_loss + vf_coef * vf                        # ppo.py:199-200
    return dict(policy_loss=policy_loss, value_loss=vf, composite_loss=composite,
                meanent=meanent, clipfrac=clipfrac)


# --------------------------------------------------------------------------------------------
# optimiser: torch.optim.Adam (single-tensor path, no amsgrad / weight decay), the optimiser both
# shipped configs use (models_dir/atari_ppo/config.yaml:76-78; drl/utils/optimization.py:29-35).
# --------------------------------------------------------------------------------------------
def adam_step(p: np.ndarray, g: np.ndarray, m: np.ndarray, v: np.ndarray, step: int, lr: float,
              beta1: float, beta2: float, eps: float) -> None:
    """In-place; `step` is the 1-based count AFTER this update."""
    p32, g32 = tc.from_numpy(p), tc.from_numpy(g)
    m32, v32 = tc.from_numpy(m), tc.from_numpy(v)
    m32.lerp_(g32, 1 - beta1)
    v32.mul_(beta2).addcmul_(g32, g32, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v32.sqrt() / (bc2 ** 0.5)).add_(eps)
    p32.addcdiv_(m32, denom, value=-(lr / bc1))


# --------------------------------------------------------------------------------------------
# the whole minibatch step / optimize loop with the reference's per-rank (= per-env) semantics
# --------------------------------------------------------------------------------------------
class Hyper:
    """PPO hyper-parameters (constructor kwargs of drl/algos/ppo.py:33-67 that reach the math)."""

    def __init__(self, reward_weights=None, clip_param=0.2, ent_coef=0.01, vf_coef=1.0,
                 vf_clipping=False, vf_simple_weighting=True, vf_loss='MSELoss',
                 lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-5):
        self.reward_weights = dict(reward_weights or {'extrinsic': 1.0})
        self.clip_param, self.ent_coef, self.vf_coef = clip_param, ent_coef, vf_coef
        self.vf_clipping, self.vf_simple_weighting, self.vf_loss = vf_clipping, vf_simple_weighting, vf_loss
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps


def annotate(params: Dict[str, np.ndarray], obs_u8: np.ndarray, actions: np.ndarray,
             reward_keys: Sequence[str]):
    """No-grad forward over a batch of frames (abstract.py:383-411): logits, logp(a), entropy, V."""
    p = {k: tc.from_numpy(v) for k, v in params.items()}
    with tc.no_grad():
        feat = trunk_forward(p, scale_obs(tc.from_numpy(obs_u8)))
        logits, values = heads_forward(p, feat, reward_keys)
        logp, ent = categorical_logprob_entropy(logits, tc.from_numpy(actions))
    return (logits.numpy(), logp.numpy(), ent.numpy(), {k: v.numpy() for k, v in values.items()})


def minibatch_losses_and_grads(params: Dict[str, np.ndarray], envs: List[dict], hp: Hyper,
                               want_grads: bool = True, keep: Optional[dict] = None, masks: Optional[List[dict]] = None):
    """One learner step over `len(envs)` ranks.  `envs[n]` holds the minibatch of rank/env n:
    obs uint8[B,84,84,4], actions[B], logprobs[B], advantages{k}[B], vpreds{k}[B], vtargs{k}[B].
    Loss per rank = ppo.py:172-207; gradients = mean over ranks (DDP, launch.py:310 / ppo.py:233);
    metrics = mean over ranks of the per-rank scalars (global_means, ppo.py:267-268)."""
    rk = list(hp.reward_weights.keys())
    p = {k: tc.from_numpy(v.copy()).requires_grad_(want_grads) for k, v in params.items()}
    per_rank = []
    for n, e in enumerate(envs):
        k2 = {} if keep is not None else None
        feat = trunk_forward(p, scale_obs(tc.from_numpy(e['observations'])), k2, None if masks is None else masks[n])
        logits, values = heads_forward(p, feat, rk)
        logp, ent = categorical_logprob_entropy(logits, tc.from_numpy(e['actions']))
        if keep is not None:
            k2.update(logits=logits, values=values, logp=logp, ent=ent)
            keep.setdefault('per_env', []).append(k2)
        per_rank.append(ppo_losses(
            logp, ent, values, tc.from_numpy(e['logprobs']),
            {k: tc.from_numpy(e['advantages'][k]) for k in rk},
            {k: tc.from_numpy(e['vpreds'][k]) for k in rk},
            {k: tc.from_numpy(e['vtargs'][k]) for k in rk},
            hp.reward_weights, hp.clip_param, hp.ent_coef, hp.vf_coef, hp.vf_clipping,
            hp.vf_simple_weighting, hp.vf_loss))
    metrics = {k: sum(r[k] for r in per_rank) / len(per_rank) for k in per_rank[0]}
    grads = None
    if want_grads:
        metrics['composite_loss'].backward()
        grads = {k: v.grad.numpy() for k, v in p.items()}
    return {k: float(v.detach()) for k, v in metrics.items()}, grads


def make_synthetic_env_rollout(rank: int, seed: int, T: int, num_actions: int,
                               reward_keys: Sequence[str]) -> dict:
    """SURVEY.md §8(d) synthetic inputs for one env/rank, seeded `10000*rank+seed`
    (drl/utils/launch.py:33-45): uint8 frames uniform 0..255, actions uniform [0,A), rewards in
    {-1,0,1}, dones Bernoulli(0.01).  T+1 frames/actions (annotate consumes T+1, abstract.py:388)."""
    rs = np.random.RandomState(10000 * rank + seed)
    ro = dict(
        observations=rs.randint(0, 256, size=(T + 1, 84, 84, 4), dtype=np.uint8),
        actions=rs.randint(0, num_actions, size=(T + 1,)).astype(np.int64),
        dones=(rs.uniform(size=(T,)) < 0.01).astype(F32),
        rewards={})
    for k in reward_keys:
        if k == 'extrinsic':
            ro['rewards'][k] = rs.randint(-1, 2, size=(T,)).astype(F32)
        else:
            ro['rewards'][k] = np.abs(rs.standard_normal((T,))).astype(F32)
    return ro


def collect_from_raw(params, raw: dict, T: int, credit: Dict[str, dict], standardize_adv: bool) -> dict:
    """annotate (abstract.py:371-411) + credit_assignment (abstract.py:413-445) on a raw rollout
    of T+1 frames; returns the T-length rollout dict PPO.optimize consumes."""
    rk = list(credit.keys())
    _, logp, ent, vals = annotate(params, raw['observations'], raw['actions'], rk)
    out = dict(observations=raw['observations'][:T], actions=raw['actions'][:T],
               logprobs=logp[:T].copy(), entropies=ent[:T].copy(), dones=raw['dones'][:T],
               rewards={k: raw['rewards'][k][:T] for k in rk},
               vpreds={}, advantages={}, vtargs={})
    for k in rk:
        c = credit[k]
        adv, vt = credit_assignment(T, 0, raw['rewards'][k], vals[k], raw['dones'],
                                    c['gamma'], c['lambda_'], c['use_dones'], standardize_adv)
        out['vpreds'][k] = vals[k][:T].copy()
        out['advantages'][k] = adv
        out['vtargs'][k] = vt
    return out


def slice_env(ro: dict, idx: np.ndarray) -> dict:
    """slice_nested_tensor (drl/utils/nested.py:6-12) for one env's rollout."""
    out = {}
    for k, v in ro.items():
        out[k] = slice_env(v, idx) if isinstance(v, dict) else v[idx]
    return out


def optimize(params: Dict[str, np.ndarray], rollouts: List[dict], perms: List[List[List[np.ndarray]]],
             hp: Hyper, adam_state: Optional[dict] = None, record: Optional[list] = None) -> dict:
    """PPO.optimize (ppo.py:242-281) for `len(rollouts)` ranks sharing one network.
    perms[epoch][rank] = list of index blocks (the sampler output, samplers.py:60-63).
    Mutates `params` in place; returns the per-epoch full-rollout metrics (ppo.py:265-268)."""
    keys = list(params.keys())
    if adam_state is None:
        adam_state = dict(step=0, m={k: np.zeros_like(params[k]) for k in keys},
                          v={k: np.zeros_like(params[k]) for k in keys})
    epoch_metrics = []
    for ep in range(len(perms)):
        nblocks = len(perms[ep][0])
        for b in range(nblocks):
            mbs = [slice_env(rollouts[r], perms[ep][r][b]) for r in range(len(rollouts))]
            metrics, grads = minibatch_losses_and_grads(params, mbs, hp)
            adam_state['step'] += 1
            for k in keys:
                adam_step(params[k], grads[k], adam_state['m'][k], adam_state['v'][k],
                          adam_state['step'], hp.lr, hp.beta1, hp.beta2, hp.eps)
            if record is not None:
                record.append(dict(metrics=metrics, grads=grads))
        m, _ = minibatch_losses_and_grads(params, rollouts, hp, want_grads=False)
        epoch_metrics.append(m)
    return dict(epoch_metrics=epoch_metrics, adam_state=adam_state)
