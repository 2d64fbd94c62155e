"""Generate `tests/golden/*.npz` by EXECUTING THE UNMODIFIED REFERENCE in this container.

Run here (the GPU box has no /root/reference):   python -m oracle.make_golden

What is pinned (all from reference code, never from the oracle):
  gae.npz        GAE.call               drl/algos/common/credit_assignment.py:195-214 (gae_extra.npz: extra_steps > 0)
  standardize.npz standardize           drl/utils/stats.py:4-17
  losses.npz     ppo_* loss functions   drl/algos/ppo.py:310-417 (+ composite, ppo.py:195-200)
  ppo_2rank.npz  annotate + credit_assignment + PPO.optimize run as TWO gloo ranks with
                 DDP(Agent(NatureCNN)) and torch.optim.Adam — the reference's deployment shape
                 (script.py:112-116, launch.py:310) — on SURVEY §8(d) synthetic inputs.
  normalizer.npz Normalizer.update/forward  drl/envs/wrappers/stateful/normalize.py:105-162
  rnd.npz        RND teacher/student forward, predictor loss + grads
                 drl/envs/wrappers/stateful/intrinsic/random_network_distillation.py:22-87,225-233
  nstep_bellman.npz NStepAdvantageEstimator.call, SimpleDiscreteBellmanOptimalityOperator.call
                 drl/algos/common/credit_assignment.py:225-249, 276-305
  reward_norm.npz ReturnAcc.update/forward (the reward normaliser's accumulator)
                 drl/envs/wrappers/stateful/normalize_reward.py:20-100
  dqn.npz        TD loss + autograd gradients through the reference's Agent(NatureCNN, SimpleDiscreteActionValueHead(Dueling)) and
                 SimpleDiscreteBellmanOptimalityOperator targets (pins oracle/dqn_oracle.py)
  dueling.npz    DuelingArchitecture.forward through SimpleDiscreteActionValueHead, and the probabilities of
                 EpsilonGreedyCategoricalPolicyHead  (dueling.py:50-57, action_value_heads.py:124-140,
                 policy_heads.py:197-226)
  actor_rollout.npz RolloutManager.generate with the reference Agent on seeded synthetic environments: actions drawn,
                 rewards / dones / frames recorded, extra_steps carry-over, episode metadata (rollout.py:105-357)
Large tensors are stored as (sum, l2, strided sample) summaries to keep fixtures small.
"""
import os
import sys

import numpy as np
import torch as tc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader, ppo_oracle as po  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
F32 = np.float32


def summary(x: np.ndarray, n: int = 2048) -> np.ndarray:
    """[sum, l2, sample...] of a big tensor (float64 accumulators)."""
    f = np.asarray(x, np.float64).ravel()
    stride = max(1, f.size // n)
    return np.concatenate([[f.sum(), np.sqrt((f * f).sum())], f[::stride][:n]])


# ---------------------------------------------------------------------------------------------
def gen_gae():
    from drl.algos.common.credit_assignment import GAE
    out = {}
    cases = [(2, 0.99, 0.95, True), (16, 0.99, 0.95, True), (128, 0.99, 0.95, True),
             (128, 0.999, 0.95, True), (128, 0.99, 0.95, False), (256, 0.99, 0.95, True),
             (37, 0.9, 0.5, True)]
    rs = np.random.RandomState(1234)
    for i, (T, g, l, ud) in enumerate(cases):
        r = rs.randint(-1, 2, size=(T,)).astype(F32) + (rs.standard_normal(T) * 0.1).astype(F32)
        v = rs.standard_normal(T + 1).astype(F32)
        d = (rs.uniform(size=T) < 0.05).astype(F32)
        op = GAE(gamma=g, use_dones=ud, lambda_=l)
        adv = op(rollout_len=T, extra_steps=0, rewards=tc.from_numpy(r), vpreds=tc.from_numpy(v),
                 dones=tc.from_numpy(d)).numpy()
        out.update({f'c{i}_meta': np.array([T, g, l, float(ud)]), f'c{i}_r': r, f'c{i}_v': v,
                    f'c{i}_d': d, f'c{i}_adv': adv})
    out['ncases'] = np.array(len(cases))
    np.savez_compressed(os.path.join(GOLD, 'gae.npz'), **out)


def gen_gae_extra():
    """GAE.call with extra_steps > 0 (the scan runs over T + n - 1 steps and the first T are returned)."""
    from drl.algos.common.credit_assignment import GAE
    out = {}
    cases = [(16, 2, 0.99, 0.95, True), (128, 4, 0.999, 0.9, False), (33, 1, 0.9, 0.5, True)]
    rs = np.random.RandomState(99)
    for i, (T, extra, g, l, ud) in enumerate(cases):
        Tn = T + extra
        r = (rs.randint(-1, 2, size=Tn) + rs.standard_normal(Tn) * 0.1).astype(F32)
        v = rs.standard_normal(Tn + 1).astype(F32)
        d = (rs.uniform(size=Tn) < 0.05).astype(F32)
        adv = GAE(gamma=g, use_dones=ud, lambda_=l)(rollout_len=T, extra_steps=extra, rewards=tc.from_numpy(r),
                                                    vpreds=tc.from_numpy(v), dones=tc.from_numpy(d)).numpy()
        out.update({f'c{i}_meta': np.array([T, extra, g, l, float(ud)]), f'c{i}_r': r, f'c{i}_v': v, f'c{i}_d': d,
                    f'c{i}_adv': adv})
    out['ncases'] = np.array(len(cases))
    np.savez_compressed(os.path.join(GOLD, 'gae_extra.npz'), **out)


def gen_standardize():
    from drl.utils.stats import standardize
    rs = np.random.RandomState(7)
    out = {}
    for i, T in enumerate([10, 128, 256, 1000]):
        x = (rs.standard_normal(T) * 3 + 1.5).astype(F32)
        out[f'x{i}'] = x
        out[f'y{i}'] = standardize(tc.from_numpy(x)).numpy()
    out['n'] = np.array(4)
    np.savez_compressed(os.path.join(GOLD, 'standardize.npz'), **out)


def gen_losses():
    from drl.algos.ppo import (ppo_policy_entropy_bonus, ppo_policy_surrogate_objective,
                               ppo_vf_loss)
    from drl.algos.common.losses import get_loss
    rs = np.random.RandomState(99)
    out = {}
    cases = [(32, ['extrinsic'], {'extrinsic': 1.0}, 0.2, 0.01, 1.0, False, True, 'MSELoss'),
             (64, ['extrinsic'], {'extrinsic': 1.0}, 0.1, 0.01, 0.5, True, True, 'MSELoss'),
             (32, ['extrinsic', 'intrinsic_rnd'], {'extrinsic': 2.0, 'intrinsic_rnd': 1.0},
              0.1, 0.001, 0.5, False, True, 'MSELoss'),
             (32, ['extrinsic', 'intrinsic_rnd'], {'extrinsic': 2.0, 'intrinsic_rnd': 1.0},
              0.1, 0.001, 0.5, True, False, 'MSELoss'),
             (48, ['extrinsic'], {'extrinsic': 1.0}, 0.2, 0.0, 1.0, True, True, 'SmoothL1Loss')]
    for i, (B, rk, rw, clip, entc, vfc, vclip, simple, crit) in enumerate(cases):
        lp_old = np.log(rs.uniform(0.05, 0.5, B)).astype(F32)
        lp_new = (lp_old + rs.standard_normal(B) * 0.2).astype(F32)
        ent = rs.uniform(0.1, 1.7, B).astype(F32)
        adv = {k: rs.standard_normal(B).astype(F32) for k in rk}
        v_old = {k: rs.standard_normal(B).astype(F32) for k in rk}
        v_new = {k: (v_old[k] + rs.standard_normal(B) * 0.3).astype(F32) for k in rk}
        vt = {k: rs.standard_normal(B).astype(F32) for k in rk}
        t = lambda a: tc.from_numpy(a)  # noqa: E731
        e = ppo_policy_entropy_bonus(entropies=t(ent), ent_coef=entc)
        p = ppo_policy_surrogate_objective(
            logprobs_new=t(lp_new), logprobs_old=t(lp_old),
            advantages={k: t(adv[k]) for k in rk}, clip_param=clip, reward_weights=rw)
        vv = ppo_vf_loss(
            vpreds_new={k: t(v_new[k]) for k in rk}, vpreds_old={k: t(v_old[k]) for k in rk},
            vtargs={k: t(vt[k]) for k in rk}, clip_param=clip, vf_loss_criterion=get_loss(crit),
            vf_loss_clipping=vclip, vf_simple_weighting=simple, reward_weights=rw)
        policy_loss = -(p['policy_surrogate_objective'] + e['policy_entropy_bonus'])
        composite = policy_loss + vfc * vv['vf_loss']
        out[f'c{i}_hyper'] = np.array([clip, entc, vfc, float(vclip), float(simple),
                                       float(crit == 'SmoothL1Loss')])
        out[f'c{i}_keys'] = np.array(rk)
        out[f'c{i}_weights'] = np.array([rw[k] for k in rk], F32)
        out[f'c{i}_lp_old'], out[f'c{i}_lp_new'], out[f'c{i}_ent'] = lp_old, lp_new, ent
        for k in rk:
            out[f'c{i}_adv_{k}'], out[f'c{i}_vold_{k}'] = adv[k], v_old[k]
            out[f'c{i}_vnew_{k}'], out[f'c{i}_vt_{k}'] = v_new[k], vt[k]
        out[f'c{i}_out'] = np.array([float(policy_loss), float(vv['vf_loss']), float(composite),
                                     float(e['state_averaged_entropy']), float(p['clipfrac'])], F32)
    out['ncases'] = np.array(len(cases))
    np.savez_compressed(os.path.join(GOLD, 'losses.npz'), **out)


# ---------------------------------------------------------------------------------------------
PPO2 = dict(T=16, B=8, A=6, epochs=2, seed=0, world=2, clip=0.2, ent=0.01, vf=1.0,
            lr=1e-3, betas=(0.9, 0.999), eps=1e-5, gamma=0.99, lam=0.95, param_seed=11)


def _ppo_rank(rank, world, port, cfg, rkeys, rweights, vclip, tag, retq):
    sys.path.insert(0, ROOT)
    from oracle import ref_loader as rl, ppo_oracle as po2
    rl.load()
    import gym
    from drl.envs.wrappers.stateless.abstract import RewardSpec
    from drl.agents.integration.agent import Agent
    from drl.agents.preprocessing import ToChannelMajor
    from drl.agents.architectures import NatureCNN, Linear
    from drl.agents.heads import CategoricalPolicyHead, SimpleValueHead
    from drl.algos.common.credit_assignment import GAE
    from drl.algos.ppo import PPO
    from torch.nn.parallel import DistributedDataParallel as DDP

    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    tc.distributed.init_process_group('gloo', rank=rank, world_size=world)
    tc.set_num_threads(2)
    T, B, A = cfg['T'], cfg['B'], cfg['A']

    class SynthEnv(gym.core.Env):
        def __init__(self):
            self.observation_space = gym.spaces.Box(0., 1., (84, 84, 4), np.float32)
            self.action_space = gym.spaces.Discrete(A)
            self._rs = np.random.RandomState(5)

        @property
        def reward_spec(self):
            return RewardSpec(keys=['extrinsic_raw'] + list(rkeys))

        def reset(self):
            return self._rs.uniform(size=(84, 84, 4)).astype(np.float32)

        def step(self, a):
            r = {k: 0.0 for k in ['extrinsic_raw'] + list(rkeys)}
            return self.reset(), r, False, {}

    preds = {'policy': CategoricalPolicyHead(512, A, Linear, {}, None, None)}
    for k in rkeys:
        preds[f'value_{k}'] = SimpleValueHead(512, Linear, {}, None, None)
    agent = Agent([ToChannelMajor()], NatureCNN(4, None, None), preds)
    init = po2.init_params(A, rkeys, seed=cfg['param_seed'])
    sd = {k: tc.from_numpy(v.copy()) for k, v in init.items()}
    agent.load_state_dict(sd, strict=True)
    assert list(dict(agent.named_parameters()).keys()) == po2.param_keys(rkeys)
    net = DDP(agent)
    opt = tc.optim.Adam(net.parameters(), lr=cfg['lr'], betas=cfg['betas'], eps=cfg['eps'])
    np.random.seed(10000 * rank + cfg['seed'])        # launch.py:33-45 seeding rule
    tc.manual_seed(10000 * rank + cfg['seed'])
    credit = {k: GAE(gamma=(cfg['gamma'] if k == 'extrinsic' else 0.99), use_dones=(k == 'extrinsic'),
                     lambda_=cfg['lam']) for k in rkeys}
    import tempfile
    tmp = tempfile.mkdtemp()
    algo = PPO(rank=rank, world_size=world, rollout_len=T, extra_steps=0,
               credit_assignment_ops=credit, stats_window_len=100, non_learning_steps=0,
               max_steps=10 ** 9, checkpoint_frequency=10 ** 9, checkpoint_dir=tmp, log_dir=tmp,
               media_dir=tmp, global_step=1, env=SynthEnv(), policy_net=net, policy_optimizer=opt,
               policy_scheduler=None, value_net=None, value_optimizer=None, value_scheduler=None,
               standardize_adv=True, opt_epochs=cfg['epochs'], learner_batch_size=B,
               clip_param_init=cfg['clip'], clip_param_final=cfg['clip'],
               ent_coef_init=cfg['ent'], ent_coef_final=cfg['ent'], vf_loss_cls='MSELoss',
               vf_loss_coef=cfg['vf'], vf_loss_clipping=vclip, vf_simple_weighting=True,
               use_pcgrad=False, reward_weights=(rweights if len(rkeys) > 1 else None))

    raw = po2.make_synthetic_env_rollout(rank, cfg['seed'], T, A, rkeys)
    rollout = dict(
        observations=po2.scale_obs(tc.from_numpy(raw['observations'])),   # float32 NHWC, F4
        actions=tc.from_numpy(raw['actions']),
        rewards={k: tc.from_numpy(raw['rewards'][k]) for k in rkeys},
        dones=tc.from_numpy(raw['dones']))
    rollout = algo.annotate(rollout, no_grad=True)
    rollout = algo.credit_assignment(rollout)

    rec = {}
    for k in rkeys:
        rec[f'adv_{k}'] = rollout['advantages'][k].numpy().copy()
        rec[f'vtarg_{k}'] = rollout['vtargs'][k].numpy().copy()
        rec[f'vpred_{k}'] = rollout['vpreds'][k].numpy().copy()
    rec['logp'] = rollout['logprobs'].numpy().copy()
    rec['ent'] = rollout['entropies'].numpy().copy()

    # record the permutations the sampler draws, the DDP-averaged grads of the first step, and
    # the per-epoch global metrics
    perms = []
    orig_perm = np.random.permutation

    def rec_perm(n):
        p = orig_perm(n)
        perms.append(np.asarray(p).copy())
        return p
    np.random.permutation = rec_perm
    steps = []
    orig_step = opt.step

    def rec_step(*a, **k):
        if len(steps) < 2:
            steps.append({n: p.grad.detach().numpy().copy() for n, p in agent.named_parameters()})
        return orig_step(*a, **k)
    opt.step = rec_step
    scalars = []
    if rank == 0:
        algo._writer.add_scalar = lambda tag, scalar_value, global_step: scalars.append((tag, float(scalar_value)))
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        algo.optimize(rollout)
    np.random.permutation = orig_perm

    rec['perms'] = np.stack(perms)                      # [epochs, T]
    for si, g in enumerate(steps):
        for n, v in g.items():
            rec[f'grad{si}::{n}'] = v if v.size <= 4096 else summary(v)
    for n, p in agent.named_parameters():
        v = p.detach().numpy()
        rec[f'final::{n}'] = v.copy() if v.size <= 4096 else summary(v)
    if rank == 0:
        names = ['policy_loss', 'value_loss', 'composite_loss', 'meanent', 'clipfrac']
        em = np.zeros((cfg['epochs'], 5), np.float64)
        for stag, val in scalars:
            ep, nm = stag.split('/')
            em[int(ep.split('_')[1]), names.index(nm)] = val
        rec['epoch_metrics'] = em
    np.savez_compressed(os.path.join(GOLD, f'_{tag}_rank{rank}.npz'), **rec)
    tc.distributed.destroy_process_group()


def gen_ppo(tag, rkeys, rweights, vclip, port):
    import torch.multiprocessing as mp
    cfg = dict(PPO2)
    mp.spawn(_ppo_rank, args=(cfg['world'], port, cfg, rkeys, rweights, vclip, tag, None),
             nprocs=cfg['world'], join=True)
    out = {'cfg_T': cfg['T'], 'cfg_B': cfg['B'], 'cfg_A': cfg['A'], 'cfg_epochs': cfg['epochs'],
           'cfg_seed': cfg['seed'], 'cfg_world': cfg['world'], 'cfg_param_seed': cfg['param_seed'],
           'cfg_hyper': np.array([cfg['clip'], cfg['ent'], cfg['vf'], cfg['lr'], cfg['betas'][0],
                                  cfg['betas'][1], cfg['eps'], cfg['gamma'], cfg['lam'], float(vclip)]),
           'cfg_rkeys': np.array(list(rkeys)),
           'cfg_rweights': np.array([rweights[k] for k in rkeys], np.float64)}
    for r in range(cfg['world']):
        f = os.path.join(GOLD, f'_{tag}_rank{r}.npz')
        d = np.load(f)
        for k in d.files:
            out[f'r{r}::{k}'] = d[k]
        os.remove(f)
    np.savez_compressed(os.path.join(GOLD, f'{tag}.npz'), **out)


# ---------------------------------------------------------------------------------------------
def gen_normalizer():
    from drl.envs.wrappers.stateful.normalize import Normalizer
    from oracle import rnd_oracle as ro
    items, x = ro.normalizer_fixture_inputs()
    nz = Normalizer([84, 84, 1], -5, 5)
    for it in items:
        nz.update(tc.from_numpy(it))
    y = nz(tc.from_numpy(x)).numpy()
    np.savez_compressed(os.path.join(GOLD, 'normalizer.npz'), y=summary(y),
                        steps=np.array(float(nz.steps)), m1=summary(nz.moment1.numpy()),
                        m2=summary(nz.moment2.numpy()))


def gen_rnd():
    from drl.envs.wrappers.stateful.intrinsic.random_network_distillation import (
        _TeacherNetwork, _StudentNetwork)
    from drl.envs.wrappers.stateful.normalize import Normalizer
    from oracle import rnd_oracle as ro
    teacher, student = _TeacherNetwork([84, 84, 1]), _StudentNetwork([84, 84, 1], 1)
    # LAPACK-dependent orthogonal init replaced by the seeded numpy init the tests can rebuild
    tinit, sinit = ro.init_net(False, 100), ro.init_net(True, 200)
    assert [n for n, _ in teacher.named_parameters()] == ro.TEACHER_KEYS
    assert [n for n, _ in student.named_parameters()] == ro.STUDENT_KEYS
    teacher.load_state_dict({k: tc.from_numpy(v.copy()) for k, v in tinit.items()})
    student.load_state_dict({k: tc.from_numpy(v.copy()) for k, v in sinit.items()})
    items, obs_u8 = ro.rnd_fixture_inputs()
    nz = Normalizer([84, 84, 1], -5, 5)
    for it in items:
        nz.update(tc.from_numpy(it))
    obs = po.scale_obs(tc.from_numpy(obs_u8))[:, :, :, -1:]
    normalized = nz(obs)
    y, yh = teacher(normalized), student(normalized)
    per = tc.square(y - yh).mean(dim=-1)
    loss = per.mean(dim=0)
    loss.backward()
    out = dict(y=summary(y.detach().numpy()), yh=summary(yh.detach().numpy()),
               per=per.detach().numpy(), loss=np.array(float(loss)))
    for n, p in student.named_parameters():
        g = p.grad.numpy()
        out[f'sgrad::{n}'] = g if g.size <= 1024 else summary(g, 512)
    np.savez_compressed(os.path.join(GOLD, 'rnd.npz'), **out)


def gen_nstep_bellman():
    from drl.algos.common.credit_assignment import NStepAdvantageEstimator, SimpleDiscreteBellmanOptimalityOperator
    out = {}
    rs = np.random.RandomState(4321)
    cases = [(16, 0, 0.99, True, 4), (128, 2, 0.99, True, 6), (128, 4, 0.999, False, 18), (37, 1, 0.9, True, 3)]
    for i, (T, extra, g, ud, A) in enumerate(cases):
        n = extra + 1
        r = (rs.randint(-1, 2, size=T + n - 1) + rs.standard_normal(T + n - 1) * 0.1).astype(F32)
        v = rs.standard_normal(T + n).astype(F32)
        d = (rs.uniform(size=T + n - 1) < 0.1).astype(F32)
        q = rs.standard_normal((T + n, A)).astype(F32)
        qt = rs.standard_normal((T + n, A)).astype(F32)
        q[3] = q[3, 0]                                              # ties: argmax must take the first maximum
        qt[5, 1:3] = qt[5].max() + 1.0
        adv = NStepAdvantageEstimator(gamma=g, use_dones=ud)(
            rollout_len=T, extra_steps=extra, rewards=tc.from_numpy(r), vpreds=tc.from_numpy(v), dones=tc.from_numpy(d)).numpy()
        bb = {}
        for dq in (False, True):
            bb[dq] = SimpleDiscreteBellmanOptimalityOperator(gamma=g, use_dones=ud, double_q=dq)(
                rollout_len=T, extra_steps=extra, rewards=tc.from_numpy(r), qpreds=tc.from_numpy(q),
                tgt_qpreds=tc.from_numpy(qt), dones=tc.from_numpy(d)).numpy()
        out.update({f'c{i}_meta': np.array([T, extra, g, float(ud), A]), f'c{i}_r': r, f'c{i}_v': v, f'c{i}_d': d,
                    f'c{i}_q': q, f'c{i}_qt': qt, f'c{i}_adv': adv, f'c{i}_bb': bb[False], f'c{i}_bb_dq': bb[True]})
    out['ncases'] = np.array(len(cases))
    np.savez_compressed(os.path.join(GOLD, 'nstep_bellman.npz'), **out)


def gen_reward_norm():
    """ReturnAcc driven step by step exactly as NormalizeRewardWrapper.step does (:186-188): reward as a 0-dim
    float32 tensor, done as bool; moments and normalised probes recorded after every window update."""
    from drl.envs.wrappers.stateful.normalize_reward import ReturnAcc
    out = {}
    cases = [(0.99, True, 2600), (0.99, False, 2100), (0.9, True, 700)]
    rs = np.random.RandomState(77)
    probes = np.array([-3.0, -1.0, -0.25, 0.0, 0.5, 1.0, 40.0], F32)
    for i, (gamma, use_dones, steps) in enumerate(cases):
        acc = ReturnAcc(gamma, -10., 10., use_dones)
        r = (rs.randint(-1, 2, size=steps).astype(F32) + (rs.standard_normal(steps) * 0.05).astype(F32)).astype(F32)
        d = rs.uniform(size=steps) < 0.01
        trace, first = [], [float(acc(tc.tensor(1.0).unsqueeze(0)))]          # forward before any statistics -> 0
        for t in range(steps):
            before = float(acc.steps)
            acc.update(tc.tensor(r[t]).float(), bool(d[t]))
            if float(acc.steps) != before:
                trace.append([t, float(acc.steps), float(acc.moment1), float(acc.moment2)] +
                             [float(acc(tc.tensor(x).float().unsqueeze(0))) for x in probes])
        out.update({f'c{i}_meta': np.array([gamma, float(use_dones), steps, acc.trace_length]), f'c{i}_r': r,
                    f'c{i}_d': d.astype(F32), f'c{i}_trace': np.array(trace, np.float64),
                    f'c{i}_first': np.array(first)})
    out['probes'] = probes
    out['ncases'] = np.array(len(cases))
    np.savez_compressed(os.path.join(GOLD, 'reward_norm.npz'), **out)


def gen_dueling():
    from drl.agents.architectures.stateless.dueling import DuelingArchitecture
    from drl.agents.heads.action_value_heads import SimpleDiscreteActionValueHead
    from drl.agents.heads.policy_heads import EpsilonGreedyCategoricalPolicyHead
    out = {}
    cases = [(6, 16, 0.1, 'orth'), (18, 33, 0.05, 'orth'), (4, 8, 1.0, 'orth'), (18, 4, 0.0, 'zeros')]
    for i, (A, nb, eps, init) in enumerate(cases):
        tc.manual_seed(100 + i)
        w_init = (lambda w: tc.nn.init.orthogonal_(w, gain=2 ** 0.5)) if init == 'orth' else tc.nn.init.zeros_
        b_init = (lambda b: tc.nn.init.normal_(b, std=0.1)) if init == 'orth' else tc.nn.init.zeros_
        head = SimpleDiscreteActionValueHead(num_features=512, num_actions=A, head_architecture_cls=DuelingArchitecture,
                                             head_architecture_cls_args={'widening': 1}, w_init=w_init, b_init=b_init)
        feats = tc.relu(tc.randn(nb, 512))
        with tc.no_grad():
            q = head(feats)
            pol = EpsilonGreedyCategoricalPolicyHead(action_value_head=head)
            pol.epsilon = eps
            dist = pol(feats)
        sd = head._action_value_head.state_dict()
        for k, v in sd.items():
            out[f'c{i}_p_{k}'] = v.numpy()
        out.update({f'c{i}_meta': np.array([A, nb, eps]), f'c{i}_feat': feats.numpy(), f'c{i}_q': q.numpy(),
                    f'c{i}_probs': dist.probs.numpy()})
    out['ncases'] = np.array(len(cases))
    np.savez_compressed(os.path.join(GOLD, 'dueling.npz'), **out)


def gen_actor_rollout():
    """The reference RolloutManager (rollout.py:227-357) + Agent, one manager per synthetic environment under
    torch.manual_seed(env_seed(e)): two consecutive generate() calls (extra_steps > 0: carry-over) per environment."""
    from oracle import actor_oracle as ao
    from drl.algos.common.rollout import RolloutManager
    from drl.agents.integration.agent import Agent
    from drl.agents.preprocessing import ToChannelMajor
    from drl.agents.architectures import NatureCNN, Linear
    from drl.agents.heads import CategoricalPolicyHead, SimpleValueHead
    import gym
    c = ao.CFG
    A = c['A']
    preds = {'policy': CategoricalPolicyHead(512, A, Linear, {}, None, None),
             'value_extrinsic': SimpleValueHead(512, Linear, {}, None, None)}
    agent = Agent([ToChannelMajor()], NatureCNN(4, None, None), preds)
    init = po.init_params(A, ['extrinsic'], seed=c['param_seed'])
    agent.load_state_dict({k: tc.from_numpy(v.copy()) for k, v in init.items()}, strict=True)

    class Env(ao.SynthActorEnv, gym.core.Env):
        def __init__(self, e):
            ao.SynthActorEnv.__init__(self, e, A, as_float=True)
            self.observation_space = gym.spaces.Box(0., 1., (84, 84, 4), np.float32)
            self.action_space = gym.spaces.Discrete(A)

        @property
        def reward_spec(self):
            from drl.envs.wrappers.stateless.abstract import RewardSpec
            return RewardSpec(keys=['extrinsic_raw', 'extrinsic'])

    out = {'cfg': np.array([c['E'], c['T'], c['extra_steps'], A, c['rollouts'], c['param_seed']])}
    for e in range(c['E']):
        tc.manual_seed(ao.env_seed(e))
        env = Env(e)
        mgr = RolloutManager(env, agent, c['T'], c['extra_steps'])
        for k in range(c['rollouts']):
            res = mgr.generate()
            obs = np.rint(res['observations'].numpy() * 255.0).astype(np.uint8)
            out[f'r{k}_e{e}_actions'] = res['actions'].numpy().astype(np.int64)
            out[f'r{k}_e{e}_rew'] = res['rewards']['extrinsic'].numpy()
            out[f'r{k}_e{e}_raw'] = res['rewards']['extrinsic_raw'].numpy()
            out[f'r{k}_e{e}_dones'] = res['dones'].numpy()
            out[f'r{k}_e{e}_obs_sum'] = ao.frame_checksums(obs)
            for key, vals in res['metadata'].items():
                out[f'r{k}_e{e}_md_{key}'] = np.asarray(vals, np.float64)
    # FrameStackWrapper (frame_stack.py:14-59) over a single-frame environment: a reset, five steps, a reset, three steps
    from drl.envs.wrappers.stateless.frame_stack import FrameStackWrapper

    class OneFrame(ao.SingleFrameEnv, gym.core.Env):
        def __init__(self):
            ao.SingleFrameEnv.__init__(self, 9, A, as_float=False)
            self.observation_space = gym.spaces.Box(0, 255, (84, 84, 1), np.uint8)
            self.action_space = gym.spaces.Discrete(A)
    fs = FrameStackWrapper(OneFrame(), num_stack=4)
    stacks = [fs.reset()] + [fs.step(k % A)[0] for k in range(5)] + [fs.reset()] + [fs.step(k % A)[0] for k in range(3)]
    out['framestack_sums'] = ao.frame_checksums(np.stack(stacks).astype(np.uint8))
    np.savez_compressed(os.path.join(GOLD, 'actor_rollout.npz'), **out)


def gen_dqn():
    """The DQN blocks of the reference — Agent(NatureCNN, SimpleDiscreteActionValueHead(DuelingArchitecture)) built like
    launch.py:268-282 and SimpleDiscreteBellmanOptimalityOperator — under the TD loss of oracle/dqn_oracle.py, gradients by torch
    autograd THROUGH THE REFERENCE MODULES.  Pins oracle/dqn_oracle.py (the restatement the GPU tests compare with)."""
    from drl.agents.architectures.stateless.dueling import DuelingArchitecture
    from drl.agents.architectures.stateless.nature_cnn import NatureCNN
    from drl.agents.heads.action_value_heads import SimpleDiscreteActionValueHead
    from drl.agents.integration.agent import Agent
    from drl.agents.preprocessing.vision import ToChannelMajor
    from drl.algos.common.credit_assignment import SimpleDiscreteBellmanOptimalityOperator
    from oracle import dqn_oracle as do
    out = {}
    cases = [(6, 5, 0.99, True, True, 'SmoothL1Loss'), (18, 4, 0.9, True, False, 'MSELoss'), (4, 3, 0.99, False, True, 'SmoothL1Loss')]
    for i, (A, nb, gamma, use_dones, double_q, loss) in enumerate(cases):
        def build(seed):
            head = SimpleDiscreteActionValueHead(num_features=512, num_actions=A, head_architecture_cls=DuelingArchitecture,
                                                 head_architecture_cls_args={'widening': 1}, w_init=None, b_init=None)
            agent = Agent([ToChannelMajor()], NatureCNN(4, None, None), {'action_value_extrinsic': head})
            agent.load_state_dict({k: tc.from_numpy(v.copy()) for k, v in do.init_params(A, seed).items()}, strict=True)
            return agent
        online, target = build(40 + i), build(70 + i)
        obs, nxt, actions, rewards, dones, weights = do.make_case(500 + i, nb, A)
        x, xn = po.scale_obs(tc.from_numpy(obs)), po.scale_obs(tc.from_numpy(nxt))
        with tc.no_grad():
            qn_on = online(xn, predict=['action_value_extrinsic'])['action_value_extrinsic']
            qn_tg = target(xn, predict=['action_value_extrinsic'])['action_value_extrinsic']
        op = SimpleDiscreteBellmanOptimalityOperator(gamma=gamma, use_dones=use_dones, double_q=double_q)
        y = tc.stack([op(rollout_len=1, extra_steps=0, rewards=tc.tensor([rewards[j]]), qpreds=tc.stack([qn_on[j], qn_on[j]]),
                         tgt_qpreds=tc.stack([qn_tg[j], qn_tg[j]]), dones=tc.tensor([dones[j]]))[0] for j in range(nb)])
        # DuelingArchitecture.forward subtracts the mean IN PLACE from a ReLU output (dueling.py:53), which autograd refuses
        # ("modified by an inplace operation": the shipped head was never trained).  The differentiable pass therefore calls the
        # reference's own sub-modules and forms `adv - mean(adv) + val` out of place; the no-grad passes above use forward() itself.
        head = online._predictors['action_value_extrinsic']._action_value_head
        feats = online._architecture(online._preprocessing(x))
        proj = head._proj(feats)
        adv = head._advantage_stream(proj)
        q = (adv - adv.mean(dim=-1, keepdim=True)) + head._value_stream(proj)
        with tc.no_grad():
            assert tc.equal(q.detach(), online(x, predict=['action_value_extrinsic'])['action_value_extrinsic'])
        q_taken = q.gather(1, tc.from_numpy(actions).unsqueeze(1)).squeeze(1)
        crit = getattr(tc.nn, loss)(reduction='none')                                     # drl/algos/common/losses.py:6-18
        total = (tc.from_numpy(weights) * crit(q_taken, y)).mean()
        online.zero_grad()
        total.backward()
        grads = {k: v.grad.numpy() for k, v in online.named_parameters()}
        out.update({f'c{i}_meta': np.array([A, nb, gamma, float(use_dones), float(double_q), float(loss == 'SmoothL1Loss'), 40 + i, 70 + i, 500 + i]),
                    f'c{i}_q': q.detach().numpy(), f'c{i}_y': y.numpy(), f'c{i}_loss': np.array(float(total.detach())),
                    f'c{i}_qn_on': qn_on.numpy(), f'c{i}_qn_tg': qn_tg.numpy()})
        out.update({f'c{i}_{k}': v for k, v in do.summarize_grads(grads, list(grads.keys())).items()})
    out['ncases'] = np.array(len(cases))
    np.savez_compressed(os.path.join(GOLD, 'dqn.npz'), **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref_loader.load()
    gen_gae()
    gen_gae_extra()
    gen_standardize()
    gen_losses()
    gen_normalizer()
    gen_rnd()
    gen_reward_norm()
    gen_nstep_bellman()
    gen_dueling()
    gen_dqn()
    gen_actor_rollout()
    gen_ppo('ppo_2rank', ['extrinsic'], {'extrinsic': 1.0}, False, 29611)
    gen_ppo('ppo_2rank_rnd', ['extrinsic', 'intrinsic_rnd'], {'extrinsic': 2.0, 'intrinsic_rnd': 1.0},
            True, 29612)
    print('golden vectors written to', GOLD)


if __name__ == '__main__':
    main()
