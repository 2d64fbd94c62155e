"""Pins the CPU oracle (oracle/*.py) to the UNMODIFIED reference: every function is checked
against `tests/golden/*.npz` (outputs of /root/reference executed by oracle/make_golden.py) and
against the reference's own known-answer tests, transcribed with their file:line.
CPU only; runs everywhere."""
import os

import numpy as np
import pytest
import torch as tc

from oracle import ppo_oracle as po, rnd_oracle as ro

F32 = np.float32


from _util import summary, close_to_summary  # noqa: E402,F401


# ---- GAE ------------------------------------------------------------------------------------
def test_gae_bit_exact_vs_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, 'gae.npz'))
    for i in range(int(g['ncases'])):
        T, gamma, lam, ud = g[f'c{i}_meta']
        adv = po.gae(int(T), 0, g[f'c{i}_r'], g[f'c{i}_v'], g[f'c{i}_d'], float(gamma), float(lam), bool(ud))
        # same fp32 operation order as credit_assignment.py:205-213 -> identical bits
        assert np.array_equal(adv.view(np.uint32), g[f'c{i}_adv'].view(np.uint32)), f'case {i}'


def test_gae_known_answers():
    """drl/algos/common/test_credit_assignment.py:7-61 (3 cases)."""
    gamma, lam = 0.99, 0.95
    r = np.array([0., 0.], F32)
    v = np.array([0., 0., 1.], F32)
    exp1 = np.array([r[0] + gamma * v[1] + lam * (r[0] + gamma * r[1] + gamma ** 2 * v[2]), r[1] + gamma * v[2]], F32)
    np.testing.assert_allclose(po.gae(2, 0, r, v, np.array([0., 0.], F32), gamma, lam, True), exp1, rtol=1.3e-6, atol=1e-5)
    d2 = np.array([1., 0.], F32)
    exp2 = np.array([0.0, r[1] + gamma * v[2]], F32)
    np.testing.assert_allclose(po.gae(2, 0, r, v, d2, gamma, lam, True), exp2, rtol=1.3e-6, atol=1e-5)
    np.testing.assert_allclose(po.gae(2, 0, r, v, d2, gamma, lam, False), exp1, rtol=1.3e-6, atol=1e-5)


def test_standardize(golden_dir):
    g = np.load(os.path.join(golden_dir, 'standardize.npz'))
    for i in range(int(g['n'])):
        np.testing.assert_array_equal(po.standardize(g[f'x{i}']), g[f'y{i}'])
    y = po.standardize(np.arange(10, dtype=F32))          # drl/utils/test_stats.py:10-14
    np.testing.assert_allclose(y.mean(), 0.0, atol=1e-5)
    np.testing.assert_allclose(y.std(ddof=1), 1.0, rtol=1.3e-6)


# ---- losses ---------------------------------------------------------------------------------
def _loss_case(g, i):
    clip, entc, vfc, vclip, simple, sl1 = g[f'c{i}_hyper']
    rk = [str(k) for k in g[f'c{i}_keys']]
    rw = {k: float(w) for k, w in zip(rk, g[f'c{i}_weights'])}
    t = lambda a: tc.from_numpy(np.asarray(a))  # noqa: E731
    out = po.ppo_losses(
        t(g[f'c{i}_lp_new']), t(g[f'c{i}_ent']), {k: t(g[f'c{i}_vnew_{k}']) for k in rk},
        t(g[f'c{i}_lp_old']), {k: t(g[f'c{i}_adv_{k}']) for k in rk},
        {k: t(g[f'c{i}_vold_{k}']) for k in rk}, {k: t(g[f'c{i}_vt_{k}']) for k in rk},
        rw, float(clip), float(entc), float(vfc), bool(vclip), bool(simple),
        'SmoothL1Loss' if sl1 else 'MSELoss')
    return np.array([float(out[k]) for k in ('policy_loss', 'value_loss', 'composite_loss', 'meanent', 'clipfrac')], F32)


def test_losses_vs_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, 'losses.npz'))
    for i in range(int(g['ncases'])):
        np.testing.assert_allclose(_loss_case(g, i), g[f'c{i}_out'], rtol=1e-6, atol=1e-7, err_msg=f'case {i}')


def test_losses_known_answers():
    """drl/algos/test_ppo.py:7-14, 17-41, 44-88."""
    t = tc.tensor
    z = tc.zeros(6)
    lp_new = tc.log(t([0.105, 0.08, 0.24, 0.095, 0.12, 0.16]))
    lp_old = tc.log(t([0.10, 0.10, 0.20, 0.10, 0.10, 0.20]))
    adv = t([1.0, 1.0, 1.0, -1.0, -1.0, -1.0])
    out = po.ppo_losses(lp_new, z, {'extrinsic': z}, lp_old, {'extrinsic': adv}, {'extrinsic': z},
                        {'extrinsic': z}, {'extrinsic': 1.0}, 0.10, 0.0, 0.0, False, True)
    exp = np.mean([1.05, 0.80, 1.10, -0.95, -1.20, -0.90])
    np.testing.assert_allclose(-float(out['policy_loss']), exp, rtol=1.3e-6, atol=1e-5)
    ent = t([0.0, 1.0, 2.0, 3.0])
    z4 = tc.zeros(4)
    out = po.ppo_losses(z4, ent, {'extrinsic': z4}, z4, {'extrinsic': z4}, {'extrinsic': z4},
                        {'extrinsic': z4}, {'extrinsic': 1.0}, 0.1, 0.01, 0.0, False, True)
    np.testing.assert_allclose(float(out['meanent']), 1.5)
    np.testing.assert_allclose(-float(out['policy_loss']), 0.015, rtol=1e-6)
    v_new = t([1.05, 2.4, 1.2, -1.05, -2.4, -1.2])
    v_old = t([1.0, 2.0, 1.0, -1.0, -2.0, -1.0])
    vt = t([1.5, 1.5, 1.5, -1.5, -1.5, -1.5])
    for clipv, vals in ((False, [1.05, 2.4, 1.2, -1.05, -2.4, -1.2]), (True, [1.05, 2.4, 1.1, -1.05, -2.4, -1.1])):
        out = po.ppo_losses(z, z, {'extrinsic': v_new}, z, {'extrinsic': z}, {'extrinsic': v_old},
                            {'extrinsic': vt}, {'extrinsic': 1.0}, 0.10, 0.0, 1.0, clipv, True)
        exp = np.mean([(a - b) ** 2 for a, b in zip(vals, [1.5] * 3 + [-1.5] * 3)])
        np.testing.assert_allclose(float(out['value_loss']), exp, rtol=1.3e-6, atol=1e-5)


# ---- full 2-rank PPO run ----------------------------------------------------------------------
@pytest.mark.parametrize('tag', ['ppo_2rank', 'ppo_2rank_rnd'])
def test_ppo_optimize_vs_reference_2rank_ddp(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, f'{tag}.npz'))
    T, B, A, epochs = int(g['cfg_T']), int(g['cfg_B']), int(g['cfg_A']), int(g['cfg_epochs'])
    seed, world = int(g['cfg_seed']), int(g['cfg_world'])
    clip, ent, vf, lr, b1, b2, eps, gamma, lam, vclip = [float(x) for x in g['cfg_hyper']]
    rk = [str(k) for k in g['cfg_rkeys']]
    rw = {k: float(w) for k, w in zip(rk, g['cfg_rweights'])}
    params = po.init_params(A, rk, seed=int(g['cfg_param_seed']))
    credit = {k: dict(gamma=(gamma if k == 'extrinsic' else 0.99), lambda_=lam, use_dones=(k == 'extrinsic')) for k in rk}
    rollouts = []
    for r in range(world):
        raw = po.make_synthetic_env_rollout(r, seed, T, A, rk)
        ro_ = po.collect_from_raw(params, raw, T, credit, standardize_adv=True)
        np.testing.assert_allclose(ro_['logprobs'], g[f'r{r}::logp'], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(ro_['entropies'], g[f'r{r}::ent'], rtol=1e-5, atol=1e-6)
        for k in rk:
            np.testing.assert_allclose(ro_['vpreds'][k], g[f'r{r}::vpred_{k}'], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(ro_['advantages'][k], g[f'r{r}::adv_{k}'], rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(ro_['vtargs'][k], g[f'r{r}::vtarg_{k}'], rtol=1e-4, atol=1e-5)
        rollouts.append(ro_)
    # sampler stream: RandomState(10000*rank+seed): ctor resample + one per epoch (samplers.py:54-63)
    perms = []
    for ep in range(epochs):
        perms.append([])
    for r in range(world):
        rs = np.random.RandomState(10000 * r + seed)
        rs.permutation(T)                       # IndependentSampler.__init__ -> resample()
        for ep in range(epochs):
            blocks = po.minibatch_indices(T, B, rs)
            np.testing.assert_array_equal(np.concatenate(blocks), g[f'r{r}::perms'][ep])
            perms[ep].append(blocks)
    hp = po.Hyper(rw, clip, ent, vf, bool(vclip), True, 'MSELoss', lr, b1, b2, eps)
    record = []
    res = po.optimize(params, rollouts, perms, hp, record=record)
    keys = po.param_keys(rk)
    for si in range(2):
        for k in keys:
            close_to_summary(record[si]['grads'][k], g[f'r0::grad{si}::{k}'], rtol=2e-3, atol=1e-7)
    for k in keys:
        close_to_summary(params[k], g[f'r0::final::{k}'], rtol=1e-3, atol=2e-5)
    em = g['r0::epoch_metrics']
    names = ['policy_loss', 'value_loss', 'composite_loss', 'meanent', 'clipfrac']
    for ep in range(epochs):
        got = np.array([res['epoch_metrics'][ep][n] for n in names])
        np.testing.assert_allclose(got[:4], em[ep][:4], rtol=2e-3, atol=2e-5)
        assert abs(got[4] - em[ep][4]) <= 2.0 / T + 1e-6     # clipfrac: discontinuous indicator


# ---- normaliser + RND ---------------------------------------------------------------------------
def test_normalizer_vs_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, 'normalizer.npz'))
    items, x = ro.normalizer_fixture_inputs()
    nz = ro.Normalizer([84, 84, 1])
    for it in items:
        nz.update(it)
    assert nz.steps == float(g['steps'])
    close_to_summary(nz.m1, g['m1'], rtol=1e-6, always=True)
    close_to_summary(nz.m2, g['m2'], rtol=1e-6, always=True)
    close_to_summary(nz.apply(x), g['y'], rtol=1e-5, always=True)


def test_normalizer_known_answers():
    """drl/envs/wrappers/stateful/test_normalize.py:41-78 style: moments of a 2-item stream, clip."""
    nz = ro.Normalizer([2], -1.0, 1.0)
    nz.update(np.array([1.0, 2.0], F32))
    nz.update(np.array([3.0, 6.0], F32))
    np.testing.assert_allclose(nz.m1, [2.0, 4.0])
    np.testing.assert_allclose(nz.m2, [5.0, 20.0])
    y = nz.apply(np.array([[2.5, 100.0]], F32))
    np.testing.assert_allclose(y, [[0.5, 1.0]], rtol=1e-5)


def test_rnd_vs_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, 'rnd.npz'))
    items, obs_u8 = ro.rnd_fixture_inputs()
    nz = ro.Normalizer([84, 84, 1])
    for it in items:
        nz.update(it)
    loss, per, y, yh, grads = ro.rnd_learn_loss_and_grads(ro.init_net(False, 100), ro.init_net(True, 200), nz, obs_u8)
    np.testing.assert_allclose(loss, float(g['loss']), rtol=1e-5)
    np.testing.assert_allclose(per, g['per'], rtol=1e-5)
    close_to_summary(y, g['y'], rtol=1e-5, always=True)
    close_to_summary(yh, g['yh'], rtol=1e-5, always=True)
    for k in ro.STUDENT_KEYS:
        gk = grads[k]
        gold = g[f'sgrad::{k}']
        s = summary(gk, 512) if gk.size > 1024 else np.asarray(gk, np.float64)
        np.testing.assert_allclose(s, gold, rtol=1e-4, atol=1e-7 + 1e-5 * np.abs(gold).max())


def test_reward_normalizer_oracle_vs_reference_golden(golden_dir):
    """ReturnAcc (normalize_reward.py:20-100): window updates land on the same steps, moments within float32
    rounding of torch's mean (1e-6 rel), normalised probes likewise; forward() is 0 before any statistics."""
    from oracle import reward_norm_oracle as ro
    g = np.load(os.path.join(golden_dir, 'reward_norm.npz'))
    probes = g['probes']
    for i in range(int(g['ncases'])):
        gamma, use_dones, steps, L = g[f'c{i}_meta']
        acc = ro.ReturnAcc(float(gamma), -10.0, 10.0, bool(use_dones))
        assert acc.trace_length == int(L)
        assert float(acc.forward(1.0)) == float(g[f'c{i}_first'][0]) == 0.0
        r, d, trace = g[f'c{i}_r'], g[f'c{i}_d'], g[f'c{i}_trace']
        got = []
        for t in range(int(steps)):
            before = float(acc.steps)
            acc.update(r[t], d[t] != 0)
            if float(acc.steps) != before:
                got.append([t, float(acc.steps), float(acc.moment1), float(acc.moment2)] + [float(acc.forward(x)) for x in probes])
        got = np.array(got)
        assert got.shape == trace.shape
        np.testing.assert_array_equal(got[:, :2], trace[:, :2])
        np.testing.assert_allclose(got[:, 2:], trace[:, 2:], rtol=2e-6, atol=1e-7)


def test_nstep_and_bellman_oracle_vs_reference_golden(golden_dir):
    """credit_assignment.py:225-249, 276-305: bit-exact, plus the expected values of the reference's own tests
    (test_credit_assignment.py:64-170)."""
    from oracle import credit_oracle as co
    g = np.load(os.path.join(golden_dir, 'nstep_bellman.npz'))
    for i in range(int(g['ncases'])):
        T, extra, gam, ud, A = g[f'c{i}_meta']
        T, extra = int(T), int(extra)
        np.testing.assert_array_equal(co.nstep_advantages(T, extra, g[f'c{i}_r'], g[f'c{i}_v'], g[f'c{i}_d'], float(gam), bool(ud)),
                                      g[f'c{i}_adv'])
        for dq, key in ((False, 'bb'), (True, 'bb_dq')):
            np.testing.assert_array_equal(
                co.bellman_backups(T, extra, g[f'c{i}_r'], g[f'c{i}_q'], g[f'c{i}_qt'], g[f'c{i}_d'], float(gam), bool(ud), dq),
                g[f'c{i}_{key}'])
    # reference test vectors
    r, v = np.array([1., 2., 3., 4.], np.float32), np.array([0., 0., 0., 0., 5.], np.float32)
    adv = co.nstep_advantages(2, 2, r, v, np.zeros(4, np.float32), 0.99, True)
    np.testing.assert_allclose(adv, [1 + .99 * 2 + .99 ** 2 * 3, 2 + .99 * 3 + .99 ** 2 * 4 + .99 ** 3 * 5], rtol=1e-6)
    q = np.array([[0., 0., 0.], [0., 0., 1.], [0., 0., 1.]], np.float32)
    qt = np.array([[0., 0., 0.], [1., 0., 0.], [0., 1., 0.]], np.float32)
    z = np.zeros(2, np.float32)
    np.testing.assert_allclose(co.bellman_backups(2, 0, z, q, qt, z, 0.99, True, False), [0.99, 0.99], rtol=1e-6)
    np.testing.assert_allclose(co.bellman_backups(2, 0, z, q, qt, z, 0.99, True, True), [0.0, 0.0], atol=1e-7)


def test_gae_extra_steps_oracle_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'gae_extra.npz'))
    for i in range(int(g['ncases'])):
        T, extra, gam, lam, ud = g[f'c{i}_meta']
        adv = po.gae(int(T), int(extra), g[f'c{i}_r'], g[f'c{i}_v'], g[f'c{i}_d'], float(gam), float(lam), bool(ud))
        np.testing.assert_array_equal(adv, g[f'c{i}_adv'])


def test_dueling_and_epsilon_greedy_oracle_vs_reference_golden(golden_dir):
    """dueling.py:50-57 (through SimpleDiscreteActionValueHead) and policy_heads.py:219-224: the oracle reproduces the
    reference's Q-values to float32 rounding (matmul order) and its normalised action probabilities."""
    from oracle import dueling_oracle as do
    g = np.load(os.path.join(golden_dir, 'dueling.npz'))
    for i in range(int(g['ncases'])):
        A, nb, eps = g[f'c{i}_meta']
        params = {k: g[f'c{i}_p_{k}'] for k in do.PARAM_KEYS}
        q = do.dueling_forward(params, g[f'c{i}_feat'])
        assert q.shape == (int(nb), int(A))
        np.testing.assert_allclose(q, g[f'c{i}_q'], rtol=1e-5, atol=2e-6)
        pr = do.epsilon_greedy_probs(g[f'c{i}_q'], float(eps))
        pr = pr / pr.sum(-1, keepdims=True)                       # Categorical(probs=...) normalises
        np.testing.assert_allclose(pr, g[f'c{i}_probs'], rtol=0, atol=1.2e-7)
    # the reference's own test: all-zero weights give all-zero Q (test_dueling.py:7-21) = golden case 3
    assert not g['c3_q'].any()
    # ties: the first maximum is the greedy action (torch.argmax)
    pr = do.epsilon_greedy_probs(np.array([[1., 3., 3., 0.]], np.float32), 0.2)
    np.testing.assert_allclose(pr, [[0.05, 0.85, 0.05, 0.05]], rtol=1e-6)


def test_actor_rollout_oracle_vs_reference_rollout_manager_golden(golden_dir):
    """oracle/actor_oracle.py (the RolloutManager / Rollout / Categorical.sample restatement) against the golden written by
    the unmodified reference RolloutManager + Agent (oracle/make_golden.py::gen_actor_rollout): bit-exact."""
    from oracle import actor_oracle as ao, ppo_oracle as po
    g = np.load(os.path.join(golden_dir, 'actor_rollout.npz'))
    E, T, extra, A, n_rollouts, pseed = [int(v) for v in g['cfg']]
    params = po.init_params(A, ['extrinsic'], seed=pseed)
    for e in range(E):
        ros = ao.rollouts_oracle(params, e, T, extra, A, n_rollouts, ao.env_seed(e))
        for k, ro in enumerate(ros):
            np.testing.assert_array_equal(ro['actions'], g[f'r{k}_e{e}_actions'])
            np.testing.assert_array_equal(ro['rewards'], g[f'r{k}_e{e}_rew'])
            np.testing.assert_array_equal(ro['raw'], g[f'r{k}_e{e}_raw'])
            np.testing.assert_array_equal(ro['dones'], g[f'r{k}_e{e}_dones'])
            np.testing.assert_array_equal(ao.frame_checksums(ro['observations']), g[f'r{k}_e{e}_obs_sum'])
            for key in ('ep_len', 'ep_ret', 'ep_len_raw', 'ep_ret_raw'):
                np.testing.assert_allclose(np.asarray(ro['metadata'][key], np.float64), g[f'r{k}_e{e}_md_{key}'], rtol=1e-12, atol=1e-9)
    # the FrameStackWrapper restatement against the reference's wrapper (reset, 5 steps, reset, 3 steps)
    fs = ao.HostFrameStack(ao.SingleFrameEnv(9, A, False))
    stacks = [fs.reset()] + [fs.step(k % A)[0] for k in range(5)] + [fs.reset()] + [fs.step(k % A)[0] for k in range(3)]
    np.testing.assert_array_equal(ao.frame_checksums(np.stack(stacks)), g['framestack_sums'])


def test_dqn_oracle_vs_reference_modules_golden(golden_dir):
    """oracle/dqn_oracle.py (NatureCNN + dueling head + double-Q Bellman targets + TD loss, torch autograd) against
    tests/golden/dqn.npz = the same loss through the reference's OWN Agent / SimpleDiscreteActionValueHead(DuelingArchitecture) /
    SimpleDiscreteBellmanOptimalityOperator (oracle/make_golden.py::gen_dqn): Q values, targets, loss and every gradient tensor."""
    from oracle import dqn_oracle as do
    g = np.load(os.path.join(golden_dir, 'dqn.npz'))
    for i in range(int(g['ncases'])):
        A, nb, gamma, ud, dq, sl1, s_on, s_tg, s_case = g[f'c{i}_meta']
        A, nb = int(A), int(nb)
        p, pt = do.init_params(A, int(s_on)), do.init_params(A, int(s_tg))
        obs, nxt, actions, rewards, dones, weights = do.make_case(int(s_case), nb, A)
        res, grads = do.loss_and_grads(p, pt, obs, nxt, actions, rewards, dones, weights, float(gamma), bool(ud), bool(dq),
                                       'SmoothL1Loss' if sl1 else 'MSELoss')
        np.testing.assert_allclose(res['q'], g[f'c{i}_q'], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(res['y'], g[f'c{i}_y'], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(res['loss'], float(g[f'c{i}_loss']), rtol=1e-5)
        summ = do.summarize_grads(grads, list(grads.keys()))
        assert sum(float(g[f'c{i}_n_{k}']) > 0 for k in grads) >= 14            # the case exercises (nearly) every tensor
        for k, v in summ.items():
            ref = g[f'c{i}_{k}']
            np.testing.assert_allclose(v, ref, rtol=1e-4, atol=1e-6 * max(float(np.abs(ref).max()), 1e-6), err_msg=f'case {i} {k}')
