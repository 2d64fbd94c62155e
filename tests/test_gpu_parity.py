"""Parity of the CUDA path (through the C-ABI) against the CPU oracle on identical seeded inputs,
the committed golden vectors, and size-independent properties at BASELINE sizes.
All tests need a B200:  python -m pytest tests -m gpu
Tolerances are stated per test; integer/index work is bit-exact."""
import ctypes
import os

import numpy as np
import pytest
import torch as tc

from oracle import ppo_oracle as po, rnd_oracle as ro, sumtree_oracle as so

pytestmark = pytest.mark.gpu
F32 = np.float32


def dev():
    return tc.device('cuda', 0)


def cu(a):
    return tc.as_tensor(np.ascontiguousarray(a)).to(dev())


# ---- GAE / standardize -------------------------------------------------------------------------------
def test_gae_vs_reference_golden(golden_dir):
    from pytorch_drl_b200 import ops
    g = np.load(os.path.join(golden_dir, 'gae.npz'))
    for i in range(int(g['ncases'])):
        T, gamma, lam, ud = g[f'c{i}_meta']
        T = int(T)
        adv, vt = ops.gae(cu(g[f'c{i}_r'][None]), cu(g[f'c{i}_v'][None]), cu(g[f'c{i}_d'][None]),
                          float(gamma), float(lam), bool(ud), False, T)
        ref = g[f'c{i}_adv']
        # warp-parallel affine scan re-associates the recurrence: fp32 tolerance, not bit-exact
        np.testing.assert_allclose(adv.cpu().numpy()[0], ref, rtol=2e-5, atol=2e-6, err_msg=f'case {i}')
        np.testing.assert_allclose(vt.cpu().numpy()[0], ref + g[f'c{i}_v'][:T], rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize('N,T,use_dones,std', [(1024, 128, True, True), (64, 256, True, True),
                                               (33, 128, False, True), (7, 37, True, False), (5, 1000, True, True)])
def test_gae_batched_vs_oracle(N, T, use_dones, std):
    from pytorch_drl_b200 import ops
    rs = np.random.RandomState(N + T)
    ld = (T + 1 + 3) // 4 * 4
    r = np.zeros((N, ld), F32); v = np.zeros((N, ld), F32); d = np.zeros((N, ld), F32)
    r[:, :T] = rs.randint(-1, 2, size=(N, T)); v[:, :T + 1] = rs.standard_normal((N, T + 1))
    d[:, :T] = rs.uniform(size=(N, T)) < 0.01
    adv, vt = ops.gae(cu(r), cu(v), cu(d), 0.99, 0.95, use_dones, std, T)
    adv, vt = adv.cpu().numpy(), vt.cpu().numpy()
    for e in list(range(min(N, 6))) + [N - 1]:
        a_ref, vt_ref = po.credit_assignment(T, 0, r[e, :T], v[e, :T + 1], d[e, :T], 0.99, 0.95, use_dones, std)
        np.testing.assert_allclose(adv[e], a_ref, rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(vt[e], vt_ref, rtol=2e-5, atol=2e-5)
    if std:   # property at full size: per-env zero mean / unit unbiased std
        np.testing.assert_allclose(adv.mean(1), 0, atol=1e-5)
        np.testing.assert_allclose(adv.std(1, ddof=1), 1, rtol=1e-4)


def test_standardize_golden(golden_dir):
    from pytorch_drl_b200.utils import standardize
    g = np.load(os.path.join(golden_dir, 'standardize.npz'))
    for i in range(int(g['n'])):
        np.testing.assert_allclose(standardize(cu(g[f'x{i}'])).cpu().numpy(), g[f'y{i}'], rtol=1e-5, atol=1e-6)


def test_gae_dropin_signature():
    from pytorch_drl_b200.algos import GAE
    op = GAE(gamma=0.99, use_dones=True, lambda_=0.95)        # test_credit_assignment.py:7-61
    adv = op(rollout_len=2, extra_steps=0, rewards=cu(np.array([0., 0.], F32)),
             vpreds=cu(np.array([0., 0., 1.], F32)), dones=cu(np.array([1., 0.], F32)))
    np.testing.assert_allclose(adv.cpu().numpy(), [0.0, 0.99], rtol=1.3e-6, atol=1e-5)


# ---- fused losses -------------------------------------------------------------------------------------
def _random_loss_case(rs, B, A, rk, rw, clip, entc, vfc, vclip, simple, crit='MSELoss'):
    logits = (rs.standard_normal((B, A)) * 1.5).astype(F32)
    actions = rs.randint(0, A, size=B).astype(np.int64)
    values = {k: rs.standard_normal(B).astype(F32) for k in rk}
    lt = tc.from_numpy(logits)
    logp_new, _ = po.categorical_logprob_entropy(lt, tc.from_numpy(actions))
    lp_old = (logp_new.numpy() + rs.standard_normal(B) * 0.3).astype(F32)
    adv = {k: rs.standard_normal(B).astype(F32) for k in rk}
    v_old = {k: (values[k] + rs.standard_normal(B) * 0.3).astype(F32) for k in rk}
    vt = {k: rs.standard_normal(B).astype(F32) for k in rk}
    return dict(logits=logits, actions=actions, values=values, lp_old=lp_old, adv=adv, v_old=v_old, vt=vt)


def _oracle_loss(c, rk, rw, clip, entc, vfc, vclip, simple, crit):
    logits = tc.from_numpy(c['logits']).requires_grad_(True)
    vals = {k: tc.from_numpy(c['values'][k]).requires_grad_(True) for k in rk}
    logp, ent = po.categorical_logprob_entropy(logits, tc.from_numpy(c['actions']))
    out = po.ppo_losses(logp, ent, vals, tc.from_numpy(c['lp_old']), {k: tc.from_numpy(c['adv'][k]) for k in rk},
                        {k: tc.from_numpy(c['v_old'][k]) for k in rk}, {k: tc.from_numpy(c['vt'][k]) for k in rk},
                        rw, clip, entc, vfc, vclip, simple, crit)
    out['composite_loss'].backward()
    return ({k: float(v) for k, v in out.items()}, logits.grad.numpy(),
            np.stack([vals[k].grad.numpy() for k in rk], 1), logp.detach().numpy(), ent.detach().numpy())


@pytest.mark.parametrize('B,A,rk,rw,clip,entc,vfc,vclip,simple,crit', [
    (32, 6, ['extrinsic'], {'extrinsic': 1.0}, 0.2, 0.01, 1.0, False, True, 'MSELoss'),
    (4096, 6, ['extrinsic'], {'extrinsic': 1.0}, 0.1, 0.01, 0.5, True, True, 'MSELoss'),
    (1000, 18, ['extrinsic', 'intrinsic_rnd'], {'extrinsic': 2.0, 'intrinsic_rnd': 1.0}, 0.1, 0.001, 0.5, True, False, 'MSELoss'),
    (257, 18, ['extrinsic', 'intrinsic_rnd'], {'extrinsic': 2.0, 'intrinsic_rnd': 1.0}, 0.1, 0.001, 0.5, False, True, 'SmoothL1Loss'),
    (64, 32, ['extrinsic'], {'extrinsic': 1.0}, 0.2, 0.0, 1.0, True, True, 'SmoothL1Loss')])
def test_fused_loss_kernel_vs_oracle(B, A, rk, rw, clip, entc, vfc, vclip, simple, crit):
    from pytorch_drl_b200 import ops
    rs = np.random.RandomState(B + A)
    c = _random_loss_case(rs, B, A, rk, rw, clip, entc, vfc, vclip, simple, crit)
    ref, dl_ref, dv_ref, logp_ref, ent_ref = _oracle_loss(c, rk, rw, clip, entc, vfc, vclip, simple, crit)
    hyper = ops.make_hyper(A, [rw[k] for k in rk], clip, entc, vfc, vclip, simple, crit)
    keep = [cu(c['actions']), cu(c['lp_old'])] + [cu(c['adv'][k]) for k in rk] + [cu(c['v_old'][k]) for k in rk] + [cu(c['vt'][k]) for k in rk]
    K = len(rk)
    batch = ops.make_batch(keep[0], keep[1], keep[2:2 + K], keep[2 + K:2 + 2 * K], keep[2 + 2 * K:])
    losses, logp, ent, dl, dv = ops.ppo_loss(cu(c['logits']), cu(np.stack([c['values'][k] for k in rk], 1)), batch, hyper)
    got = dict(zip(ops.LOSS_NAMES, losses.cpu().tolist()))
    for n in ('policy_loss', 'value_loss', 'composite_loss', 'meanent'):
        np.testing.assert_allclose(got[n], ref[n], rtol=2e-5, atol=2e-6, err_msg=n)
    assert abs(got['clipfrac'] - ref['clipfrac']) <= 1.0 / B + 1e-6
    np.testing.assert_allclose(logp.cpu().numpy(), logp_ref, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ent.cpu().numpy(), ent_ref, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dl.cpu().numpy(), dl_ref, rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(dv.cpu().numpy(), dv_ref, rtol=1e-4, atol=1e-7)


def test_fused_loss_reference_known_answers():
    """drl/algos/test_ppo.py:17-41 and :44-88 through the kernel (logits chosen so that
    log-softmax reproduces the test's probabilities)."""
    from pytorch_drl_b200 import ops
    p_new = np.array([0.105, 0.08, 0.24, 0.095, 0.12, 0.16])
    logits = np.log(np.stack([p_new, 1 - p_new], 1)).astype(F32)
    lp_old = np.log(np.array([0.10, 0.10, 0.20, 0.10, 0.10, 0.20])).astype(F32)
    adv = np.array([1, 1, 1, -1, -1, -1], F32)
    v_new = np.array([1.05, 2.4, 1.2, -1.05, -2.4, -1.2], F32)
    v_old = np.array([1.0, 2.0, 1.0, -1.0, -2.0, -1.0], F32)
    vt = np.array([1.5, 1.5, 1.5, -1.5, -1.5, -1.5], F32)
    keep = [cu(np.zeros(6, np.int64)), cu(lp_old), cu(adv), cu(v_old), cu(vt)]
    batch = ops.make_batch(keep[0], keep[1], [keep[2]], [keep[3]], [keep[4]])
    for vclip, exp_v in ((False, [1.05, 2.4, 1.2, -1.05, -2.4, -1.2]), (True, [1.05, 2.4, 1.1, -1.05, -2.4, -1.1])):
        hyper = ops.make_hyper(2, [1.0], 0.10, 0.0, 1.0, vclip, True)
        losses, *_ = ops.ppo_loss(cu(logits), cu(v_new[:, None]), batch, hyper, want_grad=False)
        got = losses.cpu().numpy()
        np.testing.assert_allclose(-got[0], np.mean([1.05, 0.80, 1.10, -0.95, -1.20, -0.90]), rtol=2e-6, atol=1e-5)
        np.testing.assert_allclose(got[1], np.mean([(a - b) ** 2 for a, b in zip(exp_v, [1.5] * 3 + [-1.5] * 3)]), rtol=2e-6, atol=1e-5)


# ---- Adam ---------------------------------------------------------------------------------------------
def test_adam_vs_torch_semantics():
    from pytorch_drl_b200 import ops
    rs = np.random.RandomState(5)
    n = 1687719
    p = rs.standard_normal(n).astype(F32); m = np.zeros(n, F32); v = np.zeros(n, F32)
    P, M, V = cu(p), cu(m), cu(v)
    for step in range(1, 4):
        g = (rs.standard_normal(n) * 0.01).astype(F32)
        po.adam_step(p, g, m, v, step, 1e-3, 0.9, 0.999, 1e-5)
        ops.adam_step(P, cu(g), M, V, step, 1e-3, (0.9, 0.999), 1e-5)
    np.testing.assert_allclose(P.cpu().numpy(), p, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(M.cpu().numpy(), m, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(V.cpu().numpy(), v, rtol=1e-6, atol=1e-12)


# ---- sum-tree: bit-exact vs the C oracle -----------------------------------------------------------------
@pytest.mark.parametrize('cap,batch', [(64, 20), (1 << 20, 512), (1 << 20, 3000)])
def test_sumtree_bit_exact(cap, batch):
    from pytorch_drl_b200.replay import SumTree
    rs = np.random.RandomState(cap % 1000 + batch)
    ref = so.SumTree(cap)
    gpu = SumTree(cap)
    init = rs.uniform(0, 1, size=cap).astype(F32)
    ref.tree[cap:] = init
    ref.rebuild()
    gpu.set_leaves(cu(init))
    assert np.array_equal(gpu.tree.cpu().numpy().view(np.uint32), ref.tree.view(np.uint32))
    for it in range(3):
        idx = rs.randint(0, cap, size=batch).astype(np.int64)
        if it == 1:
            idx[: batch // 4] = idx[0]                 # collisions: the highest j wins
        pr = rs.uniform(0.001, 5.0, size=batch).astype(F32)
        ref.update(idx, pr)
        gpu.update(cu(idx), cu(pr))
        assert np.array_equal(gpu.tree.cpu().numpy().view(np.uint32), ref.tree.view(np.uint32)), f'update {it}'
        uni = rs.uniform(size=batch).astype(F32)
        u_ref = ref.stratified_u(uni)
        u_gpu = gpu.stratified_u(cu(uni))
        assert np.array_equal(u_gpu.cpu().numpy().view(np.uint32), u_ref.view(np.uint32))
        i_ref, p_ref = ref.sample(u_ref)
        i_gpu, p_gpu = gpu.sample(u_gpu)
        assert np.array_equal(i_gpu.cpu().numpy(), i_ref)           # sampled transitions: bit-exact
        assert np.array_equal(p_gpu.cpu().numpy().view(np.uint32), p_ref.view(np.uint32))


def test_sumtree_edge_cases():
    from pytorch_drl_b200.replay import SumTree
    t = SumTree(8)
    t.update(cu(np.zeros(0, np.int64)), cu(np.zeros(0, F32)))        # empty batch
    t.update(cu(np.array([3], np.int64)), cu(np.array([2.0], F32)))
    idx, pr = t.sample(cu(np.array([0.0, 1.999, 5.0], F32)))            # u >= total falls on the last massive leaf
    assert idx.cpu().tolist() == [3, 3, 3] and pr.cpu().tolist() == [2.0, 2.0, 2.0]
    with pytest.raises(ValueError):
        SumTree(12)


# ---- normaliser ----------------------------------------------------------------------------------------
def test_normalizer_vs_reference_golden(golden_dir):
    from pytorch_drl_b200 import ops
    from _util import close_to_summary
    g = np.load(os.path.join(golden_dir, 'normalizer.npz'))
    items, x = ro.normalizer_fixture_inputs()
    m1 = tc.zeros(84 * 84, device=dev()); m2 = tc.zeros(84 * 84, device=dev())
    steps = ops.normalizer_update(m1, m2, cu(items), 0.0)
    assert steps == float(g['steps'])
    close_to_summary(m1.cpu().numpy(), g['m1'], rtol=1e-6, always=True)
    close_to_summary(m2.cpu().numpy(), g['m2'], rtol=1e-6, always=True)
    y = ops.normalizer_apply(cu(x), m1, m2)
    close_to_summary(y.cpu().numpy(), g['y'], rtol=1e-5, always=True)


# ---- NatureCNN + PPO step -------------------------------------------------------------------------------
def _build(precision, A, rk, max_batch, param_seed=11):
    from pytorch_drl_b200.engine import LearnerEngine
    eng = LearnerEngine(A, [f'value_{k}' for k in rk], max_batch, precision)
    params = po.init_params(A, rk, seed=param_seed)
    eng.load_named({k: tc.from_numpy(v) for k, v in params.items()})
    return eng, params


TOL = {   # relative L2 error of a whole tensor vs the fp32 oracle; abs/rel tolerance on the loss scalars
    # fp32 = CUDA-core engine: differs from torch-CPU only by summation order.
    'fp32': dict(act=2e-5, grad=2e-4, loss=2e-5),
    # bf16 = tcgen05 engine: bf16 operands (eps 2^-9 = 2e-3 per element), fp32 accumulation.  Activations
    # stay within 1.5e-2; a gradient tensor of this 24-sample minibatch is a small difference of large
    # sums, which amplifies the operand rounding to a few percent of its norm (direction is kept:
    # cosine > 0.997, checked below).
    'bf16': dict(act=1.5e-2, grad=8e-2, loss=3e-3),
}


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def _precisions():
    return ['fp32', 'bf16']


@pytest.mark.parametrize('precision', _precisions())
def test_forward_vs_oracle(precision):
    A, rk, nb = 6, ['extrinsic'], 37
    eng, params = _build(precision, A, rk, 64)
    rs = np.random.RandomState(3)
    obs = rs.randint(0, 256, size=(50, 84, 84, 4), dtype=np.uint8)
    rows = rs.permutation(50)[:nb].astype(np.int64)
    actions = rs.randint(0, A, size=50).astype(np.int64)
    logits, values, logp, ent = eng.forward(cu(obs), cu(rows), actions=cu(actions))
    keep = {}
    p = {k: tc.from_numpy(v) for k, v in params.items()}
    with tc.no_grad():
        feat = po.trunk_forward(p, po.scale_obs(tc.from_numpy(obs[rows])), keep)
        l_ref, v_ref = po.heads_forward(p, feat, rk)
        lp_ref, e_ref = po.categorical_logprob_entropy(l_ref, tc.from_numpy(actions[rows]))
    t = TOL[precision]
    for i, name in enumerate(['act1', 'act2', 'act3']):
        got = eng.activation(i + 1, nb).float().cpu().numpy().reshape((nb,) + tuple(keep[name].shape[2:]) + (keep[name].shape[1],))
        ref = keep[name].permute(0, 2, 3, 1).numpy()
        assert rel_l2(got, ref) < t['act'], (name, rel_l2(got, ref))
    assert rel_l2(eng.activation(4, nb).cpu().numpy().reshape(nb, 512), keep['feat'].numpy()) < t['act']
    assert rel_l2(logits.cpu().numpy(), l_ref.numpy()) < t['act']
    assert rel_l2(values.cpu().numpy()[:, 0], v_ref['extrinsic'].numpy()) < t['act']
    np.testing.assert_allclose(logp.cpu().numpy(), lp_ref.numpy(), rtol=0, atol=t['loss'])
    np.testing.assert_allclose(ent.cpu().numpy(), e_ref.numpy(), rtol=0, atol=t['loss'])


def _bf16_round(x):
    return tc.from_numpy(np.asarray(x, np.float32)).to(tc.bfloat16).float().numpy()


def _bf16_ulp(x):
    return 2.0 ** (np.floor(np.log2(np.maximum(np.abs(x), 1e-30))) - 7)          # bf16: 8 significant bits


def test_conv1_integer_forward_exactness():
    """conv1 of the tcgen05 engine runs on the raw bytes (u8 x s8 integer MMAs, 16-bit fixed-point weights, s32 sums;
    umma_conv1_i8.cuh).  (1) Impulse responses: one 255 pixel per frame, channel 31 carries the code 1 + ci*64 + ky*8 + kx as
    its weight -> every output is that code to within half a weight step (tap geometry of the overlapping-row descriptors, both
    weight halves), zero elsewhere.
    (2) Range extremes: all 256 weights of a channel at +-max with all pixels 255 is the largest |256 S_hi + S_lo| the s32
    combine can meet -> equals the fp64 convolution after bf16 rounding.  (3) Random frames: act1 within HALF a bf16 ulp
    + 2e-4 of the fp64 oracle (weight quantisation 2^-15 of the channel maximum), i.e. bf16 output rounding is the only
    visible error.  nature_cnn.py:27-29, scale_observations.py:35."""
    A, rk = 6, ['extrinsic']
    eng, params = _build('bf16', A, rk, 64)
    wname = [k for k in params if params[k].shape == (32, 4, 8, 8)][0]
    bname = [k for k in params if params[k].shape == (32,)][0]
    code = (1 + np.arange(256)).reshape(4, 8, 8).astype(F32)
    w = np.stack([code * (co + 1) / 32.0 for co in range(32)]).astype(F32)
    eng.load_named({k: tc.from_numpy({wname: w, bname: np.zeros(32, F32)}.get(k, v)) for k, v in params.items()})
    pix = [(0, 0, 0), (0, 0, 3), (0, 5, 1), (3, 0, 2), (5, 7, 0), (40, 41, 1), (83, 83, 3), (17, 62, 2), (4, 4, 0), (7, 7, 3),
           (83, 0, 1), (0, 83, 2), (79, 79, 0), (80, 3, 3)]
    obs = np.zeros((len(pix), 84, 84, 4), np.uint8)
    for i, (y, x, c) in enumerate(pix):
        obs[i, y, x, c] = 255
    eng.forward(cu(obs), cu(np.arange(len(pix), dtype=np.int64)))
    got = eng.activation(1, len(pix)).float().cpu().numpy().reshape(len(pix), 20, 20, 32)
    for i, (y, x, c) in enumerate(pix):
        want = np.zeros((20, 20), F32)
        for oy in range(20):
            for ox in range(20):
                ky, kx = y - 4 * oy, x - 4 * ox
                if 0 <= ky < 8 and 0 <= kx < 8:
                    want[oy, ox] = code[c, ky, kx]
        # weight step of channel 31: 256 / 32639 (|q| <= 127 * 256 + 127) -> half a step + half a bf16 ulp; codes differ by >= 1
        assert np.all(np.abs(got[i, :, :, 31] - want) <= 0.5 * 256 / 32639 + 0.5 * _bf16_ulp(want)), f'impulse {(y, x, c)}'
        assert np.all((got[i, :, :, 31] != 0) == (want != 0))
    # (2) extremes of the s32 range
    rs = np.random.RandomState(5)
    w = np.zeros((32, 4, 8, 8), F32)
    w[0] = 0.75; w[1] = -0.75; w[2] = 0.75 * rs.choice([-1.0, 1.0], size=(4, 8, 8)); w[3] = rs.uniform(-0.75, 0.75, (4, 8, 8))
    w[4:] = rs.standard_normal((28, 4, 8, 8)) * 0.05
    bias = rs.uniform(-0.1, 0.1, 32).astype(F32)
    bias[1] = 200.0                                           # keeps the most negative sum visible through the ReLU
    eng.load_named({k: tc.from_numpy({wname: w, bname: bias}.get(k, v)) for k, v in params.items()})
    obs = np.full((3, 84, 84, 4), 255, np.uint8)
    obs[1] = rs.randint(0, 2, size=(84, 84, 4)) * 255
    obs[2] = rs.randint(0, 256, size=(84, 84, 4))
    eng.forward(cu(obs), cu(np.arange(3, dtype=np.int64)))
    got = eng.activation(1, 3).float().cpu().numpy().reshape(3, 20, 20, 32)
    x64 = tc.from_numpy(obs).permute(0, 3, 1, 2).double() * float(np.float32(1.0 / 255.0))
    ref = tc.relu(tc.nn.functional.conv2d(x64, tc.from_numpy(w).double(), tc.from_numpy(bias).double(), stride=4)).permute(0, 2, 3, 1).numpy()
    ulp = _bf16_ulp(ref)
    wmax = np.abs(w).reshape(32, -1).max(1)
    assert np.all(np.abs(got - ref) <= 0.5 * ulp + 256 * wmax * 2.0 ** -15 + 1e-6), np.abs(got - ref).max()
    assert got[0, 0, 0, 0] == _bf16_round(0.75 * 256 + bias[0]) and got[0, 0, 0, 1] == _bf16_round(200.0 - 0.75 * 256)
    # (3) random frames with the seeded NatureCNN weights
    eng, params = _build('bf16', A, rk, 64)
    obs = rs.randint(0, 256, size=(40, 84, 84, 4), dtype=np.uint8)
    rows = rs.permutation(40)[:33].astype(np.int64)
    eng.forward(cu(obs), cu(rows))
    got = eng.activation(1, 33).float().cpu().numpy().reshape(33, 20, 20, 32)
    x64 = tc.from_numpy(obs[rows]).permute(0, 3, 1, 2).double() * float(np.float32(1.0 / 255.0))
    ref = tc.relu(tc.nn.functional.conv2d(x64, tc.from_numpy(params[wname]).double(), tc.from_numpy(params[bname]).double(),
                                          stride=4)).permute(0, 2, 3, 1).numpy()
    assert np.all(np.abs(got - ref) <= 0.5 * _bf16_ulp(ref) + 2e-4), np.abs(got - ref).max()


def _ppo_case(A, rk, n_envs, T, B, seed=0):
    params = po.init_params(A, rk, seed=11)
    credit = {k: dict(gamma=0.99, lambda_=0.95, use_dones=(k == 'extrinsic')) for k in rk}
    raws = [po.make_synthetic_env_rollout(r, seed, T, A, rk) for r in range(n_envs)]
    rollouts = [po.collect_from_raw(params, raw, T, credit, True) for raw in raws]
    return params, raws, rollouts, credit


@pytest.mark.parametrize('precision', _precisions())
@pytest.mark.parametrize('A,rk,rw,vclip', [(6, ['extrinsic'], {'extrinsic': 1.0}, False),
                                           (18, ['extrinsic', 'intrinsic_rnd'], {'extrinsic': 2.0, 'intrinsic_rnd': 1.0}, True)])
def test_ppo_step_losses_and_grads_vs_oracle(precision, A, rk, rw, vclip):
    from pytorch_drl_b200 import ops
    from pytorch_drl_b200.algos import RolloutStorage
    n_envs, T, B = 3, 16, 8
    params, raws, rollouts, credit = _ppo_case(A, rk, n_envs, T, B)
    eng, _ = _build(precision, A, rk, n_envs * B)
    st = RolloutStorage(n_envs, T, rk, dev())
    d = st.as_dict()
    for e in range(n_envs):
        st.load_env(e, raws[e]['observations'], raws[e]['actions'], raws[e]['rewards'], raws[e]['dones'])
        d['logprobs'][e, :T] = cu(rollouts[e]['logprobs'])
        for k in rk:
            d['advantages'][k][e, :T] = cu(rollouts[e]['advantages'][k])
            d['vpreds'][k][e, :T] = cu(rollouts[e]['vpreds'][k])
            d['vtargs'][k][e, :T] = cu(rollouts[e]['vtargs'][k])
    rs = np.random.RandomState(9)
    blocks = [rs.permutation(T)[:B] for _ in range(n_envs)]
    rows = np.concatenate([b + e * st.Tp for e, b in enumerate(blocks)]).astype(np.int64)
    hp = po.Hyper(rw, 0.1, 0.01, 0.5, vclip, True)
    ref_m, ref_g = po.minibatch_losses_and_grads(params, [po.slice_env(rollouts[e], blocks[e]) for e in range(n_envs)], hp)
    hyper = ops.make_hyper(A, [rw[k] for k in rk], 0.1, 0.01, 0.5, vclip, True)
    batch = ops.make_batch(d['actions'], d['logprobs'], [d['advantages'][k] for k in rk],
                           [d['vpreds'][k] for k in rk], [d['vtargs'][k] for k in rk])
    losses = eng.ppo_step(d['observations'], cu(rows), batch, hyper).cpu().numpy()
    t = TOL[precision]
    got = dict(zip(ops.LOSS_NAMES, losses))
    for n in ('policy_loss', 'value_loss', 'composite_loss', 'meanent'):
        np.testing.assert_allclose(got[n], ref_m[n], rtol=t['loss'], atol=t['loss'], err_msg=n)
    assert abs(got['clipfrac'] - ref_m['clipfrac']) <= 2.0 / (n_envs * B) + 1e-6
    gv = eng.named_views(eng.grads)
    worst = {}
    for k in po.param_keys(rk):
        worst[k] = rel_l2(gv[k].cpu().numpy(), ref_g[k])
    short = lambda k: '.'.join(k.split('.')[-4:])  # noqa: E731
    bad = {short(k): round(v, 4) for k, v in worst.items() if not v < t['grad']}
    assert not bad, (bad, {short(k): round(v, 4) for k, v in worst.items()})
    for k in po.param_keys(rk):
        a, b = gv[k].cpu().numpy().ravel().astype(np.float64), ref_g[k].ravel().astype(np.float64)
        if np.linalg.norm(b) > 0:
            assert a @ b / (np.linalg.norm(a) * np.linalg.norm(b)) > 0.997, k
    # metrics-only pass agrees with the step's losses
    m2 = eng.ppo_step(d['observations'], cu(rows), batch, hyper, grads=False).cpu().numpy()
    np.testing.assert_allclose(m2, losses, rtol=1e-5, atol=1e-6)


# Bars of `PPO.optimize` against the unmodified reference's 2-rank DDP run, per precision mode (measured values in
# profiles/r02_precision.json; DESIGN.md §4 "Precision modes" explains the bf16 figures):
#   annotate  logp / vtarg / advantage of the T+1-frame no-grad forward              (abs / rel)
#   param     final parameter tensors after opt_epochs x T/B Adam steps: relative error of the summary vector
#   update    the same error divided by how far the reference moved that tensor (the 12 Adam updates) — the
#             bf16 engine's ReLU masks differ from the fp32 ones for ~0.1 % of the activations, which under Adam's
#             first steps (update ~ lr * sign(g)) shows as up to ~0.2 of the update norm
#   metric    per-epoch policy / value / composite loss and mean entropy (rel + abs)
GOLDEN_TOL = {
    'fp32': dict(logp=(1e-4, 1e-5), adv=(2e-3, 2e-4), vtarg=(1e-4, 1e-5), param_rtol=2e-3, param_atol=5e-5, update=1e-3,
                 metric=(5e-3, 5e-5)),
    'bf16': dict(logp=(0, 5e-4), adv=(5e-3, 2e-3), vtarg=(1e-3, 1e-3), param_rtol=None, param_atol=None, update=0.35,
                 metric=(1.5e-2, 4e-2)),
}


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
@pytest.mark.parametrize('tag', ['ppo_2rank', 'ppo_2rank_rnd'])
def test_ppo_optimize_vs_reference_2rank_golden(golden_dir, precision, tag):
    """The drop-in PPO.optimize on one GPU with envs_per_rank=2 against the UNMODIFIED reference run
    as 2 gloo ranks (tests/golden/ppo_2rank*.npz): final parameters, first-step grads, epoch metrics —
    for BOTH engines: the CUDA-core fp32 engine and the tcgen05 bf16 engine that bench.py times."""
    from pytorch_drl_b200 import agents, algos
    from _util import close_to_summary
    g = np.load(os.path.join(golden_dir, f'{tag}.npz'))
    T, B, A, epochs = int(g['cfg_T']), int(g['cfg_B']), int(g['cfg_A']), int(g['cfg_epochs'])
    seed, world = int(g['cfg_seed']), int(g['cfg_world'])
    clip, ent, vf, lr, b1, b2, eps, gamma, lam, vclip = [float(x) for x in g['cfg_hyper']]
    rk = [str(k) for k in g['cfg_rkeys']]
    rw = {k: float(w) for k, w in zip(rk, g['cfg_rweights'])}
    preds = {'policy': agents.CategoricalPolicyHead(512, A)}
    for k in rk:
        preds[f'value_{k}'] = agents.SimpleValueHead(512)
    agent = agents.Agent([agents.ToChannelMajor()], agents.NatureCNN(4), preds)
    agent.load_state_dict({k: tc.from_numpy(v) for k, v in po.init_params(A, rk, int(g['cfg_param_seed'])).items()})
    agent.fuse(max_batch=world * (T + 1), precision=precision)
    opt = tc.optim.Adam(agent.parameters(), lr=lr, betas=(b1, b2), eps=eps)
    credit = {k: algos.GAE(gamma=(gamma if k == 'extrinsic' else 0.99), use_dones=(k == 'extrinsic'), lambda_=lam) for k in rk}
    algo = algos.PPO(rank=0, world_size=1, rollout_len=T, extra_steps=0, credit_assignment_ops=credit,
                     global_step=1, max_steps=10 ** 9, policy_net=agent, policy_optimizer=opt,
                     standardize_adv=True, opt_epochs=epochs, learner_batch_size=B, clip_param_init=clip,
                     clip_param_final=clip, ent_coef_init=ent, ent_coef_final=ent, vf_loss_coef=vf,
                     vf_loss_clipping=bool(vclip), reward_weights=(rw if len(rk) > 1 else None),
                     envs_per_rank=world, sampler_seed=seed)
    st = algos.RolloutStorage(world, T, rk, dev())
    for r in range(world):
        raw = po.make_synthetic_env_rollout(r, seed, T, A, rk)
        st.load_env(r, raw['observations'], raw['actions'], raw['rewards'], raw['dones'])
    rollout = algo.credit_assignment(algo.annotate(st.as_dict()))
    tol = GOLDEN_TOL[precision]
    for r in range(world):
        np.testing.assert_allclose(rollout['logprobs'][r, :T].cpu().numpy(), g[f'r{r}::logp'], rtol=tol['logp'][0], atol=tol['logp'][1])
        for k in rk:
            np.testing.assert_allclose(rollout['advantages'][k][r, :T].cpu().numpy(), g[f'r{r}::adv_{k}'], rtol=tol['adv'][0], atol=tol['adv'][1])
            np.testing.assert_allclose(rollout['vtargs'][k][r, :T].cpu().numpy(), g[f'r{r}::vtarg_{k}'], rtol=tol['vtarg'][0], atol=tol['vtarg'][1])
    init = po.init_params(A, rk, int(g['cfg_param_seed']))
    algo.optimize(rollout)
    final = {k: v.detach().cpu().numpy() for k, v in agent.state_dict().items()}
    from _util import summary
    for k in po.param_keys(rk):
        gold = g[f'r0::final::{k}']
        if tol['param_rtol'] is not None:
            close_to_summary(final[k], gold, rtol=tol['param_rtol'], atol=tol['param_atol'])
        # error relative to the distance the reference moved the tensor (sum / norm entries of the summary left out)
        big = final[k].size > 4096
        got_s = summary(final[k])[2:] if big else np.asarray(final[k], np.float64).ravel()
        init_s = summary(init[k])[2:] if big else np.asarray(init[k], np.float64).ravel()
        gold_s = np.asarray(gold, np.float64).ravel()[2:] if big else np.asarray(gold, np.float64).ravel()
        moved = np.linalg.norm(gold_s - init_s)
        assert np.linalg.norm(got_s - gold_s) <= tol['update'] * moved + 1e-7, (k, np.linalg.norm(got_s - gold_s) / max(moved, 1e-30))
    em = g['r0::epoch_metrics']
    for ep in range(epochs):
        got = np.array([algo.last_epoch_metrics[ep][n] for n in algos.ppo.METRIC_NAMES])
        np.testing.assert_allclose(got[:4], em[ep][:4], rtol=tol['metric'][0], atol=tol['metric'][1])
        assert abs(got[4] - em[ep][4]) <= 2.0 / T + 1e-6


# ---- RND predictor ---------------------------------------------------------------------------------------
# fp32 = CUDA-core engine (summation order only); bf16 = tcgen05 engine (bf16 operands incl. the normalised frame, fp32 accumulation)
RND_TOL = {'fp32': dict(per=2e-4, loss=2e-4, grad=2e-4), 'bf16': dict(per=2e-2, loss=1e-2, grad=0.1)}     # 8-sample fixture: see TOL['bf16']


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_rnd_learn_vs_reference_golden(golden_dir, precision):
    """Predictor loss, per-sample intrinsic reward and student gradients against the UNMODIFIED
    reference networks (tests/golden/rnd.npz) and the oracle, for both engines (bars in RND_TOL)."""
    from pytorch_drl_b200.envs import RandomNetworkDistillation
    g = np.load(os.path.join(golden_dir, 'rnd.npz'))
    items, obs_u8 = ro.rnd_fixture_inputs()
    rnd = RandomNetworkDistillation({'lr': 1e-4, 'eps': 1e-5}, world_size=32, max_batch=16, precision=precision)
    teacher, student = ro.init_net(False, 100), ro.init_net(True, 200)
    rnd.load({k: tc.from_numpy(v) for k, v in teacher.items()}, {k: tc.from_numpy(v) for k, v in student.items()})
    rnd.observe(cu(items))
    rnd._sync_normalizers_global(); rnd._sync_normalizers_local()
    nz = ro.Normalizer([84, 84, 1])
    for it in items:
        nz.update(it)
    loss_ref, per_ref, y_ref, yh_ref, grads_ref = ro.rnd_learn_loss_and_grads(teacher, student, nz, obs_u8)
    np.testing.assert_allclose(loss_ref, float(g['loss']), rtol=1e-5)
    t = RND_TOL[precision]
    obs = cu(obs_u8)
    err = rnd.intrinsic_reward(obs).cpu().numpy()
    np.testing.assert_allclose(err, g['per'], rtol=t['per'])
    before = rnd.student_params.clone()
    rows = tc.arange(obs_u8.shape[0], device=dev(), dtype=tc.int64)
    loss = rnd.learn_rows(obs, rows)
    np.testing.assert_allclose(float(loss.item()), float(g['loss']), rtol=t['loss'])
    gv = rnd.student_views(rnd.student_grads)
    worst = {k: rel_l2(gv[k].cpu().numpy(), grads_ref[k]) for k in ro.STUDENT_KEYS}
    assert max(worst.values()) < t['grad'], worst
    # one Adam step moved the student (lr 1e-4, first step = lr * sign)
    delta = (rnd.student_params - before).abs().max().item()
    assert 0.5e-4 < delta < 1.5e-4


@pytest.mark.parametrize('nb', [1, 2, 7, 148, 297, 301])
def test_rnd_bf16_engine_ragged_batches_vs_fp32_engine(nb):
    """The tcgen05 RND engine packs sample PAIRS along the channels: odd batches (a phantom second sample), fewer pairs than
    SMs, more than one pair per CTA, row-index gathers.  Rewards, loss and every student gradient tensor must agree with the
    CUDA-core fp32 engine (itself checked against the reference golden); a second learn step checks the repacked weights."""
    from pytorch_drl_b200.envs import RandomNetworkDistillation
    g = tc.Generator(device=dev())
    g.manual_seed(500 + nb)
    frames = tc.randint(0, 256, (nb + 5, 84, 84, 4), device=dev(), dtype=tc.uint8, generator=g)
    rows = tc.randperm(nb + 5, device=dev(), generator=g)[:nb].contiguous()
    items = tc.rand((6, 84, 84, 1), device=dev(), generator=g)
    teacher, student = ro.init_net(False, 101), ro.init_net(True, 201)
    out = {}
    for precision in ('fp32', 'bf16'):
        rnd = RandomNetworkDistillation({'lr': 1e-3, 'eps': 1e-5}, world_size=32, max_batch=nb, precision=precision)
        rnd.load({k: tc.from_numpy(v) for k, v in teacher.items()}, {k: tc.from_numpy(v) for k, v in student.items()})
        rnd.observe(items)
        rnd._sync_normalizers_global(); rnd._sync_normalizers_local()
        err = rnd.intrinsic_reward(frames, rows).cpu().numpy()
        loss = float(rnd.learn_rows(frames, rows).item())
        grads = {k: v.cpu().numpy().copy() for k, v in rnd.student_views(rnd.student_grads).items()}
        loss2 = float(rnd.learn_rows(frames, rows).item())          # after one Adam step: exercises the operand repack
        out[precision] = (err, loss, grads, loss2)
    (e32, l32, g32, l32b), (e16, l16, g16, l16b) = out['fp32'], out['bf16']
    np.testing.assert_allclose(e16, e32, rtol=3e-2)
    np.testing.assert_allclose([l16, l16b], [l32, l32b], rtol=1e-2)
    assert l32b < l32                                               # the step reduced the predictor loss
    for k in g32:
        assert np.isfinite(g16[k]).all(), k
        assert rel_l2(g16[k], g32[k]) < (0.15 if nb < 16 else 8e-2), (nb, k, rel_l2(g16[k], g32[k]))     # single-digit batches: see TOL['bf16']


def test_rnd_learn_wrapper_semantics():
    """learn(mb): row dropping by world/32 (random_network_distillation.py:203-207), no update before
    non_learning_steps, normaliser sync aliasing."""
    from pytorch_drl_b200.envs import RandomNetworkDistillation
    from pytorch_drl_b200.utils import MinibatchView
    rs = np.random.RandomState(0)
    obs = cu(rs.randint(0, 256, size=(1, 16, 84, 84, 4), dtype=np.uint8))
    rnd = RandomNetworkDistillation({'lr': 1e-4, 'eps': 1e-5}, world_size=8, non_learning_steps=4, max_batch=16)
    rnd.observe(cu(rs.uniform(0, 1, size=(3, 84, 84, 1)).astype(F32)))
    mb = MinibatchView({'observations': obs}, tc.arange(16, device=dev(), dtype=tc.int64))
    p0 = rnd.student_params.clone()
    rnd.learn(mb)                                   # steps (3) < non_learning_steps (4): no predictor update
    assert tc.equal(rnd.student_params, p0) and rnd._synced_normalizer.steps == 3.0
    assert rnd._unsynced_normalizer.moment1 is rnd._synced_normalizer.moment1      # aliasing quirk reproduced
    rnd.observe(cu(rs.uniform(0, 1, size=(2, 84, 84, 1)).astype(F32)))
    rnd.learn(mb)                                   # the check reads the SYNCED steps (still 3, :224): no update yet
    assert tc.equal(rnd.student_params, p0) and rnd._synced_normalizer.steps == 5.0
    np.random.seed(0)
    rnd.learn(mb)                                   # now 5 >= 4: keeps int(8/32 * 16) = 4 rows and updates
    assert not tc.equal(rnd.student_params, p0) and rnd.adam_steps == 1


PAIR_DEFAULT = 11     # library default of debug flag 7 (conv kernels on CTA pairs)


def test_alternative_bf16_engines_stay_in_parity():
    """The kernel variants still selectable through drl_debug_set (fc GEMMs with / without CTA pairs and 2 / 4 epilogue
    warpgroups; conv kernels on single CTAs, CTA pairs, sample-spanning tiles) must keep producing the default gradients."""
    from pytorch_drl_b200 import _lib, ops
    from pytorch_drl_b200.algos import RolloutStorage
    L = _lib.lib()
    A, rk, rw = 6, ['extrinsic'], {'extrinsic': 1.0}
    n_envs, T, B = 3, 16, 8
    params, raws, rollouts, credit = _ppo_case(A, rk, n_envs, T, B)
    st = RolloutStorage(n_envs, T, rk, dev())
    d = st.as_dict()
    for e in range(n_envs):
        st.load_env(e, raws[e]['observations'], raws[e]['actions'], raws[e]['rewards'], raws[e]['dones'])
        d['logprobs'][e, :T] = cu(rollouts[e]['logprobs'])
        for k in rk:
            for name in ('advantages', 'vpreds', 'vtargs'):
                d[name][k][e, :T] = cu(rollouts[e][name][k])
    rs = np.random.RandomState(9)
    blocks = [rs.permutation(T)[:B] for _ in range(n_envs)]
    rows = cu(np.concatenate([b + e * st.Tp for e, b in enumerate(blocks)]).astype(np.int64))
    hp = po.Hyper(rw, 0.1, 0.01, 0.5, False, True)
    _, ref_g = po.minibatch_losses_and_grads(params, [po.slice_env(rollouts[e], blocks[e]) for e in range(n_envs)], hp)
    hyper = ops.make_hyper(A, [1.0], 0.1, 0.01, 0.5, False, True)
    batch = ops.make_batch(d['actions'], d['logprobs'], [d['advantages']['extrinsic']], [d['vpreds']['extrinsic']],
                           [d['vtargs']['extrinsic']])
    try:
        for f6, f7 in ((0, 0), (2, 0), (1, 7), (5, 3), (5, 11)):
            L.drl_debug_set(6, f6)      # fc GEMMs: bits 0-1 CTA pairs (0 none, 1 forward, 2 forward + data gradient), bit 2 four epilogue warpgroups in the data gradient
            L.drl_debug_set(7, f7)      # conv kernels on CTA pairs: bit 0 conv2 forward, bit 1 conv3 dgrad, bit 2 conv2 dgrad, bit 3 conv3 dgrad with sample-spanning tiles
            flags = (f6, f7)
            eng, _ = _build('bf16', A, rk, n_envs * B)
            eng.ppo_step(d['observations'], rows, batch, hyper)
            gv = eng.named_views(eng.grads)
            for k in po.param_keys(rk):
                assert rel_l2(gv[k].cpu().numpy(), ref_g[k]) < TOL['bf16']['grad'], (flags, k)
    finally:
        L.drl_debug_set(6, 5)
        L.drl_debug_set(7, PAIR_DEFAULT)


# ---- reward normaliser -------------------------------------------------------------------
def test_reward_normalizer_vs_reference_golden(golden_dir):
    """One environment per golden case, fed in blocks of 128 steps like a rollout: the window updates must land
    on the same steps and give the reference's moments and normalised probes (normalize_reward.py:20-100)."""
    from pytorch_drl_b200.envs import BatchedRewardNormalizer
    g = np.load(os.path.join(golden_dir, 'reward_norm.npz'))
    probes = g['probes']
    for i in range(int(g['ncases'])):
        gamma, use_dones, steps, L = g[f'c{i}_meta']
        rn = BatchedRewardNormalizer(1, float(gamma), bool(use_dones), device=dev())
        assert rn.trace_length == int(L)
        r, d, trace = g[f'c{i}_r'], g[f'c{i}_d'], g[f'c{i}_trace']
        assert float(rn.normalize(cu(np.ones(1, np.float32)))[0]) == 0.0          # no statistics yet
        got, t0 = [], 0
        while t0 < int(steps):
            n = min(128, int(steps) - t0)
            for t in range(t0, t0 + n):                                              # step granularity to catch the update step
                before = float(rn.steps[0])
                rn.step_block(cu(r[None, t:t + 1]), cu(d[None, t:t + 1]))
                if float(rn.steps[0]) != before:
                    rn.synced.copy_(tc.stack([rn.steps[0], rn.moment1[0], rn.moment2[0]]))
                    got.append([t, float(rn.steps[0]), float(rn.moment1[0]), float(rn.moment2[0])] +
                               rn.normalize(cu(probes)).cpu().tolist())
            t0 += n
        got = np.array(got)
        assert got.shape == trace.shape
        np.testing.assert_array_equal(got[:, :2], trace[:, :2])
        np.testing.assert_allclose(got[:, 2:], trace[:, 2:], rtol=2e-6, atol=1e-7)


def test_reward_normalizer_blocks_vs_oracle():
    """64 environments x 3 rollouts of 700 steps with a learn() between them, whole [N, T] blocks per call, against
    the oracle's batched restatement (moments of every environment and every normalised reward)."""
    from oracle import reward_norm_oracle as rno
    from pytorch_drl_b200.envs import BatchedRewardNormalizer
    N, T, gamma = 64, 700, 0.9
    rs = np.random.RandomState(3)
    ref = rno.BatchedRewardNormalizer(N, gamma, True)
    rn = BatchedRewardNormalizer(N, gamma, True, device=dev())
    for it in range(3):
        r = (rs.randint(-1, 2, size=(N, T)) + rs.standard_normal((N, T)) * 0.05).astype(np.float32)
        d = (rs.uniform(size=(N, T)) < 0.02).astype(np.float32)
        want = ref.step_block(r, d)
        got = rn.step_block(cu(r), cu(d)).cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-7)
        for name in ('steps', 'moment1', 'moment2'):
            np.testing.assert_allclose(getattr(rn, name).cpu().numpy(), np.array([getattr(a, name) for a in ref.unsynced]),
                                       rtol=2e-6, atol=1e-7, err_msg=name)
        ref.learn()
        rn.learn()
        np.testing.assert_allclose(rn.synced.cpu().numpy(), [ref.synced.steps, ref.synced.moment1, ref.synced.moment2], rtol=2e-6)


# ---- n-step advantages / Bellman backups ----------------------------------------------------
def test_nstep_and_bellman_vs_reference_golden(golden_dir):
    """Bit-exact against the reference's NStepAdvantageEstimator / SimpleDiscreteBellmanOptimalityOperator, single
    trajectory (the reference's call shape) and 3 identical envs per launch."""
    from pytorch_drl_b200.algos.credit_assignment import NStepAdvantageEstimator, SimpleDiscreteBellmanOptimalityOperator
    g = np.load(os.path.join(golden_dir, 'nstep_bellman.npz'))
    for i in range(int(g['ncases'])):
        T, extra, gam, ud, A = g[f'c{i}_meta']
        T, extra = int(T), int(extra)
        r, v, d, q, qt = (cu(g[f'c{i}_{k}']) for k in ('r', 'v', 'd', 'q', 'qt'))
        op = NStepAdvantageEstimator(gamma=float(gam), use_dones=bool(ud))
        np.testing.assert_array_equal(op(rollout_len=T, extra_steps=extra, rewards=r, vpreds=v, dones=d).cpu().numpy(), g[f'c{i}_adv'])
        many = op(rollout_len=T, extra_steps=extra, rewards=r.repeat(3, 1), vpreds=v.repeat(3, 1), dones=d.repeat(3, 1)).cpu().numpy()
        for e in range(3):
            np.testing.assert_array_equal(many[e], g[f'c{i}_adv'])
        for dq, key in ((False, 'bb'), (True, 'bb_dq')):
            bop = SimpleDiscreteBellmanOptimalityOperator(gamma=float(gam), use_dones=bool(ud), double_q=dq)
            got = bop(rollout_len=T, extra_steps=extra, rewards=r, qpreds=q, tgt_qpreds=qt, dones=d).cpu().numpy()
            np.testing.assert_array_equal(got, g[f'c{i}_{key}'])
            got3 = bop(rollout_len=T, extra_steps=extra, rewards=r.repeat(3, 1), qpreds=q.repeat(3, 1, 1),
                       tgt_qpreds=qt.repeat(3, 1, 1), dones=d.repeat(3, 1)).cpu().numpy()
            np.testing.assert_array_equal(got3[2], g[f'c{i}_{key}'])


def test_gae_extra_steps_vs_reference_golden(golden_dir):
    """GAE.call with extra_steps > 0 (credit_assignment.py:204-214); same fp32 tolerance as the T-step case (the
    warp-parallel affine scan re-associates the recurrence)."""
    from pytorch_drl_b200.algos.credit_assignment import GAE
    g = np.load(os.path.join(golden_dir, 'gae_extra.npz'))
    for i in range(int(g['ncases'])):
        T, extra, gam, lam, ud = g[f'c{i}_meta']
        op = GAE(gamma=float(gam), use_dones=bool(ud), lambda_=float(lam))
        adv = op(rollout_len=int(T), extra_steps=int(extra), rewards=cu(g[f'c{i}_r']), vpreds=cu(g[f'c{i}_v']), dones=cu(g[f'c{i}_d']))
        np.testing.assert_allclose(adv.cpu().numpy(), g[f'c{i}_adv'], rtol=2e-5, atol=2e-6)


# ---- error behaviour at the boundary (ValueError for argument errors, RuntimeError otherwise) -----
def test_error_behaviour_matches_reference_conventions():
    import ctypes
    from pytorch_drl_b200 import _lib, ops
    from pytorch_drl_b200.algos import IndependentSampler
    from pytorch_drl_b200.algos.credit_assignment import GAE, get_credit_assignment_op
    from pytorch_drl_b200.engine import LearnerEngine
    L = _lib.lib()
    # shapes that do not fit: ValueError (credit_assignment.py:25 / samplers.py:51 style)
    with pytest.raises(ValueError):
        ops.gae(cu(np.zeros((2, 8), np.float32)), cu(np.zeros((2, 8), np.float32)), cu(np.zeros((2, 8), np.float32)),
                0.99, 0.95, True, False, 8)                                      # vpreds needs T+1 entries
    with pytest.raises(ValueError):
        get_credit_assignment_op('NoSuchOp', {})
    with pytest.raises((AssertionError, ValueError)):
        IndependentSampler(10, 4)                                                 # T % B != 0 (samplers.py:51)
    # C-ABI status codes, called raw
    assert L.drl_gae_f32(None, None, None, 1, 8, 8, 9, 8, ctypes.c_float(0.99), ctypes.c_float(0.95), 1, 0, None, None, None) == 2
    h = ctypes.c_void_p()
    assert L.drl_net_create(ctypes.byref(h), 0, 0, 6, 1, 1) == 1                  # max_batch 0: bad shape
    assert L.drl_net_create(ctypes.byref(h), 99, 64, 6, 1, 1) == 5                # no such device: NO_DEVICE, not a fallback
    assert L.drl_rnd_create(ctypes.byref(h), 0, 64, 2, 1) == 6                    # RND widening != 1: UNSUPPORTED, loudly
    assert L.drl_rnd_create(ctypes.byref(h), 0, 64, 1, 7) == 2                    # unknown precision: bad argument
    eng = LearnerEngine(6, ['value_extrinsic'], 8, 'bf16')
    obs = tc.zeros((16, 84, 84, 4), dtype=tc.uint8, device=dev())
    logits, values = tc.empty((16, 6), device=dev()), tc.empty((16, 1), device=dev())
    # one C call may not exceed the handle's max_batch (the Python engine chunks for the caller)
    assert L.drl_net_forward(eng._h, _lib.ptr(obs), None, 16, _lib.ptr(logits), _lib.ptr(values), None, None, None, None) == 1
    assert L.drl_net_forward(eng._h, None, None, 8, _lib.ptr(logits), _lib.ptr(values), None, None, None, None) == 2
    lg, vl, _, _ = eng.forward(obs)
    assert lg.shape == (16, 6) and vl.shape == (16, 1)
    assert L.drl_status_string(6).decode() != ''


# ---- ragged / boundary batch sizes: the tcgen05 engine against the (oracle-checked) fp32 engine ------------------
@pytest.mark.parametrize('nb', [1, 2, 3, 127, 129, 149, 257, 297, 300, 1000])
def test_bf16_engine_ragged_batches_vs_fp32_engine(nb):
    """Sizes around every tiling boundary of the tcgen05 kernels: fewer samples than SMs, one sample more than the
    SM count (CTAs with 2 and with 1 samples; odd sample-pair tails), non-multiples of the 128- and 256-row GEMM
    tiles of the fc layer and its CTA pairs.  Activations, logits, losses and every gradient tensor must agree
    with the CUDA-core fp32 engine (itself checked against the oracle) within the bf16 bars."""
    from pytorch_drl_b200 import ops
    A, rk = 6, ['extrinsic']
    g = tc.Generator(device=dev())
    g.manual_seed(100 + nb)
    obs = tc.randint(0, 256, (nb, 84, 84, 4), device=dev(), dtype=tc.uint8, generator=g)
    rows = tc.randperm(nb, device=dev(), generator=g)
    actions = tc.randint(0, A, (nb,), device=dev(), generator=g)
    logp_old = tc.full((nb,), -float(np.log(A)), device=dev())
    adv = tc.randn(nb, device=dev(), generator=g)
    vold = tc.randn(nb, device=dev(), generator=g) * 0.1
    vtarg = tc.randn(nb, device=dev(), generator=g)
    hyper = ops.make_hyper(A, [1.0], 0.1, 0.01, 0.5, False, True)
    batch = ops.make_batch(actions, logp_old, [adv], [vold], [vtarg])
    out = {}
    for precision in ('fp32', 'bf16'):
        eng, _ = _build(precision, A, rk, nb)
        losses = eng.ppo_step(obs, rows, batch, hyper).cpu().numpy()
        grads = {k: v.cpu().numpy().copy() for k, v in eng.named_views(eng.grads).items()}
        logits, values, _, _ = eng.forward(obs, rows)
        out[precision] = (losses, grads, logits.cpu().numpy(), values.cpu().numpy())
        del eng
    (l32, g32, lg32, v32), (l16, g16, lg16, v16) = out['fp32'], out['bf16']
    t = TOL['bf16']
    assert np.isfinite(l16).all() and all(np.isfinite(v).all() for v in g16.values())
    assert rel_l2(lg16, lg32) < t['act'] and rel_l2(v16, v32) < 2 * t['act']
    np.testing.assert_allclose(l16[:4], l32[:4], rtol=5 * t['loss'], atol=5 * t['loss'])
    for k in g32:
        # single-digit batches make some gradient tensors tiny differences of large sums: direction + norm check
        a, b = g16[k].ravel().astype(np.float64), g32[k].ravel().astype(np.float64)
        if np.linalg.norm(b) > 0:
            assert a @ b / (np.linalg.norm(a) * np.linalg.norm(b)) > 0.99, (nb, k)
            assert rel_l2(g16[k], g32[k]) < 0.15, (nb, k, rel_l2(g16[k], g32[k]))


def test_bf16_engine_reuse_across_batch_sizes():
    """One network handle, learner steps of different sizes back to back (workspaces, ReLU-mask buffers and TMEM
    are reused): each must equal what a fresh handle of exactly that size computes."""
    from pytorch_drl_b200 import ops
    A, rk = 6, ['extrinsic']
    g = tc.Generator(device=dev())
    g.manual_seed(7)
    big = 300
    obs = tc.randint(0, 256, (big, 84, 84, 4), device=dev(), dtype=tc.uint8, generator=g)
    actions = tc.randint(0, A, (big,), device=dev(), generator=g)
    logp_old = tc.full((big,), -float(np.log(A)), device=dev())
    adv, vold, vtarg = (tc.randn(big, device=dev(), generator=g) for _ in range(3))
    hyper = ops.make_hyper(A, [1.0], 0.1, 0.01, 0.5, False, True)
    batch = ops.make_batch(actions, logp_old, [adv], [vold], [vtarg])
    shared, _ = _build('bf16', A, rk, big)
    for nb in (300, 1, 149, 300, 2):
        rows = tc.arange(nb, device=dev(), dtype=tc.int64)
        l_shared = shared.ppo_step(obs, rows, batch, hyper).cpu().numpy()
        g_shared = shared.grads.cpu().numpy().copy()
        fresh, _ = _build('bf16', A, rk, nb)
        l_fresh = fresh.ppo_step(obs, rows, batch, hyper).cpu().numpy()
        g_fresh = fresh.grads.cpu().numpy()
        np.testing.assert_allclose(l_shared, l_fresh, rtol=1e-5, atol=1e-6)
        assert np.linalg.norm(g_shared - g_fresh) <= 1e-4 * max(np.linalg.norm(g_fresh), 1e-30), nb   # atomics order only
        del fresh


def test_bf16_engine_arithmetic_with_its_own_relu_masks():
    """Separates the two error sources of the tcgen05 engine's gradient (DESIGN.md "Precision modes"): at 512 samples the
    per-tensor distance to the fp32 oracle is 5-7 % — almost all of it the ~0.1 % of ReLU units that switch when activations are
    rounded to bf16.  Against the oracle evaluated WITH THE DEVICE'S OWN ReLU masks (z * mask instead of relu(z), masks read back
    from act1 / act2 / act3 / features of the same step) only the arithmetic error of the bf16 operands is left: every parameter
    tensor within 1e-2 (measured 0.2-0.54 %), the whole gradient within 5e-3."""
    from pytorch_drl_b200 import ops
    A, rk, nb = 6, ['extrinsic'], 512
    eng, params = _build('bf16', A, rk, nb)
    rs = np.random.RandomState(77)
    obs = rs.randint(0, 256, size=(nb, 84, 84, 4), dtype=np.uint8)
    actions = rs.randint(0, A, size=nb).astype(np.int64)
    logp_old = np.full(nb, -np.log(A), F32)
    adv, vold, vtarg = (rs.standard_normal(nb).astype(F32) for _ in range(3))
    hyper = ops.make_hyper(A, [1.0], 0.1, 0.01, 0.5, False, True)
    d_act, d_lp, d_adv, d_vold, d_vt = cu(actions), cu(logp_old), cu(adv), cu(vold), cu(vtarg)     # the batch struct holds raw pointers
    batch = ops.make_batch(d_act, d_lp, [d_adv], [d_vold], [d_vt])
    eng.ppo_step(cu(obs), cu(np.arange(nb, dtype=np.int64)), batch, hyper)
    got = {k: v.cpu().numpy().astype(np.float64) for k, v in eng.named_views(eng.grads).items()}
    shapes = {'act1': (nb, 20, 20, 32), 'act2': (nb, 9, 9, 64), 'act3': (nb, 7, 7, 64)}
    masks = {k: (eng.activation(i + 1, nb).float().cpu().reshape(shp) > 0).float().permute(0, 3, 1, 2).contiguous()
             for i, (k, shp) in enumerate(shapes.items())}
    # the learner step keeps its features in bf16 only; a forward with the same (not yet updated) weights reproduces them in fp32
    masks['feat'] = (eng.features(cu(obs)).cpu() > 0).float()
    env = dict(observations=obs, actions=actions, logprobs=logp_old, advantages={'extrinsic': adv}, vpreds={'extrinsic': vold},
               vtargs={'extrinsic': vtarg})
    hp = po.Hyper({'extrinsic': 1.0}, 0.1, 0.01, 0.5, False, True)
    _, g_masked = po.minibatch_losses_and_grads(params, [env], hp, masks=[masks])
    _, g_plain = po.minibatch_losses_and_grads(params, [env], hp)
    per_masked = {k: rel_l2(got[k], g_masked[k]) for k in g_masked}
    per_plain = {k: rel_l2(got[k], g_plain[k]) for k in g_plain}
    flat = lambda d: np.concatenate([np.asarray(d[k], np.float64).ravel() for k in g_masked])
    print('bf16 gradient rel-L2 per tensor, oracle with the device masks:', {k[-30:]: round(v, 4) for k, v in per_masked.items()})
    print('bf16 gradient rel-L2 per tensor, plain oracle:               ', {k[-30:]: round(v, 4) for k, v in per_plain.items()})
    assert max(per_plain.values()) < 0.12, per_plain
    assert max(per_masked.values()) < 1e-2, per_masked
    assert rel_l2(flat(got), flat(g_masked)) < 5e-3


def test_full_size_step_properties():
    """BASELINE.json config[1] size (1024 envs x 32 samples = 32 768 per step), where the oracle is too slow:
    size-independent properties of the learner step on the tcgen05 engine —
      * permutation invariance: the same minibatch in another row order gives the same losses and gradients,
      * additivity: every loss is a mean over samples, so grad(all) = (grad(first half) + grad(second half)) / 2,
      * consistency with the oracle-checked fp32 engine on a 512-sample slice of the same data."""
    from pytorch_drl_b200 import ops
    A, rk, nb = 6, ['extrinsic'], 32768
    g = tc.Generator(device=dev())
    g.manual_seed(2024)
    obs = tc.empty((nb, 84, 84, 4), device=dev(), dtype=tc.uint8)
    for s in range(0, nb, 4096):
        obs[s:s + 4096] = tc.randint(0, 256, (4096, 84, 84, 4), device=dev(), dtype=tc.uint8, generator=g)
    actions = tc.randint(0, A, (nb,), device=dev(), generator=g)
    logp_old = tc.full((nb,), -float(np.log(A)), device=dev())
    adv, vold, vtarg = (tc.randn(nb, device=dev(), generator=g) for _ in range(3))
    hyper = ops.make_hyper(A, [1.0], 0.1, 0.01, 0.5, False, True)
    batch = ops.make_batch(actions, logp_old, [adv], [vold], [vtarg])
    eng, _ = _build('bf16', A, rk, nb)

    def step(rows):
        losses = eng.ppo_step(obs, rows, batch, hyper).cpu().numpy().astype(np.float64)
        return losses, eng.grads.cpu().numpy().astype(np.float64)

    ident = tc.arange(nb, device=dev(), dtype=tc.int64)
    l_all, g_all = step(ident)
    assert np.isfinite(l_all).all() and np.isfinite(g_all).all() and np.linalg.norm(g_all) > 0
    l_perm, g_perm = step(tc.randperm(nb, device=dev(), generator=g))
    np.testing.assert_allclose(l_perm, l_all, rtol=1e-4, atol=1e-6)
    assert np.linalg.norm(g_perm - g_all) <= 2e-4 * np.linalg.norm(g_all)
    l_a, g_a = step(ident[:nb // 2])
    l_b, g_b = step(ident[nb // 2:])
    np.testing.assert_allclose(0.5 * (l_a + l_b)[:4], l_all[:4], rtol=1e-4, atol=1e-6)
    assert np.linalg.norm(0.5 * (g_a + g_b) - g_all) <= 2e-4 * np.linalg.norm(g_all)
    sub = ident[1000:1000 + 4096]
    l_sub, g_sub = step(sub)
    names = eng.names
    offs = eng.offsets
    del eng
    ref, _ = _build('fp32', A, rk, 4096)
    l_ref = ref.ppo_step(obs, sub, batch, hyper).cpu().numpy()
    g_ref = ref.grads.cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(l_sub[:4], l_ref[:4], rtol=5 * TOL['bf16']['loss'], atol=5 * TOL['bf16']['loss'])
    assert g_sub @ g_ref / (np.linalg.norm(g_sub) * np.linalg.norm(g_ref)) > 0.995
    # rel-L2 per parameter tensor at 4 096 samples (measured 0.5-7 %: with random-sign advantages the minibatch gradient is
    # a sum of cancelling per-sample terms, so the ~0.1 % of ReLU masks that differ between bf16 and fp32 activations
    # show as sqrt(2 f) of the norm and do NOT average out with the batch size — DESIGN.md §4 "Precision modes")
    per = {n: rel_l2(g_sub[offs[i]:offs[i + 1]], g_ref[offs[i]:offs[i + 1]]) for i, n in enumerate(names)}
    print('bf16 vs fp32 engine, gradient rel-L2 per tensor at 4096 samples:', {k.split('._network')[0][-24:] + k[-9:]: round(v, 4) for k, v in per.items()})
    assert rel_l2(g_sub, g_ref) < 2e-2, rel_l2(g_sub, g_ref)
    assert max(per.values()) < 0.12, per


# ---- differentiable Agent.forward: custom torch losses on top of the fused network ------------------
@pytest.mark.parametrize('precision', _precisions())
def test_agent_forward_is_differentiable_like_the_reference(precision):
    """loss.backward() through Agent.forward (Categorical + value heads) must give the gradients torch autograd gives
    on the oracle model for the same custom loss, accumulate like autograd, and survive zero_grad(set_to_none=True)."""
    from pytorch_drl_b200 import agents
    A, rk, nb = 6, ['extrinsic', 'intrinsic_rnd'], 24
    params = po.init_params(A, rk, seed=21)
    preds = {'policy': agents.CategoricalPolicyHead(512, A)}
    for k in rk:
        preds[f'value_{k}'] = agents.SimpleValueHead(512)
    agent = agents.Agent([agents.ToChannelMajor()], agents.NatureCNN(4), preds)
    agent.load_state_dict({k: tc.from_numpy(v) for k, v in params.items()})
    agent.fuse(max_batch=16, precision=precision)                       # smaller than the batch: chunked backward
    rs = np.random.RandomState(5)
    obs_u8 = rs.randint(0, 256, size=(nb, 84, 84, 4)).astype(np.uint8)
    actions = rs.randint(0, A, size=nb)
    target = rs.standard_normal(nb).astype(np.float32)

    def custom_loss(pi_logits, v_ext, v_int):
        dist = tc.distributions.Categorical(logits=pi_logits)
        a = tc.as_tensor(actions, device=pi_logits.device)
        tgt = tc.as_tensor(target, device=pi_logits.device)
        return -(dist.log_prob(a) * tgt).mean() - 0.03 * dist.entropy().mean() + ((v_ext - tgt) ** 2).mean() + v_int.abs().mean()

    # oracle: plain torch autograd on the reference-layout parameters
    p_ref = {k: tc.from_numpy(v.copy()).requires_grad_(True) for k, v in params.items()}
    feat = po.trunk_forward(p_ref, po.scale_obs(tc.from_numpy(obs_u8)))
    lg, vals = po.heads_forward(p_ref, feat, rk)
    custom_loss(lg, vals['extrinsic'], vals['intrinsic_rnd']).backward()
    g_ref = {k: v.grad.numpy() for k, v in p_ref.items()}

    def run():
        out = agent(tc.from_numpy(obs_u8).to(dev()), ['policy', 'value_extrinsic', 'value_intrinsic_rnd'])
        loss = custom_loss(out['policy'].logits, out['value_extrinsic'], out['value_intrinsic_rnd'])
        loss.backward()
        return float(loss.detach())

    opt = tc.optim.SGD(agent.parameters(), lr=0.0)
    opt.zero_grad(set_to_none=True)
    run()
    t = TOL[precision]
    got = {k: p.grad.cpu().numpy().copy() for k, p in agent.named_parameters()}
    for k in g_ref:
        assert rel_l2(got[k], g_ref[k]) < t['grad'], (k, rel_l2(got[k], g_ref[k]))
    run()                                                               # second backward without zero_grad: accumulates
    for k, p in agent.named_parameters():
        assert rel_l2(p.grad.cpu().numpy(), 2.0 * got[k]) < 1e-4, k
    opt.zero_grad(set_to_none=True)
    run()
    for k, p in agent.named_parameters():
        assert rel_l2(p.grad.cpu().numpy(), got[k]) < 1e-4, k
    with tc.no_grad():                                                  # no graph, no backward kernels
        out = agent(tc.from_numpy(obs_u8).to(dev()), ['policy'])
        assert not out['policy'].logits.requires_grad


# ---- extra_steps > 0 in the PPO rollout path (abstract.py:413-445 with n = extra_steps + 1) ------------------------
@pytest.mark.parametrize('op_name', ['GAE', 'NStepAdvantageEstimator'])
def test_ppo_credit_assignment_with_extra_steps(op_name):
    from pytorch_drl_b200 import agents, algos
    from pytorch_drl_b200.algos.credit_assignment import NStepAdvantageEstimator
    from oracle import credit_oracle as co
    A, rk, n_envs, T, extra = 6, ['extrinsic'], 3, 16, 2
    n = extra + 1
    agent = agents.Agent([agents.ToChannelMajor()], agents.NatureCNN(4),
                         {'policy': agents.CategoricalPolicyHead(512, A), 'value_extrinsic': agents.SimpleValueHead(512)})
    agent.load_state_dict({k: tc.from_numpy(v) for k, v in po.init_params(A, rk, 3).items()})
    agent.fuse(max_batch=n_envs * (T + n), precision='fp32')
    op = algos.GAE(gamma=0.99, use_dones=True, lambda_=0.9) if op_name == 'GAE' else NStepAdvantageEstimator(gamma=0.99, use_dones=True)
    algo = algos.PPO(rank=0, world_size=1, rollout_len=T, extra_steps=extra, credit_assignment_ops={'extrinsic': op},
                     policy_net=agent, policy_optimizer=tc.optim.Adam(agent.parameters(), lr=1e-3), standardize_adv=True,
                     opt_epochs=1, learner_batch_size=8, envs_per_rank=n_envs, sampler_seed=0)
    st = algos.RolloutStorage(n_envs, T, rk, dev(), extra_steps=extra)
    assert st.Tp >= T + n
    rs = np.random.RandomState(0)
    obs = rs.randint(0, 256, size=(n_envs, T + n, 84, 84, 4)).astype(np.uint8)
    act = rs.randint(0, A, size=(n_envs, T + n))
    rew = (rs.randint(-1, 2, size=(n_envs, T + n - 1)) + rs.standard_normal((n_envs, T + n - 1)) * 0.1).astype(np.float32)
    don = (rs.uniform(size=(n_envs, T + n - 1)) < 0.1).astype(np.float32)
    for e in range(n_envs):
        st.load_env(e, obs[e], act[e], {'extrinsic': rew[e]}, don[e])
    rollout = algo.credit_assignment(algo.annotate(st.as_dict()))
    v = rollout['vpreds']['extrinsic'].cpu().numpy()
    # the annotate pass covered all T + n frames: value predictions equal a direct forward of those frames
    _, vals, _, _ = agent.engine.forward(tc.from_numpy(obs.reshape(-1, 84, 84, 4)).to(dev()))
    np.testing.assert_allclose(v[:, :T + n], vals.cpu().numpy().reshape(n_envs, T + n), rtol=1e-5, atol=1e-6)
    for e in range(n_envs):
        if op_name == 'GAE':
            adv = po.gae(T, extra, rew[e], v[e, :T + n], don[e], 0.99, 0.9, True)
        else:
            adv = co.nstep_advantages(T, extra, rew[e], v[e, :T + n], don[e], 0.99, True)
        np.testing.assert_allclose(rollout['vtargs']['extrinsic'][e, :T].cpu().numpy(), adv + v[e, :T], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(rollout['advantages']['extrinsic'][e, :T].cpu().numpy(), po.standardize(adv), rtol=2e-4, atol=2e-5)


# ---- DQN head blocks ---------------------------------------------------------------------
def test_dueling_head_and_epsilon_greedy_vs_reference_golden(golden_dir):
    """The drop-in modules load the reference's state_dict unchanged; Q-values match the reference to float32
    rounding, the epsilon-greedy probabilities bit for bit against the oracle's float32 expression."""
    from oracle import dueling_oracle as do
    from pytorch_drl_b200 import _lib, agents
    from pytorch_drl_b200._lib import current_stream, ptr
    g = np.load(os.path.join(golden_dir, 'dueling.npz'))
    for i in range(int(g['ncases'])):
        A, nb, eps = int(g[f'c{i}_meta'][0]), int(g[f'c{i}_meta'][1]), float(g[f'c{i}_meta'][2])
        head = agents.SimpleDiscreteActionValueHead(num_features=512, num_actions=A, head_architecture_cls=agents.DuelingArchitecture,
                                                    head_architecture_cls_args={'widening': 1})
        head._action_value_head.load_state_dict({k: tc.from_numpy(g[f'c{i}_p_{k}']) for k in do.PARAM_KEYS}, strict=True)
        head.to(dev())
        feats = cu(g[f'c{i}_feat'])
        q = head(feats)
        np.testing.assert_allclose(q.cpu().numpy(), g[f'c{i}_q'], rtol=1e-5, atol=2e-6)
        pol = agents.EpsilonGreedyCategoricalPolicyHead(action_value_head=head)
        pol.epsilon = eps
        dist = pol(feats)
        np.testing.assert_allclose(dist.probs.cpu().numpy(), g[f'c{i}_probs'], rtol=0, atol=1.2e-7)
        # raw kernel on the REFERENCE's q: bit-identical to the float32 expression, argmax = first maximum
        qref = cu(g[f'c{i}_q'])
        probs, greedy = tc.empty_like(qref), tc.empty(nb, dtype=tc.int64, device=dev())
        assert _lib.lib().drl_epsilon_greedy_probs_f32(ptr(qref), nb, A, eps, ptr(probs), ptr(greedy), current_stream()) == 0
        np.testing.assert_array_equal(probs.cpu().numpy(), do.epsilon_greedy_probs(g[f'c{i}_q'], eps))
        np.testing.assert_array_equal(greedy.cpu().numpy(), np.argmax(g[f'c{i}_q'], axis=-1))
    # ties, larger batch than CTAs x warps, error behaviour
    qt = cu(np.array([[1., 3., 3., 0.], [2., 2., 2., 2.]], np.float32))
    probs, greedy = tc.empty_like(qt), tc.empty(2, dtype=tc.int64, device=dev())
    assert _lib.lib().drl_epsilon_greedy_probs_f32(ptr(qt), 2, 4, 0.2, ptr(probs), ptr(greedy), current_stream()) == 0
    assert greedy.cpu().tolist() == [1, 0]
    assert _lib.lib().drl_epsilon_greedy_probs_f32(ptr(qt), 2, 4, 1.5, ptr(probs), ptr(greedy), current_stream()) != 0
    big = tc.relu(tc.randn(5000, 512, device=dev()))
    params = {k: g[f'c1_p_{k}'] for k in do.PARAM_KEYS}
    head = agents.SimpleDiscreteActionValueHead(512, 18, agents.DuelingArchitecture, {'widening': 1})
    head._action_value_head.load_state_dict({k: tc.from_numpy(v) for k, v in params.items()})
    head.to(dev())
    np.testing.assert_allclose(head(big).cpu().numpy(), do.dueling_forward(params, big.cpu().numpy()), rtol=1e-5, atol=3e-6)
    with pytest.raises(RuntimeError):
        head(tc.zeros(2, 512))                                   # CPU features: no fallback
    with pytest.raises(NotImplementedError):
        agents.SimpleDiscreteActionValueHead(512, 6, agents.Linear, {})


# ---- the reference's own known-answer tests for the agent stack, run through the fused path -------------------
@pytest.mark.parametrize('precision', _precisions())
def test_reference_known_answers_zero_weights(precision):
    """test_nature_cnn.py:7-22 (all-zero weights and biases: zero features), integration/test_agent.py:10-57 and
    test_launch.py:310-318 (zero-initialised heads: value 0, uniform policy, log_prob = log(1/A)) and
    heads/test_policy_heads.py:15-27 — on frames that are NOT zero, so any leak of the input would show."""
    from pytorch_drl_b200 import agents
    A, nb = 6, 32
    z = tc.nn.init.zeros_
    ag = agents.Agent([agents.ToChannelMajor()], agents.NatureCNN(4, z, z),
                      {'policy': agents.CategoricalPolicyHead(512, A, w_init=z, b_init=z),
                       'value_extrinsic': agents.SimpleValueHead(512, w_init=z, b_init=z)})
    ag.fuse(max_batch=nb, precision=precision, device=dev())
    g = tc.Generator(device=dev())
    g.manual_seed(5)
    obs = tc.randint(0, 256, (nb, 84, 84, 4), device=dev(), dtype=tc.uint8, generator=g)
    with tc.no_grad():
        pred = ag(obs, ['policy', 'value_extrinsic'])
    tc.testing.assert_close(pred['value_extrinsic'], tc.zeros(nb, device=dev()), rtol=1e-4, atol=1e-4)
    acts = tc.zeros(nb, dtype=tc.int64, device=dev())
    tc.testing.assert_close(pred['policy'].log_prob(acts), tc.log(tc.full((nb,), 1.0 / A, device=dev())), rtol=1e-4, atol=1e-4)
    tc.testing.assert_close(pred['policy'].entropy(), tc.full((nb,), float(np.log(A)), device=dev()), rtol=1e-4, atol=1e-4)
    # zero trunk, non-zero head bias: features are exactly zero, so logits == bias for every sample
    with tc.no_grad():
        dict(ag.named_parameters())['_predictors.policy._policy_head._network.0.bias'].copy_(tc.arange(A, dtype=tc.float32) * 0.25)
    ag.engine.sync_params()
    with tc.no_grad():
        pred = ag(obs, ['policy'])
    ref = tc.log_softmax(tc.arange(A, dtype=tc.float32, device=dev()) * 0.25, 0)
    tc.testing.assert_close(pred['policy'].logits, ref.expand(nb, A), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('precision', _precisions())
def test_qagent_trunk_features_and_dueling_q_vs_oracle(precision, golden_dir):
    """DQN acting / target evaluation through the fused trunk: QAgent = NatureCNN on the engine (drl_net_features) +
    dueling action-value head + epsilon-greedy policy, against the oracle trunk (nature_cnn.py:27-36) followed by the
    oracle dueling head with the reference's own head weights (golden case 0)."""
    from oracle import dueling_oracle as do
    from pytorch_drl_b200 import agents
    A, nb = 6, 45
    g = np.load(os.path.join(golden_dir, 'dueling.npz'))
    hp = {k: g[f'c0_p_{k}'] for k in do.PARAM_KEYS}
    trunk = po.init_params(A, ['extrinsic'], seed=21)
    head = agents.SimpleDiscreteActionValueHead(512, A, agents.DuelingArchitecture, {'widening': 1})
    head._action_value_head.load_state_dict({k: tc.from_numpy(v) for k, v in hp.items()})
    qa = agents.QAgent([agents.ToChannelMajor()], agents.NatureCNN(4), {'action_value': head})
    qa.load_state_dict({k: tc.from_numpy(trunk[k]) for k in po.TRUNK_KEYS}, strict=False)
    qa.fuse(max_batch=32, precision=precision, device=dev())          # 45 samples > max_batch: the engine chunks
    rs = np.random.RandomState(2)
    obs = rs.randint(0, 256, size=(nb, 84, 84, 4), dtype=np.uint8)
    feat_ref = po.trunk_forward({k: tc.from_numpy(trunk[k]) for k in po.TRUNK_KEYS}, po.scale_obs(tc.from_numpy(obs))).numpy()
    q_ref = do.dueling_forward(hp, feat_ref)
    feats = qa.engine.features(cu(obs))
    t = TOL[precision]
    assert rel_l2(feats.cpu().numpy(), feat_ref) < t['act']
    pol = qa._predictors['policy']
    pol.epsilon = 0.1
    out = qa(cu(obs), ['action_value', 'policy'])
    q = out['action_value'].cpu().numpy()
    assert q.shape == (nb, A) and rel_l2(q, q_ref) < 2 * t['act']
    probs = out['policy'].probs.cpu().numpy()
    np.testing.assert_allclose(probs.sum(-1), 1.0, rtol=1e-6)
    np.testing.assert_array_equal(probs.argmax(-1), q.argmax(-1))         # greedy action = argmax of the SAME q
    np.testing.assert_allclose(np.sort(probs, -1)[:, -1], 0.9 + 0.1 / A, rtol=1e-5)
    # float32 frames pre-scaled by 1/255 (the reference's rollout format) map back to the same bytes
    out2 = qa(po.scale_obs(tc.from_numpy(obs)).to(dev()), ['action_value'])
    np.testing.assert_array_equal(out2['action_value'].cpu().numpy(), q)


# ---- DQN learner step around the replay buffer (BASELINE config "Double/Dueling DQN with prioritized replay"; §8 f4) ----
def _dqn_setup(precision, A, nb, capacity, seeds=(40, 70), **kw):
    from oracle import dqn_oracle as do
    from pytorch_drl_b200 import agents
    from pytorch_drl_b200.algos.dqn import DQN
    from pytorch_drl_b200.replay import PrioritizedReplayBuffer
    params = [do.init_params(A, s) for s in seeds]
    qas = []
    for p in params:
        head = agents.SimpleDiscreteActionValueHead(512, A, agents.DuelingArchitecture, {'widening': 1})
        qa = agents.QAgent([agents.ToChannelMajor()], agents.NatureCNN(4), {'action_value_extrinsic': head})
        qa.load_state_dict({k: tc.from_numpy(v) for k, v in p.items()}, strict=False)
        qa.fuse(max_batch=nb, precision=precision, device=dev())
        qas.append(qa)
    replay = PrioritizedReplayBuffer(capacity, 0.6, 1e-6, device=dev())
    dqn = DQN(qas[0], qas[1], replay, batch_size=nb, **kw)
    # the constructor copies online -> target; the parity cases want a DIFFERENT target network
    dqn.target.engine.load_named({**{k: tc.zeros_like(v) for k, v in dqn.target.engine.named_views().items()},
                                  **{k: tc.from_numpy(params[1][k]) for k in po.TRUNK_KEYS}})
    for k, v in dqn.target_head.views().items():
        v.copy_(tc.from_numpy(params[1][do.HEAD_PREFIX + k]))
    return dqn, params


@pytest.mark.parametrize('precision', _precisions())
@pytest.mark.parametrize('A,gamma,use_dones,double_q,loss', [(6, 0.99, True, True, 'SmoothL1Loss'), (18, 0.9, False, False, 'MSELoss')])
def test_dqn_loss_and_grads_vs_oracle(precision, A, gamma, use_dones, double_q, loss):
    """Q(s, .), double-Q Bellman targets, importance-weighted TD loss and its gradient w.r.t. EVERY parameter (dueling head through
    drl_dueling_loss_backward_f32, trunk through drl_net_backward_features) against oracle/dqn_oracle.py, which is pinned to the
    reference's own modules + torch autograd (tests/golden/dqn.npz).  fp32 engine: tight; tcgen05 engine: the bf16 bars of TOL."""
    from oracle import dqn_oracle as do
    nb = 24
    dqn, (p_on, p_tg) = _dqn_setup(precision, A, nb, 64, gamma=gamma, use_dones=use_dones, double_q=double_q, loss=loss)
    obs, nxt, actions, rewards, dones, weights = do.make_case(11, nb, A)
    frames = np.concatenate([obs, nxt])                                    # ring rows: s_i = i, s'_i = nb + i
    rows, nrows = np.arange(nb, dtype=np.int64), np.arange(nb, 2 * nb, dtype=np.int64)
    dqn.loss_and_grads(cu(frames), cu(rows), cu(nrows), cu(actions), cu(rewards), cu(dones), cu(weights))
    t = TOL[precision]
    qn_on = dqn.q_values(dqn.online, dqn.head, cu(frames), cu(nrows)).cpu().numpy()
    qn_tg = dqn.q_values(dqn.target, dqn.target_head, cu(frames), cu(nrows)).cpu().numpy()
    assert rel_l2(qn_on, do.q_values(p_on, nxt)) < 2 * t['act'] and rel_l2(qn_tg, do.q_values(p_tg, nxt)) < 2 * t['act']
    y_dev = dqn.y.cpu().numpy()
    # the targets are EXACTLY the reference operator applied to the device's own Q(s', .)
    np.testing.assert_array_equal(y_dev, do.bellman_targets(rewards, dones, qn_on, qn_tg, gamma, use_dones, double_q))
    res, grads = do.loss_and_grads(p_on, p_tg, obs, nxt, actions, rewards, dones, weights, gamma, use_dones, double_q, loss,
                                   y_override=None if precision == 'fp32' else y_dev)
    assert rel_l2(dqn.q.cpu().numpy(), res['q']) < 2 * t['act']
    np.testing.assert_allclose(y_dev, res['y'], rtol=0, atol=2 * t['act'] * max(1.0, float(np.abs(res['y']).max())))
    np.testing.assert_allclose(dqn.td.cpu().numpy(), res['td'], rtol=0, atol=2 * t['act'] * max(1.0, float(np.abs(res['q']).max())))
    assert abs(float(dqn.loss.item()) - res['loss']) < max(t['loss'], 20 * t['act']) * max(1.0, abs(res['loss']))
    got = {**{k: v for k, v in dqn.online.engine.named_views(dqn.online.engine.grads).items() if k in po.TRUNK_KEYS},
           **{do.HEAD_PREFIX + k: v for k, v in dqn.head.views(dqn.head_grads).items()}}

    def check(gref_all, bar, min_cos):
        for k, gref in gref_all.items():
            gdev = got[k].cpu().numpy().astype(np.float64)
            if np.linalg.norm(gref) == 0:
                assert np.linalg.norm(gdev) == 0, k
                continue
            err = rel_l2(gdev, gref)
            cos = float((gdev * gref).sum() / (np.linalg.norm(gdev) * np.linalg.norm(gref)))
            worst[0] = max(worst[0], err / bar)
            assert err < bar and cos > min_cos, (k, err, cos)
    worst = [0.0]
    if precision == 'fp32':
        check(grads, t['grad'], 0.999999)                                   # the whole chain against the oracle
    else:
        # bf16 trunk: a few ReLU units of the 50 / 25-wide head flip with the rounded features and move a 24-sample head gradient
        # by tens of percent, so the chain is checked link by link: the fp32 head kernels against the oracle AT THE DEVICE'S
        # FEATURES (tight), then the bf16 trunk against the oracle's VJP of the device's d(features) (the bf16 bars of TOL).
        feat_dev = dqn.online.engine.features(cu(frames), cu(rows)).cpu().numpy()
        hres, hgrads, dfeat_ref = do.head_loss_and_grads(p_on, feat_dev, y_dev, actions, weights, loss)
        assert rel_l2(dqn.q.cpu().numpy(), hres['q']) < 1e-5 and abs(float(dqn.loss.item()) - hres['loss']) < 1e-5 * max(1.0, hres['loss'])
        check(hgrads, 2e-4, 0.999999)
        assert rel_l2(dqn.dfeat.cpu().numpy(), dfeat_ref) < 2e-4
        check(do.trunk_vjp(p_on, obs, dqn.dfeat.cpu().numpy()), 0.12, 0.99)
    print(f'dqn gradients [{precision}, A={A}]: worst error / bar = {worst[0]:.3f}')
    # the engine's own (unused) policy / value heads receive no gradient
    for k, v in dqn.online.engine.named_views(dqn.online.engine.grads).items():
        if k not in po.TRUNK_KEYS:
            assert float(v.abs().max()) == 0.0, k


def test_dqn_learner_step_updates_network_priorities_and_target():
    """One whole learner step: prioritized sample of 32 of 256 stored transitions, update, |td| priorities back into the sum-tree
    (bit-exact vs the buffer's own update path), Adam on trunk + head against torch.optim.Adam semantics, target sync."""
    from oracle import dqn_oracle as do
    A, nb, cap = 6, 32, 256
    dqn, _ = _dqn_setup('fp32', A, nb, cap, gamma=0.99, lr=1e-3, target_update_interval=3, beta=0.5)
    dqn.sync_target()
    rs = np.random.RandomState(3)
    obs, _, actions, rewards, dones, _ = do.make_case(12, cap, A)
    dqn.replay.add(cu(obs), cu(actions), cu(rewards), cu(dones))
    dqn.replay.update_priorities(cu(np.arange(cap, dtype=np.int64)), cu(rs.uniform(0.1, 2.0, cap).astype(F32)))
    uni = cu(rs.uniform(size=nb).astype(F32))
    p0 = dqn.online.engine.params.clone()
    h0 = dqn.head.flat.clone()
    tree0 = dqn.replay.index.tree.tree.clone()
    expect = dqn.replay.sample(nb, 0.5, uni)                                # same uniform numbers -> the same batch
    loss = dqn.learner_step(uni)
    tc.cuda.synchronize()
    assert np.isfinite(float(loss.item())) and dqn.steps == 1 and dqn.online.engine.adam_steps == 1
    # first Adam step: every parameter with a non-zero gradient moves by lr * sign(g) (torch.optim.Adam, bias-corrected)
    for params, before, grads in ((dqn.online.engine.params, p0, dqn.online.engine.grads), (dqn.head.flat, h0, dqn.head_grads)):
        g = grads.cpu().numpy()
        step = (params - before).cpu().numpy()
        nzm = np.abs(g) > 1e-6
        assert nzm.sum() > 1000 if params is dqn.online.engine.params else nzm.sum() > 100
        np.testing.assert_allclose(step[nzm], -1e-3 * np.sign(g[nzm]), rtol=2e-2)
        assert np.all(step[g == 0] == 0)
    # priorities of the sampled transitions = (|td| + eps)^alpha through the buffer's own kernel; everything else untouched
    idx = expect['idx']
    want = tree0.clone()
    from pytorch_drl_b200.replay import SumTree
    ref_tree = SumTree(cap, dev())
    ref_tree.tree.copy_(want)
    ref_tree.update(idx, (dqn.td.abs() + 1e-6).pow(0.6).float())
    assert tc.equal(dqn.replay.index.tree.tree, ref_tree.tree)
    # the target network follows every 3rd step
    assert not tc.equal(dqn.target.engine.params, dqn.online.engine.params)
    dqn.learner_step()
    dqn.learner_step()
    assert dqn.steps == 3 and tc.equal(dqn.target.engine.params, dqn.online.engine.params) and tc.equal(dqn.target_head.flat, dqn.head.flat)
    # module parameters ARE the flat buffers (state_dict sees the update)
    sd = dqn.online.state_dict()
    assert tc.equal(sd[do.HEAD_PREFIX + '_proj.0.bias'], dqn.head.views()['_proj.0.bias'])


# ---- checkpoint / resume: complete optimizer + RND state in the reference's formats --------------------
def _ppo_for_resume(precision, seed=0):
    from pytorch_drl_b200 import agents, algos
    A, rk, T, B, n_envs = 6, ['extrinsic'], 16, 8, 2
    agent = agents.Agent([agents.ToChannelMajor()], agents.NatureCNN(4),
                         {'policy': agents.CategoricalPolicyHead(512, A), 'value_extrinsic': agents.SimpleValueHead(512)})
    agent.load_state_dict({k: tc.from_numpy(v) for k, v in po.init_params(A, rk, 11).items()})
    agent.fuse(max_batch=n_envs * (T + 1), precision=precision)
    opt = tc.optim.Adam(agent.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-5)
    sched = tc.optim.lr_scheduler.LambdaLR(opt, lambda e: 1.0 / (1 + e))

    def make_algo(global_step=1):
        return algos.PPO(rank=0, world_size=1, rollout_len=T, extra_steps=0,
                         credit_assignment_ops={'extrinsic': algos.GAE(gamma=0.99, use_dones=True, lambda_=0.95)},
                         global_step=global_step, max_steps=10 ** 9, policy_net=agents.DDPShim(agent), policy_optimizer=opt,
                         policy_scheduler=sched, opt_epochs=2, learner_batch_size=B, envs_per_rank=n_envs, sampler_seed=seed)

    def rollout(i):
        st = algos.RolloutStorage(n_envs, T, rk, dev())
        for r in range(n_envs):
            raw = po.make_synthetic_env_rollout(r, 100 + i, T, A, rk)
            st.load_env(r, raw['observations'], raw['actions'], raw['rewards'], raw['dones'])
        return st.as_dict()
    return agent, opt, sched, make_algo, rollout


def test_checkpoint_resume_continues_the_trajectory(tmp_path):
    """Adam moments / step count live in the engine's flat buffers but are exposed as the torch optimizer's `state`, so the
    reference's persist / resume protocol (ppo.py:295-307, launch.py:377-383: load_state_dict on nets + optimizers BEFORE the
    algorithm is built) carries the complete state: a resumed run continues like the uninterrupted one.  Also: the policy net
    is saved under DDP's `module.` keys, and `optimizer.load_state_dict` in the middle of a run is adopted."""
    from pytorch_drl_b200.utils import checkpoint as ck
    agent, opt, sched, make_algo, rollout = _ppo_for_resume('fp32')
    algo = make_algo()
    algo.optimize(algo.credit_assignment(algo.annotate(rollout(0))))
    import copy
    sd = copy.deepcopy(opt.state_dict())          # state_dict() hands out the live (aliasing) tensors, like torch's own
    assert len(sd['state']) == 12 and float(sd['state'][0]['step']) == 4.0          # 2 epochs x 16 / 8 steps
    eng = agent.engine
    assert tc.equal(sd['state'][6]['exp_avg'].reshape(-1), eng.exp_avg[eng.offsets[6]:eng.offsets[7]])
    ckdir = str(tmp_path / 'ck')
    ck.save_checkpoints(ckdir, algo.checkpointables, steps=32)
    assert sorted(os.listdir(ckdir)) == ['policy_net_32.pth', 'policy_optimizer_32.pth', 'policy_scheduler_32.pth']
    assert all(k.startswith('module.') for k in tc.load(os.path.join(ckdir, 'policy_net_32.pth')))
    # uninterrupted continuation
    algo.optimize(algo.credit_assignment(algo.annotate(rollout(1))))
    want = {k: v.detach().clone() for k, v in agent.state_dict().items()}
    want_lr = opt.param_groups[0]['lr']
    # resumed: fresh objects, checkpoint loaded before the algorithm is built (the reference's order)
    agent2, opt2, sched2, make_algo2, _ = _ppo_for_resume('fp32')
    from pytorch_drl_b200 import agents
    loaded = ck.maybe_load_checkpoints(ckdir, {'policy_net': agents.DDPShim(agent2), 'policy_optimizer': opt2,
                                               'policy_scheduler': sched2}, map_location='cpu')
    assert loaded == 32
    algo2 = make_algo2(global_step=1)
    assert agent2.engine.adam_steps == 4
    # same permutation stream as the uninterrupted run's second optimize: replay the first call's draws
    for s in algo2._samplers:
        for _ in range(2):
            s.resample()
    algo2.optimize(algo2.credit_assignment(algo2.annotate(rollout(1))))
    assert abs(opt2.param_groups[0]['lr'] - want_lr) < 1e-12
    for k, v in agent2.state_dict().items():
        assert rel_l2(v.cpu().numpy(), want[k].cpu().numpy()) < 1e-5, k
    # mid-run load_state_dict on the live optimizer: the engine adopts it at its next step
    opt2.load_state_dict(sd)
    algo2.learner_step(__import__('pytorch_drl_b200').utils.MinibatchView(
        algo2.credit_assignment(algo2.annotate(rollout(2))), tc.arange(16, device=dev(), dtype=tc.int64)))
    assert agent2.engine.adam_steps == 5 and float(opt2.state_dict()['state'][0]['step']) == 5.0


def test_rnd_checkpointables_restore_the_predictor(tmp_path):
    """random_network_distillation.py:175-184: normaliser, teacher, student and optimizer round-trip through files; the restored
    wrapper gives the same intrinsic reward and takes the same next predictor step."""
    from pytorch_drl_b200.envs import RandomNetworkDistillation
    from pytorch_drl_b200.utils import checkpoint as ck
    g = tc.Generator(device=dev())
    g.manual_seed(1)
    frames = tc.randint(0, 256, (24, 84, 84, 4), device=dev(), dtype=tc.uint8, generator=g)
    rows = tc.arange(24, device=dev(), dtype=tc.int64)
    a = RandomNetworkDistillation({'lr': 1e-3, 'eps': 1e-5}, world_size=32, max_batch=24, seed=3)
    a.observe(tc.rand((5, 84, 84, 1), device=dev(), generator=g))
    a._sync_normalizers_global(); a._sync_normalizers_local()
    for _ in range(3):
        a.learn_rows(frames, rows)
    d = str(tmp_path / 'rnd')
    ck.save_checkpoints(d, a.checkpointables, steps=7)
    assert sorted(os.listdir(d)) == ['rnd_observation_normalizer_7.pth', 'rnd_optimizer_7.pth', 'rnd_student_net_7.pth', 'rnd_teacher_net_7.pth']
    assert list(tc.load(os.path.join(d, 'rnd_student_net_7.pth')).keys())[0] == 'module._network.0._network.1.weight'
    b = RandomNetworkDistillation({'lr': 1e-3, 'eps': 1e-5}, world_size=32, max_batch=24, seed=99)
    assert ck.maybe_load_checkpoints(d, b.checkpointables) == 7
    b._sync_normalizers_local()
    assert b.adam_steps == 3 and b._synced_normalizer.steps == 5.0
    np.testing.assert_allclose(b.intrinsic_reward(frames).cpu().numpy(), a.intrinsic_reward(frames).cpu().numpy(), rtol=1e-6)
    la, lb = float(a.learn_rows(frames, rows).item()), float(b.learn_rows(frames, rows).item())
    assert abs(la - lb) <= 1e-6 * abs(la)
    assert rel_l2(b.student_params.cpu().numpy(), a.student_params.cpu().numpy()) < 1e-6


# ---- prioritized replay (BASELINE config 4; parity unpinned: the reference has no sum-tree / replay) ------------
@pytest.mark.parametrize('cap,n', [(1 << 20, 2049), (1 << 20, 5000), (1 << 20, 70000), (1 << 12, 9000), (256, 3000)])
def test_sumtree_large_batches_multi_cta_bit_exact(cap, n):
    """Batches beyond one CTA's 2048 slots: 128 CTAs (one per depth-7 subtree) + the 7-level top pass must leave exactly the
    tree sequential in-order writes leave (oracle/sumtree.c), duplicates (highest j wins), out-of-range indices ignored."""
    from pytorch_drl_b200.replay import SumTree
    rs = np.random.RandomState(cap % 97 + n)
    t, ref = SumTree(cap), so.SumTree(cap)
    base = rs.uniform(0.0, 1.0, size=cap).astype(F32)
    t.set_leaves(cu(base)); ref.tree[cap:] = base; ref.rebuild()
    idx = rs.randint(0, cap, size=n).astype(np.int64)
    idx[rs.randint(0, n, size=n // 4)] = idx[rs.randint(0, n, size=n // 4)]        # plenty of duplicates
    idx[::17] = rs.randint(0, min(cap, 300), size=idx[::17].size)                    # hot leaves: many writers each
    pr = rs.uniform(0.01, 5.0, size=n).astype(F32)
    bad = idx.copy()
    bad[5], bad[n // 2] = -3, cap + 7                                                # ignored by the kernel
    keep = (bad >= 0) & (bad < cap)
    ref.update(bad[keep], pr[keep])
    t.update(cu(bad), cu(pr))
    assert np.array_equal(t.tree.cpu().numpy().view(np.uint32), ref.tree.view(np.uint32))
    with pytest.raises(ValueError):
        t.update(cu(bad), cu(pr), validate=True)


def test_prioritized_replay_buffer_vs_oracle():
    """Ring writes with wrap-around, max-priority insertion, the fused stratified sample + importance weights kernel against
    oracle/sumtree.c (indices / priorities bit-exact, weights through powf: rtol 2e-6), gathered transitions = ring rows,
    priority updates, empty-tree / zero-priority guards."""
    from pytorch_drl_b200.replay import PrioritizedReplayBuffer, PrioritizedReplayIndex
    cap, alpha, eps, beta = 1 << 10, 0.6, 1e-6, 0.4
    rs = np.random.RandomState(2)
    buf = PrioritizedReplayBuffer(cap, alpha, eps, frame_shape=(8, 8, 4))
    ref = so.SumTree(cap)
    frames_ref = np.zeros((cap, 8, 8, 4), np.uint8)
    act_ref, rew_ref = np.zeros(cap, np.int64), np.zeros(cap, F32)
    # empty buffer: weights 0, no NaN
    e = buf.sample(16, beta, cu(np.full(16, 0.5, F32)))
    assert tc.isfinite(e['weights']).all() and float(e['weights'].abs().sum()) == 0.0
    cursor, open_slot = 0, None
    for n in (700, 500):                                           # the second add wraps around
        fr = rs.randint(0, 256, size=(n, 8, 8, 4)).astype(np.uint8)
        ac, rw, dn = rs.randint(0, 6, size=n).astype(np.int64), rs.standard_normal(n).astype(F32), (rs.uniform(size=n) < 0.1).astype(F32)
        slots = buf.add(cu(fr), cu(ac), cu(rw), cu(dn)).cpu().numpy()
        want = (np.arange(n) + cursor) % cap
        assert np.array_equal(slots, want)
        frames_ref[want], act_ref[want], rew_ref[want] = fr, ac, rw
        # max priority for the new transitions and for the one that was waiting for its successor; the newest one stays at 0
        upd = want if open_slot is None else np.concatenate([[open_slot], want])
        pr = np.ones(len(upd), F32)
        pr[-1] = 0.0
        ref.update(upd, pr)
        cursor = (cursor + n) % cap
        open_slot = (cursor - 1) % cap
    assert buf.count == cap and buf.cursor == 1200 % cap
    assert float(buf.index.tree.leaves[open_slot]) == 0.0        # the newest transition waits for its successor: never sampled
    assert np.array_equal(buf.index.tree.tree.cpu().numpy().view(np.uint32), ref.tree.view(np.uint32))
    # TD-error priorities for a batch, then a sampled batch
    upd = rs.permutation(cap)[:300].astype(np.int64)
    td = rs.standard_normal(300).astype(F32)
    buf.update_priorities(cu(upd), cu(td))
    pr_dev = buf.index.tree.leaves[cu(upd)].cpu().numpy()                     # (|td| + eps)^alpha as the device computed it
    np.testing.assert_allclose(pr_dev, (np.abs(td.astype(np.float64)) + eps) ** alpha, rtol=2e-6)
    ref.update(upd, pr_dev)                                                   # same priorities -> same tree, bit for bit
    assert np.array_equal(buf.index.tree.tree.cpu().numpy().view(np.uint32), ref.tree.view(np.uint32))
    uni = rs.uniform(size=512).astype(F32)
    s = buf.sample(512, beta, cu(uni), gather=True)
    i_ref, p_ref, w_ref = ref.per_sample(uni, float(cap), beta)
    idx = s['idx'].cpu().numpy()
    assert np.array_equal(idx, i_ref) and np.array_equal(s['priorities'].cpu().numpy(), p_ref)
    np.testing.assert_allclose(s['weights'].cpu().numpy(), w_ref, rtol=2e-6)
    assert float(s['weights'].max()) == 1.0
    assert np.array_equal(s['observations'].cpu().numpy(), frames_ref[idx])
    assert np.array_equal(s['next_observations'].cpu().numpy(), frames_ref[(idx + 1) % cap])
    assert np.array_equal(s['actions'].cpu().numpy(), act_ref[idx]) and np.array_equal(s['rewards'].cpu().numpy(), rew_ref[idx])
    # the unfused path gives the same indices
    i2, p2 = buf.index.tree.sample(buf.index.tree.stratified_u(cu(uni)))
    assert np.array_equal(i2.cpu().numpy(), i_ref)
    # zero-priority leaves are never given an infinite weight
    ix = PrioritizedReplayIndex(64)
    ix.add(cu(np.arange(8, dtype=np.int64)))
    ix.tree.update(cu(np.array([3], np.int64)), cu(np.array([0.0], F32)))
    _, _, w = ix.sample(32, 1.0, cu(rs.uniform(size=32).astype(F32)))
    assert tc.isfinite(w).all()


# ---- actor side: batched policy sampling + the rollout ring ---------------------------------------------
@pytest.mark.parametrize('A', [6, 18])
def test_sample_actions_with_supplied_noise_equals_torch_multinomial_draw(A):
    """argmax(probs / q): bit-exact vs the oracle's restatement of torch.multinomial (index work); log-probabilities 1e-6."""
    from oracle import actor_oracle as ao
    from pytorch_drl_b200 import _lib
    L = _lib.lib()
    rs = np.random.RandomState(A)
    n = 5000
    logits = (rs.standard_normal((n, A)) * 2).astype(F32)
    noise = tc.empty(n, A).exponential_(1, generator=tc.Generator().manual_seed(3)).numpy()
    acts = tc.empty(n, dtype=tc.int64, device=dev())
    logp = tc.empty(n, dtype=tc.float32, device=dev())
    d_logits, d_noise = cu(logits), cu(noise)
    st = L.drl_sample_actions_f32(d_logits.data_ptr(), n, A, d_noise.data_ptr(), 0, 0, acts.data_ptr(), logp.data_ptr(), None)
    assert st == 0
    ref = ao.sample_actions(logits, noise)
    got = acts.cpu().numpy()
    # expf on the device vs the host differs in the last ulp: a draw can only differ where the two best scores tie to 1e-6
    p = tc.softmax(tc.from_numpy(logits), -1).numpy() / noise
    top2 = np.sort(p, axis=1)[:, -2:]
    near_tie = (top2[:, 1] - top2[:, 0]) <= 1e-5 * top2[:, 1]
    assert ((got == ref) | near_tie).all() and near_tie.sum() < 5
    ref_logp = tc.log_softmax(tc.from_numpy(logits), -1).numpy()[np.arange(n), got]
    np.testing.assert_allclose(logp.cpu().numpy(), ref_logp, rtol=1e-5, atol=2e-6)
    # argument errors like the rest of the ABI
    assert L.drl_sample_actions_f32(None, n, A, None, 0, 0, acts.data_ptr(), None, None) == 2
    assert L.drl_sample_actions_f32(d_logits.data_ptr(), n, 0, None, 0, 0, acts.data_ptr(), None, None) == 1


def test_sample_actions_in_kernel_philox_follows_the_policy_distribution():
    from pytorch_drl_b200 import _lib
    L = _lib.lib()
    A, n = 6, 400000
    row = np.array([0.3, -1.2, 2.0, 0.0, 1.1, -3.0], F32)
    logits = cu(np.tile(row, (n, 1)))
    p = np.exp(row - row.max()); p /= p.sum()

    def draw(seed, counter):
        acts = tc.empty(n, dtype=tc.int64, device=dev())
        assert L.drl_sample_actions_f32(logits.data_ptr(), n, A, None, seed, counter, acts.data_ptr(), None, None) == 0
        return acts.cpu().numpy()
    a0, a0b, a1, a2 = draw(5, 0), draw(5, 0), draw(5, 1), draw(6, 0)
    assert (a0 == a0b).all()                                         # a pure function of (seed, counter, row)
    assert (a0 != a1).mean() > 0.3 and (a0 != a2).mean() > 0.3       # new step / new seed -> new draws
    freq = np.bincount(a0, minlength=A) / n
    sigma = np.sqrt(p * (1 - p) / n)
    assert (np.abs(freq - p) < 5 * sigma + 1e-6).all(), (freq, p)
    # rows are independent: neighbouring rows agree no more often than sum p^2
    assert abs((a0[1:] == a0[:-1]).mean() - (p * p).sum()) < 5e-3


def _fused_agent(precision, A, max_batch, seed):
    from pytorch_drl_b200 import agents
    agent = agents.Agent([agents.ToChannelMajor()], agents.NatureCNN(4),
                         {'policy': agents.CategoricalPolicyHead(512, A), 'value_extrinsic': agents.SimpleValueHead(512)})
    agent.load_state_dict({k: tc.from_numpy(v) for k, v in po.init_params(A, ['extrinsic'], seed).items()})
    return agent.fuse(max_batch=max_batch, precision=precision)


def test_batched_rollout_manager_vs_reference_rollout_manager_golden(golden_dir):
    """All environments of a rank through ONE BatchedRolloutManager == one reference RolloutManager per environment (the
    golden was written by the unmodified reference): actions drawn, rewards, dones, frames, carry-over, episode metadata.
    Integer / index / byte work: bit-exact.  The policy runs on the fp32 engine (a bf16 logit can flip a draw)."""
    from oracle import actor_oracle as ao
    from pytorch_drl_b200 import algos
    g = np.load(os.path.join(golden_dir, 'actor_rollout.npz'))
    E, T, extra, A, n_rollouts, pseed = [int(v) for v in g['cfg']]
    agent = _fused_agent('fp32', A, 64, pseed)
    envs = [ao.SynthActorEnv(e, A, as_float=(e == 1)) for e in range(E)]      # one env hands out float frames in [0, 1]
    gens = [tc.Generator().manual_seed(ao.env_seed(e)) for e in range(E)]
    mgr = algos.BatchedRolloutManager(envs, agent, T, extra,
                                      noise=lambda step: tc.cat([ao.exponential_noise(gens[e], A) for e in range(E)]))
    steps = T + extra
    for k in range(n_rollouts):
        res = mgr.generate()
        md = res['metadata']
        obs = res['observations'].cpu().numpy()
        for e in range(E):
            np.testing.assert_array_equal(res['actions'][e, :steps + 1].cpu().numpy(), g[f'r{k}_e{e}_actions'])
            np.testing.assert_array_equal(res['rewards']['extrinsic'][e, :steps].cpu().numpy(), g[f'r{k}_e{e}_rew'])
            np.testing.assert_array_equal(res['dones'][e, :steps].cpu().numpy(), g[f'r{k}_e{e}_dones'])
            np.testing.assert_array_equal(ao.frame_checksums(obs[e, :steps + 1]), g[f'r{k}_e{e}_obs_sum'])
        for key in ('ep_len', 'ep_ret', 'ep_len_raw', 'ep_ret_raw'):
            want = np.concatenate([g[f'r{k}_e{e}_md_{key}'] for e in range(E)])
            np.testing.assert_allclose(np.asarray(md[key], np.float64), want, rtol=1e-12, atol=1e-9)


def test_batched_rollout_manager_vs_oracle_more_envs_no_extra_steps():
    """17 environments, extra_steps = 0, three rollouts, against the oracle's restatement of the manager (pinned to the
    reference by the golden above)."""
    from oracle import actor_oracle as ao
    from pytorch_drl_b200 import algos
    E, T, A, n_rollouts, pseed = 17, 5, 6, 3, 4
    agent = _fused_agent('fp32', A, 32, pseed)
    params = po.init_params(A, ['extrinsic'], pseed)
    envs = [ao.SynthActorEnv(e, A, as_float=False) for e in range(E)]
    gens = [tc.Generator().manual_seed(ao.env_seed(e)) for e in range(E)]
    mgr = algos.BatchedRolloutManager(envs, agent, T, 0,
                                      noise=lambda step: tc.cat([ao.exponential_noise(gens[e], A) for e in range(E)]))
    want = [ao.rollouts_oracle(params, e, T, 0, A, n_rollouts, ao.env_seed(e)) for e in range(E)]
    for k in range(n_rollouts):
        res = mgr.generate()
        obs = res['observations'].cpu().numpy()
        for e in range(E):
            np.testing.assert_array_equal(res['actions'][e, :T + 1].cpu().numpy(), want[e][k]['actions'], err_msg=f'env {e}')
            np.testing.assert_array_equal(res['rewards']['extrinsic'][e, :T].cpu().numpy(), want[e][k]['rewards'])
            np.testing.assert_array_equal(res['dones'][e, :T].cpu().numpy(), want[e][k]['dones'])
            np.testing.assert_array_equal(obs[e, :T + 1], want[e][k]['observations'])


def test_device_rollout_record_report_like_the_reference_rollout():
    """DeviceRollout.record / record_nexts / report (rollout.py:159-224) from host AND device sources against a numpy
    emulation of the reference's per-environment buffers, including the extra_steps carry-over and the erased remainder."""
    from pytorch_drl_b200.algos import DeviceRollout
    N, T, extra = 5, 4, 3
    steps = T + extra
    r = DeviceRollout(N, T, extra, ['extrinsic', 'intrinsic'])
    rs = np.random.RandomState(0)
    ref = dict(obs=np.zeros((N, steps + 1, 84, 84, 4), np.uint8), act=np.zeros((N, steps + 1), np.int64),
               ext=np.zeros((N, steps), F32), intr=np.zeros((N, steps), F32), done=np.zeros((N, steps), F32))
    for k in range(3):
        start = 0 if k == 0 else extra
        for t in range(start, steps):
            o = rs.randint(0, 256, size=(N, 84, 84, 4)).astype(np.uint8)
            a = rs.randint(0, 6, size=N)
            re, ri, d = rs.standard_normal(N).astype(F32), rs.uniform(size=N).astype(F32), (rs.uniform(size=N) < 0.3).astype(F32)
            ref['obs'][:, t], ref['act'][:, t], ref['ext'][:, t], ref['intr'][:, t], ref['done'][:, t] = o, a, re, ri, d
            if t % 2:       # device sources
                r.record(t, cu(o), cu(a.astype(np.int64)), {'extrinsic': cu(re), 'intrinsic': cu(ri)}, cu(d))
            else:           # host sources (numpy / lists)
                r.record(t, o, a.tolist(), {'extrinsic': re, 'intrinsic': tc.from_numpy(ri)}, d)
        o, a = rs.randint(0, 256, size=(N, 84, 84, 4)).astype(np.uint8), rs.randint(0, 6, size=N)
        ref['obs'][:, steps], ref['act'][:, steps] = o, a
        r.record_nexts(o, a)
        res = r.report()
        np.testing.assert_array_equal(res['observations'][:, :steps + 1].cpu().numpy(), ref['obs'])
        np.testing.assert_array_equal(res['actions'][:, :steps + 1].cpu().numpy(), ref['act'])
        np.testing.assert_array_equal(res['rewards']['extrinsic'][:, :steps].cpu().numpy(), ref['ext'])
        np.testing.assert_array_equal(res['rewards']['intrinsic'][:, :steps].cpu().numpy(), ref['intr'])
        np.testing.assert_array_equal(res['dones'][:, :steps].cpu().numpy(), ref['done'])
        keep = {n: v[:, T:steps].copy() for n, v in ref.items()}          # rollout.py:213-223
        for n in ref:
            ref[n][:, :extra] = keep[n]
        cur = r.storage.data                                              # the store now being recorded into
        np.testing.assert_array_equal(cur['observations'][:, :extra].cpu().numpy(), ref['obs'][:, :extra])
        np.testing.assert_array_equal(cur['actions'][:, :extra].cpu().numpy(), ref['act'][:, :extra])
        np.testing.assert_array_equal(cur['dones'][:, :extra].cpu().numpy(), ref['done'][:, :extra])
    with pytest.raises(ValueError):
        r.record(0, np.zeros((N, 84, 84, 3), np.uint8), [0] * N, {'extrinsic': [0.] * N, 'intrinsic': [0.] * N}, [0.] * N)
    with pytest.raises(KeyError):
        r.record(0, np.zeros((N, 84, 84, 4), np.uint8), [0] * N, {'extrinsic': [0.] * N}, [0.] * N)


@pytest.mark.parametrize('precision', _precisions())
def test_training_loop_collects_through_the_batched_manager(precision, tmp_path):
    """abstract.py:211-221 end to end on the device: collect (BatchedRolloutManager.generate + annotate + credit assignment)
    -> optimize -> persist, two iterations, in-kernel Philox draws; the policy must have moved and a checkpoint must exist."""
    from oracle import actor_oracle as ao
    from pytorch_drl_b200 import agents, algos
    A, T, B, E = 6, 8, 4, 6
    agent = _fused_agent(precision, A, E * (T + 1), 3)
    before = {k: v.clone() for k, v in agent.state_dict().items()}
    opt = tc.optim.Adam(agent.parameters(), lr=1e-3, eps=1e-5)
    algo = algos.PPO(rank=0, world_size=1, rollout_len=T, extra_steps=0,
                     credit_assignment_ops={'extrinsic': algos.GAE(gamma=0.99, use_dones=True, lambda_=0.95)},
                     global_step=0, max_steps=2 * E * T, checkpoint_frequency=E * T, checkpoint_dir=str(tmp_path),
                     policy_net=agents.DDPShim(agent), policy_optimizer=opt, opt_epochs=2, learner_batch_size=B,
                     envs_per_rank=E, sampler_seed=0)
    algo.rollout_mgr = algos.BatchedRolloutManager([ao.SynthActorEnv(e, A, as_float=False) for e in range(E)], agent, T, 0, seed=11)
    algo.training_loop()
    assert algo._global_step == 2 * E * T
    moved = max(float((agent.state_dict()[k].cpu() - before[k].cpu()).abs().max()) for k in before)
    assert moved > 1e-5
    assert any(f.endswith('.pth') for f in os.listdir(tmp_path))


def test_batched_rollout_manager_intrinsic_and_normalised_rewards():
    """The RND reward and the reward normaliser inside the batched manager (random_network_distillation.py:186-201,
    normalize_reward.py:171-194) equal the already-pinned batched pieces applied to the recorded rollout afterwards:
    reward of the NEW frame of each step, raw value fed to the running return statistics, normalised value recorded."""
    from oracle import actor_oracle as ao
    from pytorch_drl_b200 import algos
    from pytorch_drl_b200.envs import BatchedRewardNormalizer, RandomNetworkDistillation
    E, T, A = 4, 6, 6
    agent = _fused_agent('fp32', A, 32, 9)

    def make_rnd():
        rnd = RandomNetworkDistillation({'lr': 1e-4, 'eps': 1e-5}, world_size=1, max_batch=64, precision='fp32', seed=3)
        rnd.observe(tc.rand((16, 84, 84, 1), device=dev(), generator=tc.Generator(device=dev()).manual_seed(1)))
        rnd._sync_normalizers_global(); rnd._sync_normalizers_local()
        return rnd

    def make_nz():
        nz = BatchedRewardNormalizer(E, 0.5, False)          # trace length 10: the 40 warm-up steps fill the return statistics
        nz.step_block(cu(np.random.RandomState(2).uniform(size=(E, 40)).astype(F32)), cu(np.zeros((E, 40), F32)))
        nz.learn()
        return nz
    rnd, nz = make_rnd(), make_nz()
    mgr = algos.BatchedRolloutManager([ao.SynthActorEnv(e, A, as_float=False) for e in range(E)], agent, T, 0, seed=5, intrinsic=rnd,
                                      reward_normalizers={'intrinsic_rnd': nz})
    res = mgr.generate()
    assert sorted(res['rewards'].keys()) == ['extrinsic', 'intrinsic_rnd']
    steps_before = float(rnd._unsynced_normalizer.steps)
    rnd2, nz2 = make_rnd(), make_nz()
    rows_next = (tc.arange(E, device=dev())[:, None] * mgr.rollout.Tp + tc.arange(1, T + 1, device=dev())[None, :]).reshape(-1)
    # where an episode ended the store holds the reset frame: the reward was taken on the terminal frame step() returned
    envs2 = [ao.SynthActorEnv(e, A, as_float=False) for e in range(E)]
    acts, dn = res['actions'].cpu().numpy(), res['dones'].cpu().numpy()
    stepped = np.zeros((E, T, 84, 84, 4), np.uint8)
    for e, env in enumerate(envs2):
        env.reset()
        for t in range(T):
            stepped[e, t] = env.step(acts[e, t])[0]
            if dn[e, t]:
                env.reset()
    assert dn[:, :T].sum() > 0
    raw = rnd2.intrinsic_reward(cu(stepped.reshape(E * T, 84, 84, 4))).view(E, T)
    np.testing.assert_array_equal(res['observations'].view(-1, 84, 84, 4)[rows_next].cpu().numpy()[dn[:, :T].reshape(-1) == 0],
                                  stepped.reshape(E * T, 84, 84, 4)[dn[:, :T].reshape(-1) == 0])
    want = nz2.step_block(raw, res['dones'][:, :T].contiguous())
    got = res['rewards']['intrinsic_rnd'][:, :T]
    np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=1e-6, atol=1e-7)
    assert float(got.abs().max()) > 0
    np.testing.assert_allclose(nz.moment1.cpu().numpy(), nz2.moment1.cpu().numpy(), rtol=1e-6)
    assert steps_before == float(rnd2._unsynced_normalizer.steps) + E * T      # every new frame was observed once


def test_device_frame_stack_equals_reference_frame_stack_wrapper(golden_dir):
    """drl_rollout_stack_frames == FrameStackWrapper (frame_stack.py:50-59): byte work, bit-exact vs the golden written by the
    reference's wrapper (environment 0) and vs the oracle's restatement (all environments, in-place and slot-to-slot)."""
    from oracle import actor_oracle as ao
    from pytorch_drl_b200.algos import DeviceRollout
    g = np.load(os.path.join(golden_dir, 'actor_rollout.npz'))
    N, A = 3, 6
    r = DeviceRollout(N, 12, 0, ['extrinsic'])
    hosts = [ao.HostFrameStack(ao.SingleFrameEnv(9 + e, A, False)) for e in range(N)]
    singles = [ao.SingleFrameEnv(9 + e, A, False) for e in range(N)]
    want, t = [], 0
    want.append(np.stack([h.reset() for h in hosts]))
    r.stack_frames(0, 0, np.stack([s.reset() for s in singles]), np.ones(N, np.uint8))
    for k in list(range(5)) + ['reset'] + list(range(3)):
        if k == 'reset':
            want.append(np.stack([h.reset() for h in hosts]))
            r.stack_frames(t, t + 1, np.stack([s.reset() for s in singles]), np.ones(N, np.uint8))
        else:
            want.append(np.stack([h.step(k % A)[0] for h in hosts]))
            r.stack_frames(t, t + 1, np.stack([s.step(k % A)[0] for s in singles]))
        t += 1
    got = r.storage.data['observations'][:, :t + 1].cpu().numpy()
    np.testing.assert_array_equal(got, np.stack(want, axis=1))
    np.testing.assert_array_equal(ao.frame_checksums(got[0]), g['framestack_sums'])
    # mode 2 leaves an environment untouched; a float frame in [0, 1] is taken as rint(255 x)
    before = r.storage.data['observations'][:, t].cpu().numpy().copy()
    newest = np.stack([s.step(0)[0] for s in singles])
    r.stack_frames(t, t, newest.astype(np.float32) / np.float32(255.0), np.array([2, 0, 2], np.uint8))
    after = r.storage.data['observations'][:, t].cpu().numpy()
    np.testing.assert_array_equal(after[[0, 2]], before[[0, 2]])
    np.testing.assert_array_equal(after[1], np.concatenate([before[1][..., 1:], newest[1]], axis=-1))


@pytest.mark.parametrize('with_intrinsic', [False, True])
def test_batched_rollout_manager_device_frame_stack_equals_host_frame_stack(with_intrinsic):
    """frame_stack=True (single frames over PCIe, stacks built on the device) gives the rollout of the same environments
    behind a host-side FrameStackWrapper, bit for bit: frames, draws, rewards (incl. the RND reward on terminal frames), dones."""
    from oracle import actor_oracle as ao
    from pytorch_drl_b200 import algos
    from pytorch_drl_b200.envs import RandomNetworkDistillation
    E, T, extra, A = 5, 6, 1, 6
    agent = _fused_agent('fp32', A, 64, 2)
    out = []
    for device_stack in (False, True):
        rnd = None
        if with_intrinsic:
            rnd = RandomNetworkDistillation({'lr': 1e-4, 'eps': 1e-5}, world_size=1, max_batch=64, precision='fp32', seed=3)
            rnd.observe(tc.rand((16, 84, 84, 1), device=dev(), generator=tc.Generator(device=dev()).manual_seed(1)))
            rnd._sync_normalizers_global(); rnd._sync_normalizers_local()
        envs = [ao.SingleFrameEnv(20 + e, A, as_float=False) for e in range(E)]
        if not device_stack:
            envs = [ao.HostFrameStack(e) for e in envs]
        mgr = algos.BatchedRolloutManager(envs, agent, T, extra, seed=9, intrinsic=rnd, frame_stack=device_stack)
        res = [mgr.generate() for _ in range(2)]
        out.append([{k: (v.clone() if isinstance(v, tc.Tensor) else {kk: vv.clone() for kk, vv in v.items()} if k == 'rewards' else v)
                     for k, v in ro.items() if k in ('observations', 'actions', 'rewards', 'dones', 'metadata')} for ro in res])
    steps = T + extra
    for a, b in zip(*out):
        assert a['dones'][:, :steps].sum() > 0
        np.testing.assert_array_equal(a['observations'][:, :steps + 1].cpu().numpy(), b['observations'][:, :steps + 1].cpu().numpy())
        np.testing.assert_array_equal(a['actions'][:, :steps + 1].cpu().numpy(), b['actions'][:, :steps + 1].cpu().numpy())
        np.testing.assert_array_equal(a['dones'][:, :steps].cpu().numpy(), b['dones'][:, :steps].cpu().numpy())
        for k in a['rewards']:
            np.testing.assert_array_equal(a['rewards'][k][:, :steps].cpu().numpy(), b['rewards'][k][:, :steps].cpu().numpy())
        assert a['metadata'] == b['metadata']
