"""How far is each precision mode of the learner step from the reference?  (run on a B200)

  1. `PPO.optimize` against the UNMODIFIED reference's 2-rank DDP run (tests/golden/ppo_2rank*.npz): worst relative
     error of the final parameters (summary vectors) and of the per-epoch metrics, per precision.
  2. Gradient of one learner step at 512 / 4096 / 32768 samples: rel-L2 and cosine per parameter tensor of each
     tensor-core mode against the CUDA-core fp32 engine (itself within 2e-4 of torch-CPU).

    python scripts/precision_report.py [--out gpurun_out/precision.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch as tc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from oracle import ppo_oracle as po  # noqa: E402  (the checker; this script is test infrastructure)
from _util import summary  # noqa: E402


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def golden_run(tag, precision):
    from pytorch_drl_b200 import agents, algos
    dev = tc.device('cuda', 0)
    g = np.load(os.path.join(ROOT, 'tests', 'golden', f'{tag}.npz'))
    T, B, A, epochs = int(g['cfg_T']), int(g['cfg_B']), int(g['cfg_A']), int(g['cfg_epochs'])
    seed, world = int(g['cfg_seed']), int(g['cfg_world'])
    clip, ent, vf, lr, b1, b2, eps, gamma, lam, vclip = [float(x) for x in g['cfg_hyper']]
    rk = [str(k) for k in g['cfg_rkeys']]
    rw = {k: float(w) for k, w in zip(rk, g['cfg_rweights'])}
    preds = {'policy': agents.CategoricalPolicyHead(512, A)}
    for k in rk:
        preds[f'value_{k}'] = agents.SimpleValueHead(512)
    agent = agents.Agent([agents.ToChannelMajor()], agents.NatureCNN(4), preds)
    init = po.init_params(A, rk, int(g['cfg_param_seed']))
    agent.load_state_dict({k: tc.from_numpy(v) for k, v in init.items()})
    agent.fuse(max_batch=world * (T + 1), precision=precision)
    opt = tc.optim.Adam(agent.parameters(), lr=lr, betas=(b1, b2), eps=eps)
    credit = {k: algos.GAE(gamma=(gamma if k == 'extrinsic' else 0.99), use_dones=(k == 'extrinsic'), lambda_=lam) for k in rk}
    algo = algos.PPO(rank=0, world_size=1, rollout_len=T, extra_steps=0, credit_assignment_ops=credit,
                     global_step=1, max_steps=10 ** 9, policy_net=agent, policy_optimizer=opt,
                     standardize_adv=True, opt_epochs=epochs, learner_batch_size=B, clip_param_init=clip,
                     clip_param_final=clip, ent_coef_init=ent, ent_coef_final=ent, vf_loss_coef=vf,
                     vf_loss_clipping=bool(vclip), reward_weights=(rw if len(rk) > 1 else None),
                     envs_per_rank=world, sampler_seed=seed)
    st = algos.RolloutStorage(world, T, rk, dev)
    for r in range(world):
        raw = po.make_synthetic_env_rollout(r, seed, T, A, rk)
        st.load_env(r, raw['observations'], raw['actions'], raw['rewards'], raw['dones'])
    rollout = algo.credit_assignment(algo.annotate(st.as_dict()))
    out = {'annotate': {}}
    for r in range(world):
        out['annotate'][f'logp_r{r}_maxabs'] = float(np.abs(rollout['logprobs'][r, :T].cpu().numpy() - g[f'r{r}::logp']).max())
        for k in rk:
            out['annotate'][f'adv_{k}_r{r}_relL2'] = rel_l2(rollout['advantages'][k][r, :T].cpu().numpy(), g[f'r{r}::adv_{k}'])
            out['annotate'][f'vtarg_{k}_r{r}_relL2'] = rel_l2(rollout['vtargs'][k][r, :T].cpu().numpy(), g[f'r{r}::vtarg_{k}'])
    algo.optimize(rollout)
    final = {k: v.detach().cpu().numpy() for k, v in agent.state_dict().items()}
    out['final_params'] = {}
    for k in po.param_keys(rk):
        gold = g[f'r0::final::{k}']
        big = final[k].size > 4096
        s = summary(final[k]) if big else np.asarray(final[k], np.float64)
        s0 = summary(init[k]) if big else np.asarray(init[k], np.float64)
        gold = gold.reshape(s.shape)
        # error of the parameter itself, and relative to how far the reference moved it (the update)
        upd = np.linalg.norm((gold - s0)[2:] if big else gold - s0)
        err = np.linalg.norm((s - gold)[2:] if big else s - gold)
        out['final_params'][k] = {'relL2': float(err / max(np.linalg.norm(gold[2:] if big else gold), 1e-30)),
                                  'err_over_update': float(err / max(upd, 1e-30)),
                                  'maxabs': float(np.abs((s - gold)[2:] if big else s - gold).max())}
    em = g['r0::epoch_metrics']
    got = np.array([[algo.last_epoch_metrics[ep][n] for n in algos.ppo.METRIC_NAMES] for ep in range(epochs)])
    out['epoch_metrics'] = {'got': got.tolist(), 'ref': em[:, :5].tolist(),
                            'max_rel_first4': float((np.abs(got[:, :4] - em[:, :4]) / np.maximum(np.abs(em[:, :4]), 1e-3)).max()),
                            'max_abs_first4': float(np.abs(got[:, :4] - em[:, :4]).max()),
                            'max_abs_clipfrac': float(np.abs(got[:, 4] - em[:, 4]).max())}
    return out


def grad_report(precisions, sizes):
    from pytorch_drl_b200 import ops
    from pytorch_drl_b200.engine import LearnerEngine
    dev = tc.device('cuda', 0)
    A, rk = 6, ['extrinsic']
    nbmax = max(sizes)
    g = tc.Generator(device=dev)
    g.manual_seed(2024)
    obs = tc.empty((nbmax, 84, 84, 4), device=dev, dtype=tc.uint8)
    for s in range(0, nbmax, 4096):
        m = min(4096, nbmax - s)
        obs[s:s + m] = tc.randint(0, 256, (m, 84, 84, 4), device=dev, dtype=tc.uint8, generator=g)
    actions = tc.randint(0, A, (nbmax,), device=dev, generator=g)
    logp_old = tc.full((nbmax,), -float(np.log(A)), device=dev)
    adv, vold, vtarg = (tc.randn(nbmax, device=dev, generator=g) for _ in range(3))
    hyper = ops.make_hyper(A, [1.0], 0.1, 0.01, 0.5, False, True)
    batch = ops.make_batch(actions, logp_old, [adv], [vold], [vtarg])
    params = po.init_params(A, rk, seed=11)
    res = {}
    for nb in sizes:
        rows = tc.arange(nb, device=dev, dtype=tc.int64)
        got = {}
        for prec in ['fp32'] + list(precisions):
            eng = LearnerEngine(A, ['value_extrinsic'], nb, prec)
            eng.load_named({k: tc.from_numpy(v) for k, v in params.items()})
            losses = eng.ppo_step(obs, rows, batch, hyper).cpu().numpy().astype(np.float64)
            got[prec] = (losses, {k: v.cpu().numpy().astype(np.float64) for k, v in eng.named_views(eng.grads).items()})
            del eng
            tc.cuda.empty_cache()
        r = {}
        for prec in precisions:
            per = {}
            for k, ref in got['fp32'][1].items():
                a, b = got[prec][1][k].ravel(), ref.ravel()
                per['.'.join(k.split('.')[-3:])] = {'relL2': rel_l2(a, b), 'cos': float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-30))}
            flat_a = np.concatenate([v.ravel() for v in got[prec][1].values()])
            flat_b = np.concatenate([v.ravel() for v in got['fp32'][1].values()])
            r[prec] = {'tensors': per, 'all_relL2': rel_l2(flat_a, flat_b),
                       'loss_abs_err': np.abs(got[prec][0] - got['fp32'][0]).tolist()}
        res[str(nb)] = r
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'precision.json'))
    ap.add_argument('--precisions', default='bf16')
    ap.add_argument('--sizes', default='512,4096,32768')
    a = ap.parse_args()
    precs = a.precisions.split(',')
    rep = {'golden': {}, 'grads_vs_fp32_engine': None}
    for tag in ('ppo_2rank', 'ppo_2rank_rnd'):
        for prec in ['fp32'] + precs:
            rep['golden'][f'{tag}/{prec}'] = golden_run(tag, prec)
    rep['grads_vs_fp32_engine'] = grad_report(precs, [int(s) for s in a.sizes.split(',')])
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(rep, open(a.out, 'w'), indent=1)
    for k, v in rep['golden'].items():
        worst = max(v['final_params'].items(), key=lambda kv: kv[1]['err_over_update'])
        print(k, 'final-param worst err/update', round(worst[1]['err_over_update'], 4), worst[0].split('.')[-3:],
              'worst relL2', round(max(x['relL2'] for x in v['final_params'].values()), 6),
              'metrics max rel', round(v['epoch_metrics']['max_rel_first4'], 5), 'max abs', round(v['epoch_metrics']['max_abs_first4'], 6))
    for nb, r in rep['grads_vs_fp32_engine'].items():
        for prec, x in r.items():
            print(nb, prec, 'all', round(x['all_relL2'], 5), {k: round(t['relL2'], 4) for k, t in x['tensors'].items()})


if __name__ == '__main__':
    main()
