"""Why does the tensor-core engine's minibatch gradient differ from fp32 by ~5 % of its norm, independent of the batch size?
CPU-only emulation on the oracle network (test infrastructure: imports oracle/): operands (weights and layer inputs) are
rounded to `m` explicit mantissa bits with a straight-through gradient — bf16 = 7 bits, tf32 = 10, a 2-term bf16 split = 15 —
and the gradient of a PPO-like loss with random-sign advantages / value targets is compared with plain fp32.  A second
run keeps the rounding but takes every ReLU mask from the fp32 forward: what is left is the smooth part of the error.

    python scripts/precision_emulation.py [--samples 512]
"""
import argparse
import os
import sys

import numpy as np
import torch as tc
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ppo_oracle as po  # noqa: E402


def rnd_to(x, m):
    """round-to-nearest-even to m explicit mantissa bits, straight-through gradient"""
    if m >= 23:
        return x
    i = x.detach().contiguous().view(tc.int32)
    drop = 23 - m
    lsb = (i >> drop) & 1
    r = ((i + ((1 << (drop - 1)) - 1) + lsb) >> drop) << drop
    return x + (r.view(tc.float32) - x.detach())


def forward(p, obs, m, masks=None, keep_masks=None):
    x = obs.permute(0, 3, 1, 2)
    K = po.TRUNK_KEYS
    specs = [(K[0], K[1], 4), (K[2], K[3], 2), (K[4], K[5], 1)]
    out_masks = []
    for li, (w, b, s) in enumerate(specs):
        pre = F.conv2d(rnd_to(x, m) if li else x, rnd_to(p[w], m), p[b], stride=s)   # conv1 reads exact uint8 values
        mk = (pre > 0).float() if masks is None else masks[li]
        out_masks.append((pre > 0).float().detach())
        x = pre * mk
    pre = F.linear(rnd_to(x.flatten(1), m), rnd_to(p[K[6]], m), p[K[7]])
    mk = (pre > 0).float() if masks is None else masks[3]
    out_masks.append((pre > 0).float().detach())
    feat = pre * mk
    if keep_masks is not None:
        keep_masks.extend(out_masks)
    return feat


def grads(params, obs, actions, adv, vtarg, m, masks=None, keep_masks=None):
    p = {k: tc.from_numpy(v.copy()).requires_grad_(True) for k, v in params.items()}
    feat = forward(p, obs, m, masks, keep_masks)
    logits, values = po.heads_forward(p, feat, ['extrinsic'])
    logp, ent = po.categorical_logprob_entropy(logits, actions)
    ratio = tc.exp(logp - np.log(1.0 / 6))
    loss = -(tc.min(adv * ratio, adv * tc.clip(ratio, 0.9, 1.1))).mean() - 0.01 * ent.mean() + 0.5 * ((values['extrinsic'] - vtarg) ** 2).mean()
    loss.backward()
    return {k: v.grad.numpy().astype(np.float64) for k, v in p.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--samples', type=int, default=512)
    a = ap.parse_args()
    rs = np.random.RandomState(0)
    nb = a.samples
    params = po.init_params(6, ['extrinsic'], seed=11)
    obs = po.scale_obs(tc.from_numpy(rs.randint(0, 256, size=(nb, 84, 84, 4), dtype=np.uint8)))
    actions = tc.from_numpy(rs.randint(0, 6, size=nb))
    adv = tc.from_numpy(rs.standard_normal(nb).astype(np.float32))
    vtarg = tc.from_numpy(rs.standard_normal(nb).astype(np.float32))
    masks32 = []
    g32 = grads(params, obs, actions, adv, vtarg, 23, keep_masks=masks32)
    names = {'bf16 (7 bits)': 7, 'tf32 (10 bits)': 10, 'bf16 x 2 split (15 bits)': 15}
    short = lambda k: k.replace('_architecture._network.', 'trunk.').replace('_predictors.', '').replace('._network.0', '')  # noqa: E731
    for name, m in names.items():
        mk = []
        g = grads(params, obs, actions, adv, vtarg, m, keep_masks=mk)
        flips = [float((a_ != b_).float().mean()) for a_, b_ in zip(mk, masks32)]
        gm = grads(params, obs, actions, adv, vtarg, m, masks=masks32)
        rel = {short(k): round(float(np.linalg.norm(g[k] - g32[k]) / np.linalg.norm(g32[k])), 5) for k in g32 if k.endswith('weight')}
        relm = {short(k): round(float(np.linalg.norm(gm[k] - g32[k]) / np.linalg.norm(g32[k])), 5) for k in g32 if k.endswith('weight')}
        print(f'{name}: ReLU-mask flip fraction per layer {[round(f, 5) for f in flips]}')
        print(f'   gradient rel-L2 vs fp32                 {rel}')
        print(f'   same rounding, fp32 ReLU masks           {relm}')


if __name__ == '__main__':
    main()
