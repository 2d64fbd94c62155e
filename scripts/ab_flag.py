"""A/B check of a kernel-variant switch (drl_debug_set): one learner step on identical inputs with flag <key> = <v0> and = <v1>,
relative difference of the whole gradient vector and of the losses (atomics-order noise ~ 5e-7 means identical arithmetic).
Usage: python scripts/ab_flag.py <key> <v0> <v1> [nb]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch as tc  # noqa: E402

from pytorch_drl_b200 import _lib, ops  # noqa: E402
from pytorch_drl_b200.engine import LearnerEngine  # noqa: E402

key, v0, v1 = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
nb = int(sys.argv[4]) if len(sys.argv) > 4 else 777
L = _lib.lib()
A = 6
dev = tc.device('cuda', 0)
g = tc.Generator(device=dev)
g.manual_seed(3)
obs = tc.randint(0, 256, (nb, 84, 84, 4), device=dev, dtype=tc.uint8, generator=g)
rows = tc.arange(nb, device=dev)
actions = tc.randint(0, A, (nb,), device=dev, generator=g)
lp = tc.full((nb,), -float(np.log(A)), device=dev)
adv, vold, vt = (tc.randn(nb, device=dev, generator=g) for _ in range(3))
hyper = ops.make_hyper(A, [1.0], 0.1, 0.01, 0.5, False, True)
batch = ops.make_batch(actions, lp, [adv], [vold], [vt])
res = []
for v in (v0, v1):
    L.drl_debug_set(key, v)
    eng = LearnerEngine(A, ['value_extrinsic'], nb, 'bf16')
    gg = tc.Generator(device=dev)
    gg.manual_seed(0)
    eng.params.copy_(tc.randn(eng.nparams, device=dev, generator=gg) * 0.02)
    eng.sync_params()
    losses = eng.ppo_step(obs, rows, batch, hyper).clone()
    tc.cuda.synchronize()
    res.append((eng.grads.clone(), losses))
    del eng
print('grads rel diff', float((res[1][0] - res[0][0]).norm() / res[0][0].norm()), ' losses max abs diff',
      float((res[1][1] - res[0][1]).abs().max()))
