"""Experiment behind tma_conv2s_kernel: do TMA tensor-map loads with 128B swizzle work into shared-memory destinations that
are 128-byte but not 1024-byte aligned?  Runs one learner step with (a) the default 1024-byte aligned staging buffers,
(b) buffers packed at the box size (debug flag 0 = 2), (c) CTA pairs, (d) CTA pairs + sample-spanning tiles, and prints the
relative difference of the whole gradient vector (atomics-order noise ~ 5e-7 means identical).  Usage: python scripts/try_unaligned_tma.py"""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch as tc
from pytorch_drl_b200 import _lib, ops
from pytorch_drl_b200.engine import LearnerEngine
L = _lib.lib()
A, nb = 6, 777
dev = tc.device('cuda', 0)
g = tc.Generator(device=dev); g.manual_seed(3)
obs = tc.randint(0, 256, (nb, 84, 84, 4), device=dev, dtype=tc.uint8, generator=g)
rows = tc.arange(nb, device=dev)
actions = tc.randint(0, A, (nb,), device=dev, generator=g)
lp = tc.full((nb,), -float(np.log(A)), device=dev)
adv, vold, vt = (tc.randn(nb, device=dev, generator=g) for _ in range(3))
hyper = ops.make_hyper(A, [1.0], 0.1, 0.01, 0.5, False, True)
batch = ops.make_batch(actions, lp, [adv], [vold], [vt])
res = []
for flag0, flag7 in ((0, 0), (2, 0), (0, 3), (0, 11)):
    L.drl_debug_set(0, flag0); L.drl_debug_set(7, flag7)
    eng = LearnerEngine(A, ['value_extrinsic'], nb, 'bf16')
    gg = tc.Generator(device=dev); gg.manual_seed(0)
    eng.params.copy_(tc.randn(eng.nparams, device=dev, generator=gg) * 0.02); eng.sync_params()
    eng.ppo_step(obs, rows, batch, hyper)
    tc.cuda.synchronize()
    res.append(eng.grads.clone())
    del eng
for i in range(1, 4):
    d = (res[i] - res[0]).norm() / res[0].norm()
    print('variant', i, 'rel diff vs aligned single-CTA', float(d))
