"""Fused PPO loss kernel (drl_ppo_loss_fwd_bwd) alone on cuda:0 at a size larger than L2.
Usage: python scripts/loss_rate.py [samples]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch as tc  # noqa: E402

from pytorch_drl_b200 import _lib, ops  # noqa: E402
from pytorch_drl_b200._lib import current_stream, ptr  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
A = 6
dev = tc.device('cuda', 0)
L = _lib.lib()
logits, values = tc.randn((n, A), device=dev), tc.randn((n, 1), device=dev)
acts = tc.randint(0, A, (n,), device=dev)
lp = tc.log_softmax(logits, -1).gather(1, acts[:, None]).squeeze(1) + 0.05 * tc.randn((n,), device=dev)
advs, vold, vtg = (tc.randn((n,), device=dev) for _ in range(3))
hyper = ops.make_hyper(A, [1.0], 0.2, 0.01, 1.0, False, True)
batch = ops.make_batch(acts, lp, [advs], [vold], [vtg])
losses, lpo, eno = tc.empty(5, device=dev), tc.empty(n, device=dev), tc.empty(n, device=dev)
dl, dvv = tc.empty_like(logits), tc.empty_like(values)
res = {}
for flag in [0]:

    def call():
        assert L.drl_ppo_loss_fwd_bwd(ptr(logits), ptr(values), None, n, ctypes.byref(batch), ctypes.byref(hyper),
                                      ptr(losses), ptr(lpo), ptr(eno), ptr(dl), ptr(dvv), current_stream()) == 0
    for _ in range(3):
        call()
    tc.cuda.synchronize()
    e0, e1 = tc.cuda.Event(enable_timing=True), tc.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        call()
    e1.record()
    tc.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    res[flag] = (losses.clone(), dl.clone())
    print(f'{n} samples {ms * 1e3:.1f} us = {(8 * A + 40) * n / ms / 1e6:.0f} GB/s')
if len(res) == 2:
    print('max |d losses|', float((res[0][0] - res[1][0]).abs().max()), ' dlogits equal:', bool(tc.equal(res[0][1], res[1][1])))
