"""RND predictor throughput on cuda:0: learn steps and batched intrinsic reward, with the per-launch-site breakdown.
Usage: python scripts/rnd_rate.py [samples] [bf16|fp32]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch as tc  # noqa: E402

from pytorch_drl_b200.envs import RandomNetworkDistillation  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
precision = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
dev = tc.device('cuda', 0)
rnd = RandomNetworkDistillation({'lr': 1e-4, 'betas': (0.9, 0.999), 'eps': 1e-8}, world_size=1, max_batch=nb, device=dev, precision=precision)
rnd._synced_normalizer.moment2.fill_(1.0)
obs = tc.randint(0, 256, (nb, 84, 84, 4), device=dev, dtype=tc.uint8)
rows = tc.arange(nb, device=dev, dtype=tc.int64)
for name, fn in (('learn', lambda: rnd.learn_rows(obs, rows)), ('intrinsic_reward', lambda: rnd.intrinsic_reward(obs, rows))):
    for _ in range(3):
        fn()
    tc.cuda.synchronize()
    e0, e1 = tc.cuda.Event(enable_timing=True), tc.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    tc.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f'rnd[{precision}] {name}: {nb} samples in {ms:.3f} ms = {nb / ms * 1e3 / 1e6:.2f} M samples/s')
    # per-launch-site breakdown (CUDA events recorded by the library)
    import ctypes
    from pytorch_drl_b200 import _lib
    L = _lib.lib()
    L.drl_profile_enable(1)
    nbuf = ctypes.create_string_buffer(48 * 64)
    msb, cnt, ns = (ctypes.c_float * 64)(), (ctypes.c_int * 64)(), ctypes.c_int()
    L.drl_profile_read(64, nbuf, msb, cnt, ctypes.byref(ns))
    fn()
    tc.cuda.synchronize()
    L.drl_profile_read(64, nbuf, msb, cnt, ctypes.byref(ns))
    L.drl_profile_enable(0)
    sites = {nbuf.raw[48 * i:48 * i + 48].split(b'\0')[0].decode(): (round(msb[i], 3), cnt[i]) for i in range(ns.value)}
    print('   sites (total ms, launches):', sites)
