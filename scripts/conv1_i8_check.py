"""Integer conv1 forward (umma_conv1_i8.cuh) against the fp32 torch oracle: act1 rel-L2 on random frames with shuffled row
indices, and the conv1_fwd time at 32 768 samples.
Usage: python scripts/conv1_i8_check.py [nb_timing]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch as tc  # noqa: E402

from oracle import ppo_oracle as po  # noqa: E402
from pytorch_drl_b200 import _lib  # noqa: E402
from pytorch_drl_b200.engine import LearnerEngine  # noqa: E402
from bench import SiteProfile  # noqa: E402

L = _lib.lib()
dev = tc.device('cuda', 0)
A, rk = 6, ['extrinsic']
params = po.init_params(A, rk, seed=11)
rs = np.random.RandomState(3)
for nstore, nb in ((50, 37), (400, 333), (3, 1)):
    obs = rs.randint(0, 256, size=(nstore, 84, 84, 4), dtype=np.uint8)
    rows = rs.permutation(nstore)[:nb].astype(np.int64)
    keep = {}
    p = {k: tc.from_numpy(v) for k, v in params.items()}
    with tc.no_grad():
        po.trunk_forward(p, po.scale_obs(tc.from_numpy(obs[rows])), keep)
    ref = keep['act1'].permute(0, 2, 3, 1).numpy()
    for flag in (1,):
        eng = LearnerEngine(A, ['value_extrinsic'], max(nb, 64), 'bf16')
        eng.load_named({k: tc.from_numpy(v) for k, v in params.items()})
        eng.forward(tc.from_numpy(obs).to(dev), tc.from_numpy(rows).to(dev))
        tc.cuda.synchronize()
        got = eng.activation(1, nb).float().cpu().numpy().reshape(ref.shape)
        err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        print(f'nb {nb} act1 rel-L2 vs fp32 oracle {err:.3e}  max abs {np.abs(got - ref).max():.3e}', flush=True)
        del eng

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
g = tc.Generator(device=dev)
g.manual_seed(1)
obs = tc.randint(0, 256, (nb, 84, 84, 4), device=dev, dtype=tc.uint8, generator=g)
rows = tc.randperm(nb, device=dev, generator=g)
for flag in (1, 1):
    eng = LearnerEngine(A, ['value_extrinsic'], nb, 'bf16')
    eng.load_named({k: tc.from_numpy(v) for k, v in params.items()})
    for _ in range(3):
        eng.forward(obs, rows)
    sp = SiteProfile(L)
    sp.start()
    for _ in range(10):
        eng.forward(obs, rows)
    tc.cuda.synchronize()
    prof = sp.stop()
    print('conv1_fwd ms', prof.get('conv1_fwd'), flush=True)
    del eng
