"""Engineering experiment: which smem-descriptor base-offset convention do shifted (non-1024B-aligned)
128B-swizzled operands need?  Runs the bf16 forward + ppo step in each mode and prints errors vs the
oracle.  Usage: python scripts/try_base_offset.py"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch as tc
from oracle import ppo_oracle as po
from pytorch_drl_b200 import _lib
from pytorch_drl_b200.engine import LearnerEngine

L = _lib.lib()
A, rk, nb = 6, ['extrinsic'], 37
params = po.init_params(A, rk, seed=11)
rs = np.random.RandomState(3)
obs = rs.randint(0, 256, size=(nb, 84, 84, 4), dtype=np.uint8)
keep = {}
p = {k: tc.from_numpy(v) for k, v in params.items()}
with tc.no_grad():
    po.trunk_forward(p, po.scale_obs(tc.from_numpy(obs)), keep)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


for use_tma, base in ((0, 0), (1, 1), (1, 0)):
    L.drl_debug_set(1, use_tma)
    L.drl_debug_set(0, base)
    eng = LearnerEngine(A, ['value_extrinsic'], 64, 'bf16')
    eng.load_named({k: tc.from_numpy(v) for k, v in params.items()})
    eng.forward(tc.from_numpy(obs).cuda())
    tc.cuda.synchronize()
    out = []
    for i, name in enumerate(['act1', 'act2', 'act3']):
        ref = keep[name].permute(0, 2, 3, 1).numpy()
        got = eng.activation(i + 1, nb).float().cpu().numpy().reshape(ref.shape)
        out.append(f'{name}={rel(got, ref):.4f}')
    print(f'tma_conv={use_tma} base_offset={base}:', ' '.join(out), flush=True)
    del eng
