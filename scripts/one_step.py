"""Runs a few full-size learner steps (32 768 samples) on cuda:0 — the target of the ncu captures
whose summaries live under profiles/.  Usage: python scripts/one_step.py [steps] [precision] [debug flags "6=1,7=3"] [ppo|rnd]
`rnd`: RND predictor steps (RandomNetworkDistillation.learn_rows) on the same frames instead of PPO learner steps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch as tc  # noqa: E402

from pytorch_drl_b200 import ops  # noqa: E402
from pytorch_drl_b200.algos import RolloutStorage  # noqa: E402
from pytorch_drl_b200.engine import LearnerEngine  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
precision = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
n_envs, T, B, A = 1024, 32, 32, 6
if len(sys.argv) > 3:
    from pytorch_drl_b200 import _lib
    for kv in filter(None, sys.argv[3].split(',')):
        _lib.lib().drl_debug_set(int(kv.split('=')[0]), int(kv.split('=')[1]))
dev = tc.device('cuda', 0)
eng = LearnerEngine(A, ['value_extrinsic'], n_envs * B, precision)
g = tc.Generator(device=dev)
g.manual_seed(0)
eng.params.copy_(tc.randn(eng.nparams, device=dev, generator=g) * 0.02)
eng.sync_params()
st = RolloutStorage(n_envs, T, ['extrinsic'], dev)
d = st.as_dict()
for e0 in range(0, n_envs, 128):
    d['observations'][e0:e0 + 128] = tc.randint(0, 256, (128, st.Tp, 84, 84, 4), device=dev, dtype=tc.uint8, generator=g)
d['actions'][:] = tc.randint(0, A, (n_envs, st.Tp), device=dev, generator=g)
d['logprobs'][:] = -np.log(A)
d['advantages']['extrinsic'][:] = tc.randn((n_envs, st.Tp), device=dev, generator=g)
d['vtargs']['extrinsic'][:] = tc.randn((n_envs, st.Tp), device=dev, generator=g)
hyper = ops.make_hyper(A, [1.0], 0.2, 0.01, 1.0, False, True)
batch = ops.make_batch(d['actions'], d['logprobs'], [d['advantages']['extrinsic']], [d['vpreds']['extrinsic']],
                       [d['vtargs']['extrinsic']])
rs = np.random.RandomState(0)
if len(sys.argv) > 4 and sys.argv[4] == 'rnd':
    from pytorch_drl_b200.envs import RandomNetworkDistillation
    rnd = RandomNetworkDistillation({'lr': 1e-4, 'eps': 1e-5}, world_size=32, max_batch=n_envs * B, device=dev, precision=precision)
    rnd._synced_normalizer.moment2.fill_(1.0)
    for s in range(steps):
        rows = np.concatenate([rs.permutation(T)[:B] + e * st.Tp for e in range(n_envs)]).astype(np.int64)
        loss = rnd.learn_rows(d['observations'], tc.as_tensor(rows, device=dev))
    tc.cuda.synchronize()
    print('rnd loss', float(loss.item()))
    sys.exit(0)
for s in range(steps):
    rows = np.concatenate([rs.permutation(T)[:B] + e * st.Tp for e in range(n_envs)]).astype(np.int64)
    losses = eng.ppo_step(d['observations'], tc.as_tensor(rows, device=dev), batch, hyper)
    eng.adam_step(1e-3, (0.9, 0.999), 1e-5)
tc.cuda.synchronize()
print('losses', losses.cpu().tolist())
