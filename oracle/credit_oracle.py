"""TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product).

CPU restatement (numpy float32, the reference's operation order) of
    NStepAdvantageEstimator.call                  drl/algos/common/credit_assignment.py:225-249
    SimpleDiscreteBellmanOptimalityOperator.call  drl/algos/common/credit_assignment.py:276-305
Pinned to the unmodified reference by tests/golden/nstep_bellman.npz (oracle/make_golden.py::gen_nstep_bellman)
and by the expected values of the reference's own tests (drl/algos/common/test_credit_assignment.py:64-170)."""
import numpy as np

F32 = np.float32


def _backup(R, rewards, dones, t, n, gamma, use_dones):
    g = F32(gamma)
    for s in reversed(range(n)):                                   # R = r + nonterminal * gamma * R   (:243-246, :299-303)
        nt_g = F32(F32(1.0) - dones[t + s]) * g if use_dones else g
        R = F32(rewards[t + s] + F32(nt_g * R))
    return R


def nstep_advantages(rollout_len, extra_steps, rewards, vpreds, dones, gamma, use_dones):
    T, n = rollout_len, extra_steps + 1
    rewards, vpreds, dones = rewards.astype(F32), vpreds.astype(F32), dones.astype(F32)
    adv = np.zeros(T, F32)
    for t in range(T):
        adv[t] = F32(_backup(vpreds[t + n], rewards, dones, t, n, gamma, use_dones) - vpreds[t])
    return adv


def bellman_backups(rollout_len, extra_steps, rewards, qpreds, tgt_qpreds, dones, gamma, use_dones, double_q):
    T, n = rollout_len, extra_steps + 1
    rewards, dones = rewards.astype(F32), dones.astype(F32)
    qpreds, tgt_qpreds = qpreds.astype(F32), tgt_qpreds.astype(F32)
    out = np.zeros(T, F32)
    for t in range(T):
        a = int(np.argmax(qpreds[t + n] if double_q else tgt_qpreds[t + n]))     # first maximum, like torch.argmax
        out[t] = _backup(tgt_qpreds[t + n, a], rewards, dones, t, n, gamma, use_dones)
    return out
