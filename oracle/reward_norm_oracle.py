"""TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs, never by the product).

CPU restatement of the reference's reward normaliser
    ReturnAcc                 drl/envs/wrappers/stateful/normalize_reward.py:20-100
    NormalizeRewardWrapper    drl/envs/wrappers/stateful/normalize_reward.py:103-210   (step / learn semantics)
for N environments at once, in numpy float32 with the reference's operation order.  Pinned to the unmodified
reference by tests/golden/reward_norm.npz (oracle/make_golden.py::gen_reward_norm).

Per environment the accumulator keeps the pending (reward, done) pairs; whenever 2*L of them are buffered
(L = int(5 / (1 - gamma)), :27) it computes the discounted returns of the window backwards from a zero bootstrap
(:60-68), folds mean and mean-square of the FIRST L returns into the running moments with weight L (:70-88) and
drops those L pairs.  Normalising a reward divides it by the std of returns (shift=False, :97) and clips to
[-10, 10] (:121); it returns 0 while no statistics exist (:98-99)."""
import numpy as np

F32 = np.float32


class ReturnAcc:
    def __init__(self, gamma: float, clip_low: float = -10.0, clip_high: float = 10.0, use_dones: bool = True):
        self.gamma, self.use_dones = gamma, use_dones
        self.clip_low, self.clip_high = clip_low, clip_high
        self.trace_length = int(5 / (1 - gamma))                      # normalize_reward.py:27
        self.pending = []                                             # list of (float32 reward, bool done)
        self.steps, self.moment1, self.moment2 = F32(0), F32(0), F32(0)

    def _unload(self):                                                # normalize_reward.py:60-68
        L = self.trace_length
        ret = F32(0)
        returns = np.zeros(2 * L, F32)
        g32 = F32(self.gamma)
        for t in reversed(range(2 * L)):
            r, d = self.pending[t]
            if self.use_dones and d:
                ret = F32(r)                                          # mask = 0: r + 0 * gamma * ret
            else:
                ret = F32(F32(r) + F32(g32 * ret))                    # python-scalar gamma is applied in float32
            returns[t] = ret
        self.pending = self.pending[L:]
        return returns[:L]

    def _update_from_returns(self, returns):                          # normalize_reward.py:70-88
        L = F32(returns.shape[0])
        steps = F32(self.steps + L)
        keep = F32(F32(steps - L) / steps)
        w = F32(L / steps)
        mean = F32(np.mean(returns.astype(np.float64)))
        msq = F32(np.mean(np.square(returns).astype(np.float64)))
        self.moment1 = F32(F32(self.moment1 * keep) + F32(w * mean))
        self.moment2 = F32(F32(self.moment2 * keep) + F32(w * msq))
        self.steps = steps

    def update(self, r, d):                                           # normalize_reward.py:90-95
        self.pending.append((F32(r), bool(d)))
        if len(self.pending) >= 2 * self.trace_length:
            self._update_from_returns(self._unload())

    def forward(self, r, eps: float = 1e-8):                          # normalize_reward.py:97-100 + normalize.py:150-162
        if self.steps == 0:
            return F32(0)
        var = F32(self.moment2 - F32(self.moment1 * self.moment1))
        y = F32(F32(r) * F32(1.0 / np.sqrt(F32(var + F32(eps)))))
        return F32(min(max(y, F32(self.clip_low)), F32(self.clip_high)))


class BatchedRewardNormalizer:
    """N environments of one rank between two `learn()` calls: `step` semantics of NormalizeRewardWrapper
    (:171-194) applied to whole [N, T] blocks; `learn` = global then local sync (:196-210, 144-156), with the
    global mean taken over all N environments (world_size = N in the reference's rank = env model)."""

    def __init__(self, n_envs: int, gamma: float, use_dones: bool):
        self.unsynced = [ReturnAcc(gamma, -10.0, 10.0, use_dones) for _ in range(n_envs)]
        self.synced = ReturnAcc(gamma, -10.0, 10.0, use_dones)

    def step_block(self, rewards: np.ndarray, dones: np.ndarray) -> np.ndarray:
        N, T = rewards.shape
        out = np.zeros((N, T), F32)
        for e in range(N):
            acc = self.unsynced[e]
            for t in range(T):
                out[e, t] = self.synced.forward(rewards[e, t])
                acc.update(rewards[e, t], dones[e, t] != 0)
        return out

    def learn(self):
        n = F32(len(self.unsynced))
        for name in ('steps', 'moment1', 'moment2'):
            tot = F32(0)
            for a in self.unsynced:
                tot = F32(tot + getattr(a, name))
            setattr(self.synced, name, F32(tot / n))
        for a in self.unsynced:
            a.steps, a.moment1, a.moment2 = self.synced.steps, self.synced.moment1, self.synced.moment2
