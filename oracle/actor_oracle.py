"""CPU oracle for the actor side (batched policy sampling + rollout recording).

TEST INFRASTRUCTURE ONLY — see the header of `oracle/ppo_oracle.py` for who may import this.
Pinned by `tests/golden/actor_rollout.npz`, which `oracle/make_golden.py::gen_actor_rollout` writes by running the UNMODIFIED
reference `RolloutManager` (drl/algos/common/rollout.py:227-357) with the reference `Agent` on the synthetic environments
below, one manager per environment, each under `torch.manual_seed(env_seed(e))`.

Citations relative to /root/reference/.
"""
from typing import Dict, List

import numpy as np
import torch as tc

F32 = np.float32

CFG = dict(E=3, T=6, extra_steps=2, A=6, rollouts=2, param_seed=21)


def env_seed(e: int) -> int:
    return 10000 * e + 7          # the launch.py:33-45 seeding rule (10000 * rank + seed)


class SynthActorEnv:
    """Deterministic stand-in for a wrapped Atari environment: uint8 frame stacks from a seeded stream, rewards that depend
    on the action taken (a wrong draw changes everything after it), episode ends and `ale.lives` on some steps.
    `as_float=True` returns the frames as float32 in [0, 1] (what the reference's network consumes)."""

    def __init__(self, e: int, num_actions: int, as_float: bool):
        self._rs = np.random.RandomState(1000 + e)
        self._A, self._as_float, self._e = num_actions, as_float, e
        self.frames: List[np.ndarray] = []          # every uint8 frame handed out, in order

    @property
    def reward_spec(self):
        class _Spec:                                  # RewardSpec(keys=...) of drl/envs/wrappers/stateless/abstract.py
            keys = ['extrinsic_raw', 'extrinsic']
        return _Spec()

    def _frame(self):
        f = self._rs.randint(0, 256, size=(84, 84, 4)).astype(np.uint8)
        self.frames.append(f)
        return (f.astype(F32) / F32(255.0)) if self._as_float else f

    def reset(self):
        return self._frame()

    def step(self, a):
        a = int(a)
        assert 0 <= a < self._A
        raw = float(np.round((a + 1) * self._rs.uniform(), 3))
        r = {'extrinsic_raw': raw, 'extrinsic': float(np.sign(raw - 1.5))}
        done = bool(self._rs.uniform() < 0.25)
        info = {'ale.lives': int(self._rs.randint(0, 2))} if self._e % 2 == 0 else {}
        return self._frame(), r, done, info


def exponential_noise(gen: tc.Generator, num_actions: int) -> tc.Tensor:
    """The q ~ Exp(1) variates `torch.multinomial(probs[1, A], 1)` draws (aten: `q = empty_like(probs).exponential_(1)`,
    result = argmax(probs / q)) — `Categorical.sample()` of rollout.py:285-287."""
    return tc.empty(1, num_actions).exponential_(1, generator=gen)


def sample_actions(logits: np.ndarray, noise: np.ndarray) -> np.ndarray:
    """argmax_a softmax(logits)[a] / q[a] per row (policy_heads.py: Categorical(logits=...).sample())."""
    p = tc.softmax(tc.from_numpy(np.asarray(logits, F32)), dim=-1)
    return tc.argmax(p / tc.from_numpy(np.asarray(noise, F32)), dim=-1).numpy()


def frame_checksums(frames_u8: np.ndarray) -> np.ndarray:
    """One int64 per frame: position-weighted byte sum (order-sensitive)."""
    f = np.asarray(frames_u8, np.uint8).reshape(frames_u8.shape[0], -1).astype(np.int64)
    w = (np.arange(f.shape[1], dtype=np.int64) % 251) + 1
    return (f * w).sum(axis=1)


def golden_layout() -> Dict[str, str]:
    return {'r{k}_e{e}_actions': '[T+extra+1] int64', 'r{k}_e{e}_rew': '[T+extra] f32 (extrinsic)', 'r{k}_e{e}_raw': 'extrinsic_raw',
            'r{k}_e{e}_dones': '[T+extra] f32', 'r{k}_e{e}_obs_sum': '[T+extra+1] int64 checksums of rint(255 * frame)',
            'r{k}_e{e}_md_<key>': 'metadata lists'}


def rollouts_oracle(params: Dict[str, np.ndarray], e: int, T: int, extra_steps: int, num_actions: int, rollouts: int,
                    seed: int) -> List[dict]:
    """RolloutManager (rollout.py:227-357) + Rollout (:105-224) for environment `e`, restated: reset, draw a_0, then per step
    `env.step(a_t)`, record (o_t, a_t, r_t, d_t) at slot t, reset on done, draw a_{t+1}; `initial` pre-generates `extra_steps`
    steps; every report keeps T+extra+1 frames / actions and carries the last `extra_steps` steps to the front."""
    from oracle import ppo_oracle as po
    p = {k: tc.from_numpy(np.asarray(v)) for k, v in params.items()}
    gen = tc.Generator().manual_seed(seed)
    env = SynthActorEnv(e, num_actions, as_float=False)

    def choose(o_u8):
        with tc.no_grad():
            feat = po.trunk_forward(p, po.scale_obs(tc.from_numpy(o_u8[None])))
            logits, _ = po.heads_forward(p, feat, [])
        return int(sample_actions(logits.numpy(), exponential_noise(gen, num_actions).numpy())[0])

    steps = T + extra_steps
    obs = np.zeros((steps + 1, 84, 84, 4), np.uint8)
    act = np.zeros(steps + 1, np.int64)
    rew, raw, dones = np.zeros(steps, F32), np.zeros(steps, F32), np.zeros(steps, F32)
    md = {k: [] for k in ('ep_len', 'ep_ret', 'ep_len_raw', 'ep_ret_raw')}
    present = dict(ep_len=0, ep_ret=0., ep_len_raw=0, ep_ret_raw=0.)
    o_t = env.reset()
    a_t = choose(o_t)
    out = []

    def run(start, end):
        nonlocal o_t, a_t
        for t in range(start, end):
            o_tp1, r_t, done_t, info_t = env.step(a_t)
            obs[t], act[t], rew[t], raw[t], dones[t] = o_t, a_t, r_t['extrinsic'], r_t['extrinsic_raw'], float(done_t)
            present['ep_len'] += 1; present['ep_ret'] += r_t['extrinsic']
            present['ep_len_raw'] += 1; present['ep_ret_raw'] += r_t['extrinsic_raw']
            if done_t:
                for k in ('ep_len', 'ep_ret'):
                    md[k].append(present[k]); present[k] = type(present[k])(0)
                if (info_t['ale.lives'] == 0) if 'ale.lives' in info_t else True:
                    for k in ('ep_len_raw', 'ep_ret_raw'):
                        md[k].append(present[k]); present[k] = type(present[k])(0)
                o_tp1 = env.reset()
            o_t, a_t = o_tp1, choose(o_tp1)

    if extra_steps:
        run(0, extra_steps)
    for _ in range(rollouts):
        run(extra_steps, steps)
        obs[steps], act[steps] = o_t, a_t
        out.append(dict(observations=obs.copy(), actions=act.copy(), rewards=rew.copy(), raw=raw.copy(), dones=dones.copy(),
                        metadata={k: list(v) for k, v in md.items()}))
        for k in md:
            md[k] = []
        keep = (obs[T:steps].copy(), act[T:steps].copy(), rew[T:steps].copy(), raw[T:steps].copy(), dones[T:steps].copy())
        obs[:], act[:], rew[:], raw[:], dones[:] = 0, 0, 0, 0, 0
        if extra_steps:
            obs[:extra_steps], act[:extra_steps], rew[:extra_steps], raw[:extra_steps], dones[:extra_steps] = keep
    return out


class SingleFrameEnv(SynthActorEnv):
    """SynthActorEnv handing out SINGLE frames [84, 84, 1] (an environment chain without its FrameStackWrapper)."""

    def _frame(self):
        f = self._rs.randint(0, 256, size=(84, 84, 1)).astype(np.uint8)
        self.frames.append(f)
        return (f.astype(F32) / F32(255.0)) if self._as_float else f


class HostFrameStack:
    """FrameStackWrapper with num_stack frames (drl/envs/wrappers/stateless/frame_stack.py:50-59), restated: reset fills the
    window with copies of the first frame, step appends the newest one; the observation is their concatenation on the last
    axis, oldest first."""

    def __init__(self, env, num_stack: int = 4):
        self.env, self._n, self._frames = env, num_stack, []

    @property
    def reward_spec(self):
        return self.env.reward_spec

    def reset(self):
        o = self.env.reset()
        self._frames = [o] * self._n
        return np.concatenate(self._frames, axis=-1)

    def step(self, a):
        o, r, d, info = self.env.step(a)
        self._frames = self._frames[1:] + [o]
        return np.concatenate(self._frames, axis=-1), r, d, info
