"""Actor side of the learner path: the device-resident rollout ring and the batched rollout manager.

Reference: drl/algos/common/rollout.py — `MetadataManager` (:29-102), `Rollout` (:105-224, one environment, float32 host
tensors) and `RolloutManager` (:227-357, one batch-1 policy forward + `Categorical.sample()` per environment step).  Here one
process owns `n_envs` environments: frames go from pinned host memory straight into their slot of the uint8 store
[n_envs, Tp, 84, 84, 4] the learner trains on (`drl_rollout_record`), the policy is evaluated for all environments in one
call that reads those slots by row number (`drl_net_act`), and `report()` hands the filled store to the learner while the
actors continue in the second store of the ring (the `extra_steps` carry-over is one `drl_rollout_copy_steps` per array).

Stepping the environments themselves is host Python and stays whatever the caller's environments are: any object with
`reset() -> frame` and `step(a) -> (frame, reward_dict, done, info)`, the reference's wrapper protocol.
"""
import copy
import ctypes
from typing import Any, Callable, Dict, List, Mapping, Optional, Sequence

import numpy as np
import torch as tc

from .. import _lib
from .._lib import check
from ..ops import current_stream, ptr
from .ppo import RolloutStorage

OBS_SHAPE = (84, 84, 4)
INTRINSIC_KEY = 'intrinsic_rnd'       # random_network_distillation.py:131


class MetadataManager:
    """rollout.py:29-102 (running episode statistics: `present` accumulates, `present_done` moves it to `pasts`)."""

    def __init__(self, defaults: Mapping[str, Any]):
        self._defaults = dict(defaults)
        self._present = copy.deepcopy(self._defaults)
        self._pasts = {k: [] for k in self._defaults}

    @property
    def keys(self):
        return self._defaults.keys()

    @property
    def present(self) -> Dict[str, Any]:
        return self._present

    @property
    def pasts(self) -> Dict[str, List[Any]]:
        return self._pasts

    def update_present(self, deltas: Mapping[str, Any]) -> None:
        assert set(deltas.keys()) <= set(self._defaults.keys())
        for k, v in deltas.items():
            self._present[k] += v

    def present_done(self, *keys: str) -> None:
        for k in keys:
            self._pasts[k].append(self._present[k])
            self._present[k] = copy.deepcopy(self._defaults[k])

    def past_done(self) -> None:
        self._pasts = {k: [] for k in self._defaults}


def _frames_u8(frames) -> np.ndarray:
    """uint8 frames as they are; float frames in [0, 1] (DeepmindWrapper scale=True, deepmind.py:57-59) back to 0..255."""
    a = np.asarray(frames)
    if a.dtype == np.uint8:
        return a
    return np.rint(a.astype(np.float32) * 255.0).clip(0, 255).astype(np.uint8)


class DeviceRollout:
    """`Rollout` (rollout.py:105-224) for `n_envs` environments on the device: two `RolloutStorage`s used alternately, so
    that the rollout `report()` returned stays valid (and is trained on) while the next one is being recorded."""

    def __init__(self, n_envs: int, rollout_len: int, extra_steps: int, reward_keys: Sequence[str], device=None):
        if not tc.cuda.is_available():
            raise RuntimeError('DeviceRollout needs a CUDA device; there is no CPU fallback')
        self.device = tc.device('cuda', tc.cuda.current_device()) if device is None else tc.device(device)
        self._L = _lib.lib()
        self.n_envs, self.T, self.extra_steps = int(n_envs), int(rollout_len), int(extra_steps)
        self.timesteps = self.T + self.extra_steps
        self.reward_keys = list(reward_keys)
        if len(self.reward_keys) > 4:
            raise ValueError('at most 4 reward streams')
        self._ring = [RolloutStorage(n_envs, rollout_len, self.reward_keys, self.device, extra_steps) for _ in range(2)]
        self._cur = 0
        self.Tp = self._ring[0].Tp
        # pinned staging for one time step of host data: frames + the scalars of a step, read by the device directly
        self._h_obs = tc.empty((n_envs,) + OBS_SHAPE, dtype=tc.uint8).pin_memory()
        self._h_act = tc.empty(n_envs, dtype=tc.int64).pin_memory()
        self._h_done = tc.empty(n_envs, dtype=tc.float32).pin_memory()
        self._h_rew = {k: tc.empty(n_envs, dtype=tc.float32).pin_memory() for k in self.reward_keys}
        self._h_obs_np = self._h_obs.numpy()
        self._h_new = tc.empty((n_envs, 84, 84), dtype=tc.uint8).pin_memory()        # single newest frames (stack_frames)
        self._h_mask = tc.empty(n_envs, dtype=tc.uint8).pin_memory()
        self._d_new = tc.empty((n_envs, 84, 84), dtype=tc.uint8, device=self.device)  # the copy engine moves them at full PCIe rate
        self._staged = tc.cuda.Event()
        self._staged.record()

    @property
    def storage(self) -> RolloutStorage:
        return self._ring[self._cur]

    def staging_frames(self) -> np.ndarray:
        """The pinned [n_envs, 84, 84, 4] uint8 staging buffer, free to be overwritten: environments that render into it and
        then call `write_frames(t, staging_frames_array)` save the host-side copy of the frames."""
        self._staged.synchronize()
        return self._h_obs_np

    def rows_at(self, t: int) -> tc.Tensor:
        """Flat row numbers of time step `t` of every environment (what the kernels take as `row_idx`)."""
        return tc.arange(self.n_envs, device=self.device, dtype=tc.int64) * self.Tp + int(t)

    def _record(self, t: int, obs=None, actions=None, rewards: Optional[Mapping[str, Any]] = None, dones=None) -> None:
        d = self.storage.data
        self._staged.synchronize()                     # the previous step's staged values have been consumed
        obs_p, on_host = None, 0
        if obs is not None:
            if isinstance(obs, tc.Tensor) and obs.is_cuda:
                if obs.dtype != tc.uint8 or tuple(obs.shape) != (self.n_envs,) + OBS_SHAPE or not obs.is_contiguous():
                    raise ValueError('device frames must be a contiguous uint8 [n_envs, 84, 84, 4] tensor')
                obs_p = ptr(obs)
            else:
                if obs is not self._h_obs_np:           # else: rendered in place (staging_frames())
                    a = _frames_u8(obs.numpy() if isinstance(obs, tc.Tensor) else obs)
                    if a.shape != (self.n_envs,) + OBS_SHAPE:
                        raise ValueError(f'frames must be [n_envs, 84, 84, 4], got {a.shape}')
                    self._h_obs_np[...] = a
                obs_p, on_host = ptr(self._h_obs), 1

        def stage(buf: tc.Tensor, value):
            if value is None:
                return None
            if isinstance(value, tc.Tensor) and value.is_cuda:
                if value.dtype != buf.dtype or value.numel() != self.n_envs or not value.is_contiguous():
                    raise ValueError(f'device values must be contiguous {buf.dtype} [n_envs]')
                return ptr(value)
            buf.numpy()[...] = np.asarray(value.cpu() if isinstance(value, tc.Tensor) else value).reshape(self.n_envs)
            return ptr(buf)

        a_p, d_p = stage(self._h_act, actions), stage(self._h_done, dones)
        K = len(self.reward_keys)
        stores = (ctypes.c_void_p * max(K, 1))(*[ptr(d['rewards'][k]) for k in self.reward_keys])
        r_ps = None
        if rewards is not None:
            missing = [k for k in self.reward_keys if k not in rewards]
            if missing:
                raise KeyError(f'reward keys {missing} missing from the step reward')
            r_ps = (ctypes.c_void_p * max(K, 1))(*[stage(self._h_rew[k], rewards[k]) for k in self.reward_keys])
        check(self._L.drl_rollout_record(ptr(d['observations']), ptr(d['actions']), ptr(d['dones']), stores, K, self.n_envs,
                                         self.Tp, int(t), obs_p, a_p, d_p, r_ps, on_host, current_stream()), 'rollout_record')
        self._staged.record()

    # ---- the reference's Rollout interface, batched over the environments of this rank ---------------------------
    def record(self, t: int, o_t, a_t, r_t: Mapping[str, Any], d_t) -> None:
        """rollout.py:159-179 for every environment at once: o_t [n_envs,84,84,4], a_t [n_envs], r_t {key: [n_envs]}, d_t [n_envs]."""
        assert 0 <= t <= self.timesteps - 1
        self._record(t, o_t, a_t, r_t, d_t)

    def record_nexts(self, o_Tp1, a_Tp1) -> None:
        """rollout.py:181-195: the observation and action after the last step, into the rightmost slot."""
        self._record(self.timesteps, o_Tp1, a_Tp1)

    # ---- piecewise writers used by BatchedRolloutManager (a frame is uploaded once, into its slot) -----------------
    def write_frames(self, t: int, frames) -> None:
        self._record(t, obs=frames)

    def write_actions(self, t: int, actions) -> None:
        self._record(t, actions=actions)

    def write_outcome(self, t: int, rewards: Mapping[str, Any], dones) -> None:
        self._record(t, rewards=rewards, dones=dones)

    def staging_newest(self):
        """Pinned ([n_envs, 84, 84] uint8 newest frames, [n_envs] uint8 modes) for `stack_frames`, free to be overwritten."""
        self._staged.synchronize()
        return self._h_new.numpy(), self._h_mask.numpy()

    def stack_frames(self, t_src: int, t_dst: int, newest=None, modes=None) -> None:
        """FrameStackWrapper (frame_stack.py:50-59, num_stack = 4) on the device: slot `t_dst` = slot `t_src` shifted by one
        channel + `newest` [n_envs, 84, 84] as the last channel; `modes` [n_envs]: 1 = the environment was reset (four copies of
        its newest frame), 2 = leave it untouched.  None = the pinned buffers of `staging_newest()` were filled in place."""
        new_np, mask_np = self.staging_newest()
        if newest is not None and newest is not new_np:
            if isinstance(newest, tc.Tensor) and newest.is_cuda:
                raise ValueError('stack_frames takes host frames (the environments render on the host)')
            new_np[...] = _frames_u8(newest).reshape(self.n_envs, 84, 84)
        if modes is not None and modes is not mask_np:
            mask_np[...] = np.asarray(modes, dtype=np.uint8).reshape(self.n_envs)
        elif modes is None and newest is not None:
            mask_np[...] = 0
        d = self.storage.data
        self._d_new.copy_(self._h_new, non_blocking=True)
        check(self._L.drl_rollout_stack_frames(ptr(d['observations']), self.n_envs, self.Tp, int(t_src), int(t_dst), ptr(self._d_new),
                                               ptr(self._h_mask), current_stream()), 'rollout_stack_frames')
        self._staged.record()

    def report(self) -> Dict[str, Any]:
        """rollout.py:197-224: returns the filled store (a `RolloutStorage` dict: uint8 frames and per-step scalars laid out
        [n_envs, Tp, ...]) and continues in the other store of the ring, whose first `extra_steps` steps receive the last
        `extra_steps` steps of this one.  The returned tensors stay untouched until the `report()` after next."""
        done, nxt = self._ring[self._cur], self._ring[1 - self._cur]
        if self.extra_steps > 0:
            st = current_stream()
            src, dst, n = done.data, nxt.data, self.extra_steps

            def carry(a: tc.Tensor, b: tc.Tensor, step_bytes: int):
                check(self._L.drl_rollout_copy_steps(ptr(b), ptr(a), self.n_envs, self.Tp, self.Tp, step_bytes, 0, self.T, n, st),
                      'rollout_copy_steps')
            carry(src['observations'], dst['observations'], 84 * 84 * 4)
            carry(src['actions'], dst['actions'], 8)
            carry(src['dones'], dst['dones'], 4)
            for k in self.reward_keys:
                carry(src['rewards'][k], dst['rewards'][k], 4)
        self._cur = 1 - self._cur
        return done.as_dict()


class BatchedRolloutManager:
    """`RolloutManager` (rollout.py:227-357) over a list of environments: one `drl_net_act` per time step for all of them.

    envs         list of environments (reset() -> frame; step(a) -> (frame, reward dict, done, info)); frames are uint8
                 [84,84,4] (DeepmindWrapper scale=False) or float in [0,1]
    rollout_net  a fused `agents.Agent` (or its `DDPShim`)
    noise        optional callable (step_number) -> float32 [n_envs, A] tensor of Exp(1) variates for the draw (the variates
                 `torch.multinomial` would have drawn); default: generated in the kernel from (`seed`, step number)
    intrinsic    optional `envs.RandomNetworkDistillation`: adds reward key 'intrinsic_rnd' (:131) from the new frames, batched
                 (random_network_distillation.py:186-201)
    reward_normalizers  optional {reward key: envs.BatchedRewardNormalizer} applied after that (normalize_reward.py:171-194)
    frame_stack  True: the environments return SINGLE frames ([84,84] or [84,84,1]; no FrameStackWrapper in their chain) and
                 the 4-frame stacks are built on the device (frame_stack.py:50-59): 7 KB per environment step cross PCIe
                 instead of 28 KB
    """

    def __init__(self, envs: Sequence[Any], rollout_net, rollout_len: int, extra_steps: int,
                 reward_keys: Optional[Sequence[str]] = None, noise: Optional[Callable[[int], tc.Tensor]] = None, seed: int = 0,
                 intrinsic=None, reward_normalizers: Optional[Mapping[str, Any]] = None, device=None, frame_stack: bool = False):
        assert rollout_len > 0
        assert extra_steps >= 0
        self._envs = list(envs)
        self._net = getattr(rollout_net, 'module', rollout_net)
        self._engine = self._net._require_engine()
        self._rollout_len, self._extra_steps = rollout_len, extra_steps
        self._noise, self._seed, self._step_no = noise, int(seed), 0
        self._intrinsic, self._reward_normalizers = intrinsic, dict(reward_normalizers or {})
        self._frame_stack = bool(frame_stack)
        self._L = _lib.lib()
        n = len(self._envs)
        if n > self._engine.max_batch:
            raise ValueError(f'{n} environments but the engine was fused for max_batch={self._engine.max_batch}')
        keys = list(reward_keys) if reward_keys is not None else self._get_reward_keys()
        self._stored_keys = [k for k in keys if k != 'extrinsic_raw']      # what the learner's RolloutStorage keeps
        self._rollout = DeviceRollout(n, rollout_len, extra_steps, self._stored_keys, device or self._engine.device)
        self._actions = tc.empty(n, dtype=tc.int64, device=self._rollout.device)
        self._h_actions = tc.empty(n, dtype=tc.int64).pin_memory()
        self._metadata = [MetadataManager({'ep_len': 0, 'ep_ret': 0., 'ep_len_raw': 0, 'ep_ret_raw': 0.}) for _ in range(n)]
        # o_0 into slot 0 and a_0 for it (rollout.py:253-254)
        self._t = 0
        first = np.stack([_frames_u8(e.reset()) for e in self._envs])
        if self._frame_stack:
            self._rollout.stack_frames(0, 0, first, np.ones(n, np.uint8))
        else:
            self._rollout.write_frames(0, first)
        self._a_t = self._choose_actions(0)
        self.generate(initial=True)

    def _get_reward_keys(self) -> List[str]:
        spec = getattr(self._envs[0], 'reward_spec', None)
        if spec is None:                                                     # rollout.py:271-276
            raise ValueError('Reward spec cannot be None.\nWrap environment with RewardToDict wrapper to fix.')
        keys = list(spec.keys)
        if self._intrinsic is not None and INTRINSIC_KEY not in keys:
            keys.append(INTRINSIC_KEY)
        return keys

    @property
    def rollout(self) -> DeviceRollout:
        return self._rollout

    def _choose_actions(self, t: int) -> np.ndarray:
        """rollout.py:280-288 for all environments: policy forward on the frames of slot `t` + one draw each; the actions
        are written to slot `t` of the store and returned to the host for `env.step`."""
        r, d = self._rollout, self._rollout.storage.data
        noise = None
        if self._noise is not None:
            noise = self._noise(self._step_no).to(r.device, tc.float32).contiguous()
            if tuple(noise.shape) != (r.n_envs, self._engine.A):
                raise ValueError('noise must be [n_envs, num_actions]')
        st = current_stream()
        check(self._L.drl_net_act(self._engine._h, ptr(d['observations']), ptr(r.rows_at(t)), r.n_envs, ptr(noise), self._seed,
                                  self._step_no, ptr(self._actions), None, None, None, st), 'net_act')
        check(self._L.drl_rollout_record(None, ptr(d['actions']), None, None, 0, r.n_envs, r.Tp, int(t), None, ptr(self._actions),
                                         None, None, 0, st), 'rollout_record')
        self._h_actions.copy_(self._actions, non_blocking=True)
        tc.cuda.current_stream().synchronize()
        self._step_no += 1
        return self._h_actions.numpy().copy()

    @tc.no_grad()
    def generate(self, initial: bool = False) -> Optional[Dict[str, Any]]:
        """rollout.py:294-357: steps every environment `rollout_len` times (`extra_steps` times when `initial`), returns the
        device rollout dict + 'metadata' (the episode statistics of all environments, concatenated in environment order)."""
        if initial:
            if self._extra_steps == 0:
                return None
            start_t, end_t = 0, self._extra_steps
        else:
            start_t, end_t = self._extra_steps, self._rollout_len + self._extra_steps
        r, n = self._rollout, len(self._envs)
        for t in range(start_t, end_t):
            # pinned staging: each frame is copied once on the host, once to the device
            if self._frame_stack:
                frames, modes = r.staging_newest()
                modes[...] = 0
            else:
                frames, modes = r.staging_frames(), None
            rewards = {k: np.zeros(n, dtype=np.float32) for k in self._stored_keys}
            dones = np.zeros(n, dtype=np.float32)
            resets: List[Any] = []
            for e, env in enumerate(self._envs):
                o_tp1, r_t, done_t, info_t = env.step(self._a_t[e])
                md = self._metadata[e]
                md.update_present({'ep_len': 1, 'ep_ret': r_t['extrinsic'], 'ep_len_raw': 1, 'ep_ret_raw': r_t['extrinsic_raw']})
                if done_t:
                    md.present_done('ep_len', 'ep_ret')
                    was_real_done = info_t['ale.lives'] == 0 if 'ale.lives' in info_t else True   # rollout.py:330-336
                    if was_real_done:
                        md.present_done('ep_len_raw', 'ep_ret_raw')
                    if self._intrinsic is not None:           # the intrinsic reward is taken on the frame step() returned
                        resets.append((e, _frames_u8(env.reset())))
                    else:
                        o_tp1 = env.reset()
                        if modes is not None:
                            modes[e] = 1
                frames[e] = _frames_u8(o_tp1).reshape(frames.shape[1:])
                dones[e] = float(done_t)
                for k in self._stored_keys:
                    if k in r_t:
                        rewards[k][e] = float(r_t[k])
            if self._frame_stack:
                r.stack_frames(t, t + 1)                        # o_{t+1} = shift(o_t) + newest frame, built in its slot
            else:
                r.write_frames(t + 1, frames)                   # o_{t+1}: uploaded once, straight into its slot
            step_rewards: Dict[str, Any] = dict(rewards)
            if self._intrinsic is not None:                     # on the NEW frames, like the wrapper's step()
                obs = r.storage.data['observations']
                rows = r.rows_at(t + 1)
                step_rewards[INTRINSIC_KEY] = self._intrinsic.intrinsic_reward(obs, rows)
                self._intrinsic.observe(obs.view(-1, 84, 84, 4)[rows][..., -1:].float() * float(np.float32(1.0 / 255.0)))  # :191-192
                if resets and self._frame_stack:                # then the finished environments continue from their reset frame
                    frames, modes = r.staging_newest()
                    modes[...] = 2
                    for e, f in resets:
                        frames[e], modes[e] = f.reshape(84, 84), 1
                    r.stack_frames(t + 1, t + 1)
                elif resets:
                    idx = tc.as_tensor([e for e, _ in resets], device=r.device, dtype=tc.int64)
                    obs.view(-1, 84, 84, 4)[rows[idx]] = tc.from_numpy(np.stack([f for _, f in resets])).to(r.device)
            for k, nz in self._reward_normalizers.items():
                raw = tc.as_tensor(step_rewards[k], device=r.device, dtype=tc.float32).reshape(n, 1)
                step_rewards[k] = nz.step_block(raw, tc.as_tensor(dones, device=r.device).reshape(n, 1)).reshape(n).contiguous()
            r.write_outcome(t, step_rewards, dones)
            self._a_t = self._choose_actions(t + 1)             # a_{t+1} into slot t+1 (the last one is record_nexts)
        if initial:
            return None
        results = dict(r.report())
        results['metadata'] = {k: [v for md in self._metadata for v in md.pasts[k]] for k in self._metadata[0].keys}
        for md in self._metadata:
            md.past_done()
        # the frame / action of slot `timesteps` open the next rollout at slot `extra_steps` (rollout.py:341-342 keeps
        # o_t / a_t across calls; report() already carried slots [T, T+extra_steps) over)
        prev, cur = results, r.storage.data
        for name, nbytes in (('observations', 84 * 84 * 4), ('actions', 8)):
            check(self._L.drl_rollout_copy_steps(ptr(cur[name]), ptr(prev[name]), n, r.Tp, r.Tp, nbytes, self._extra_steps,
                                                 r.timesteps, 1, current_stream()), 'rollout_copy_steps')
        return results
