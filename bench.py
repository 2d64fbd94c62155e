#!/usr/bin/env python
"""bench.py — PPO update samples/sec on synthetic 84x84x4 rollouts (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path on the host cores

A "step" is one learner step of PPO.optimize (drl/algos/ppo.py:246-257): gather a minibatch of
B samples from each env of the GPU, NatureCNN forward, fused losses, backward, gradient
all-reduce across GPUs, Adam.  Workload at N=1: BASELINE config[1] — 1024 envs x 128 steps per
GPU, B=32 per env => 32 768 samples per step per GPU (weak scaling: every GPU owns 1024 envs).
Rank 0 prints ONE JSON line:
  value      device-resident learner steps (frames already in HBM), CUDA events, max over ranks
  sustained  the same steps for ~1 s (power-capped clocks) — the training rate; `roofline.step` is computed from it
  e2e        whole `PPO.optimize_host` calls: the uint8 rollout is uploaded from pinned host memory once per call
             (chunked, overlapped with the annotate forward), then GAE, opt_epochs x T/B learner steps, the per-epoch
             metrics passes and the device-to-host read of the metrics — the same work the reference arm times
  roofline   dominant kernel (per-launch CUDA events inside the timed region) + `step`: whole-step tensor roofline
  extras     (N=1) the other BASELINE configs and side measurements: config0 (8 envs), config3 (RND + PPO), sweep
             (4k..64k samples per step), optimize_call, precision modes, GAE / loss / sum-tree
  allreduce_check (N>1) the bucketed gradient all-reduce and the 5-float metrics mean verified against all-gather + sum
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch as tc  # noqa: E402

A, T, B = 6, 128, 32
REWARD_KEYS = ['extrinsic']
MACS = {  # algorithmic multiply-accumulates per sample
    'conv1_fwd': 3276800, 'conv2_fwd': 2654208, 'conv3_fwd': 1806336, 'fc_fwd': 1605632,
    'conv1_wgrad': 3276800, 'conv2_wgrad': 2654208, 'conv3_wgrad': 1806336, 'fc_wgrad': 1605632,
    'conv2_dgrad': 2654208, 'conv3_dgrad': 1806336, 'fc_dgrad': 1605632,
}
BYTES = {  # algorithmic HBM bytes per sample: operands read once + result written once
    'conv1_fwd': 28224 + 25600, 'conv1_wgrad': 28224 + 25600,
    'conv2_fwd': 25600 + 10368, 'conv2_dgrad': 10368 + 25600, 'conv2_wgrad': 25600 + 10368,
    'conv3_fwd': 10368 + 6272, 'conv3_dgrad': 6272 + 10368, 'conv3_wgrad': 10368 + 6272,
    'fc_fwd': 6272 + 2048, 'fc_dgrad': 1024 + 6272, 'fc_wgrad': 6272 + 1024,
}
STEP_FLOPS_PER_SAMPLE = 49525760          # fwd 18.69 + bwd 30.83 MFLOP (A=6, K=1), SURVEY §8d
RND_FLOPS_PER_SAMPLE = 16.61e6            # teacher fwd + student fwd + student bwd, SURVEY §8d


def peaks():
    """Roofline denominators: the driver-written MEASURED_PEAKS.json (copy bandwidth; cuBLAS bf16 GEMM as a burst at
    full clocks and sustained under the power cap), else the fallback numbers of B200_PROFILING.md."""
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        burst = float(d.get('bf16_tflops', 1590.0))
        return dict(tflops_burst=burst, tflops_sustained=float(d.get('bf16_tflops_sustained', burst)),
                    hbm=float(d.get('hbm_gbs', 6650.0)), src='measured (MEASURED_PEAKS.json)')
    return dict(tflops_burst=1590.0, tflops_sustained=1400.0, hbm=6650.0, src='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""

    def __init__(self, gpu_index: int):
        self.rows, self.proc = [], None
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits',
                                          '-lms', '20', '-i', str(gpu_index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(',')]))

    def wait_ready(self, timeout=3.0):
        """nvidia-smi needs 0.1-0.3 s to start: block until its first sample so that it is already looping when the
        timed region begins."""
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.01)

    def in_window(self, t0, t1):
        return [r for (t, r) in list(self.rows) if t0 <= t <= t1]

    def stop(self, t0, t1, load_windows=()):
        """Samples inside the timed region [t0, t1]; when the region is shorter than nvidia-smi's sampling period the
        samples of `load_windows` (the same learner steps running untimed right before / after it) are used and the
        result says so."""
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.05)
        self.proc.terminate()
        rows, window = self.in_window(t0, t1), 'timed region'
        if len(rows) < 2:
            for (w0, w1) in load_windows:
                rows = rows + self.in_window(w0, w1)
            window = 'timed region + the same steps running untimed around it'
        sm = [float(r[0]) for r in rows if r and r[0].replace('.', '', 1).isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace('.', '', 1).isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(rows), 'window': window}


# ------------------------------------------------------------------------------------------------------
# CPU side: the reference's own implementation (oracle/_ref, vendored at build time) or the oracle port
# ------------------------------------------------------------------------------------------------------
def cpu_port_step_rate(n_envs: int, seconds: float, warmup: int = 1, max_steps: int = 2000, steps=None):
    """Times the CPU restatement of the reference's learner step (oracle port: torch-CPU fp32 autograd
    through the same conv/linear/Adam ops the reference calls, all host threads) on a bounded sample:
    `n_envs` ranks x B samples per step.  Returns (samples/s, steps, threads, seconds)."""
    from oracle import ppo_oracle as po
    tc.set_num_threads(max(1, os.cpu_count() or 1))
    rs = np.random.RandomState(0)
    params = po.init_params(A, REWARD_KEYS, seed=11)
    hp = po.Hyper({'extrinsic': 1.0}, 0.2, 0.01, 1.0, False, True)
    mbs = []
    for e in range(n_envs):
        mbs.append(dict(observations=rs.randint(0, 256, size=(B, 84, 84, 4), dtype=np.uint8),
                        actions=rs.randint(0, A, size=B).astype(np.int64),
                        logprobs=np.full(B, -np.log(A), np.float32),
                        advantages={'extrinsic': rs.standard_normal(B).astype(np.float32)},
                        vpreds={'extrinsic': np.zeros(B, np.float32)},
                        vtargs={'extrinsic': rs.standard_normal(B).astype(np.float32)}))
    keys = list(params.keys())
    m = {k: np.zeros_like(params[k]) for k in keys}
    v = {k: np.zeros_like(params[k]) for k in keys}

    def one(step):
        _, grads = po.minibatch_losses_and_grads(params, mbs, hp)
        for k in keys:
            po.adam_step(params[k], grads[k], m[k], v[k], step, 1e-3, 0.9, 0.999, 1e-5)

    for i in range(warmup):
        one(i + 1)
    t0 = time.perf_counter()
    n = 0
    while True:
        one(warmup + n + 1)
        n += 1
        el = time.perf_counter() - t0
        if steps is not None:
            if n >= steps:
                break
        elif el >= seconds or n >= max_steps:
            break
    return n * n_envs * B / el, n, tc.get_num_threads(), el


def cpu_baseline(seconds: float, warmup: int = 1, steps=None):
    """The reference's CPU path on this host: the UNMODIFIED reference when a copy travels with the repository (oracle/_ref),
    else the oracle port.  Returns the `cpu_baseline` object of the JSON line."""
    from oracle import ref_bench
    if ref_bench.reference_root() is not None:
        n = steps if steps is not None else max(2, int(seconds / 1.0))
        port = 20000 + (os.getpid() % 20000)
        r = ref_bench.run(warmup, n, T=T, B=B, A=A, epochs=3, port=port)
        return {'value': r['rate'], 'unit': 'samples/s', 'cores': r['cores'], 'kind': 'reference', 'seconds': r['seconds'],
                'steps': r['steps'],
                'sample': f"{r['steps']} x (annotate + credit_assignment + PPO.optimize) of the unmodified reference: {r['world']} gloo "
                          f"ranks (envs) x {r['threads_per_rank']} threads, T={T}, B={B}, 3 epochs incl. the per-epoch metrics "
                          f"passes = {r['samples_per_step']} update samples per step; os.cpu_count={os.cpu_count()}, torch {r['torch']}"}
    rate, n, threads, el = cpu_port_step_rate(8, seconds, warmup=warmup, steps=steps)
    return {'value': rate, 'unit': 'samples/s', 'cores': threads, 'kind': 'port', 'seconds': el, 'steps': n,
            'sample': f'{n} learner steps of 8 envs x {B} samples (oracle port of the reference CPU path, torch-CPU fp32, '
                      f'os.cpu_count={os.cpu_count()}; no copy of the reference under oracle/_ref)'}


def run_reference(args, rank):
    """`--impl reference`: rank 0 times the reference's own CPU implementation with all host threads; other ranks exit."""
    if rank != 0:
        return
    cb = cpu_baseline(0.0, warmup=max(1, min(args.warmup, 2)), steps=max(1, args.steps))
    line = {
        'impl': 'reference', 'metric': 'PPO update samples/sec (84x84x4 frames)', 'value': cb['value'], 'unit': 'samples/s',
        'n_gpus': args.gpus, 'steps': cb['steps'], 'warmup': max(1, min(args.warmup, 2)), 'ms_per_step': 1e3 * cb['seconds'] / cb['steps'],
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'PPO NatureCNN synthetic 84x84x4 rollout, 1024 envs x 128 steps, B=32 per env',
                   'sample': cb['sample']},
        'cpu_baseline': {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
        'e2e': {'value': cb['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------------
def make_agent(num_actions, reward_keys, max_batch, precision, dev):
    from pytorch_drl_b200 import agents
    o2, o01 = (lambda w: tc.nn.init.orthogonal_(w, gain=2 ** 0.5)), (lambda w: tc.nn.init.orthogonal_(w, gain=0.01))
    preds = {'policy': agents.CategoricalPolicyHead(512, num_actions, w_init=o01, b_init=tc.nn.init.zeros_)}
    for k in reward_keys:
        preds[f'value_{k}'] = agents.SimpleValueHead(512, w_init=o01, b_init=tc.nn.init.zeros_)
    agent = agents.Agent([agents.ToChannelMajor()], agents.NatureCNN(4, o2, tc.nn.init.zeros_), preds)
    agent.fuse(max_batch=max_batch, precision=precision, device=dev)
    return agent


def fill_synthetic(d, st, n_envs, num_actions, reward_keys, dev, seed, obs=True):
    """SURVEY §8d synthetic inputs, generated on the device."""
    g = tc.Generator(device=dev)
    g.manual_seed(seed)
    if obs:
        for e0 in range(0, n_envs, 64):
            e1 = min(n_envs, e0 + 64)
            d['observations'][e0:e1] = tc.randint(0, 256, (e1 - e0, st.Tp, 84, 84, 4), device=dev, dtype=tc.uint8, generator=g)
    d['actions'][:] = tc.randint(0, num_actions, (n_envs, st.Tp), device=dev, generator=g)
    for k in reward_keys:
        if k == 'extrinsic':
            d['rewards'][k][:] = tc.randint(-1, 2, (n_envs, st.Tp), device=dev, generator=g).float()
        else:
            d['rewards'][k][:] = tc.rand((n_envs, st.Tp), device=dev, generator=g)
    d['dones'][:] = (tc.rand((n_envs, st.Tp), device=dev, generator=g) < 0.01).float()


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    tc.cuda.synchronize()
    e0, e1 = tc.cuda.Event(enable_timing=True), tc.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    tc.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


class SiteProfile:
    """Per-launch-site CUDA-event times recorded by the library (drl_profile_*)."""

    def __init__(self, L):
        self.L = L
        self.nbuf = ctypes.create_string_buffer(48 * 96)
        self.ms, self.cnt, self.ns = (ctypes.c_float * 96)(), (ctypes.c_int * 96)(), ctypes.c_int()

    def start(self):
        self.L.drl_profile_enable(1)
        self.L.drl_profile_read(96, self.nbuf, self.ms, self.cnt, ctypes.byref(self.ns))

    def stop(self):
        self.L.drl_profile_read(96, self.nbuf, self.ms, self.cnt, ctypes.byref(self.ns))
        self.L.drl_profile_enable(0)
        return {self.nbuf.raw[48 * i:48 * i + 48].split(b'\0')[0].decode(): (self.ms[i] / max(1, self.cnt[i]), self.cnt[i])
                for i in range(self.ns.value)}


# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--envs', type=int, default=1024, help='envs per GPU')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-sustained', action='store_true', help='skip the ~1 s power-capped leg')
    ap.add_argument('--no-extras', action='store_true', help='skip the other configs / side measurements')
    ap.add_argument('--debug-flags', default='', help='kernel-variant experiments, e.g. "6=1,7=3" (drl_debug_set)')
    ap.add_argument('--comm', default='nccl', choices=['p2p-overlap', 'p2p', 'p2p-1', 'nccl-overlap', 'nccl', 'nccl-1'],
                    help='gradient exchange variant at N > 1: NCCL or NVLink peer-memory kernels, after the backward pass (2 buckets / -1: one) or overlapped with it')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if world != args.gpus and world > 1:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}')
    assert args.warmup >= 3 or args.steps <= 2, 'timing rule: W >= 3'

    import torch.distributed as dist
    from pytorch_drl_b200 import _lib, algos
    from pytorch_drl_b200.comm import Communicator
    from pytorch_drl_b200.utils.nested import MinibatchView
    tc.cuda.set_device(local_rank)
    dev = tc.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    L = _lib.lib()
    for kv in filter(None, args.debug_flags.split(',')):
        k, v = kv.split('=')
        L.drl_debug_set(int(k), int(v))
    n_envs = args.envs
    nb = n_envs * B
    extras_on = world == 1 and not args.no_extras
    max_batch = max(nb, 65536) if extras_on else nb         # the sweep goes up to 64k samples per step

    # ---- model + algo through the public drop-in API ----
    tc.manual_seed(0)
    agent = make_agent(A, REWARD_KEYS, max_batch, args.precision, dev)
    opt = tc.optim.Adam(agent.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-5)
    comm = None
    if world > 1:
        comm = Communicator(rank, world, local_rank, p2p_elems=(1 << 22) if args.comm.startswith('p2p') else 0)
        comm.overlap = args.comm.endswith('overlap')
        comm.single_bucket = args.comm.endswith('-1')
    algo = algos.PPO(rank=rank, world_size=world, rollout_len=T, extra_steps=0,
                     credit_assignment_ops={'extrinsic': algos.GAE(gamma=0.99, use_dones=True, lambda_=0.95)},
                     global_step=1, max_steps=10 ** 7, policy_net=agent, policy_optimizer=opt, standardize_adv=True,
                     opt_epochs=3, learner_batch_size=B, clip_param_init=0.2, clip_param_final=0.0,
                     ent_coef_init=0.01, ent_coef_final=0.01, vf_loss_coef=1.0, envs_per_rank=n_envs, comm=comm,
                     sampler_seed=0)

    # ---- synthetic rollout, generated on the device; annotate + GAE through the API ----
    st = algos.RolloutStorage(n_envs, T, REWARD_KEYS, dev)
    d = st.as_dict()
    fill_synthetic(d, st, n_envs, A, REWARD_KEYS, dev, 10000 * rank)
    rollout = algo.credit_assignment(algo.annotate(d))
    tc.cuda.synchronize()

    allreduce_check = run_allreduce_check(comm, agent.engine, dev, rank, world) if world > 1 else None

    rs = np.random.RandomState(1234 + rank)

    def step_rows():
        blocks = [rs.permutation(T)[:B] for _ in range(n_envs)]
        return algo._minibatch_rows(rollout, blocks)

    total_steps = args.warmup + args.steps
    rows_list = [step_rows() for _ in range(total_steps)]
    hyper, batch = algo._hyper(), algo._batch(rollout)

    def barrier():
        if world > 1:
            dist.barrier()
        tc.cuda.synchronize()

    def allmax(x):
        if world > 1:
            t = tc.tensor([x], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    # ---- device-resident timing (value) ----
    clocks = ClockSampler(local_rank) if rank == 0 else None     # started early: nvidia-smi needs ~0.2 s to its first sample
    if clocks is not None:
        clocks.wait_ready()
    t_warm0 = time.time()
    for i in range(args.warmup):
        algo.learner_step(MinibatchView(rollout, rows_list[i]), hyper, batch)
    barrier()
    prof = SiteProfile(L)
    prof.start()
    launches0 = L.drl_launch_count()
    ev0, ev1 = tc.cuda.Event(enable_timing=True), tc.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record()
    for i in range(args.steps):
        algo.learner_step(MinibatchView(rollout, rows_list[args.warmup + i]), hyper, batch)
    ev1.record()
    barrier()
    t_wall1 = time.time()
    elapsed_ms = allmax(ev0.elapsed_time(ev1))
    launches = L.drl_launch_count() - launches0
    sites = prof.stop()                                            # per-launch-site CUDA-event times of the timed region
    ms_per_step = elapsed_ms / args.steps
    value = nb * world / (ms_per_step * 1e-3)
    clk = clocks.stop(t_wall0, t_wall1, [(t_warm0, t_wall0)]) if clocks else None

    # ---- the same steps for ~1 s: the board reaches its power cap and lowers the SM clock (MEASURED_PEAKS.json's
    #      "sustained" regime): the training rate ----
    sustained = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(1000.0 / ms_per_step))       # same count on every rank (ms_per_step is the max over ranks)
        clocks2 = ClockSampler(local_rank) if rank == 0 else None
        if clocks2 is not None:
            clocks2.wait_ready()
        barrier()
        s0, s1 = tc.cuda.Event(enable_timing=True), tc.cuda.Event(enable_timing=True)
        tw0 = time.time()
        s0.record()
        for i in range(n_sus):
            algo.learner_step(MinibatchView(rollout, rows_list[args.warmup + i % args.steps]), hyper, batch)
        s1.record()
        barrier()
        tw1 = time.time()
        el = allmax(s0.elapsed_time(s1))
        sustained = {'steps': n_sus, 'ms_per_step': el / n_sus, 'value': nb * world / (el / n_sus * 1e-3), 'unit': 'samples/s',
                     'clocks': clocks2.stop(tw0, tw1) if clocks2 else None}

    # ---- end to end: PPO.optimize_host on pinned host rollouts ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(algo, st, rollout, n_envs, dev, world, args, allmax, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    capped = bool(clk and 'sw_power_cap' in (clk.get('reasons') or []))
    pk['tflops'] = pk['tflops_sustained'] if capped else pk['tflops_burst']
    pk['src'] += ', bf16 ' + ('sustained (power cap active)' if capped else 'burst (full clocks, no power cap in the timed region)')
    step_tf = STEP_FLOPS_PER_SAMPLE * nb / (ms_per_step * 1e-3) / 1e12
    step_roof = {'bound': 'tensor', 'unit': 'TFLOP/s', 'flops_per_sample': STEP_FLOPS_PER_SAMPLE,
                 'burst': {'achieved': step_tf, 'peak': pk['tflops'], 'frac': step_tf / pk['tflops']}}
    if sustained:
        cap2 = bool(sustained['clocks'] and 'sw_power_cap' in (sustained['clocks'].get('reasons') or []))
        stf = STEP_FLOPS_PER_SAMPLE * sustained['value'] / world / 1e12
        speak = pk['tflops_sustained'] if cap2 else pk['tflops_burst']
        sustained['step_achieved_tflops'] = stf
        sustained['step_frac'] = stf / speak
        sustained['tensor_peak'] = 'sustained' if cap2 else 'burst'
        step_roof['sustained'] = {'achieved': stf, 'peak': speak, 'frac': stf / speak}
    # the grading quantity of SURVEY §8(d): whole learner step against the tensor roofline, sustained regime when measured
    lead = step_roof.get('sustained', step_roof['burst'])
    step_roof.update({'achieved': lead['achieved'], 'peak': lead['peak'], 'frac': lead['frac'],
                      'regime': 'sustained' if 'sustained' in step_roof else 'burst'})
    # dominant kernel of the step and its roofline
    dom = max((k for k in sites if k in MACS), key=lambda k: sites[k][0] * 1.0, default=None)
    roof = None
    per_kernel = {k: round(v[0], 4) for k, v in sites.items()}
    if dom:
        flops = 2.0 * MACS[dom] * nb
        ach = flops / (sites[dom][0] * 1e-3) / 1e12
        gbs = BYTES[dom] * nb / (sites[dom][0] * 1e-3) / 1e9
        # `traffic` = dram read+write bytes per launch from the committed ncu --set full capture of the same kernel at the same size
        traffic = None
        for name in ('r02_kernel_traffic.json', 'r01_kernel_traffic.json'):
            tp = os.path.join(ROOT, 'profiles', name)
            if os.path.exists(tp):
                tj = json.load(open(tp))
                if tj.get('samples_per_launch') == nb and dom in tj.get('kernels', {}):
                    traffic = tj['kernels'][dom]['dram_bytes']
                    break
        tensor = {'bound': 'tensor', 'achieved': ach, 'peak': pk['tflops'], 'unit': 'TFLOP/s', 'frac': ach / pk['tflops']}
        hbm = {'bound': 'hbm', 'achieved': gbs, 'peak': pk['hbm'], 'unit': 'GB/s', 'frac': gbs / pk['hbm']}
        first, other = (hbm, tensor) if hbm['frac'] > tensor['frac'] else (tensor, hbm)     # the roofline that binds this kernel first
        roof = dict(first)
        roof.update({'kernel': dom, 'traffic': traffic, 'peak_source': pk['src'], 'other': other,
                     'algorithmic_bytes_per_launch': BYTES[dom] * nb, 'algorithmic_flops_per_launch': flops,
                     'avg_launch_ms': sites[dom][0], 'kernel_share_of_step': sites[dom][0] / ms_per_step,
                     'step': step_roof, 'step_frac': step_roof['frac'], 'per_kernel_ms': per_kernel})
    extras = None
    if extras_on:
        extras = run_extras(algo, agent, rollout, st, n_envs, dev, ms_per_step, pk, hyper, batch)
    cpu = None
    if world == 1 and not args.no_cpu:
        cb = cpu_baseline(args.cpu_seconds)
        cpu = {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    line = {
        'metric': 'PPO update samples/sec (84x84x4 frames)', 'value': value, 'unit': 'samples/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': args.precision, 'data': 'synthetic',
        'config': {'workload': f'PPO NatureCNN synthetic 84x84x4 rollout, {n_envs} envs x {T} steps per GPU, '
                               f'learner step = {B} samples x {n_envs} envs = {nb} samples/GPU, A={A}, K=1',
                   'parallelism': f'dp{world}', 'l2_policy': 'inputs larger than L2 (3.8 GB uint8 rollout per GPU, '
                                                             'a different random minibatch every step)'},
        'clocks': clk, 'sustained': sustained, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roof, 'cpu_baseline': cpu,
        'allreduce_check': allreduce_check, 'extras': extras,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
def run_allreduce_check(comm, eng, dev, rank, world):
    """a19 / a25 on the driver's record: the bucketed SUM all-reduce of a rank-seeded gradient-sized buffer (exactly representable
    small integers, so every summation order gives the same floats) equals all-gather + fixed-order sum bit for bit, and
    Communicator.mean_ of a 5-float vector equals the mean of the gathered vectors.  Runs before any timing."""
    import torch.distributed as dist
    n = eng.nparams
    idx = tc.arange(n, device=dev, dtype=tc.int64)

    def seeded(r):
        return (((idx * 7 + r * 13) % 257) - 128).float()
    buf = seeded(rank)
    comm.allreduce_(buf, comm.grad_buckets(eng.offsets))
    gathered = [tc.empty(n, device=dev) for _ in range(world)]
    dist.all_gather(gathered, seeded(rank))
    want = gathered[0].clone()
    for r in range(1, world):
        want += gathered[r]
    ok_grad = bool(tc.equal(buf, want))
    one = seeded(rank)
    comm.allreduce_(one)                                     # single bucket
    ok_one = bool(tc.equal(one, want))
    vec = tc.arange(5, device=dev, dtype=tc.float32) * 0.25 + rank
    got = comm.mean_(vec.clone())
    vg = [tc.empty(5, device=dev) for _ in range(world)]
    dist.all_gather(vg, vec)
    ok_mean = bool(tc.allclose(got, tc.stack(vg).sum(0) / world, rtol=0, atol=1e-6))
    flags = tc.tensor([ok_grad, ok_one, ok_mean], device=dev, dtype=tc.int32)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    ok = bool(flags.min().item())
    return {'status': 'ok' if ok else 'FAILED', 'world': world, 'elements': int(n), 'buckets': comm.grad_buckets(eng.offsets),
            'bit_identical_to_allgather_sum': bool(flags[0].item()) and bool(flags[1].item()), 'mean_5_floats': bool(flags[2].item()),
            'backend': 'p2p' if comm.p2p else 'nccl'}


# ------------------------------------------------------------------------------------------------------
def run_e2e(algo, st, rollout, n_envs, dev, world, args, allmax, barrier):
    """End to end through the public API: `PPO.optimize_host(host_rollout, storage)` — per call the uint8 rollout (frames, actions,
    rewards, dones) is copied from PINNED HOST memory to the device (chunked on a copy stream; the annotate forward of a chunk
    overlaps the upload of the next), then credit assignment, opt_epochs x T/B learner steps with the per-epoch metrics passes,
    and the metrics are read back to the host every epoch.  Update samples per call = opt_epochs * n_envs * T."""
    d = st.as_dict()
    host = {'observations': d['observations'].cpu().pin_memory(), 'actions': d['actions'].cpu().pin_memory(),
            'dones': d['dones'].cpu().pin_memory(),
            'rewards': {k: d['rewards'][k].cpu().pin_memory() for k in st.reward_keys}}
    h2d = host['observations'].numel() + sum(v.numel() * v.element_size() for v in (host['actions'], host['dones']))
    h2d += sum(v.numel() * v.element_size() for v in host['rewards'].values())
    epochs = algo._opt_epochs
    samples = epochs * n_envs * T
    n_calls = max(2, args.steps // 6)
    algo.optimize_host(host, st)                          # warm-up (allocations, first-touch)
    barrier()
    ev0, ev1 = tc.cuda.Event(enable_timing=True), tc.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(n_calls):
        algo.optimize_host(host, st)
    ev1.record()
    tc.cuda.synchronize()
    wall = time.perf_counter() - t0
    el = allmax(ev0.elapsed_time(ev1))
    assert all(np.isfinite(list(m.values())).all() for m in algo.last_epoch_metrics)
    steps_per_call = epochs * (T // B)
    out = {'value': samples * world / (el / n_calls * 1e-3), 'unit': 'samples/s', 'h2d_bytes_per_step': int(h2d),
           'd2h_bytes_per_step': 20 * epochs, 'ms_per_step': el / n_calls, 'wall_ms_per_step': 1e3 * wall / n_calls,
           'step': f'one PPO.optimize_host call = upload + annotate + GAE + {steps_per_call} learner steps + {epochs} metrics passes',
           'calls': n_calls, 'update_samples_per_call': samples,
           'h2d_GBps': h2d / (el / n_calls * 1e-3) / 1e9,
           'api': 'PPO.optimize_host(pinned host rollout, RolloutStorage)'}
    # Secondary figure, NOT the reference's strictly on-policy schedule: the next rollout is uploaded (double-buffered storage)
    # while the current one is optimised — what an actor-learner pipeline that streams frames to the device during collection
    # achieves.  Every call still contains one full upload; reported beside, never instead of, the sequential number above.
    try:
        from pytorch_drl_b200 import algos
        st2 = algos.RolloutStorage(n_envs, T, st.reward_keys, dev)
        stores = [st, st2]
        pending = algo.upload_rollout(host, stores[0])
        barrier()
        p0, p1 = tc.cuda.Event(enable_timing=True), tc.cuda.Event(enable_timing=True)
        p0.record()
        for c in range(n_calls):
            cur = stores[c % 2]
            rollout_c = algo.annotate_uploaded(cur, pending)
            pending = algo.upload_rollout(host, stores[(c + 1) % 2])     # next rollout's upload overlaps this rollout's epochs
            algo.optimize(rollout_c)
        p1.record()
        tc.cuda.synchronize()
        elp = allmax(p0.elapsed_time(p1))
        out['pipelined'] = {'value': samples * world / (elp / n_calls * 1e-3), 'ms_per_step': elp / n_calls,
                            'note': 'upload of rollout k+1 overlapped with optimize of rollout k (two device storages); off-policy by one rollout, '
                                    'so it is a secondary figure'}
        del st2
    except Exception as ex:                                   # noqa: BLE001
        out['pipelined'] = {'error': repr(ex)[:200]}
    return out


# ------------------------------------------------------------------------------------------------------
def run_extras(algo, agent, rollout, st, n_envs, dev, ms_per_step, pk, hyper, batch):
    """The other BASELINE configs and the side measurements SURVEY §8(d) asks for (rank 0, N=1; a few seconds of GPU time)."""
    from pytorch_drl_b200 import _lib, algos, ops
    from pytorch_drl_b200._lib import current_stream, ptr
    from pytorch_drl_b200.envs import RandomNetworkDistillation
    from pytorch_drl_b200.replay import SumTree
    from pytorch_drl_b200.utils.nested import MinibatchView
    L = _lib.lib()
    hbm_peak = pk['hbm']
    out = {}
    g = tc.Generator(device=dev)
    g.manual_seed(7)
    nb = n_envs * B
    all_rows = st.rows(T)

    # ---- one whole PPO.optimize call on the device-resident rollout (annotate + GAE + epochs x steps + metrics passes) ----
    def whole():
        algo.optimize(algo.credit_assignment(algo.annotate(rollout)))
    ms_opt = timed(whole, iters=3, warm=1)
    out['optimize_call'] = {'ms': ms_opt, 'update_samples': 3 * n_envs * T, 'samples_per_s': 3 * n_envs * T / (ms_opt * 1e-3),
                            'what': f'annotate ({n_envs * (T + 1)} forward samples) + GAE + 3 epochs x {T // B} learner steps + 3 metrics passes, device resident',
                            'step_frac_of_tensor_peak': STEP_FLOPS_PER_SAMPLE * 3 * n_envs * T / (ms_opt * 1e-3) / 1e12 / pk['tflops_sustained']}

    # ---- sweep (BASELINE config 5): samples per learner step 4k..64k ----
    sweep = []
    for n in (4096, 8192, 16384, 32768, 65536):
        rows = all_rows[tc.randperm(all_rows.numel(), device=dev, generator=g)[:n]].contiguous()
        mb = MinibatchView(rollout, rows)
        ms = timed(lambda: algo.learner_step(mb, hyper, batch), iters=10, warm=3)
        tf = STEP_FLOPS_PER_SAMPLE * n / (ms * 1e-3) / 1e12
        sweep.append({'samples_per_step': n, 'ms_per_step': ms, 'samples_per_s': n / (ms * 1e-3), 'step_tflops': tf,
                      'step_frac_burst_peak': tf / pk['tflops_burst']})
    out['sweep'] = sweep

    # ---- config 0 (BASELINE config[0] shapes on one GPU): 8 envs, T=128/B=32 and the shipped T=256/B=64 ----
    cfg0 = []
    for (T0, B0) in ((128, 32), (256, 64)):
        a0 = make_agent(A, REWARD_KEYS, 8 * (T0 + 1), 'bf16', dev)
        o0 = tc.optim.Adam(a0.parameters(), lr=2.5e-4, eps=1e-5)
        al0 = algos.PPO(rank=0, world_size=1, rollout_len=T0, extra_steps=0,
                        credit_assignment_ops={'extrinsic': algos.GAE(gamma=0.99, use_dones=True, lambda_=0.95)}, global_step=1,
                        max_steps=10 ** 7, policy_net=a0, policy_optimizer=o0, opt_epochs=3, learner_batch_size=B0,
                        envs_per_rank=8, sampler_seed=0)
        s0 = algos.RolloutStorage(8, T0, REWARD_KEYS, dev)
        fill_synthetic(s0.as_dict(), s0, 8, A, REWARD_KEYS, dev, 99)
        r0 = al0.credit_assignment(al0.annotate(s0.as_dict()))
        h0, b0 = al0._hyper(), al0._batch(r0)
        er = al0._epoch_rows(r0)
        mbs = [MinibatchView(r0, er[i]) for i in range(er.shape[0])]
        k = [0]

        def one_step():
            al0.learner_step(mbs[k[0] % len(mbs)], h0, b0)
            k[0] += 1
        ms_gpu = timed(one_step, iters=50, warm=5)
        tc.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            one_step()
        host_ms = (time.perf_counter() - t0) / 50 * 1e3          # time to ENQUEUE a step (python + ctypes + launches), no sync
        tc.cuda.synchronize()
        ms_call = timed(lambda: al0.optimize(al0.credit_assignment(al0.annotate(r0))), iters=3, warm=1)
        cfg0.append({'envs': 8, 'T': T0, 'B': B0, 'samples_per_step': 8 * B0, 'ms_per_step': ms_gpu, 'host_enqueue_ms_per_step': host_ms,
                     'samples_per_s': 8 * B0 / (ms_gpu * 1e-3), 'optimize_call_ms': ms_call,
                     'optimize_call_samples_per_s': 3 * 8 * T0 / (ms_call * 1e-3),
                     'bound': 'launch latency: ~20 kernels per step on 1-4 CTAs each; the host enqueue time is the floor when larger'})
        del a0, al0, s0
    out['config0'] = cfg0

    # ---- config 3 (RND + PPO per models_dir/atari_rnd): A=18, K=2 (extrinsic gamma .999 / intrinsic gamma .99 no dones, weights
    #      2.0 / 1.0), clip .1, ent .001, vf_coef .5, 4 epochs, lr 1e-4; RND predictor learn once per PPO minibatch, no row drop ----
    rk3 = ['extrinsic', 'intrinsic_rnd']
    a3 = make_agent(18, rk3, nb, 'bf16', dev)
    o3 = tc.optim.Adam(a3.parameters(), lr=1e-4, eps=1e-5)
    rnd = RandomNetworkDistillation({'lr': 1e-4, 'eps': 1e-5}, world_size=32, widening=1, non_learning_steps=0, max_batch=nb, device=dev)
    rnd.observe(tc.rand((8, 84, 84, 1), device=dev, generator=g))
    rnd._sync_normalizers_global(); rnd._sync_normalizers_local()
    al3 = algos.PPO(rank=0, world_size=1, rollout_len=T, extra_steps=0,
                    credit_assignment_ops={'extrinsic': algos.GAE(gamma=0.999, use_dones=True, lambda_=0.95),
                                           'intrinsic_rnd': algos.GAE(gamma=0.99, use_dones=False, lambda_=0.95)},
                    global_step=10 ** 6, non_learning_steps=0, max_steps=5 * 10 ** 8, policy_net=a3, policy_optimizer=o3,
                    standardize_adv=False, opt_epochs=4, learner_batch_size=B, clip_param_init=0.1, clip_param_final=0.1,
                    ent_coef_init=0.001, ent_coef_final=0.001, vf_loss_coef=0.5, reward_weights={'extrinsic': 2.0, 'intrinsic_rnd': 1.0},
                    envs_per_rank=n_envs, sampler_seed=0, env=rnd)
    s3 = algos.RolloutStorage(n_envs, T, rk3, dev)
    d3 = s3.as_dict()
    d3['observations'] = rollout['observations']            # the same 3.8 GB of frames
    fill_synthetic(d3, s3, n_envs, 18, rk3, dev, 5, obs=False)
    # intrinsic rewards come from the RND networks themselves, batched over the rollout
    ms_rew = timed(lambda: rnd.intrinsic_reward(d3['observations'], all_rows), iters=3, warm=1)
    d3['rewards']['intrinsic_rnd'][:, :T] = rnd.intrinsic_reward(d3['observations'], all_rows).view(n_envs, T).clamp_(-5, 5)
    r3 = al3.credit_assignment(al3.annotate(d3))
    h3, b3 = al3._hyper(), al3._batch(r3)
    er3 = al3._epoch_rows(r3)
    mbs3 = [MinibatchView(r3, er3[i]) for i in range(er3.shape[0])]
    k3 = [0]

    def step3():
        mb = mbs3[k3[0] % len(mbs3)]
        al3._update_trainable_wrappers(mb)                   # RND.learn(mb): predictor MSE step + normaliser sync
        al3.learner_step(mb, h3, b3)
        k3[0] += 1
    ms3 = timed(step3, iters=12, warm=4)
    ms_rnd = timed(lambda: rnd.learn(mbs3[0]), iters=12, warm=2)
    ms_ppo3 = timed(lambda: al3.learner_step(mbs3[1], h3, b3), iters=12, warm=2)
    flops3 = (STEP_FLOPS_PER_SAMPLE + 3 * 2 * 512 * (12 + 1) + RND_FLOPS_PER_SAMPLE) * nb    # +12 actions, +1 value head: fwd + 2 x bwd
    out['config3'] = {'workload': f'RND + PPO (atari_rnd shapes): A=18, K=2, {n_envs} envs x {T} steps, {nb} samples per step, RND learn per minibatch (no row drop)',
                      'ms_per_step': ms3, 'samples_per_s': nb / (ms3 * 1e-3), 'ppo_step_ms': ms_ppo3, 'rnd_learn_ms': ms_rnd,
                      'rnd_learn_samples_per_s': nb / (ms_rnd * 1e-3), 'rnd_learn_tflops': RND_FLOPS_PER_SAMPLE * nb / (ms_rnd * 1e-3) / 1e12,
                      'rnd_learn_frac_burst_peak': RND_FLOPS_PER_SAMPLE * nb / (ms_rnd * 1e-3) / 1e12 / pk['tflops_burst'],
                      'step_tflops': flops3 / (ms3 * 1e-3) / 1e12, 'step_frac_burst_peak': flops3 / (ms3 * 1e-3) / 1e12 / pk['tflops_burst'],
                      'rnd_intrinsic_reward': {'samples': n_envs * T, 'ms': ms_rew, 'samples_per_s': n_envs * T / (ms_rew * 1e-3)},
                      'rnd_engine': rnd.precision_name}
    out['rnd_learn'] = {'samples': nb, 'ms': ms_rnd, 'samples_per_s': nb / (ms_rnd * 1e-3), 'engine': 'tcgen05 bf16 (csrc/umma_rnd.cu)'}
    del a3, al3, rnd, s3, d3, r3, mbs3
    tc.cuda.empty_cache()

    # ---- precision modes: fp32 CUDA-core engine vs the tcgen05 bf16 engine at 4 096 samples (rate, gradient rel-L2) ----
    npre = 4096
    rows_p = all_rows[tc.randperm(all_rows.numel(), device=dev, generator=g)[:npre]].contiguous()
    eng16 = agent.engine
    from pytorch_drl_b200.engine import LearnerEngine
    eng32 = LearnerEngine(A, ['value_extrinsic'], npre, 'fp32', dev)
    eng32.load_named({k: v.clone() for k, v in eng16.named_views().items()})
    ms16 = timed(lambda: eng16.ppo_step(rollout['observations'], rows_p, batch, hyper), iters=10, warm=3)
    g16 = eng16.grads.double().clone()
    ms32 = timed(lambda: eng32.ppo_step(rollout['observations'], rows_p, batch, hyper), iters=2, warm=1)
    g32 = eng32.grads.double()
    per = {}
    for i, name in enumerate(eng16.names):
        a_, b_ = g16[eng16.offsets[i]:eng16.offsets[i + 1]], g32[eng16.offsets[i]:eng16.offsets[i + 1]]
        per[name.replace('_architecture._network.', 'trunk.').replace('_predictors.', '').replace('._network.0', '')] = float((a_ - b_).norm() / b_.norm().clamp_min(1e-30))
    out['precision'] = {'samples': npre,
                        'fp32': {'engine': 'CUDA cores (strict parity: <= 2e-4 rel-L2 vs torch-CPU)', 'fwd_bwd_samples_per_s': npre / (ms32 * 1e-3)},
                        'bf16': {'engine': 'tcgen05, fp32 accumulate', 'fwd_bwd_samples_per_s': npre / (ms16 * 1e-3),
                                 'grad_rel_l2_vs_fp32_all_params': float((g16 - g32).norm() / g32.norm()), 'grad_rel_l2_vs_fp32_per_tensor': per},
                        'tf32': 'not built: CPU emulation (profiles/r02_precision_emulation.txt) gives 2 % gradient rel-L2 for 10-bit operands — the error is '
                                'ReLU-mask flips, sqrt(2 x flip fraction), so tf32 would not reach rtol 1e-3 either; fp32 is the parity mode'}
    del eng32

    # ---- GAE + vtarg + per-env standardisation: 20 B per (env, step) ----
    for tag, n in (('config', n_envs), ('large', 65536)):
        r = tc.randint(-1, 2, (n, T), device=dev, generator=g).float()
        v = tc.randn((n, T + 4), device=dev, generator=g)
        dn = (tc.rand((n, T), device=dev, generator=g) < 0.01).float()
        adv, vt = tc.empty((n, T), device=dev), tc.empty((n, T), device=dev)
        ms = timed(lambda: ops.gae(r, v, dn, 0.99, 0.95, True, True, T, adv, vt))
        out[f'gae_{tag}'] = {'envs': n, 'steps': T, 'us': ms * 1e3, 'env_steps_per_s': n * T / (ms * 1e-3),
                             'GBps': 20.0 * n * T / (ms * 1e-3) / 1e9, 'frac_hbm': 20.0 * n * T / (ms * 1e-3) / 1e9 / hbm_peak,
                             'bytes_per_env_step': 20}
        del r, v, dn, adv, vt
    # ---- fused clipped-surrogate / value / entropy loss forward + backward on given logits: 8A + 12 + 20K B per sample ----
    for tag, n in (('config', n_envs * B), ('large', 1 << 22)):
        logits = tc.randn((n, A), device=dev, generator=g)
        values = tc.randn((n, 1), device=dev, generator=g)
        acts = tc.randint(0, A, (n,), device=dev, generator=g)
        lp = tc.log_softmax(logits, -1).gather(1, acts[:, None]).squeeze(1) + 0.05 * tc.randn((n,), device=dev, generator=g)
        advs, vold, vtg = (tc.randn((n,), device=dev, generator=g) for _ in range(3))
        hy = ops.make_hyper(A, [1.0], 0.2, 0.01, 1.0, False, True)
        bt = ops.make_batch(acts, lp, [advs], [vold], [vtg])
        bytes_per = 8 * A + 12 + 20 + 8                     # + logp and entropy written back
        losses, lpo, eno = tc.empty(5, device=dev), tc.empty(n, device=dev), tc.empty(n, device=dev)
        dl, dvv = tc.empty_like(logits), tc.empty_like(values)

        def call():                                          # the C-ABI entry on caller-allocated outputs
            rc = L.drl_ppo_loss_fwd_bwd(ptr(logits), ptr(values), None, n, ctypes.byref(bt), ctypes.byref(hy),
                                        ptr(losses), ptr(lpo), ptr(eno), ptr(dl), ptr(dvv), current_stream())
            assert rc == 0
        ms = timed(call)
        out[f'ppo_loss_{tag}'] = {'samples': n, 'us': ms * 1e3, 'GBps': bytes_per * n / (ms * 1e-3) / 1e9,
                                  'frac_hbm': bytes_per * n / (ms * 1e-3) / 1e9 / hbm_peak, 'bytes_per_sample': bytes_per}
        del losses, lpo, eno, dl, dvv, logits, values, acts, lp, advs, vold, vtg
    # ---- per-epoch metrics pass: forward + losses over the whole rollout, no gradient (ppo.py:265-266) ----
    ms_metrics = timed(lambda: algo.compute_losses_and_metrics(rollout, no_grad=True), iters=5, warm=2)
    steps_per_epoch = T // B
    out['metrics_pass'] = {'samples': n_envs * T, 'ms': ms_metrics, 'forward_samples_per_s': n_envs * T / (ms_metrics * 1e-3)}
    out['update_with_metrics_pass'] = {
        'samples_per_s': n_envs * T / ((steps_per_epoch * ms_per_step + ms_metrics) * 1e-3),
        'note': f'one epoch = {steps_per_epoch} learner steps ({n_envs * T} update samples) + one forward-only pass over the same {n_envs * T} samples'}
    # ---- actor side: one environment step of all envs of a rank = frames H2D into the ring + policy forward
    #      + draw + actions D2H; the reference does one batch-1 forward per environment step (rollout.py:280-288) ----
    out['actor'] = run_actor_extras(agent, n_envs, dev)
    # ---- PER sum-tree (config 4): 2^20 leaves, batch 512 — dependent-access latency bound; parity unpinned (no reference code) ----
    out['sumtree'] = run_sumtree_extras(dev, g)
    # ---- config 3 of BASELINE.json: double / dueling DQN learner step on a 2^20-transition prioritized replay buffer, batch 512 ----
    out['dqn'] = run_dqn_extras(rollout, dev, g)
    return out


def run_dqn_extras(rollout, dev, g):
    """One DQN learner step = prioritized sample of 512 of 2^20 transitions (fused stratified draw + importance weights), Q(s') of the
    online and the target network and Q(s) on the frame ring by row index (no gather copy), double-Q Bellman targets + SmoothL1 TD
    loss + gradients (dueling head, trunk), Adam on both, |td| priorities back into the sum-tree.  The ring holds 2^20 uint8 stacks
    (29.6 GB): its first 131 072 slots are the benchmark's random frames, the rest zeros (kernel times do not depend on the data)."""
    from pytorch_drl_b200 import agents
    from pytorch_drl_b200.algos.dqn import DQN
    from pytorch_drl_b200.replay import PrioritizedReplayBuffer
    A, nb, cap = 6, 512, 1 << 20
    try:
        replay = PrioritizedReplayBuffer(cap, 0.6, 1e-6, device=dev)
    except tc.OutOfMemoryError:
        return {'skipped': 'no room for the 29.6 GB frame ring next to the rollout'}
    obs = rollout['observations']
    n_real = min(obs.shape[0] * obs.shape[1], cap)
    replay.frames.view(cap, -1)[:n_real] = obs.reshape(-1, obs[0, 0].numel())[:n_real]
    replay.actions.random_(0, A, generator=g)
    replay.rewards.copy_(tc.randint(-1, 2, (cap,), device=dev, generator=g).float())
    replay.dones.copy_((tc.rand(cap, device=dev, generator=g) < 0.01).float())
    replay.count, replay.index.size = cap, cap
    replay.index.tree.set_leaves(tc.rand(cap, device=dev, generator=g) + 0.01)
    qas = []
    for _ in range(2):
        head = agents.SimpleDiscreteActionValueHead(512, A, agents.DuelingArchitecture, {'widening': 1},
                                                    w_init=lambda w: tc.nn.init.orthogonal_(w, gain=2 ** 0.5), b_init=tc.nn.init.zeros_)
        qa = agents.QAgent([agents.ToChannelMajor()], agents.NatureCNN(4, lambda w: tc.nn.init.orthogonal_(w, gain=2 ** 0.5), tc.nn.init.zeros_),
                           {'action_value_extrinsic': head})
        qas.append(qa.fuse(max_batch=nb, precision='bf16', device=dev))
    dqn = DQN(qas[0], qas[1], replay, batch_size=nb, gamma=0.99, double_q=True, loss='SmoothL1Loss', lr=1e-4, target_update_interval=1000)
    ms = timed(dqn.learner_step, iters=50, warm=5)
    tc.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        dqn.learner_step()
    host_ms = (time.perf_counter() - t0) / 50 * 1e3
    tc.cuda.synchronize()
    loss = float(dqn.loss.item())
    del dqn, qas, replay
    tc.cuda.empty_cache()
    return {'workload': 'double / dueling DQN (NatureCNN trunk + DuelingArchitecture head), 2^20-transition prioritized replay, batch 512',
            'ms_per_step': ms, 'samples_per_s': nb / (ms * 1e-3), 'host_enqueue_ms_per_step': host_ms, 'final_loss': loss,
            'trunk_passes_per_step': '3 forward (target s\', online s\', online s) + 1 backward on the live activations of the last one, at 512 samples',
            'bound': 'launch latency: ~40 launches of 4-148 CTAs per step; the host enqueue time is the floor when larger',
            'parity': 'network blocks + Bellman operator pinned to the reference (tests/golden/dqn.npz); the algorithm itself has no reference '
                      'counterpart (the reference ships no DQN learner and its dueling head cannot be differentiated as written: dueling.py:53)'}


def graph_timed(fn, iters=50):
    """GPU time per call with the launches replayed from a CUDA graph (host launch overhead removed); eager timing if the
    capture is refused."""
    try:
        fn()
        tc.cuda.synchronize()
        gr = tc.cuda.CUDAGraph()
        with tc.cuda.graph(gr):
            for _ in range(iters):
                fn()
        gr.replay()
        tc.cuda.synchronize()
        e0, e1 = tc.cuda.Event(enable_timing=True), tc.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            gr.replay()
        e1.record()
        tc.cuda.synchronize()
        return e0.elapsed_time(e1) / (5 * iters), 'cuda graph replay'
    except Exception as ex:                                   # noqa: BLE001
        tc.cuda.synchronize()
        return timed(fn, iters=iters), f'eager launches ({type(ex).__name__})'


def run_actor_extras(agent, n_envs, dev):
    from pytorch_drl_b200 import _lib, algos
    from pytorch_drl_b200._lib import current_stream, ptr
    L = _lib.lib()
    eng = agent.engine
    r = algos.DeviceRollout(n_envs, 7, 0, REWARD_KEYS, dev)
    d = r.storage.data
    r.staging_frames()[...] = np.random.RandomState(3).randint(0, 256, size=(n_envs, 84, 84, 4), dtype=np.uint8)
    acts = tc.empty(n_envs, dtype=tc.int64, device=dev)
    h_acts = tc.empty(n_envs, dtype=tc.int64).pin_memory()
    rows = [r.rows_at(t) for t in range(8)]
    k = [0]

    def upload():
        r.write_frames(k[0] % 8, r.staging_frames())

    def act():
        k[0] += 1
        rc = L.drl_net_act(eng._h, ptr(d['observations']), ptr(rows[k[0] % 8]), n_envs, None, 1, k[0], ptr(acts), None, None, None,
                           current_stream())
        assert rc == 0
    ms_up = timed(upload, iters=20, warm=3)
    ms_act = timed(act, iters=50, warm=5)

    def tick():                                     # what BatchedRolloutManager does per time step, minus env.step itself
        upload()
        act()
        h_acts.copy_(acts, non_blocking=True)
        tc.cuda.current_stream().synchronize()
    for _ in range(3):
        tick()
    t0 = time.perf_counter()
    for _ in range(30):
        tick()
    ms_tick = (time.perf_counter() - t0) / 30 * 1e3
    # frame_stack mode: single newest frames (7 KB per env) read by the stacking kernel straight from pinned host memory
    newest, modes = r.staging_newest()
    newest[...] = np.random.RandomState(4).randint(0, 256, size=newest.shape, dtype=np.uint8)
    modes[...] = 0

    def stack():
        r.stack_frames(k[0] % 8, (k[0] + 1) % 8)
    ms_stack = timed(stack, iters=20, warm=3)

    def tick_fs():
        stack()
        act()
        h_acts.copy_(acts, non_blocking=True)
        tc.cuda.current_stream().synchronize()
    for _ in range(3):
        tick_fs()
    t0 = time.perf_counter()
    for _ in range(30):
        tick_fs()
    ms_tick_fs = (time.perf_counter() - t0) / 30 * 1e3
    return {'envs': n_envs, 'frame_stack_on_device': {'stack_kernel_ms': ms_stack, 'pcie_bytes_per_env_step': 7056, 'tick_wall_ms': ms_tick_fs,
                                                      'env_steps_per_s_ceiling': n_envs / (ms_tick_fs * 1e-3)},
            'frames_h2d_ms': ms_up, 'frames_h2d_GBps': n_envs * 28224 / (ms_up * 1e-3) / 1e9,
            'policy_forward_and_draw_ms': ms_act, 'forward_samples_per_s': n_envs / (ms_act * 1e-3),
            'tick_wall_ms': ms_tick, 'env_steps_per_s_ceiling': n_envs / (ms_tick * 1e-3),
            'what': 'per time step of all envs: pinned frames -> their ring slot (one strided H2D copy), drl_net_act on those rows (bf16 engine, '
                    'in-kernel Philox draw), actions D2H + sync; env.step itself (host) not included'}


def run_sumtree_extras(dev, g):
    """BASELINE config 4: 2^20-leaf sum-tree, batch 512 priority update + prioritized sample (+ a 64 k-update batch on the
    multi-CTA path), next to the dependent-access floor: levels x measured L2 latency of a pointer chase."""
    from pytorch_drl_b200 import _lib
    from pytorch_drl_b200._lib import current_stream, ptr
    from pytorch_drl_b200.replay import SumTree
    L = _lib.lib()
    cap, bs = 1 << 20, 512
    tree = SumTree(cap, device=dev)
    tree.set_leaves(tc.rand(cap, device=dev, generator=g) + 1e-3)
    idx = tc.randint(0, cap, (bs,), device=dev, generator=g)
    pr = tc.rand(bs, device=dev, generator=g) + 1e-3
    u01 = tc.rand(bs, device=dev, generator=g)
    big_idx = tc.randint(0, cap, (65536,), device=dev, generator=g)
    big_pr = tc.rand(65536, device=dev, generator=g) + 1e-3
    idx_o, pr_o, w_o = tc.empty(bs, dtype=tc.int64, device=dev), tc.empty(bs, device=dev), tc.empty(bs, device=dev)
    ms_u, how = graph_timed(lambda: tree.update(idx, pr))
    ms_s, _ = graph_timed(lambda: L.drl_per_sample(ptr(tree.tree), cap, ptr(u01), bs, float(cap), 0.4, ptr(idx_o), ptr(pr_o), ptr(w_o),
                                                   current_stream()))
    ms_b, _ = graph_timed(lambda: tree.update(big_idx, big_pr), iters=10)
    # dependent-load latency: a random cyclic permutation over 8 MB of int32 (the tree's size: L2 resident)
    n_ring = 1 << 21
    perm = tc.randperm(n_ring, device=dev, generator=g)
    ring = tc.empty(n_ring, dtype=tc.int32, device=dev)
    ring[perm] = tc.roll(perm, -1).to(tc.int32)
    cyc = tc.zeros(2, device=dev)
    L.drl_probe_dependent_latency(ptr(ring), 4096, ptr(cyc), current_stream())
    tc.cuda.synchronize()
    sm_mhz = tc.cuda.clock_rate() if hasattr(tc.cuda, 'clock_rate') else 1965
    hop_ns = float(cyc[0].item()) / (sm_mhz * 1e-3)
    return {'leaves': cap, 'batch': bs, 'update_us': ms_u * 1e3, 'sample_us': ms_s * 1e3, 'timing': how,
            'updates_per_s': bs / (ms_u * 1e-3), 'samples_per_s': bs / (ms_s * 1e-3),
            'update_64k': {'batch': 65536, 'us': ms_b * 1e3, 'updates_per_s': 65536 / (ms_b * 1e-3), 'path': '128 CTAs (depth-7 subtrees) + top pass'},
            'sample': 'fused: stratified masses + descent + importance weights (drl_per_sample)',
            'dependent_latency': {'cycles_per_level': float(cyc[0].item()), 'ns_per_level': hop_ns, 'levels': 20,
                                  'walk_floor_us': 20 * hop_ns * 1e-3,
                                  'note': 'one pointer chase over 8 MB (L2 resident); update / sample walk 20 dependent levels'},
            'parity': 'bit-exact vs oracle/sumtree.c (parity unpinned: the reference has no sum-tree, SURVEY F2)'}


if __name__ == '__main__':
    main()
