#!/usr/bin/env python
"""bench.py — PPO update samples/sec on synthetic 84x84x4 rollouts (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A "step" is one learner step of PPO.optimize (drl/algos/ppo.py:246-257): gather a minibatch of
B samples from each env of the GPU, NatureCNN forward, fused losses, backward, gradient
all-reduce across GPUs, Adam.  Workload at N=1: BASELINE config[1] — 1024 envs x 128 steps per
GPU, B=32 per env => 32 768 samples per step per GPU (weak scaling: every GPU owns 1024 envs).
Rank 0 prints ONE JSON line (see the keys in `main`).
"""
import argparse
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
def cpu_reference_step_rate(n_envs: int, seconds: float, warmup: int = 1, max_steps: int = 2000, steps=None):
    """Times the CPU restatement of the reference's learner step (oracle port: torch-CPU fp32 autograd
    through the same conv/linear/Adam ops the reference calls, all host threads) on a bounded sample:
    `n_envs` ranks x B samples per step.  Returns (samples/s, steps, threads)."""
    from oracle import ppo_oracle as po
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU arm must still use all the host threads it can
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


def run_reference(args, rank):
    """`--impl reference`: the reference's own CPU implementation of the path is pure PyTorch-CPU
    Python that cannot travel to the GPU box (/root/reference is absent there and gym/moviepy are not
    installable), so the arm times the oracle port — pinned to the reference by tests/golden — on all
    host threads.  One step = one learner step over a bounded sample (8 ranks x 32 samples, the
    reference's own CPU-runnable shape)."""
    if rank != 0:
        return
    n_envs = 8
    rate, n, threads, el = cpu_reference_step_rate(n_envs, 0.0, warmup=max(1, args.warmup), steps=args.steps)
    line = {
        'impl': 'reference', 'metric': 'PPO update samples/sec (84x84x4 frames)', 'value': rate, 'unit': 'samples/s',
        'n_gpus': args.gpus, 'steps': n, 'warmup': max(1, args.warmup), 'ms_per_step': 1e3 * el / n,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'PPO NatureCNN synthetic 84x84x4 rollout, 1024 envs x 128 steps, B=32 per env',
                   'sample': f'{n_envs} envs x {B} samples per step on the host CPU'},
        'cpu_baseline': {'value': rate, 'unit': 'samples/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{n} learner steps of {n_envs} envs x {B} samples (oracle port, torch-CPU fp32)'},
        'e2e': {'value': rate, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


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
    ap.add_argument('--no-extras', action='store_true', help='skip the GAE / loss / metrics-pass / sum-tree side measurements')
    ap.add_argument('--debug-flags', default='', help='kernel-variant experiments, e.g. "2=1,3=0" (drl_debug_set)')
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
    from pytorch_drl_b200 import _lib, agents, algos
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

    # ---- model + algo through the public drop-in API ----
    tc.manual_seed(0)
    agent = agents.Agent(
        [agents.ToChannelMajor()],
        agents.NatureCNN(4, lambda w: tc.nn.init.orthogonal_(w, gain=2 ** 0.5), tc.nn.init.zeros_),
        {'policy': agents.CategoricalPolicyHead(512, A, w_init=lambda w: tc.nn.init.orthogonal_(w, gain=0.01), b_init=tc.nn.init.zeros_),
         'value_extrinsic': agents.SimpleValueHead(512, w_init=lambda w: tc.nn.init.orthogonal_(w, gain=0.01), b_init=tc.nn.init.zeros_)})
    agent.fuse(max_batch=nb, precision=args.precision, device=dev)
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
    g = tc.Generator(device=dev)
    g.manual_seed(10000 * rank)
    chunk = 64
    for e0 in range(0, n_envs, chunk):
        e1 = min(n_envs, e0 + chunk)
        d['observations'][e0:e1] = tc.randint(0, 256, (e1 - e0, st.Tp, 84, 84, 4), device=dev, dtype=tc.uint8, generator=g)
    d['actions'][:] = tc.randint(0, A, (n_envs, st.Tp), device=dev, generator=g)
    d['rewards']['extrinsic'][:] = tc.randint(-1, 2, (n_envs, st.Tp), device=dev, generator=g).float()
    d['dones'][:] = (tc.rand((n_envs, st.Tp), device=dev, generator=g) < 0.01).float()
    rollout = algo.credit_assignment(algo.annotate(d))
    tc.cuda.synchronize()

    rs = np.random.RandomState(1234 + rank)

    def step_rows():
        blocks = [rs.permutation(T)[:B] for _ in range(n_envs)]
        return algo._minibatch_rows(rollout, blocks)

    total_steps = args.warmup + args.steps
    rows_list = [step_rows() for _ in range(total_steps)]

    def barrier():
        if world > 1:
            dist.barrier()
        tc.cuda.synchronize()

    # ---- device-resident timing (value) ----
    clocks = ClockSampler(local_rank) if rank == 0 else None     # started early: nvidia-smi needs ~0.2 s to its first sample
    if clocks is not None:
        clocks.wait_ready()
    t_warm0 = time.time()
    for i in range(args.warmup):
        algo.learner_step(MinibatchView(rollout, rows_list[i]))
    barrier()
    L.drl_profile_enable(1)
    names = (b'\0' * (48 * 64))
    import ctypes
    nbuf = ctypes.create_string_buffer(48 * 64)
    ms = (ctypes.c_float * 64)()
    cnt = (ctypes.c_int * 64)()
    ns = ctypes.c_int()
    L.drl_profile_read(64, nbuf, ms, cnt, ctypes.byref(ns))       # clear
    launches0 = L.drl_launch_count()
    ev0, ev1 = tc.cuda.Event(enable_timing=True), tc.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record()
    for i in range(args.steps):
        algo.learner_step(MinibatchView(rollout, rows_list[args.warmup + i]))
    ev1.record()
    barrier()
    t_wall1 = time.time()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = L.drl_launch_count() - launches0
    L.drl_profile_read(64, nbuf, ms, cnt, ctypes.byref(ns))       # per-launch-site CUDA-event times of the timed region
    L.drl_profile_enable(0)
    sites = {}
    for i in range(ns.value):
        nm = nbuf.raw[48 * i:48 * i + 48].split(b'\0')[0].decode()
        sites[nm] = (ms[i] / max(1, cnt[i]), cnt[i])
    if world > 1:
        t = tc.tensor([elapsed_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = nb * world / (ms_per_step * 1e-3)
    clk = clocks.stop(t_wall0, t_wall1, [(t_warm0, t_wall0)]) if clocks else None

    # ---- the same steps for ~1 s: the board reaches its power cap and lowers the SM clock (MEASURED_PEAKS.json's
    #      "sustained" regime); reported beside `value`, which is the short run at full clocks ----
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
            algo.learner_step(MinibatchView(rollout, rows_list[args.warmup + i % args.steps]))
        s1.record()
        barrier()
        tw1 = time.time()
        el = s0.elapsed_time(s1)
        if world > 1:
            t = tc.tensor([el], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        sustained = {'steps': n_sus, 'ms_per_step': el / n_sus, 'value': nb * world / (el / n_sus * 1e-3), 'unit': 'samples/s',
                     'clocks': clocks2.stop(tw0, tw1) if clocks2 else None}

    # ---- end-to-end: host minibatches -> H2D -> learner step -> D2H loss read, every step ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(algo, agent, rollout, st, n_envs, dev, world, args, rows_list)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    # tensor peak of the regime the timed region ran in: cuBLAS burst figure at full clocks, the sustained figure when
    # nvidia-smi reported the power cap (B200_PROFILING.md: burst for a kernel timed alone, sustained inside a long step)
    capped = bool(clk and 'sw_power_cap' in (clk.get('reasons') or []))
    pk['tflops'] = pk['tflops_sustained'] if capped else pk['tflops_burst']
    pk['src'] += ', bf16 ' + ('sustained (power cap active)' if capped else 'burst (full clocks, no power cap in the timed region)')
    if sustained:
        cap2 = bool(sustained['clocks'] and 'sw_power_cap' in (sustained['clocks'].get('reasons') or []))
        sustained['step_achieved_tflops'] = STEP_FLOPS_PER_SAMPLE * sustained['value'] / world / 1e12
        sustained['step_frac'] = sustained['step_achieved_tflops'] / (pk['tflops_sustained'] if cap2 else pk['tflops_burst'])
        sustained['tensor_peak'] = 'sustained' if cap2 else 'burst'
    # dominant kernel of the step and its roofline
    dom = max((k for k in sites if k in MACS), key=lambda k: sites[k][0] * 1.0, default=None)
    roof = None
    per_kernel = {k: round(v[0], 4) for k, v in sites.items()}
    if dom:
        flops = 2.0 * MACS[dom] * nb
        ach = flops / (sites[dom][0] * 1e-3) / 1e12
        gbs = BYTES[dom] * nb / (sites[dom][0] * 1e-3) / 1e9
        # report against the roofline that binds first for this kernel (higher fraction); the other one is kept
        # beside it.  `traffic` = dram read+write bytes per launch from the committed ncu --set full capture of
        # the same kernel at the same size (profiles/r01_kernel_traffic.json), null if the size differs.
        traffic = None
        tp = os.path.join(ROOT, 'profiles', 'r01_kernel_traffic.json')
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if tj.get('samples_per_launch') == nb and dom in tj.get('kernels', {}):
                traffic = tj['kernels'][dom]['dram_bytes']
        roof = {'bound': 'tensor', 'kernel': dom, 'achieved': ach, 'peak': pk['tflops'], 'unit': 'TFLOP/s',
                'frac': ach / pk['tflops'], 'traffic': traffic, 'peak_source': pk['src'],
                'algorithmic_bytes_per_launch': BYTES[dom] * nb, 'algorithmic_flops_per_launch': flops,
                'other': {'bound': 'hbm', 'achieved': gbs, 'peak': pk['hbm'], 'unit': 'GB/s', 'frac': gbs / pk['hbm']},
                'avg_launch_ms': sites[dom][0], 'kernel_share_of_step': sites[dom][0] / ms_per_step,
                'step_achieved': STEP_FLOPS_PER_SAMPLE * nb / (ms_per_step * 1e-3) / 1e12,
                'step_frac': STEP_FLOPS_PER_SAMPLE * nb / (ms_per_step * 1e-3) / 1e12 / pk['tflops'],
                'per_kernel_ms': per_kernel}
        if roof['other']['frac'] > roof['frac']:
            o = roof['other']
            roof['other'] = {k: roof[k] for k in ('bound', 'achieved', 'peak', 'unit', 'frac')}
            roof.update(o)
    extras = None
    if world == 1 and not args.no_extras:
        extras = run_extras(algo, rollout, n_envs, dev, ms_per_step, pk['hbm'])
    cpu = None
    if world == 1 and not args.no_cpu:
        rate, n, threads, el = cpu_reference_step_rate(8, args.cpu_seconds)
        cpu = {'value': rate, 'unit': 'samples/s', 'cores': threads, 'kind': 'port',
               'sample': f'{n} learner steps of 8 envs x {B} samples in {el:.1f} s (oracle port of the reference CPU path, torch-CPU fp32, os.cpu_count={os.cpu_count()})'}
    line = {
        'metric': 'PPO update samples/sec (84x84x4 frames)', 'value': value, 'unit': 'samples/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': args.precision, 'data': 'synthetic',
        'config': {'workload': f'PPO NatureCNN synthetic 84x84x4 rollout, {n_envs} envs x {T} steps per GPU, '
                               f'learner step = {B} samples x {n_envs} envs = {nb} samples/GPU, A={A}, K=1',
                   'parallelism': f'dp{world}', 'l2_policy': 'inputs larger than L2 (3.8 GB uint8 rollout per GPU, '
                                                             'a different random minibatch every step)'},
        'clocks': clk, 'sustained': sustained, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roof, 'cpu_baseline': cpu, 'extras': extras,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _timed(fn, iters=20, warm=3):
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


def run_extras(algo, rollout, n_envs, dev, ms_per_step, hbm_peak):
    """The side measurements SURVEY §8(d) asks for next to the headline (rank 0, N=1, a few hundred ms of GPU time):
    GAE in (env x step)/s and HBM GB/s, the fused loss kernel in GB/s, the per-epoch forward-only metrics pass
    (ppo.py:265-266) and the update throughput with that pass included, the PER sum-tree at 2^20 leaves / batch 512.
    The HBM figures use inputs larger than L2 (126 MB); the config-sized calls are launch-latency bound and are
    reported as microseconds."""
    from pytorch_drl_b200 import ops
    from pytorch_drl_b200.replay import SumTree
    out = {}
    g = tc.Generator(device=dev)
    g.manual_seed(7)
    # GAE + vtarg + per-env standardisation: 20 B per (env, step) read+written once (+8 B for the second pass
    # of the standardisation, counted as algorithmic traffic only where it happens)
    for tag, n in (('config', n_envs), ('large', 65536)):
        r = tc.randint(-1, 2, (n, T), device=dev, generator=g).float()
        v = tc.randn((n, T + 4), device=dev, generator=g)
        d = (tc.rand((n, T), device=dev, generator=g) < 0.01).float()
        adv, vt = tc.empty((n, T), device=dev), tc.empty((n, T), device=dev)
        ms = _timed(lambda: ops.gae(r, v, d, 0.99, 0.95, True, True, T, adv, vt))
        out[f'gae_{tag}'] = {'envs': n, 'steps': T, 'us': ms * 1e3, 'env_steps_per_s': n * T / (ms * 1e-3),
                             'GBps': 20.0 * n * T / (ms * 1e-3) / 1e9, 'frac_hbm': 20.0 * n * T / (ms * 1e-3) / 1e9 / hbm_peak,
                             'bytes_per_env_step': 20}
        del r, v, d, adv, vt
    # fused clipped-surrogate / value / entropy loss forward + backward on given logits: 8A + 12 + 20K B per sample
    for tag, n in (('config', n_envs * B), ('large', 1 << 22)):
        logits = tc.randn((n, A), device=dev, generator=g)
        values = tc.randn((n, 1), device=dev, generator=g)
        acts = tc.randint(0, A, (n,), device=dev, generator=g)
        lp = tc.log_softmax(logits, -1).gather(1, acts[:, None]).squeeze(1) + 0.05 * tc.randn((n,), device=dev, generator=g)
        advs, vold, vtg = (tc.randn((n,), device=dev, generator=g) for _ in range(3))
        hyper = ops.make_hyper(A, [1.0], 0.2, 0.01, 1.0, False, True)
        batch = ops.make_batch(acts, lp, [advs], [vold], [vtg])
        bytes_per = 8 * A + 12 + 20 + 8                     # + logp and entropy written back
        import ctypes
        from pytorch_drl_b200 import _lib
        from pytorch_drl_b200._lib import current_stream, ptr
        L = _lib.lib()
        losses, lpo, eno = tc.empty(5, device=dev), tc.empty(n, device=dev), tc.empty(n, device=dev)
        dl, dvv = tc.empty_like(logits), tc.empty_like(values)

        def call():                                          # the C-ABI entry on caller-allocated outputs
            rc = L.drl_ppo_loss_fwd_bwd(ptr(logits), ptr(values), None, n, ctypes.byref(batch), ctypes.byref(hyper),
                                        ptr(losses), ptr(lpo), ptr(eno), ptr(dl), ptr(dvv), current_stream())
            assert rc == 0
        ms = _timed(call)
        out[f'ppo_loss_{tag}'] = {'samples': n, 'us': ms * 1e3, 'GBps': bytes_per * n / (ms * 1e-3) / 1e9,
                                  'frac_hbm': bytes_per * n / (ms * 1e-3) / 1e9 / hbm_peak, 'bytes_per_sample': bytes_per}
        del losses, lpo, eno, dl, dvv
        del logits, values, acts, lp, advs, vold, vtg
    # per-epoch metrics pass: forward + losses over the whole rollout, no gradient (ppo.py:265-266)
    ms_metrics = _timed(lambda: algo.compute_losses_and_metrics(rollout, no_grad=True), iters=5, warm=2)
    steps_per_epoch = T // B
    out['metrics_pass'] = {'samples': n_envs * T, 'ms': ms_metrics, 'forward_samples_per_s': n_envs * T / (ms_metrics * 1e-3)}
    out['update_with_metrics_pass'] = {
        'samples_per_s': n_envs * T / ((steps_per_epoch * ms_per_step + ms_metrics) * 1e-3),
        'note': f'one epoch = {steps_per_epoch} learner steps ({n_envs * T} update samples) + one forward-only pass over the same {n_envs * T} samples'}
    # PER sum-tree: 2^20 leaves, batch 512 (dependent-access latency bound; GB/s is not meaningful)
    cap, bs = 1 << 20, 512
    tree = SumTree(cap, device=dev)
    tree.set_leaves(tc.rand(cap, device=dev, generator=g) + 1e-3)
    idx = tc.randint(0, cap, (bs,), device=dev, generator=g)
    pr = tc.rand(bs, device=dev, generator=g) + 1e-3
    u01 = tc.rand(bs, device=dev, generator=g)
    ms_u = _timed(lambda: tree.update(idx, pr), iters=50)
    ms_s = _timed(lambda: tree.sample(tree.stratified_u(u01)), iters=50)
    out['sumtree'] = {'leaves': cap, 'batch': bs, 'update_us': ms_u * 1e3, 'sample_us': ms_s * 1e3,
                      'updates_per_s': bs / (ms_u * 1e-3), 'samples_per_s': bs / (ms_s * 1e-3)}
    return out


def run_e2e(algo, agent, rollout, st, n_envs, dev, world, args, rows_list):
    """Same metric through the public API with HOST minibatches: every step copies that step's
    minibatch (frames + per-sample scalars, pinned host memory) to the device, runs
    PPO.learner_step on it and reads the losses back to the host.  Copies are double-buffered on a
    side stream so that step s+1's upload overlaps step s's compute; everything is inside the timed
    region."""
    import torch.distributed as dist
    from pytorch_drl_b200.utils.nested import MinibatchView
    nb = n_envs * B
    nbuf = 2
    host, devb = [], []
    names = ['observations', 'actions', 'logprobs']
    for b in range(nbuf):
        rows = rows_list[b]
        h = {
            'observations': rollout['observations'].view(-1, 84, 84, 4)[rows].cpu().pin_memory(),
            'actions': rollout['actions'].view(-1)[rows].cpu().pin_memory(),
            'logprobs': rollout['logprobs'].view(-1)[rows].cpu().pin_memory(),
            'advantages': rollout['advantages']['extrinsic'].reshape(-1)[rows].cpu().pin_memory(),
            'vpreds': rollout['vpreds']['extrinsic'].reshape(-1)[rows].cpu().pin_memory(),
            'vtargs': rollout['vtargs']['extrinsic'].reshape(-1)[rows].cpu().pin_memory(),
        }
        host.append(h)
        devb.append({k: tc.empty(v.shape, dtype=v.dtype, device=dev) for k, v in h.items()})
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())
    ident = tc.arange(nb, device=dev, dtype=tc.int64)
    loss_host = tc.empty((args.steps + args.warmup, 5), dtype=tc.float32).pin_memory()
    copy_stream = tc.cuda.Stream(device=dev)
    main = tc.cuda.current_stream()
    ready = [tc.cuda.Event() for _ in range(nbuf)]
    freed = [tc.cuda.Event() for _ in range(nbuf)]

    def upload(i):
        b = i % nbuf
        with tc.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[b])
            for k in host[b]:
                devb[b][k].copy_(host[b][k], non_blocking=True)
            ready[b].record(copy_stream)

    def view(b):
        dd = devb[b]
        return MinibatchView({'observations': dd['observations'].view(1, nb, 84, 84, 4), 'actions': dd['actions'].view(1, nb),
                              'logprobs': dd['logprobs'].view(1, nb), 'advantages': {'extrinsic': dd['advantages'].view(1, nb)},
                              'vpreds': {'extrinsic': dd['vpreds'].view(1, nb)}, 'vtargs': {'extrinsic': dd['vtargs'].view(1, nb)}},
                             ident)

    def run(n, offset):
        upload(0)
        for i in range(n):
            b = i % nbuf
            if i + 1 < n:
                upload(i + 1)
            main.wait_event(ready[b])
            losses = algo.learner_step(view(b))
            freed[b].record(main)
            loss_host[offset + i].copy_(losses, non_blocking=True)
        tc.cuda.synchronize()

    for b in range(nbuf):
        freed[b].record(main)
    run(max(3, min(args.warmup, 5)), 0)
    if world > 1:
        dist.barrier()
    tc.cuda.synchronize()
    ev0, ev1 = tc.cuda.Event(enable_timing=True), tc.cuda.Event(enable_timing=True)
    ev0.record()
    run(args.steps, args.warmup)
    ev1.record()
    tc.cuda.synchronize()
    el = ev0.elapsed_time(ev1)
    if world > 1:
        t = tc.tensor([el], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        el = float(t.item())
    assert bool(tc.isfinite(loss_host[args.warmup:args.warmup + args.steps]).all())
    return {'value': nb * world / (el / args.steps * 1e-3), 'unit': 'samples/s', 'h2d_bytes_per_step': int(h2d),
            'd2h_bytes_per_step': 20, 'ms_per_step': el / args.steps,
            'api': 'PPO.learner_step(MinibatchView) on pinned-host minibatches, double-buffered uploads'}


if __name__ == '__main__':
    main()
