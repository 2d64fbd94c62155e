"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the header
declares, the ctypes prototypes cover the header, the host-side drop-in logic (samplers,
schedules, reward bookkeeping, Agent construction / state_dict names) behaves like the
reference's, and the product path fails loudly without a GPU (no CPU fallback)."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch as tc

import pytorch_drl_b200 as pkg
from pytorch_drl_b200 import _lib, agents, algos
from pytorch_drl_b200.engine import param_names

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'drl_b200.h')


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(drl_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_header_symbol():
    assert os.path.exists(_lib.LIB_PATH), 'build with python -m pytorch_drl_b200.build'
    out = subprocess.check_output(['nm', '-D', '--defined-only', _lib.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if ' T ' in l}
    syms = header_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if s not in exported]
    assert not missing, missing


def test_ctypes_prototypes_cover_header():
    assert sorted(_lib.PROTOTYPES.keys()) == header_symbols()
    L = _lib.lib()
    assert L.drl_abi_version() == 1
    assert L.drl_status_string(1) == b'bad shape'


def test_param_layout_matches_reference_agent():
    L = _lib.lib()
    assert L.drl_net_param_count(6, 1) == 1687719
    assert L.drl_net_param_count(18, 2) == 1693875 + 513
    import ctypes
    offs = (ctypes.c_int64 * 13)()
    assert L.drl_net_param_offsets(6, 1, offs, 13) == 0
    assert list(offs)[:9] == [0, 8192, 8224, 40992, 41056, 77920, 77984, 1683616, 1684128]
    assert L.drl_net_param_offsets(6, 1, offs, 3) == 1     # bad shape -> ValueError in the shim
    with pytest.raises(ValueError):
        _lib.check(1, 'x')
    with pytest.raises(RuntimeError):
        _lib.check(5, 'x')


def test_no_device_means_loud_failure_not_fallback():
    if tc.cuda.is_available():
        pytest.skip('GPU present')
    assert _lib.lib().drl_check_device(0) == 5
    with pytest.raises(RuntimeError):
        _lib.require_device(0)
    ag = agents.Agent([agents.ToChannelMajor()], agents.NatureCNN(4),
                      {'policy': agents.CategoricalPolicyHead(512, 6), 'value_extrinsic': agents.SimpleValueHead(512)})
    with pytest.raises(RuntimeError):
        ag.fuse(8)
    with pytest.raises(RuntimeError):
        ag(tc.zeros(1, 84, 84, 4, dtype=tc.uint8), ['policy'])
    from pytorch_drl_b200 import ops
    with pytest.raises(ValueError):
        ops.standardize(tc.arange(10.0))               # CPU tensor is rejected, never computed


def test_agent_state_dict_names_match_reference_layout():
    ag = agents.Agent([agents.ToChannelMajor()], agents.NatureCNN(4),
                      {'policy': agents.CategoricalPolicyHead(512, 6),
                       'value_extrinsic': agents.SimpleValueHead(512),
                       'value_intrinsic_rnd': agents.SimpleValueHead(512)})
    assert [n for n, _ in ag.named_parameters()] == param_names(['value_extrinsic', 'value_intrinsic_rnd'])
    with pytest.raises(ValueError):
        agents.Agent([], agents.NatureCNN(4), {'policy': agents.CategoricalPolicyHead(512, 6),
                                                 'bogus': agents.SimpleValueHead(512)})


def test_sampler_disjoint_cover_and_seed_stream():
    """drl/algos/common/test_samplers.py: blocks are disjoint and cover range(T); the stream
    equals np.random.permutation under the reference's seeding rule."""
    rs = np.random.RandomState(10000 * 3 + 0)
    s = algos.IndependentSampler(128, 32, rs)
    s.resample()
    blocks = list(iter(s))
    assert len(blocks) == 4 and all(b.shape == (32,) for b in blocks)
    assert sorted(np.concatenate(blocks).tolist()) == list(range(128))
    ref = np.random.RandomState(10000 * 3 + 0)
    ref.permutation(128)
    np.testing.assert_array_equal(np.concatenate(blocks), ref.permutation(128))
    with pytest.raises(AssertionError):
        algos.IndependentSampler(100, 32)              # samplers.py:51


def test_linear_schedule():
    s = algos.LinearSchedule(0.2, 0.0, 100)             # schedules.py:18-29
    assert s.value(0) == 0.2 and s.value(100) == 0.0 and abs(s.value(50) - 0.1) < 1e-12
    assert s.value(1000) == 0.0


def test_credit_op_registry():
    op = algos.get_credit_assignment_op('GAE', dict(gamma=0.99, lambda_=0.95, use_dones=True))
    assert isinstance(op, algos.GAE)
    with pytest.raises(ValueError):
        algos.get_credit_assignment_op('Nope', {})     # credit_assignment.py:25


def test_dqn_head_dropins_keep_reference_names_and_refuse_cpu(golden_dir):
    """SimpleDiscreteActionValueHead(DuelingArchitecture) loads the reference's state_dict with strict=True (the golden
    file holds the reference module's own tensors), mirrors its constructor errors, and never computes on the CPU."""
    from oracle import dueling_oracle as do
    g = np.load(os.path.join(golden_dir, 'dueling.npz'))
    head = agents.SimpleDiscreteActionValueHead(num_features=512, num_actions=6, head_architecture_cls=agents.DuelingArchitecture,
                                                head_architecture_cls_args={'widening': 1},
                                                w_init=lambda w: tc.nn.init.orthogonal_(w, gain=2 ** 0.5), b_init=tc.nn.init.zeros_)
    arch = head._action_value_head
    assert sorted(arch.state_dict().keys()) == sorted(do.PARAM_KEYS)
    arch.load_state_dict({k: tc.from_numpy(g[f'c0_p_{k}']) for k in do.PARAM_KEYS}, strict=True)
    assert arch.input_shape == [512] and arch.output_dim == 6 and head.num_actions == 6 and head.num_features == 512
    pol = agents.EpsilonGreedyCategoricalPolicyHead(action_value_head=head)
    with pytest.raises(AssertionError):
        pol.epsilon = 1.5                                             # policy_heads.py:185-188
    pol.epsilon = 0.1
    with pytest.raises(RuntimeError):
        head(tc.zeros(2, 512))                                        # CPU features: refused, not computed
    with pytest.raises(ValueError):
        agents.MLP(4, 4, 4, 1)                                        # mlp.py:30-31


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the arm the driver times next to ours): one JSON line with the contract's keys, run
    on the host CPU with every core, without touching CUDA."""
    import json
    import sys
    env = dict(os.environ, OMP_NUM_THREADS='1', CUDA_VISIBLE_DEVICES='')       # what torchrun would export
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                                  text=True, env=env, timeout=600)
    line = json.loads([l for l in out.splitlines() if l.startswith('{')][-1])
    assert line['impl'] == 'reference' and line['unit'] == 'samples/s' and line['higher_is_better'] is True
    assert line['value'] > 0 and line['steps'] == 1 and line['n_gpus'] >= 1
    from oracle import ref_bench
    # the UNMODIFIED reference when a copy travels with the repository (oracle/_ref, vendored by build()), else the oracle port
    assert line['cpu_baseline']['kind'] == ('reference' if ref_bench.reference_root() else 'port')
    assert 0 < line['cpu_baseline']['cores'] <= (os.cpu_count() or 1)
    assert line['e2e']['value'] == line['value'] and line['e2e']['h2d_bytes_per_step'] == 0
    assert 'workload' in line['config']


def test_nested_tensor_helpers_match_reference_tests():
    """drl/utils/test_nested.py:26-50: clone and slice (python slice and numpy index array) of a nested dict of tensors.
    Host tensors keep the reference's gather semantics; only a CUDA int64 index applied to a device rollout turns into
    a MinibatchView (tests/test_gpu_parity.py)."""
    from pytorch_drl_b200.utils.nested import clone_nested_tensor, slice_nested_tensor

    def make():
        return {'foo': tc.arange(10), 'bar': {'baz': tc.zeros(10, dtype=tc.float32)}}

    def same(a, b):
        if isinstance(a, tc.Tensor):
            tc.testing.assert_close(actual=a, expected=b)
        else:
            assert set(a.keys()) == set(b.keys())
            for k in a:
                same(a[k], b[k])

    nt = make()
    cl = clone_nested_tensor(nt)
    same(nt, cl)
    assert cl['foo'].data_ptr() != nt['foo'].data_ptr()
    same(slice_nested_tensor(make(), slice(0, 3)), {'foo': tc.arange(3), 'bar': {'baz': tc.zeros(3)}})
    same(slice_nested_tensor(make(), np.array([1, 2, 5])), {'foo': tc.tensor([1, 2, 5]), 'bar': {'baz': tc.zeros(3)}})


def test_state_dicts_round_trip_with_the_real_reference_modules():
    """f3 (checkpoint compatibility) against the UNMODIFIED reference where it is mounted (this container; skipped on
    the GPU box): the reference's Agent(NatureCNN + categorical policy + value heads) and its
    SimpleDiscreteActionValueHead(DuelingArchitecture) exchange state_dicts with the drop-ins in both directions with
    strict=True — same keys, same shapes, same tensor layouts (drl/utils/checkpointing.py:64 loads exactly these)."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip('reference tree not mounted')
    ref_loader.load()
    from drl.agents.architectures import NatureCNN as RNature, Linear as RLinear, DuelingArchitecture as RDueling
    from drl.agents.heads import CategoricalPolicyHead as RPol, SimpleValueHead as RVal
    from drl.agents.heads.action_value_heads import SimpleDiscreteActionValueHead as RQ
    from drl.agents.integration import Agent as RAgent
    from drl.agents.preprocessing import ToChannelMajor as RTCM
    w = lambda t: tc.nn.init.orthogonal_(t, gain=2 ** 0.5)       # noqa: E731
    b = tc.nn.init.zeros_
    tc.manual_seed(3)
    ref = RAgent(preprocessing=[RTCM()], architecture=RNature(img_channels=4, w_init=w, b_init=b),
                 predictors={'policy': RPol(num_features=512, num_actions=6, head_architecture_cls=RLinear,
                                            head_architecture_cls_args={}, w_init=w, b_init=b),
                             'value_extrinsic': RVal(num_features=512, head_architecture_cls=RLinear,
                                                     head_architecture_cls_args={}, w_init=w, b_init=b)})
    mine = agents.Agent([agents.ToChannelMajor()], agents.NatureCNN(4, w, b),
                        {'policy': agents.CategoricalPolicyHead(512, 6, w_init=w, b_init=b),
                         'value_extrinsic': agents.SimpleValueHead(512, w_init=w, b_init=b)})
    rsd, msd = ref.state_dict(), mine.state_dict()
    assert list(rsd.keys()) == list(msd.keys())
    assert [tuple(v.shape) for v in rsd.values()] == [tuple(v.shape) for v in msd.values()]
    mine.load_state_dict(rsd, strict=True)
    for k in rsd:
        assert tc.equal(mine.state_dict()[k], rsd[k])
    ref.load_state_dict(msd, strict=True)
    # DQN head
    rq = RQ(num_features=512, num_actions=18, head_architecture_cls=RDueling, head_architecture_cls_args={'widening': 1},
            w_init=w, b_init=b)
    mq = agents.SimpleDiscreteActionValueHead(512, 18, agents.DuelingArchitecture, {'widening': 1}, w_init=w, b_init=b)
    assert list(rq.state_dict().keys()) == list(mq.state_dict().keys())
    mq.load_state_dict(rq.state_dict(), strict=True)
    rq.load_state_dict(mq.state_dict(), strict=True)


def test_checkpoint_objects_round_trip_through_the_reference_checkpointing(tmp_path):
    """f2 / f3 against the UNMODIFIED reference (skipped where it is not mounted): the checkpointables the fused path hands to
    `save_checkpoints` (drl/utils/checkpointing.py:119-143; ppo.py:295-307; random_network_distillation.py:175-184) load into
    the reference's own DDP-wrapped networks and torch.optim.Adam with `maybe_load_checkpoints` (:85-117), and files written
    from the reference's objects load back — same kinds, same keys (incl. the `module.` prefix), same tensors."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip('reference tree not mounted')
    ref_loader.load()
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from drl.utils.checkpointing import save_checkpoints as ref_save, maybe_load_checkpoints as ref_load
    from drl.envs.wrappers.stateful.intrinsic.random_network_distillation import _StudentNetwork, _TeacherNetwork
    from drl.envs.wrappers.stateful.normalize_reward import ReturnAcc
    from pytorch_drl_b200.envs.rnd import STUDENT_KEYS, STUDENT_SHAPES, TEACHER_KEYS, TEACHER_SHAPES, _views
    from pytorch_drl_b200.utils import checkpoint as ck
    created = False
    if not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        import socket
        sock = socket.socket(); sock.bind(('127.0.0.1', 0)); free_port = sock.getsockname()[1]; sock.close()
        os.environ['MASTER_PORT'] = str(free_port)
        dist.init_process_group('gloo', rank=0, world_size=1)
        created = True
    try:
        tc.manual_seed(5)
        # ---- reference objects with some training history
        r_student, r_teacher = DDP(_StudentNetwork([84, 84, 1], 1)), DDP(_TeacherNetwork([84, 84, 1]))
        r_opt = tc.optim.Adam(r_student.parameters(), lr=1e-4, eps=1e-5)
        for _ in range(3):
            r_opt.zero_grad()
            r_student(tc.randn(2, 84, 84, 1)).square().mean().backward()
            r_opt.step()
        # ---- mine: flat buffers presented through the checkpointables
        n_s = sum(int(np.prod(s)) for s in STUDENT_SHAPES)
        n_t = sum(int(np.prod(s)) for s in TEACHER_SHAPES)
        flat_s, flat_t, m, v = tc.zeros(n_s), tc.zeros(n_t), tc.zeros(n_s), tc.zeros(n_s)
        steps = {'n': 0}
        mine = {'rnd_student_net': ck.ViewsCheckpointable(lambda: _views(flat_s, STUDENT_KEYS, STUDENT_SHAPES)),
                'rnd_teacher_net': ck.ViewsCheckpointable(lambda: _views(flat_t, TEACHER_KEYS, TEACHER_SHAPES)),
                'rnd_optimizer': ck.FlatAdamCheckpointable(STUDENT_SHAPES, m, v, lambda: steps['n'],
                                                           lambda n: steps.__setitem__('n', n),
                                                           lambda: dict(lr=1e-4, betas=(0.9, 0.999), eps=1e-5))}
        ref = {'rnd_student_net': r_student, 'rnd_teacher_net': r_teacher, 'rnd_optimizer': r_opt}
        # reference -> files -> mine (my own loader and the reference's loader agree)
        d1 = str(tmp_path / 'from_ref')
        ref_save(d1, ref, steps=128)
        assert ref_load(d1, mine, 'cpu', None) == 128
        assert ck.maybe_load_checkpoints(d1, mine) == 128
        assert list(mine['rnd_student_net'].state_dict().keys()) == list(r_student.state_dict().keys())
        for k, t in r_student.state_dict().items():
            assert tc.equal(mine['rnd_student_net'].state_dict()[k], t)
        assert steps['n'] == 3
        for i, p in enumerate(r_student.parameters()):
            assert tc.equal(mine['rnd_optimizer'].state_dict()['state'][i]['exp_avg'], r_opt.state[p]['exp_avg'])
        # mine -> files -> fresh reference objects
        d2 = str(tmp_path / 'from_mine')
        ck.save_checkpoints(d2, mine, steps=256)
        f_student, f_teacher = DDP(_StudentNetwork([84, 84, 1], 1)), DDP(_TeacherNetwork([84, 84, 1]))
        f_opt = tc.optim.Adam(f_student.parameters(), lr=1e-4, eps=1e-5)
        assert ref_load(d2, {'rnd_student_net': f_student, 'rnd_teacher_net': f_teacher, 'rnd_optimizer': f_opt}, 'cpu', None) == 256
        for (k, a), (_, b) in zip(f_student.state_dict().items(), r_student.state_dict().items()):
            assert tc.equal(a, b), k
        for p, q in zip(f_student.parameters(), r_student.parameters()):
            assert tc.equal(f_opt.state[p]['exp_avg_sq'], r_opt.state[q]['exp_avg_sq']) and float(f_opt.state[p]['step']) == 3
        # the resumed reference optimizer continues exactly like the original
        x = tc.randn(2, 84, 84, 1)
        for net, opt in ((r_student, r_opt), (f_student, f_opt)):
            opt.zero_grad()
            net(x).square().mean().backward()
            opt.step()
        for a, b in zip(f_student.parameters(), r_student.parameters()):
            assert tc.equal(a, b)
        # misaligned steps raise like the reference (checkpointing.py:114-116)
        ck.save_checkpoints(d2, {'rnd_teacher_net': mine['rnd_teacher_net']}, steps=512)
        with pytest.raises(RuntimeError):
            ck.maybe_load_checkpoints(d2, mine)
        # policy network: the fused Agent under a DDPShim has the keys of the reference's DDP(Agent)
        w = lambda t: tc.nn.init.orthogonal_(t, gain=2 ** 0.5)       # noqa: E731
        ag = agents.Agent([agents.ToChannelMajor()], agents.NatureCNN(4, w, tc.nn.init.zeros_),
                          {'policy': agents.CategoricalPolicyHead(512, 6), 'value_extrinsic': agents.SimpleValueHead(512)})
        keys = list(agents.DDPShim(ag).state_dict().keys())
        assert all(k.startswith('module.') for k in keys) and keys[0] == 'module._architecture._network.0.weight'
        ag.load_state_dict(agents.DDPShim(ag).state_dict())          # the prefixed form loads into the plain Agent
        # reward normaliser keys
        assert set(ReturnAcc(0.99, -10, 10, True).state_dict().keys()) == {'_normalizer._steps', '_normalizer._moment1', '_normalizer._moment2'}
    finally:
        if created:
            dist.destroy_process_group()


def test_register_resolves_fused_classes_through_the_reference_lookups():
    """INTEGRATION.md §3 end to end (reference mounted; skipped elsewhere): after `pytorch_drl_b200.register()` the reference's
    OWN lookup functions — script.py:55-58 `get_algo_cls`, launch.py:174-176 `get_architecture_cls`, launch.py:140-142
    `get_preprocessing`, credit_assignment.py:40-42 `get_credit_assignment_op(s)` — resolve the Fused* names to the drop-ins,
    build them from YAML-style cls_args, and the reference's own names still resolve to the reference's classes."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip('reference tree not mounted')
    drl = ref_loader.load()
    import importlib.util
    assert pkg.register() is drl
    spec = importlib.util.spec_from_file_location('ref_script', os.path.join(ref_loader.REFERENCE_ROOT, 'script.py'))
    script = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(script)                                  # defines get_algo_cls; `main` is guarded
    from drl.utils import launch
    from drl.algos.common import credit_assignment as ref_ca
    from drl.algos import PPO as RefPPO
    assert script.get_algo_cls('FusedPPO') is algos.PPO and script.get_algo_cls('PPO') is RefPPO
    assert launch.get_architecture_cls('FusedNatureCNN') is agents.NatureCNN
    assert launch.get_architecture_cls('NatureCNN') is not agents.NatureCNN
    assert isinstance(launch.get_preprocessing('FusedToChannelMajor', {}), agents.ToChannelMajor)
    ops_ = ref_ca.get_credit_assignment_ops({'extrinsic': {'cls_name': 'FusedGAE', 'cls_args': {'gamma': 0.99, 'lambda_': 0.95, 'use_dones': True}},
                                             'intrinsic_rnd': {'cls_name': 'FusedGAE', 'cls_args': {'gamma': 0.99, 'lambda_': 0.95, 'use_dones': False}}})
    assert all(isinstance(o, algos.GAE) for o in ops_.values()) and isinstance(ref_ca.get_credit_assignment_op('GAE', dict(gamma=0.9, lambda_=0.9, use_dones=True)), ref_ca.GAE)
    # the architecture / head factories of launch.py feed YAML cls_args + initialisers verbatim: the drop-ins accept them
    import importlib
    heads = importlib.import_module('drl.agents.heads')
    arch = launch.get_architecture_cls('FusedNatureCNN')(img_channels=4, w_init=lambda w: tc.nn.init.orthogonal_(w, gain=1.4142),
                                                         b_init=tc.nn.init.zeros_)
    pol = heads.FusedCategoricalPolicyHead(num_features=512, num_actions=6, head_architecture_cls=launch.get_architecture_cls('FusedLinear'),
                                           head_architecture_cls_args={}, w_init=None, b_init=None)
    val = heads.FusedSimpleValueHead(num_features=512, head_architecture_cls=launch.get_architecture_cls('FusedLinear'),
                                     head_architecture_cls_args={}, w_init=None, b_init=None)
    ag = agents.Agent([launch.get_preprocessing('FusedToChannelMajor', {})], arch, {'policy': pol, 'value_extrinsic': val})
    assert list(ag.state_dict().keys()) == param_names(['value_extrinsic'])


def test_in_tree_library_was_built_from_the_current_sources():
    """The in-tree .so travels to the GPU box as is: a stale one would silently test old kernels there."""
    import json
    from pytorch_drl_b200 import build as b
    stamp = os.path.join(b.OBJ, 'stamp.json')
    if not (os.path.exists(b.LIB) and os.path.exists(stamp)):
        pytest.skip('library not built in this tree yet')
    assert json.load(open(stamp)).get('digest') == b._digest(), 'run `python -m pytorch_drl_b200.build`: csrc/ or include/ changed since the build'
