"""ctypes binding of `include/drl_b200.h` (the C-ABI library built by `pytorch_drl_b200.build`).

There is NO CPU fallback: if the library is missing, fails to load, or no sm_100 device is
visible, every product entry point raises.  Status codes map to the exceptions the reference
raises in the same situations: bad shape/argument/alignment -> ValueError,
everything else -> RuntimeError.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int64, c_uint8,
                    c_uint64, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libdrl_b200.so')

MAX_STREAMS = 4
MAX_ACTIONS = 32
PRECISION_FP32, PRECISION_BF16 = 0, 1
NCCL_ID_BYTES = 128
IPC_HANDLE_BYTES = 64
MAX_P2P_RANKS = 8


class PpoHyper(Structure):
    _fields_ = [('num_actions', c_int), ('num_streams', c_int),
                ('reward_weights', c_float * MAX_STREAMS), ('clip_param', c_float),
                ('ent_coef', c_float), ('vf_coef', c_float), ('vf_loss_clipping', c_int),
                ('vf_simple_weighting', c_int), ('vf_loss_kind', c_int)]


class PpoBatch(Structure):
    _fields_ = [('actions', c_void_p), ('logp_old', c_void_p), ('adv', c_void_p * MAX_STREAMS),
                ('v_old', c_void_p * MAX_STREAMS), ('vtarg', c_void_p * MAX_STREAMS)]


# name -> (restype, argtypes); every symbol `include/drl_b200.h` declares
PROTOTYPES = {
    'drl_abi_version': (c_int, []),
    'drl_status_string': (c_char_p, [c_int]),
    'drl_last_cuda_error': (c_char_p, []),
    'drl_check_device': (c_int, [c_int]),
    'drl_launch_count': (ctypes.c_longlong, []),
    'drl_profile_enable': (c_int, [c_int]),
    'drl_debug_set': (c_int, [c_int, c_int]),
    'drl_profile_read': (c_int, [c_int, ctypes.c_char_p, POINTER(c_float), POINTER(c_int), POINTER(c_int)]),
    'drl_gae_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                            c_float, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'drl_standardize_f32': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'drl_ppo_loss_fwd_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, POINTER(PpoBatch),
                                     POINTER(PpoHyper), c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p]),
    'drl_net_create': (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int]),
    'drl_net_destroy': (c_int, [c_void_p]),
    'drl_net_set_comm': (c_int, [c_void_p, c_void_p]),
    'drl_net_backward': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'drl_net_backward_features': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    'drl_net_param_count': (c_int64, [c_int, c_int]),
    'drl_net_param_offsets': (c_int, [c_int, c_int, POINTER(c_int64), c_int]),
    'drl_net_set_params': (c_int, [c_void_p, c_void_p, c_void_p]),
    'drl_net_forward': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p]),
    'drl_net_features': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    'drl_net_ppo_step': (c_int, [c_void_p, c_void_p, c_void_p, c_int, POINTER(PpoBatch),
                                 POINTER(PpoHyper), c_void_p, c_void_p, c_void_p]),
    'drl_net_ppo_metrics': (c_int, [c_void_p, c_void_p, c_void_p, c_int, POINTER(PpoBatch),
                                    POINTER(PpoHyper), c_void_p, c_void_p]),
    'drl_net_activation': (c_int, [c_void_p, c_int, POINTER(c_void_p), POINTER(c_int64), POINTER(c_int)]),
    'drl_memcpy_d2d': (c_int, [c_void_p, c_void_p, ctypes.c_size_t, c_void_p]),
    'drl_copy_by_kernel': (c_int, [c_void_p, c_void_p, ctypes.c_size_t, c_void_p]),
    'drl_adam_step': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double,
                              c_double, c_double, c_int64, c_float, c_void_p]),
    'drl_net_adam_step': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double,
                                  c_double, c_double, c_int64, c_float, c_void_p]),
    'drl_normalizer_apply_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_float,
                                         c_float, c_float, c_void_p, c_void_p]),
    'drl_normalizer_update_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_double,
                                          c_void_p]),
    'drl_rnd_create': (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int]),
    'drl_rnd_destroy': (c_int, [c_void_p]),
    'drl_rnd_param_count': (c_int64, [c_int, c_int]),
    'drl_rnd_set_params': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    'drl_rnd_adam_step': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double, c_double, c_double,
                                  c_int64, c_float, c_void_p]),
    'drl_rnd_reward': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                               c_void_p]),
    'drl_rnd_learn': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p]),
    'drl_sumtree_update': (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p]),
    'drl_sumtree_rebuild': (c_int, [c_void_p, c_int64, c_void_p]),
    'drl_sumtree_sample': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    'drl_sumtree_stratified_u': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    'drl_per_sample': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    'drl_ring_gather': (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    'drl_probe_dependent_latency': (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    'drl_sample_actions_f32': (c_int, [c_void_p, c_int64, c_int, c_void_p, c_uint64, c_uint64, c_void_p, c_void_p, c_void_p]),
    'drl_net_act': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_uint64, c_uint64, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_void_p]),
    'drl_rollout_record': (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_void_p), c_int, c_int, c_int64, c_int, c_void_p,
                                   c_void_p, c_void_p, POINTER(c_void_p), c_int, c_void_p]),
    'drl_rollout_stack_frames': (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'drl_rollout_copy_steps': (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_int, c_int, c_int, c_void_p]),
    'drl_nstep_advantages_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_int64,
                                         c_float, c_int, c_void_p, c_void_p]),
    'drl_bellman_backups_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int64,
                                        c_int64, c_int64, c_float, c_int, c_int, c_void_p, c_void_p]),
    'drl_dueling_forward_f32': (c_int, [c_void_p, c_int, c_int, c_int, c_int] + [c_void_p] * 10 + [c_void_p, c_void_p]),
    'drl_dueling_loss_backward_f32': (c_int, [c_void_p, c_int, c_int, c_int, c_int] + [c_void_p] * 7 + [c_float, c_int, c_int, c_int] +
                                      [c_void_p] * 8),
    'drl_epsilon_greedy_probs_f32': (c_int, [c_void_p, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p]),
    'drl_return_acc_update_f32': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_float, c_int, c_int,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'drl_reward_norm_apply_f32': (c_int, [c_void_p, c_int64, c_void_p, c_float, c_float, c_float, c_void_p, c_void_p]),
    'drl_comm_unique_id': (c_int, [POINTER(c_uint8)]),
    'drl_comm_create': (c_int, [POINTER(c_void_p), POINTER(c_uint8), c_int, c_int, c_int]),
    'drl_comm_destroy': (c_int, [c_void_p]),
    'drl_comm_p2p_export': (c_int, [c_void_p, c_int64, POINTER(c_uint8)]),
    'drl_comm_p2p_import': (c_int, [c_void_p, POINTER(c_uint8)]),
    'drl_grad_allreduce_bucketed': (c_int, [c_void_p, c_void_p, POINTER(c_int64), c_int, c_void_p]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Loads the CUDA library; raises RuntimeError (never falls back) when it is unavailable."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: build it with `python -m pytorch_drl_b200.build` '
                '(there is no CPU fallback for the drl_b200 learner path)')
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)          # AttributeError if the ABI lost a symbol
            fn.restype, fn.argtypes = res, args
        if handle.drl_abi_version() != 1:
            raise RuntimeError('drl_b200 ABI version mismatch')
        _lib = handle
    return _lib


def check(status: int, what: str = '') -> None:
    if status == 0:
        return
    L = lib()
    msg = L.drl_status_string(status).decode()
    if status == 4:
        msg += ': ' + L.drl_last_cuda_error().decode()
    msg = f'drl_b200 {what}: {msg}' if what else f'drl_b200: {msg}'
    if status in (1, 2, 3):
        raise ValueError(msg)
    raise RuntimeError(msg)


def require_device(device_index: int = 0) -> None:
    check(lib().drl_check_device(device_index), 'device check')


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
