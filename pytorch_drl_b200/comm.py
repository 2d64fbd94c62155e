"""`Communicator`: the data-parallel plumbing that replaces the reference's
DistributedDataParallel wrapper (drl/utils/launch.py:310) and `global_mean`
(drl/algos/common/metrics.py:8-26).  One process per GPU; the NCCL communicator lives in the
C-ABI library and is bootstrapped with a unique id broadcast through torch.distributed."""
import ctypes
from typing import List, Optional

import torch as tc
import torch.distributed as dist

from . import _lib
from ._lib import check, current_stream, ptr


def shard_envs(n_envs_total: int, rank: int, world_size: int) -> range:
    """GPU `rank` owns envs [rank*n/world, (rank+1)*n/world) — the reference's rank = env model with
    several envs per process.  Requires an even split."""
    if n_envs_total % world_size != 0:
        raise ValueError('n_envs_total must be divisible by world_size')
    per = n_envs_total // world_size
    return range(rank * per, (rank + 1) * per)


def bootstrap_unique_id(rank: int, world_size: int, device_index: int = 0) -> bytes:
    """Rank 0 creates the NCCL unique id, every rank receives it through torch.distributed
    (works over gloo or nccl process groups)."""
    L = _lib.lib()
    ident = (ctypes.c_uint8 * _lib.NCCL_ID_BYTES)()
    if rank == 0:
        check(L.drl_comm_unique_id(ident), 'comm_unique_id')
    t = tc.tensor(list(ident), dtype=tc.uint8)
    if world_size > 1:
        if not dist.is_initialized():
            raise RuntimeError('torch.distributed must be initialised to bootstrap the communicator')
        if dist.get_backend() == 'nccl':
            t = t.cuda(device_index)
        dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


class Communicator:
    def __init__(self, rank: int, world_size: int, device_index: int, p2p_elems: int = 0):
        """p2p_elems > 0 additionally sets up the NVLink peer-memory all-reduce for buffers of up to that many
        floats (drl_comm_p2p_*).  Measured on 2 B200s (bench.py --comm ...): NCCL 2.86 ms/step, peer memory 2.90,
        either one overlapped with the backward pass 2.95 — at 6.75 MB per step the exchange is bound by the
        rank-to-rank synchronisation, not by bandwidth, so NCCL after the backward pass stays the default."""
        self.rank, self.world = rank, world_size
        L = _lib.lib()
        ident = (ctypes.c_uint8 * _lib.NCCL_ID_BYTES)(*bootstrap_unique_id(rank, world_size, device_index))
        h = ctypes.c_void_p()
        check(L.drl_comm_create(ctypes.byref(h), ident, rank, world_size, device_index), 'comm_create')
        self._h, self._L = h, L
        self.p2p = False
        if p2p_elems and 1 < world_size <= _lib.MAX_P2P_RANKS:
            self._enable_p2p(device_index, int(p2p_elems))

    def _enable_p2p(self, device_index: int, max_elems: int) -> None:
        """All-reduce through NVLink peer memory (drl_comm_p2p_*): export my exchange buffer's IPC handle,
        all-gather the handles through torch.distributed, import everybody's, barrier.  If any rank cannot
        (no peer access), every rank stays on NCCL."""
        mine = (ctypes.c_uint8 * _lib.IPC_HANDLE_BYTES)()
        ok = self._L.drl_comm_p2p_export(self._h, max_elems, mine) == 0
        on_gpu = dist.get_backend() == 'nccl'
        t = tc.tensor(list(mine) + [1 if ok else 0], dtype=tc.uint8)
        if on_gpu:
            t = t.cuda(device_index)
        got = [tc.empty_like(t) for _ in range(self.world)]
        dist.all_gather(got, t)
        got = [g.cpu().tolist() for g in got]
        if all(g[-1] == 1 for g in got):
            flat = [b for g in got for b in g[:-1]]
            every = (ctypes.c_uint8 * len(flat))(*flat)
            ok = self._L.drl_comm_p2p_import(self._h, every) == 0
        else:
            ok = False
        flag = tc.tensor([1 if ok else 0], dtype=tc.int32)
        if on_gpu:
            flag = flag.cuda(device_index)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)          # also the barrier the protocol needs before first use
        self.p2p = bool(int(flag.item()))
        if not self.p2p:                                      # some rank has no peer access: everybody uses NCCL
            import warnings
            self._L.drl_comm_p2p_import(self._h, None)
            warnings.warn('peer-memory all-reduce unavailable on some rank; using NCCL')

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h:
            self._L.drl_comm_destroy(h)

    @staticmethod
    def grad_buckets(offsets: List[int]) -> List[int]:
        """Two buckets in reverse-backward order of availability: trunk convs
        [0, fc.weight) and fc + heads [fc.weight, end)."""
        return [0, offsets[6], offsets[-1]]

    def allreduce_(self, flat: tc.Tensor, bucket_offsets: Optional[List[int]] = None) -> tc.Tensor:
        offs = bucket_offsets or [0, flat.numel()]
        arr = (ctypes.c_int64 * len(offs))(*offs)
        check(self._L.drl_grad_allreduce_bucketed(self._h, ptr(flat), arr, len(offs) - 1, current_stream()),
              'grad_allreduce')
        return flat

    def allreduce_grads(self, engine) -> None:
        if getattr(self, 'single_bucket', False):
            self.allreduce_(engine.grads)
        else:
            self.allreduce_(engine.grads, self.grad_buckets(engine.offsets))

    def attach(self, engine) -> None:
        """Let the engine's ppo_step exchange its gradients itself, overlapped with the backward pass (the
        role of DDP's bucket hooks, launch.py:310): the fc + heads bucket starts as soon as the fc weight
        gradient is done, the conv bucket follows the last kernel."""
        check(self._L.drl_net_set_comm(engine._h, self._h), 'net_set_comm')
        engine.comm_attached = self

    def mean_(self, vec: tc.Tensor) -> tc.Tensor:
        """global_mean of a small fp32 CUDA vector (metrics.py:8-26), one collective."""
        v = vec.contiguous()
        self.allreduce_(v)
        return v / self.world
