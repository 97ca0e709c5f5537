"""Multi-GPU plumbing (one process per GPU, torch.distributed): individuals are sharded in contiguous blocks, the model is
replicated, and each pass has exactly one exchange --

  * resolveAll:   all-gather of the per-individual results (P(G), K sampled pairs, priors) so that every rank mines the same
                  samples and sums LL in individual order;
  * exact E-pass: all-reduce (sum) of the raw 2*Pc fp64 count vector before the replicated normalisation
                  (HaploBuilder.cpp:317-331).

The collectives are installed into the host HaploModel as C callbacks (hmch_set_collectives); the compute path is unchanged.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from .host import HaploModel, lib as host_lib


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous balanced block of individuals of `rank`: the first n % world ranks get one more."""
    base, extra = divmod(n, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def _dev(t: torch.Tensor) -> torch.Tensor:
    return t.cuda() if dist.get_backend() == "nccl" else t


def allreduce_sum_(buf: np.ndarray) -> None:
    """In-place sum over ranks of a host fp64 vector (NCCL: staged through this rank's GPU)."""
    t = _dev(torch.from_numpy(buf))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if t.device.type != "cpu":
        buf[...] = t.cpu().numpy()


def allgather_blocks(send: np.ndarray, n: int, rank: int, world: int) -> np.ndarray:
    """send: this rank's records [count, elem]; returns all n records in individual order on every rank."""
    elem = send.shape[1]
    cmax = (n + world - 1) // world
    pad = np.zeros((cmax, elem), np.uint8)
    pad[:send.shape[0]] = send
    out = [_dev(torch.empty((cmax, elem), dtype=torch.uint8)) for _ in range(world)]
    dist.all_gather(out, _dev(torch.from_numpy(pad)))
    full = np.empty((n, elem), np.uint8)
    for r in range(world):
        f, c = shard_range(n, r, world)
        full[f:f + c] = out[r][:c].cpu().numpy()
    return full


_REDUCE_T = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_size_t)
_GATHER_T = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_size_t)


def attach(hm: HaploModel) -> tuple[int, int]:
    """Shard `hm` over the default process group; returns this rank's (first, count).  Call after set_panel / load_phase."""
    rank, world = dist.get_rank(), dist.get_world_size()
    n = hm.N
    first, count = shard_range(n, rank, world)

    def reduce_cb(ptr, cnt):
        buf = np.ctypeslib.as_array(ptr, shape=(cnt,))
        allreduce_sum_(buf)

    def gather_cb(send, recv, elem_bytes):
        s = np.ctypeslib.as_array(C.cast(send, C.POINTER(C.c_uint8)), shape=(count, elem_bytes)) if count else np.zeros((0, elem_bytes), np.uint8)
        full = allgather_blocks(s, n, rank, world)
        C.memmove(recv, full.ctypes.data, full.nbytes)

    hm._cb = (_REDUCE_T(reduce_cb), _GATHER_T(gather_cb))      # keep the thunks alive
    L = host_lib()
    L.hmch_set_shard.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.hmch_set_collectives.argtypes = [C.c_void_p, _REDUCE_T, _GATHER_T]
    hm._ck(L.hmch_set_shard(hm._h, first, count))
    hm._ck(L.hmch_set_collectives(hm._h, hm._cb[0], hm._cb[1]))
    return first, count
