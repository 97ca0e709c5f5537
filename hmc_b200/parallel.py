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
_REDUCE_DEV_T = C.CFUNCTYPE(None, C.c_void_p, C.c_size_t)
_GATHER_DEV_T = C.CFUNCTYPE(None, C.c_void_p, C.c_size_t)


class _DevBuf:
    """A context-owned device buffer as a zero-copy torch tensor (__cuda_array_interface__)."""

    def __init__(self, ptr: int, nbytes: int, typestr: str = "|u1", itemsize: int = 1):
        self.__cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_view(ptr: int, nbytes: int, dtype=torch.uint8, device: int = 0) -> torch.Tensor:
    if dtype == torch.float64:
        return torch.as_tensor(_DevBuf(ptr, nbytes, "<f8", 8), device=f"cuda:{device}")
    return torch.as_tensor(_DevBuf(ptr, nbytes), device=f"cuda:{device}")


def allgather_rows_(base: torch.Tensor, n: int, row_bytes: int, rank: int, world: int) -> None:
    """In-place all-gather over the individuals of a device byte tensor with one row of row_bytes per panel individual: rank r owns
    the rows of shard_range(n, r, world).  NCCL: equal shards go through one all_gather_into_tensor (in place), a ragged split
    (n % world != 0) through one broadcast per rank issued as one group.  Other backends (gloo in the tests) take the blocks one after
    the other through host copies of the device blocks."""
    nccl = dist.get_backend() == "nccl"
    if nccl and n % world == 0:
        per = (n // world) * row_bytes
        dist.all_gather_into_tensor(base[:n * row_bytes], base[rank * per:(rank + 1) * per])
        return
    works = []
    for r in range(world):
        f, c = shard_range(n, r, world)
        if not c:
            continue
        blk = base[f * row_bytes:(f + c) * row_bytes]
        if nccl:
            works.append(dist.broadcast(blk, src=r, async_op=True))
        else:
            h = blk.cpu()
            dist.broadcast(h, src=r)
            if r != rank:
                blk.copy_(h)
    for w in works:
        w.wait()


def allreduce_device_(t: torch.Tensor) -> None:
    """In-place sum over ranks of a device tensor: NCCL on the device memory itself; other backends through a host copy."""
    if dist.get_backend() == "nccl":
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    else:
        h = t.cpu()
        dist.all_reduce(h, op=dist.ReduceOp.SUM)
        t.copy_(h)


def attach(hm: HaploModel, device_collectives: bool | None = None) -> tuple[int, int]:
    """Shard `hm` over the default process group; returns this rank's (first, count).  Call after set_panel / load_phase.

    device_collectives (default: on when the backend is NCCL): the two exchanges run on the device memory of the context -- the raw
    count vector of an exact E-pass is all-reduced in place, the result rows of a resolve pass are all-gathered in place -- and the
    next pattern search reads the gathered rows where they lie.  Off: host-staged exchanges through numpy (gloo / CPU tests)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    n = hm.N
    if n < world:
        raise ValueError(f"{n} individuals cannot be sharded over {world} ranks")
    first, count = shard_range(n, rank, world)
    if device_collectives is None:
        device_collectives = dist.get_backend() == "nccl"

    def reduce_cb(ptr, cnt):
        buf = np.ctypeslib.as_array(ptr, shape=(cnt,))
        allreduce_sum_(buf)

    def gather_cb(send, recv, elem_bytes):
        s = np.ctypeslib.as_array(C.cast(send, C.POINTER(C.c_uint8)), shape=(count, elem_bytes)) if count else np.zeros((0, elem_bytes), np.uint8)
        full = allgather_blocks(s, n, rank, world)
        C.memmove(recv, full.ctypes.data, full.nbytes)

    dev_index = getattr(hm, "device_index", 0)

    def reduce_dev_cb(ptr, cnt):
        t = device_view(ptr, cnt * 8, torch.float64, dev_index)
        allreduce_device_(t)
        torch.cuda.synchronize(dev_index)

    def gather_dev_cb(ptr, row_bytes):
        t = device_view(ptr, n * row_bytes, torch.uint8, dev_index)
        allgather_rows_(t, n, row_bytes, rank, world)
        torch.cuda.synchronize(dev_index)

    hm._cb = (_REDUCE_T(reduce_cb), _GATHER_T(gather_cb), _REDUCE_DEV_T(reduce_dev_cb), _GATHER_DEV_T(gather_dev_cb))      # keep the thunks alive
    L = host_lib()
    L.hmch_set_shard.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.hmch_set_collectives.argtypes = [C.c_void_p, _REDUCE_T, _GATHER_T]
    L.hmch_set_device_collectives.argtypes = [C.c_void_p, _REDUCE_DEV_T, _GATHER_DEV_T]
    hm._ck(L.hmch_set_shard(hm._h, first, count))
    hm._ck(L.hmch_set_collectives(hm._h, hm._cb[0], hm._cb[1]))
    if device_collectives:
        hm._ck(L.hmch_set_device_collectives(hm._h, hm._cb[2], hm._cb[3]))
    else:
        hm._ck(L.hmch_set_device_collectives(hm._h, _REDUCE_DEV_T(0), _GATHER_DEV_T(0)))
    return first, count
