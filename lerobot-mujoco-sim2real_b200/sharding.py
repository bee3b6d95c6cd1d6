"""Env-level data parallelism: shard independent environments across ranks, deliver the dataset to one rank.

The step path has NO inter-GPU traffic (envs are independent).  The only exchange is the delivery of the
`[N/G, T+1, 13]` row shards to rank 0 (what `np.save` needs, [REF SOARM101/SOARM101_DataCollection.py:156]).
Control/reset random streams are keyed by the GLOBAL env index, so the dataset is bit-identical for any world size.

Two transports:
  * `SharedRows` (CUDA, all ranks on one node): rank `dst` owns the full `[N, T+1, 13]` buffer, the other ranks map it
    through CUDA IPC (`so101_shared_alloc/open`) and their `k_rollout` launches write their rows STRAIGHT into it over
    NVLink while they simulate - the gather is fused into the kernel's row writer, there is no collective, no staging
    buffer and no concatenation; ranks meet at one barrier at the end.
  * `gather_rows` (any backend; gloo in the CPU tests, NCCL across nodes): point-to-point sends of the (possibly ragged)
    shards into views of ONE preallocated `[N, ...]` tensor on `dst` - no padding, no `torch.cat`.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block `[start, stop)` of rank `rank`; the first `n_total % world` ranks get one extra."""
    if world <= 0 or not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(n_total: int, world: int) -> List[int]:
    return [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]


def dist_info() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def single_node() -> bool:
    """All ranks of the default group on this node (torchrun exports LOCAL_WORLD_SIZE)."""
    _, world = dist_info()
    return world == 1 or int(os.environ.get("LOCAL_WORLD_SIZE", "0")) == world


def gather_rows(local: torch.Tensor, n_total: int, dst: int = 0, out: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """Deliver row shards `[n_local, ...]` (rank-major, ragged by at most one row, empty shards allowed) into
    `[n_total, ...]` on rank `dst` (`out`, or a fresh tensor).  Returns None on the other ranks.  Single process:
    returns `local` (or `out` filled with it)."""
    rank, world = dist_info()
    if world == 1:
        if out is None:
            return local
        out.copy_(local)
        return out
    sizes = shard_sizes(n_total, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, expected {sizes[rank]}")
    local = local.contiguous()
    if rank == dst:
        if out is None:
            out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        assert out.is_contiguous() and out.shape[0] == n_total and out.shape[1:] == local.shape[1:]
        ops, lo = [], 0
        for r, n in enumerate(sizes):
            view = out[lo:lo + n]                 # contiguous slice of the destination: received in place
            if r == dst:
                view.copy_(local)
            elif n > 0:
                ops.append(dist.P2POp(dist.irecv, view, r))
            lo += n
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return out
    if sizes[rank] > 0:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, local, dst)]):
            req.wait()
    return None


class _DevPtr:
    """Raw device pointer as a `__cuda_array_interface__` object (zero-copy torch view of library-owned memory)."""

    def __init__(self, ptr: int, shape: Tuple[int, ...], typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class SharedRows:
    """`[n_total, *row_shape]` rows on rank `dst`, written in place by every rank's rollout kernel (CUDA IPC).

        sr = SharedRows(n_total, (T + 1, 13), torch.float64, device_index)
        env.rollout(..., out_ptr=sr.local_ptr)        # rank r writes rows [lo_r, hi_r) of the buffer on `dst`
        full = sr.finish()                            # stream sync + barrier; the [n_total, ...] tensor on dst, else None
        ...                                           # use `full` (a view of library-owned memory)
        sr.close()                                    # collective: unmap / free

    One process per GPU, all on one node (`single_node()`).  The handle exchange is one small host collective."""

    def __init__(self, n_total: int, row_shape: Tuple[int, ...], dtype: torch.dtype, device: int, dst: int = 0):
        from . import _lib
        self._lib = _lib
        self.rank, self.world = dist_info()
        self.dst, self.device, self.n_total = dst, int(device), int(n_total)
        self.row_shape, self.dtype = tuple(int(x) for x in row_shape), dtype
        esize = torch.empty((), dtype=dtype).element_size()
        self.row_bytes = esize
        for x in self.row_shape:
            self.row_bytes *= x
        self.lo, self.hi = shard_range(self.n_total, self.rank, self.world)
        self.nbytes = max(1, self.n_total * self.row_bytes)
        L = _lib.lib()
        self._base = C.c_void_p()
        self._owner = self.rank == dst
        handle = C.create_string_buffer(64)
        if self._owner:
            _lib.check(L.so101_shared_alloc(self.device, self.nbytes, C.byref(self._base), handle))
        if self.world > 1:
            box = [handle.raw if self._owner else None]
            dist.broadcast_object_list(box, src=dst)
            if not self._owner:
                _lib.check(L.so101_shared_open(self.device, box[0], C.byref(self._base)))
        self._typestr = {torch.float64: "<f8", torch.float32: "<f4", torch.int32: "<i4"}[dtype]

    @property
    def local_ptr(self) -> int:
        """where this rank's shard starts inside the buffer on `dst`"""
        return int(self._base.value) + self.lo * self.row_bytes

    def finish(self) -> Optional[torch.Tensor]:
        torch.cuda.current_stream(torch.device("cuda", self.device)).synchronize()   # my remote stores are complete
        if self.world > 1:
            dist.barrier()
        if not self._owner:
            return None
        return torch.as_tensor(_DevPtr(self._base.value, (self.n_total,) + self.row_shape, self._typestr),
                               device=torch.device("cuda", self.device))

    def close(self) -> None:
        if self._base.value is None:
            return
        L = self._lib.lib()
        if self.world > 1:
            dist.barrier()          # nobody unmaps / frees while another rank may still touch the buffer
        if self._owner:
            if self.world > 1:
                dist.barrier()      # mappers unmap first
            self._lib.check(L.so101_shared_free(self.device, self._base))
        else:
            self._lib.check(L.so101_shared_close(self.device, self._base))
            dist.barrier()
        self._base = C.c_void_p()
