"""Env-level data parallelism: shard independent environments across ranks, gather the dataset.

The step path has NO inter-GPU traffic (envs are independent).  The only exchange is the final
gather of the `[N/G, T+1, 13]` row shards to rank 0 (what `np.save` needs), done with one
`torch.distributed.gather` (NCCL on GPUs, gloo in the CPU tests).  Control/reset random streams
are keyed by the GLOBAL env index, so the gathered dataset is bit-identical for any world size.
"""
from __future__ import annotations

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


def gather_rows(local: torch.Tensor, n_total: int, dst: int = 0) -> Optional[torch.Tensor]:
    """Gather row shards `[n_local, ...]` (rank-major) into `[n_total, ...]` on rank `dst`.

    Shards may differ by one row; they are padded to the largest shard for the collective and
    trimmed afterwards.  Returns None on the other ranks.  Single process: returns `local`.
    """
    rank, world = dist_info()
    if world == 1:
        return local
    sizes = shard_sizes(n_total, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, expected {sizes[rank]}")
    nmax = max(sizes)
    send = local
    if local.shape[0] < nmax:
        pad = torch.zeros((nmax - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        send = torch.cat([local, pad], dim=0)
    send = send.contiguous()
    if rank == dst:
        bufs = [torch.empty_like(send) for _ in range(world)]
        dist.gather(send, gather_list=bufs, dst=dst)
        return torch.cat([b[: sizes[r]] for r, b in enumerate(bufs)], dim=0)
    dist.gather(send, gather_list=None, dst=dst)
    return None
