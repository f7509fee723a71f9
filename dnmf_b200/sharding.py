"""Frame sharding over GPUs (one process per GPU, torch.distributed for the plumbing).

The reference has no distributed code.  Its model has no shared learnable parameter (positions,
widths are fixed), so frames are independent units: rank r owns a contiguous
slab of frames with their video, deformation coefficients (and Adam moments) and trace columns.
No data-path collective is needed for update_motion; the only exchanges are
  * one 8-byte all-reduce (sum) of the batch loss per step, for reporting, and the global batch
    size in the 1/(B*N) scale of the MSE (Demix/dNMF.py:188);
  * when gamma_c != 0, the boundary trace columns between neighbouring slabs once per
    multiplicative sweep (temporal smoothness term of Demix/dNMF.py:145).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def frame_slab(T_global: int, world: int, rank: int) -> Tuple[int, int]:
    """(first frame, frame count) of rank's contiguous slab; slab sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(int(T_global), int(world))
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def owner_of(frame: int, T_global: int, world: int) -> int:
    base, extra = divmod(int(T_global), int(world))
    cut = extra * (base + 1)
    if frame < cut:
        return frame // (base + 1)
    return extra + (frame - cut) // max(base, 1)


def split_batch(global_ids, T_global: int, world: int, rank: int):
    """Local ids (relative to the slab) of the frames of a global minibatch that this rank owns, and
    the global batch size that scales the loss."""
    start, count = frame_slab(T_global, world, rank)
    ids = [int(i) for i in global_ids]
    mine = [i - start for i in ids if start <= i < start + count]
    return mine, len(ids)


def allreduce_loss(local_sse_over_BN: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of per-rank partial losses (each already divided by B_global*N)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(local_sse_over_BN, group=group)
    return local_sse_over_BN


def make_halo_exchange(group=None) -> Callable:
    """Returns f(first, last) -> (prev, next): the last column of the previous rank's slab and the
    first column of the next rank's, or None at the two ends of the video (edge replication,
    Demix/dNMF.py:145).  Uses one all_gather of 2*K values per sweep."""
    def exchange(first: torch.Tensor, last: torch.Tensor):
        if not (dist.is_available() and dist.is_initialized()):
            return None, None
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world == 1:
            return None, None
        mine = torch.stack((first, last)).contiguous()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine, group=group)
        prev = gathered[rank - 1][1].contiguous() if rank > 0 else None
        nxt = gathered[rank + 1][0].contiguous() if rank < world - 1 else None
        return prev, nxt
    return exchange
