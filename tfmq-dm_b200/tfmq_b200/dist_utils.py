"""Multi-GPU plumbing (torch.distributed; NCCL on the GPUs, gloo in the CPU tests).

Sampling shards by independent batches: the only collective is the one-off broadcast of rank 0's
packed int4 weights / scales / FSC tables.  Calibration is data-parallel as in the reference
(quant/calibration.py:228-389): alpha gradients are SUM-all-reduced (linklink.allreduce = SUM), here as
one flat bucket per reconstruction unit, and activation deltas are averaged as one vector."""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_range(total: int, rank_: int, world_: int) -> range:
    """Contiguous, balanced slice of `total` independent units (batches) for this rank."""
    base, extra = divmod(total, world_)
    start = rank_ * base + min(rank_, extra)
    return range(start, start + base + (1 if rank_ < extra else 0))


def shard_interval_indices(n: int, interval: int, rank_: int, world_: int) -> torch.Tensor:
    """Indices of this rank's 1/world slice inside every consecutive block of `interval` calibration
    samples (one block per timestep), reference quant/calibration.py:269-282."""
    per_rank = interval // world_
    return torch.cat([torch.arange(b + rank_ * per_rank, b + (rank_ + 1) * per_rank) for b in range(0, n, interval)])


def broadcast_tensors(tensors: Iterable[torch.Tensor], src: int = 0) -> None:
    """In-place broadcast of a set of tensors from `src` (engine constants at start-up)."""
    if world() == 1:
        return
    for t in tensors:
        if t is not None:
            dist.broadcast(t, src)


def allreduce_flat_(tensors: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """SUM-all-reduce several tensors as one flat bucket; returns views of the reduced bucket."""
    if world() == 1 or not tensors:
        return list(tensors)
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat)
    out, off = [], 0
    for t in tensors:
        out.append(flat[off:off + t.numel()].view_as(t))
        off += t.numel()
    return out


def allaverage_(values: Sequence[torch.Tensor]) -> None:
    """Average 0-dim tensors over the ranks with one all-reduce (linklink.dist_helper.allaverage)."""
    if world() == 1 or not values:
        return
    flat = torch.stack([v.detach().reshape(()) for v in values]) / world()
    dist.all_reduce(flat)
    for v, r in zip(values, flat):
        v.data.copy_(r)


def engine_constants(eng) -> List[torch.Tensor]:
    """Every device constant a sampling rank needs from rank 0."""
    out = [eng.table] if eng.table is not None else []
    for q in eng.ql.values():
        for name in ("packed", "codes", "wdelta", "wzp_f", "wzp_i32", "wsum", "bias", "w_hi", "w_lo", "w_f32", "w_oihw",
                     "h_hi", "h_lo", "h_scale"):
            t = getattr(q, name, None)
            if torch.is_tensor(t):
                out.append(t)
    for w_tf32, w_h16, b, _ in eng._plain.values():
        out += [t for t in (*w_tf32, *w_h16, b) if torch.is_tensor(t)]
    out += [t for t in eng._consts.values() if torch.is_tensor(t)]
    return out
