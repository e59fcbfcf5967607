"""Multi-GPU sharding of the hot path: volumes are independent, so the
image list is split contiguously over ranks with no data-path collective; the only exchange
is one all_gather of the per-image score table (NCCL over NVLink on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; the first n_items % world_size ranks get one extra."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_items: int, world_size: int) -> List[int]:
    return [shard_range(n_items, r, world_size)[1] - shard_range(n_items, r, world_size)[0]
            for r in range(world_size)]


def gather_scores(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """local [n_local, ...] (this rank's shard, in shard_range order) -> [n_items, ...] on every
    rank.  Shards are padded to the largest shard so a single all_gather_into_tensor suffices
    (~170 B per image: latency-bound)."""
    if not (dist.is_available() and dist.is_initialized()):
        if local.shape[0] != n_items:
            raise ValueError("single process: local shard must hold every item")
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_items, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank}: shard has {local.shape[0]} rows, expected {sizes[rank]}")
    pad = max(sizes)
    tail = local.shape[1:]
    buf = torch.zeros((pad,) + tuple(tail), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = torch.empty((world * pad,) + tuple(tail), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    out = out.view((world, pad) + tuple(tail))
    return torch.cat([out[r, : sizes[r]] for r in range(world)], dim=0)


class AsyncScoreGather:
    """The score gather on a side stream, so that a rank's next batch does not wait for the slowest
    rank's current one: `submit` orders the all_gather after the work already queued on the current
    stream and returns at once; `wait` makes the current stream wait for every gather submitted so
    far (call it before reading a table, and before the end of a timed region).  Without CUDA
    (gloo tests) it degrades to the synchronous `gather_scores`."""

    def __init__(self, device=None, group=None, keep: int = 4):
        self.group = group
        self.keep = keep
        self._tables = []   # the last `keep` results (and their inputs) stay referenced until waited for
        self.stream = None
        if device is not None and torch.device(device).type == "cuda":
            self.device = torch.device(device)
            self.stream = torch.cuda.Stream(self.device)

    def submit(self, local: torch.Tensor, n_items: int, after=None) -> torch.Tensor:
        """`after`: a CUDA event that marks `local` complete (PipelineResult.table_async); without it
        the gather is ordered behind everything queued on the current stream."""
        if self.stream is None:
            return gather_scores(local, n_items, self.group)
        main = torch.cuda.current_stream(self.device)
        if after is not None:
            self.stream.wait_event(after)
        else:
            self.stream.wait_stream(main)
        with torch.cuda.stream(self.stream):
            out = gather_scores(local, n_items, self.group)
        local.record_stream(self.stream)
        out.record_stream(main)     # allocated on the side stream, read by the caller on the main one
        self._tables.append((local, out))
        if len(self._tables) > self.keep:
            self._tables.pop(0)
        return out

    def wait(self) -> None:
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
