"""Instance sharding across the GPUs of one box.

Rows are split into contiguous blocks, one per rank (one process per GPU, torch.distributed /
NCCL over NVLink); the m-length float64 state is replicated and only its per-batch deltas (BCA)
or per-iterate confusion sums (Frank-Wolfe) are all-reduced.  No row data ever crosses a link.
torch.distributed is plumbing here: the same code runs over gloo on CPU tensors in the tests.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_rows(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row block [lo, hi) of rank `rank` (sizes differ by at most one row)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def batch_schedule(n_local_max: int, batch: int) -> int:
    """Number of commits per sweep that every rank performs (ragged shards join with empty batches)."""
    return max(1, (n_local_max + batch - 1) // batch)


class Comm:
    """Thin wrapper over a process group; group=None means a single process (no collectives)."""

    def __init__(self, group, device: Optional[torch.device] = None):
        self.group = group
        self.active = group is not None
        self.world = dist.get_world_size(group) if self.active else 1
        self.rank = dist.get_rank(group) if self.active else 0
        self.device = device
        self._n_global = {}
        self.n_allreduce = 0

    def allreduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        if self.active and self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self.n_allreduce += 1
        return t

    def allreduce_prod_(self, t: torch.Tensor) -> torch.Tensor:
        if self.active and self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.PRODUCT, group=self.group)
            self.n_allreduce += 1
        return t

    def max_int(self, v: int) -> int:
        if not (self.active and self.world > 1):
            return int(v)
        t = torch.tensor([int(v)], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return int(t.item())

    def n_global(self, n_local: int) -> int:
        if not (self.active and self.world > 1):
            return int(n_local)
        if n_local not in self._n_global:
            t = torch.tensor([int(n_local)], dtype=torch.int64, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self._n_global[n_local] = int(t.item())
        return self._n_global[n_local]

    def barrier(self):
        if self.active and self.world > 1:
            dist.barrier(group=self.group)


def make_comm(distributed, device: Optional[torch.device] = None) -> Comm:
    """distributed: False | True (default process group) | a ProcessGroup."""
    if distributed is False or distributed is None:
        return Comm(None, device)
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("distributed=True needs torch.distributed.init_process_group() first")
    group = dist.group.WORLD if distributed is True else distributed
    return Comm(group, device)
