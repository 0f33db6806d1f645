"""Instance sharding across the GPUs of one box.

Rows are split into contiguous blocks, one per rank (one process per GPU, torch.distributed /
NCCL over NVLink); the m-length float64 state is replicated and only its per-batch deltas (BCA)
or per-iterate confusion sums (Frank-Wolfe) are all-reduced.  No row data ever crosses a link.
torch.distributed is plumbing here: the same code runs over gloo on CPU tensors in the tests.
"""
from __future__ import annotations

import ctypes as C
import os
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


class _RawCuda:
    """__cuda_array_interface__ view of device memory owned by the C library."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerWindow:
    """This rank's peer-memory window plus the mappings of every other rank's window (csrc/p2p.cu).
    The commit kernels signal and read through NVLink directly; torch.distributed is only used once,
    to exchange the 64-byte CUDA IPC handles."""

    def __init__(self, ctx, comm: "Comm", payload_bytes: int, device: torch.device):
        self.ctx, self.comm, self.device = ctx, comm, device
        self.handle = C.c_void_p()
        ipc = (C.c_ubyte * 64)()
        ctx.call("xc_p2p_create", comm.world, comm.rank, int(payload_bytes), C.byref(self.handle), ipc)
        # the 64-byte IPC handles travel as one small tensor all-gather (all_gather_object pickles and takes tens of ms)
        mine = torch.frombuffer(bytearray(bytes(ipc)), dtype=torch.uint8).to(device)
        allh = torch.empty(64 * comm.world, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(allh, mine, group=comm.group)
        ok = 1
        try:
            ctx.call("xc_p2p_open", self.handle, allh.cpu().numpy().tobytes())
        except Exception:           # e.g. no peer access between two of the GPUs
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=comm.group)
        self.ok = bool(int(flag.item()))
        ptr = ctx.lib.xc_p2p_payload(self.handle)
        self.payload = torch.as_tensor(_RawCuda(int(ptr), int(payload_bytes)), device=device)

    def check(self) -> None:
        err = C.c_uint(0)
        self.ctx.call("xc_p2p_error", self.handle, C.byref(err))
        if err.value:
            raise RuntimeError(f"xcolumns_b200: peer-memory commit {err.value} timed out waiting for another rank")

    def close(self) -> None:
        if self.handle:
            self.payload = None
            self.ctx.lib.xc_p2p_destroy(self.ctx.handle, self.handle)
            self.handle = C.c_void_p()


# Windows are kept between calls: creating one costs a cudaMalloc, W - 1 cudaIpcOpenMemHandle calls and a collective
# (~0.1 s measured at 8 ranks -- a third of an end-to-end call on the 8-GPU shard).  A session borrows the window of
# its (device, group, size) key and hands it back clean.
_WINDOW_CACHE = {}


def borrow_window(ctx, comm: "Comm", payload_bytes: int, own_bytes: int, device: torch.device) -> Optional[PeerWindow]:
    """A peer window of at least payload_bytes for this communicator (None if peer access is not available).  The
    first own_bytes of the payload (this rank's delta buffers) are zero on return and every rank has passed a
    barrier after zeroing, as xc_bca_pipe_sweep requires."""
    key = (device.index, id(comm.group), comm.world, comm.rank)
    w = _WINDOW_CACHE.get(key)
    if w is not None and (w.payload is None or w.payload.numel() < payload_bytes):
        w.close()
        w = None
    if w is None:
        w = PeerWindow(ctx, comm, int(payload_bytes), device)
        if not w.ok:
            w.close()
            _WINDOW_CACHE.pop(key, None)
            return None
        _WINDOW_CACHE[key] = w
    else:
        w.comm = comm
        w.payload[:own_bytes].zero_()
        torch.cuda.current_stream(device).synchronize()
        comm.barrier()      # nobody starts pushing into a window whose owner is still clearing it
    return w


def release_windows() -> None:
    """destroy the cached windows (tests, interpreter shutdown with a live process group)"""
    for w in list(_WINDOW_CACHE.values()):
        w.close()
    _WINDOW_CACHE.clear()


def peer_commit_enabled(comm: "Comm", device: torch.device, m: int) -> bool:
    """Peer-memory commits: several ranks on CUDA devices of one box, delta vectors small enough that one
    rank reading all W of them (W * 24 m bytes over NVLink) beats a bandwidth-optimal all-reduce.
    $XCOLUMNS_B200_P2P=0 forces the NCCL all-reduce path."""
    if not (comm.active and comm.world > 1 and device.type == "cuda"):
        return False
    if os.environ.get("XCOLUMNS_B200_P2P", "1") == "0" or comm.world > 16:
        return False
    if dist.get_backend(comm.group) != "nccl":
        return False
    return 24 * m <= (2 << 20)
