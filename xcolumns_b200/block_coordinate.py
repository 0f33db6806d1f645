"""Block Coordinate Ascent on the GPU (drop-in for xcolumns/block_coordinate.py:296-801).

Two execution modes behind the reference's signature (selected with the ``mode`` keyword or
$XCOLUMNS_B200_MODE; the reference's ``**kwargs`` swallows the extra keyword):

``exact``    sequential Gauss-Seidel sweep in the reference's instance order and float64 operation
             order (csrc/bca_exact.cu).  Predictions, running state and per-sweep utilities are
             bit-comparable with the reference on tie-free inputs.  Latency bound by construction.
``batched``  block-Jacobi: the (shuffled) instance order is cut into batches; all rows of a batch
             see the same frozen state and their confusion deltas are committed together
             (csrc/bca_batched.cu).  One streaming pass over y_proba per sweep at HBM speed; final
             utilities agree with the reference within 1e-4 when batch <= n/8.
             With torch.distributed initialised and ``distributed=True`` every rank holds a row
             shard and the per-batch deltas are all-reduced (NCCL over NVLink).
``auto``     (default) exact for n <= 4096 rows (and whenever only the sequential kernels apply: greedy
             init, k = 0, Jaccard / G-mean / H-mean), batched above.
"""
from __future__ import annotations

import ctypes as C
import os
from time import time
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import numpy as np
import torch
from scipy.sparse import csr_matrix

from . import _device as dev
from . import metrics as M
from ._lib import XC_PIPE_FORK, XC_PIPE_SHUFFLE, XC_SUM_FAST, XC_SUM_ORDERED, MetricParams, PipeArgs
from .distributed import Comm, PeerWindow, borrow_window, make_comm, peer_commit_enabled
from .types import Matrix
from .utils import add_kwargs_to_signature, log_info, log_warning
from .weighted_prediction import _check_k, topk_csr_device, topk_dense_device

_AUTO_EXACT_MAX_ROWS = 4096


def _default_lag(batch_rows: int = 0, wave_rows: int = 0) -> int:
    """Commits applied this many batches late in the pipelined dense sweep (csrc/bca_batched.cu:
    xc_bca_pipe_sweep): lag + 1 batches are in flight on as many streams.  Default: 1 when a batch is at least four
    waves of the streaming kernel (two such kernels keep the GPU full: 307 k x 13 k on one GPU, 2.50 vs 2.56 ms per
    sweep), 2 for smaller batches, where a third batch in flight fills the slots a finishing batch frees while its
    commit is pending (8-GPU shard of the same matrix, 38 k rows: 0.351 vs 0.373 ms; 77 k rows: 0.628 vs 0.694 ms;
    measured on B200, profiles/r02_notes.md).  $XCOLUMNS_B200_LAG=0 restores the strict batch order, values up to 3
    are accepted."""
    env = os.environ.get("XCOLUMNS_B200_LAG")
    if env is not None:
        try:
            return max(0, min(3, int(env)))
        except ValueError:
            pass
    if batch_rows > 0 and wave_rows > 0 and batch_rows < 4 * wave_rows:
        return 2
    return 1


def _metric_params(metric_id, beta, eps, maximize, skip_tn, n_div, n_rows=None, mix=None) -> MetricParams:
    n_rows = n_div if n_rows is None else n_rows
    c1, beta2 = M.metric_c1_beta2(metric_id, beta)
    alpha, mk, mm = mix if mix is not None else (1.0, 1.0, 1.0)
    return MetricParams(metric=metric_id, maximize=int(bool(maximize)), skip_tn=int(bool(skip_tn)),
                        mix=int(mix is not None), c1=c1, beta2=beta2, eps=float(eps), n_div=float(n_div),
                        n_rows=float(n_rows), mix_alpha=alpha, mix_k=mk, mix_m=mm)


def _resolve_mode(mode: Optional[str], n: int, needs_exact: bool) -> str:
    mode = mode or os.environ.get("XCOLUMNS_B200_MODE", "auto")
    if mode not in ("auto", "exact", "batched"):
        raise ValueError("mode must be 'auto', 'exact' or 'batched'")
    if mode == "auto":
        mode = "exact" if (n <= _AUTO_EXACT_MAX_ROWS or needs_exact) else "batched"
    return mode


# ------------------------------------------------------------------------------------------
# initial prediction (block_coordinate.py:28-51)
# ------------------------------------------------------------------------------------------

def _initial_pred(y_proba, data, init_y_pred, k: int, seed, device) -> torch.Tensor:
    n, m = data.n, data.m
    if isinstance(init_y_pred, str) and init_y_pred == "top":
        if isinstance(data, dev.CsrDev):
            return topk_csr_device(data, k, None, None)[0]
        return topk_dense_device(data, k, None, None, data.code)[0]
    if isinstance(init_y_pred, str) and init_y_pred in ("random", "greedy"):
        # dense: utils.py:104-116 (numpy Generator, bit-identical).  CSR: the reference draws with
        # numba's private Mersenne state (numba_csr_functions.py:93-112), which cannot be
        # reproduced without numba; the same numpy recipe is used (documented deviation).
        rng = np.random.default_rng(seed)
        labels = np.arange(m)
        idx = np.empty((n, k), dtype=np.int32)
        for i in range(n):
            idx[i] = np.sort(rng.choice(labels, k, replace=False, shuffle=False))
        return torch.from_numpy(idx).to(device)
    if isinstance(init_y_pred, (np.ndarray, torch.Tensor, csr_matrix)):
        if tuple(init_y_pred.shape) != (n, m):
            raise ValueError(f"init_y_pred must have shape (n, m) = ({n}, {m}), but has shape {init_y_pred.shape}")
        return dev.dense_pred_to_compact(init_y_pred, k, device)
    raise ValueError(
        f"init_y_pred must be np.ndarray, Torch.tensor, csr_matrix or str in ['random', 'greedy', 'top'], but has type {type(init_y_pred)}")


# ------------------------------------------------------------------------------------------
# device session
# ------------------------------------------------------------------------------------------

class BcaSession:
    """State of one BCA run on one device: probability matrix (dense or CSR), compact prediction,
    float64 confusion state, coefficient / delta workspaces.  Used by the public functions below
    and driven directly by bench.py on HBM-resident inputs."""

    def __init__(self, data, k: int, params: MetricParams, util_params: MetricParams, aggregation: str,
                 comm: Optional[Comm] = None):
        self.data = data
        self.is_csr = isinstance(data, dev.CsrDev)
        self.device = (data.data if self.is_csr else data.t).device
        self.ctx = dev.ctx_for(self.device)
        self.k = k
        self.p = params
        self.up = util_params
        self.agg = 0 if aggregation == "mean" else 1
        self.comm = comm or Comm(None)
        self.n, self.m = data.n, data.m
        f64 = dict(dtype=torch.float64, device=self.device)
        self.state = torch.zeros((4, self.m), **f64)           # tp, fp, fn, tn
        self.state[3].fill_(-1.0)
        self.delta = torch.zeros((3, self.m), **f64)           # pending batch deltas (per-batch entry points)
        # Jaccard / G-mean / H-mean: 16-byte per-label records instead of the affine coefficient pairs
        self.use_rec = params.metric in M.RECORD_GAIN_METRICS
        clen = int(self.ctx.lib.xc_bca_coef_len(self.m))       # padded to whole coefficient tiles
        # Dense rows of one process, or of the ranks of one box with a peer window: the whole sweep is ONE C call
        # (xc_bca_sweep_dense_pipe); commits are applied `lag` batches late so consecutive batches overlap.
        if self.is_csr or self.use_rec:
            self.lag = 0
        else:   # decided on numbers every rank agrees on (ragged shards differ by a row)
            n_ref = self.comm.max_int(self.n)
            wave = self.wave_rows()
            self.lag = _default_lag(default_batch_rows(n_ref, wave), wave)
        self.pipe = (not self.is_csr) and self.comm.world == 1
        self._gb = 0                                           # batches issued so far (delta-buffer rotation)
        self.peer: Optional[PeerWindow] = None                 # peer-memory commits (sharded dense rows)
        stride = int(self.ctx.lib.xc_bca_delta_stride(self.m))
        nbuf = int(self.ctx.lib.xc_bca_pipe_buffers(self.lag))
        if (not self.is_csr) and peer_commit_enabled(self.comm, self.device, self.m):
            peer = borrow_window(self.ctx, self.comm,
                                 int(self.ctx.lib.xc_bca_window_bytes(self.m, self.lag, self.comm.world)), nbuf * stride,
                                 self.device)
            if peer is not None:
                self.peer = peer
                self.pipe = True
        self.delta_pipe = (torch.zeros(nbuf * stride, dtype=torch.uint8, device=self.device)
                           if (self.pipe and self.peer is None) else None)
        # CSR rows of one process: rows that fit one warp's registers take the reduction-based batch kernel and keep a
        # list of the labels a batch touched (folds then visit those labels only, not all m)
        self.max_row_nnz = 0
        self.touch_flag = self.touch_list = self.touch_ctl = None
        self._csr_refresh = True
        if self.is_csr and self.n > 0:
            self.max_row_nnz = int((data.indptr[1:] - data.indptr[:-1]).max().item())
            if self.comm.world == 1 and not self.use_rec and 0 < self.max_row_nnz <= 128:
                self.touch_flag = torch.zeros(self.m, dtype=torch.int32, device=self.device)
                self.touch_list = torch.empty(self.m, dtype=torch.int32, device=self.device)
                self.touch_ctl = torch.zeros(4, dtype=torch.int32, device=self.device)
        self._n_order = self.n          # rows a sweep visits (1 with the reference's normalize_conf_matrix=False quirk)
        self._inflight = False          # the pipeline's internal streams hold work this stream has not joined
        self._need_fork = True          # the host touched the state / prediction since the last pipelined sweep
        self._prev_tail_from = -1
        self._sweep_ctr = 0
        self._order2: Optional[torch.Tensor] = None
        self._snaps: List[torch.Tensor] = []
        self.colsum: Optional[torch.Tensor] = None
        # [set][coef_n | coef_s][label][B, A]: one coefficient version per batch in flight of the pipelined sweep
        self.coef = torch.zeros((self.lag + 1, 2, clen, 2), dtype=torch.float32, device=self.device)
        self.coef_n, self.coef_s = self.coef[0, 0], self.coef[0, 1]
        self.rec = torch.zeros((clen, 4), dtype=torch.float32, device=self.device) if self.use_rec else None
        self.util_buf = torch.zeros(8, **f64)
        self.pred: Optional[torch.Tensor] = None

    # -- small helpers ---------------------------------------------------------------------
    def _s(self):
        return dev.stream_ptr(self.device)

    def _sp(self, i):
        return C.c_void_p(self.state[i].data_ptr())

    def _dp(self, i):
        return C.c_void_p(self.delta[i].data_ptr())

    def zero_delta(self) -> None:
        """Only the per-batch entry points (xc_bca_coef / xc_bca_batch_*) keep pending deltas in self.delta; the
        pipelined sweep rotates and clears its own buffers on the device (no host-side clearing: a slower peer
        may still be reading this rank's window)."""
        if not self.pipe:
            self.delta.zero_()

    def join(self) -> None:
        """this stream waits for everything the pipelined sweeps still have in flight; required before the host
        (or any kernel on this stream) reads or rewrites the prediction or the state"""
        if self._inflight:
            self.ctx.call("xc_bca_pipe_join", self._s())
            self._inflight = False
        self._need_fork = True

    def close(self) -> None:
        self.join()
        if self.peer is not None:
            torch.cuda.synchronize(self.device)
            self.comm.barrier()          # every rank has left its last commit: the window can be handed back
            self.peer.check()
            self.peer = None             # (kept in the cache of xcolumns_b200.distributed for the next call)

    # -- state from the current prediction ---------------------------------------------------
    def recompute(self, order: int) -> None:
        """tp / fp / fn (and tn) from scratch (block_coordinate.py:430-436, :465-467)."""
        self.join()
        self._csr_refresh = True
        d, k = self.data, self.k
        if order == XC_SUM_ORDERED:
            if self.comm.world > 1:
                raise NotImplementedError("sequential-exact BCA runs on one GPU (replicas only)")
            if self.is_csr:
                self._recompute_csr_ordered()
            else:
                self.ctx.call("xc_confmat_dense_compact", dev.ptr(d.t), d.code, d.ld, dev.ptr(self.pred), k, d.n, d.m,
                              order, None, self._sp(0), self._sp(1), self._sp(2), self._s())
        elif self.is_csr:
            if self.colsum is None:  # sum_i eta_ij never changes: fn = colsum - tp, no pass over the rows per sweep
                self.colsum = torch.empty(self.m, dtype=torch.float64, device=self.device)
                self.ctx.call("xc_colsum_csr", dev.ptr(d.data), d.code, dev.ptr(d.indices), int(d.data.numel()), d.m,
                              dev.ptr(self.colsum), self._s())
                self.comm.allreduce_sum_(self.colsum)
            self.ctx.call("xc_confmat_csr_compact", dev.ptr(d.data), dev.ptr(d.indices), dev.ptr(d.indptr), d.code,
                          dev.ptr(self.pred), k, d.n, d.m, order, self._sp(0), self._sp(1), None, self._s())
            if self.comm.world > 1:
                self.comm.allreduce_sum_(self.state[0:2])
            torch.sub(self.colsum, self.state[0], out=self.state[2])
        else:
            if self.colsum is None:  # sum_i eta_ij never changes: one extra pass per call, not per sweep
                self.colsum = torch.empty(self.m, dtype=torch.float64, device=self.device)
                self.ctx.call("xc_colsum_dense", dev.ptr(d.t), d.code, d.n, d.m, d.ld, dev.ptr(self.colsum), self._s())
                self.comm.allreduce_sum_(self.colsum)
            self.ctx.call("xc_confmat_dense_compact", dev.ptr(d.t), d.code, d.ld, dev.ptr(self.pred), k, d.n, d.m,
                          order, dev.ptr(self.colsum), self._sp(0), self._sp(1), self._sp(2), self._s())
            if self.comm.world > 1:
                self.comm.allreduce_sum_(self.state[0:2])
                torch.sub(self.colsum, self.state[0], out=self.state[2])
        if self.p.skip_tn:
            self.state[3].fill_(-1.0)
        else:  # -tp - fp - fn + n   (confusion_matrix.py:397, n = number of rows)
            n_total = self.comm.n_global(self.n)
            self.state[3] = -self.state[0] - self.state[1] - self.state[2] + n_total

    def _recompute_csr_ordered(self) -> None:
        """Row-ordered per-label sums for CSR rows.  The probabilities never change, so their
        column-major copy (stable sort by label = every label's entries in row order) is built once;
        one warp per label then adds its column strictly in row order."""
        d, k = self.data, self.k
        if getattr(self, "_csc", None) is None:
            counts = (d.indptr[1:] - d.indptr[:-1])
            row_of = torch.repeat_interleave(torch.arange(d.n, device=self.device, dtype=torch.int32), counts)
            perm = torch.sort(d.indices, stable=True).indices
            c_ptr = torch.zeros(d.m + 1, dtype=torch.int64, device=self.device)
            c_ptr[1:] = torch.bincount(d.indices, minlength=d.m).cumsum(0)
            self._csc = (d.data[perm].contiguous(), row_of[perm].contiguous(), c_ptr)
            self._lone = torch.zeros(1, dtype=torch.int32, device=self.device)
        c_data, c_rows, c_ptr = self._csc
        self.ctx.call("xc_confmat_csc_ordered", dev.ptr(c_data), d.code, dev.ptr(c_rows), dev.ptr(c_ptr),
                      dev.ptr(d.indices), dev.ptr(d.indptr), dev.ptr(self.pred), k, d.n, d.m, self._sp(0), self._sp(1),
                      self._sp(2), dev.ptr(self._lone), self._s())
        if int(self._lone.item()):   # a predicted label that its row does not store: row-walking kernel
            self.ctx.call("xc_confmat_csr_compact", dev.ptr(d.data), dev.ptr(d.indices), dev.ptr(d.indptr), d.code,
                          dev.ptr(self.pred), k, d.n, d.m, XC_SUM_ORDERED, self._sp(0), self._sp(1), self._sp(2),
                          self._s())

    def wave_rows(self) -> int:
        """rows of one full wave of the dense batch kernel (0 for CSR: one warp per row)"""
        if self.is_csr:
            return 0
        return int(self.ctx.lib.xc_bca_wave_rows(self.ctx.handle, self.data.code, self.m))

    def permutation(self, n: int, seed: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """pseudo-random visiting order of one batched sweep (int32 [n], device)"""
        if out is None:
            out = torch.empty(n, dtype=torch.int32, device=self.device)
        self.ctx.call("xc_permutation", n, C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), dev.ptr(out), self._s())
        return out

    def utility_device(self, slot: int) -> None:
        """block_coordinate.py:54-90 on the device; result lands in util_buf[slot]."""
        self.utility_into(self.util_buf[slot:])

    def utility_into(self, out: torch.Tensor) -> None:
        """utility of the current state into out[0] (float64, device)"""
        self.join()
        self.ctx.call("xc_utility", C.byref(self.up), self.agg, self._sp(0), self._sp(1), self._sp(2), self._sp(3),
                      self.m, C.c_void_p(out.data_ptr()), self._s())

    # -- one batched sweep, whichever path the session uses ---------------------------------------
    def run_sweep(self, j: int, seed: int, shuffle: bool, batch: int, n_batches: Optional[int], full: bool,
                  util_out: torch.Tensor) -> torch.Tensor:
        """Sweep number j (>= 1) over this rank's rows: visiting order xc_permutation(n, seed) (or 0..n-1), `batch`
        rows per commit, the state after the sweep, its utility into util_out[0].  Returns a snapshot of the
        prediction as it was before the sweep (valid until three sweeps later) for restore()."""
        if self.pipe:
            d = self.data
            if self._order2 is None:
                # [order A | order B | raw order | stamps | per-block counts]; without shuffling the first n entries are the order
                self._order2 = torch.zeros(4 * self._n_order + self._n_order // 256 + 8, dtype=torch.int32,
                                           device=self.device)
                self._order2[:self._n_order] = torch.arange(self._n_order, dtype=torch.int32, device=self.device)
                self._snaps = [torch.empty_like(self.pred) for _ in range(3)]
            snap = self._snaps[j % 3]
            self._sweep_ctr += 1             # consecutive calls: order-buffer parity and the busy-row stamps rely on it
            if self._n_order != self.n:      # only the visited rows are written by the kernels: take a whole copy first
                self.join()
                snap.copy_(self.pred)
            nb = n_batches if n_batches is not None else (self._n_order + batch - 1) // batch
            flags = (XC_PIPE_SHUFFLE if shuffle else 0) | (XC_PIPE_FORK if self._need_fork else 0)
            a = PipeArgs(params=C.pointer(self.p), eta=d.t.data_ptr(), dtype=d.code, k=self.k, m=d.m, ld=d.ld,
                         n_rows=self._n_order, batch=int(batch), n_batches=int(nb), batch0=int(self._gb), lag=int(self.lag),
                         flags=flags, seed=seed & 0xFFFFFFFFFFFFFFFF, sweep=self._sweep_ctr, order=self._order2.data_ptr(),
                         coef=(self.rec if self.use_rec else self.coef).data_ptr(), pred_idx=self.pred.data_ptr(),
                         pred_snapshot=snap.data_ptr(), tp=self.state[0].data_ptr(), fp=self.state[1].data_ptr(),
                         fn=self.state[2].data_ptr(),
                         delta=self.delta_pipe.data_ptr() if self.delta_pipe is not None else None,
                         util_params=C.pointer(self.up), util_out=util_out.data_ptr(), agg=self.agg, reserved=0,
                         util_tn_rows=-1.0 if self.p.skip_tn else float(self.comm.n_global(self.n)),
                         prev_tail_from=-1 if self._need_fork else int(self._prev_tail_from))
            self.ctx.call("xc_bca_pipe_sweep", self.peer.handle if self.peer is not None else None, C.byref(a), self._s())
            self._gb += nb
            self._inflight, self._need_fork = True, False
            # where this sweep's last `lag` batches begin in its order (the next sweep's first batches avoid them)
            self._prev_tail_from = min(self._n_order, max(0, (nb - self.lag) * int(batch))) if shuffle else -1
            if full or os.environ.get("XCOLUMNS_B200_SWEEP_RECOMPUTE") == "1":
                self.recompute(XC_SUM_FAST)       # joins; the next sweep re-forks and refreshes its coefficients
                self.utility_into(util_out)
            return snap
        snap = self.pred.clone()
        order = (self.permutation(self._n_order, seed, self._order_buf()) if shuffle else self._order_buf(True))
        self.zero_delta()
        self.sweep_and_fold(order, batch, n_batches, full)
        self.utility_into(util_out)
        return snap

    def _order_buf(self, identity: bool = False) -> torch.Tensor:
        if getattr(self, "_order1", None) is None:
            self._order1 = torch.arange(self._n_order, dtype=torch.int32, device=self.device)
        elif identity and not getattr(self, "_order1_identity", True):
            torch.arange(self._n_order, dtype=torch.int32, device=self.device, out=self._order1)
        self._order1_identity = identity
        return self._order1

    def restore(self, snapshot: torch.Tensor) -> None:
        """put a prediction snapshot back (roll-back of a sweep); the state must be recomputed afterwards"""
        self.join()
        if snapshot.data_ptr() != self.pred.data_ptr():
            self.pred.copy_(snapshot)

    def sync_tn(self) -> None:
        """tn of the current tp / fp / fn (kept lazily by the pipelined sweeps)"""
        self.join()
        if self.p.skip_tn:
            self.state[3].fill_(-1.0)
        else:
            self.state[3] = -self.state[0] - self.state[1] - self.state[2] + self.comm.n_global(self.n)

    # -- sweeps -------------------------------------------------------------------------------
    def sweep_exact(self, order_dev: torch.Tensor, greedy: bool) -> None:
        d, k = self.data, self.k
        if self.is_csr:
            self.ctx.call("xc_bca_exact_sweep_csr", dev.ptr(d.data), d.code, dev.ptr(d.indices), dev.ptr(d.indptr),
                          d.n, d.m, dev.ptr(order_dev), int(order_dev.numel()), k, C.byref(self.p), int(greedy),
                          dev.ptr(self.pred), self._sp(0), self._sp(1), self._sp(2), self._sp(3), self._s())
        else:
            self.ctx.call("xc_bca_exact_sweep_dense", dev.ptr(d.t), d.code, d.n, d.m, d.ld, dev.ptr(order_dev),
                          int(order_dev.numel()), k, C.byref(self.p), int(greedy), dev.ptr(self.pred), self._sp(0),
                          self._sp(1), self._sp(2), self._sp(3), self._s())

    def finish_sweep(self, full: bool) -> None:
        """State after a batched sweep.  full: recompute it from the prediction (what the reference does at
        every sweep boundary, block_coordinate.py:465-467); otherwise fold the last batch's pending deltas
        into the running state -- the float64 sums then differ from a recomputation by rounding only
        (~1e-16 relative per commit), and the O(n k) gather pass (58 us at C3, 2 % of a sweep) is saved.
        The batched driver recomputes every 16th sweep and after every rollback."""
        if full or os.environ.get("XCOLUMNS_B200_SWEEP_RECOMPUTE") == "1":   # the switch restores the reference's cadence
            self.recompute(XC_SUM_FAST)
            return
        if self.use_rec:
            self.ctx.call("xc_bca_rec", C.byref(self.p), self._sp(0), self._sp(1), self._sp(2), self._dp(0),
                          self._dp(1), self._dp(2), self.m, dev.ptr(self.rec), self._s())
        else:
            self.ctx.call("xc_bca_coef", C.byref(self.p), self._sp(0), self._sp(1), self._sp(2), self._dp(0),
                          self._dp(1), self._dp(2), self.m, dev.ptr(self.coef_n), dev.ptr(self.coef_s), self._s())
        if not self.p.skip_tn:
            n_total = self.comm.n_global(self.n)
            self.state[3] = -self.state[0] - self.state[1] - self.state[2] + n_total

    def sweep_and_fold(self, order_dev: torch.Tensor, batch: int, n_batches: Optional[int], full: bool) -> None:
        """One block-Jacobi sweep + the state after it on the NON-pipelined paths (CSR rows, dense rows of several
        ranks without a peer window); dense rows of one process or of a box with peer windows go through
        run_sweep -> xc_bca_pipe_sweep.  Single-process CSR rows issue the whole sweep through ONE C call (small
        problems are bound by the per-call overhead of this shim).
        full: recompute the state from the prediction afterwards (block_coordinate.py:465-467)."""
        d, k = self.data, self.k
        n_loc = int(order_dev.numel())
        if self.comm.world > 1 or full or not self.is_csr:
            self.sweep_batched(order_dev, batch, n_batches)
            self.finish_sweep(full)
            return
        self.ctx.call("xc_bca_sweep_csr", C.byref(self.p), dev.ptr(d.data), d.code, dev.ptr(d.indices),
                      dev.ptr(d.indptr), d.m, dev.ptr(order_dev), n_loc, int(batch), k,
                      dev.ptr(self.rec if self.use_rec else self.coef_n), dev.ptr(self.coef_s), dev.ptr(self.pred),
                      self._sp(0), self._sp(1), self._sp(2), self._dp(0), self._dp(1), self._dp(2), self.max_row_nnz,
                      dev.ptr(self.touch_flag), dev.ptr(self.touch_list), dev.ptr(self.touch_ctl),
                      int(self._csr_refresh), self._s())
        self._csr_refresh = False
        if not self.p.skip_tn:
            self.state[3] = -self.state[0] - self.state[1] - self.state[2] + self.n

    def sweep_batched(self, order_dev: torch.Tensor, batch: int, n_batches: Optional[int] = None) -> None:
        """One block-Jacobi sweep over the (local) rows in order_dev through the per-batch entry points, `batch`
        rows per commit; with several ranks and no peer window (CSR label spaces, gloo) the deltas of every batch
        are all-reduced.  n_batches (distributed): common number of commits so every rank joins every all-reduce.
        The pending deltas of the last batch are folded by finish_sweep."""
        d, k = self.data, self.k
        n_loc = int(order_dev.numel())
        nb = n_batches if n_batches is not None else (n_loc + batch - 1) // batch
        for b in range(nb):
            lo = min(b * batch, n_loc)
            hi = min(lo + batch, n_loc)
            # fold the pending deltas into the state, refresh the gain coefficients / records
            if self.use_rec:
                self.ctx.call("xc_bca_rec", C.byref(self.p), self._sp(0), self._sp(1), self._sp(2), self._dp(0),
                              self._dp(1), self._dp(2), self.m, dev.ptr(self.rec), self._s())
            else:
                self.ctx.call("xc_bca_coef", C.byref(self.p), self._sp(0), self._sp(1), self._sp(2), self._dp(0),
                              self._dp(1), self._dp(2), self.m, dev.ptr(self.coef_n), dev.ptr(self.coef_s), self._s())
            if hi > lo:
                rows = C.c_void_p(order_dev.data_ptr() + 4 * lo)
                if self.is_csr and self.use_rec:
                    self.ctx.call("xc_bca_batch_csr_rec", C.byref(self.p), dev.ptr(d.data), d.code, dev.ptr(d.indices),
                                  dev.ptr(d.indptr), rows, hi - lo, k, dev.ptr(self.rec), self._sp(0), self._sp(1),
                                  self._sp(2), dev.ptr(self.pred), self._dp(0), self._dp(1), self._dp(2), self._s())
                elif self.is_csr:
                    self.ctx.call("xc_bca_batch_csr", dev.ptr(d.data), d.code, dev.ptr(d.indices), dev.ptr(d.indptr),
                                  rows, hi - lo, k, dev.ptr(self.coef_n), dev.ptr(self.coef_s), dev.ptr(self.pred),
                                  self._dp(0), self._dp(1), self._dp(2), self.max_row_nnz, None, None, None, self._s())
                elif self.use_rec:
                    self.ctx.call("xc_bca_batch_dense_rec", C.byref(self.p), dev.ptr(d.t), d.code, d.m, d.ld, rows,
                                  hi - lo, k, dev.ptr(self.rec), self._sp(0), self._sp(1), self._sp(2),
                                  dev.ptr(self.pred), self._dp(0), self._dp(1), self._dp(2), self._s())
                else:
                    self.ctx.call("xc_bca_batch_dense", dev.ptr(d.t), d.code, d.m, d.ld, rows, hi - lo, k,
                                  dev.ptr(self.coef_n), dev.ptr(self.coef_s), dev.ptr(self.pred), self._dp(0),
                                  self._dp(1), self._dp(2), self._s())
            if self.comm.world > 1:
                self.comm.allreduce_sum_(self.delta)


def _host_utility(binary_metric_func, aggregation: str, state: torch.Tensor, n_div) -> float:
    """The reference's own utility expression on the float64 state copied to the host
    (block_coordinate.py:54-90; note that it does NOT forward metric_kwargs)."""
    tp, fp, fn, tn = state.cpu().numpy()
    if callable(binary_metric_func):
        vals = binary_metric_func(tp / n_div, fp / n_div, fn / n_div, tn / n_div)
    else:
        vals = np.array([f(tp[i] / n_div, fp[i] / n_div, fn[i] / n_div, tn[i] / n_div)
                         for i, f in enumerate(binary_metric_func)])
    if not isinstance(vals, np.ndarray):
        raise ValueError(f"binary_metric_func must return np.ndarray, but returned {type(vals)}")
    if vals.shape != (tp.shape[0],):
        raise ValueError(f"binary_metric_func must return np.ndarray of shape {tp.shape[0]}, but returned {vals.shape}")
    if aggregation == "sum":
        return vals.sum()
    if aggregation == "mean":
        return vals.mean()
    raise ValueError(f"Unsupported utility aggregation function: {aggregation}, must be either 'mean' or 'sum'")


def default_batch_rows(n: int, wave_rows: int = 0) -> int:
    """Rows one rank commits together: about n/8 (<= n/6 keeps the reference's
    fixed point within 3e-7 for F-measures), rounded down to whole waves of the streaming kernel so
    that no launch ends with a partially filled wave; tiny inputs use n/8 as is."""
    b = max(1, (n + 7) // 8)      # rounded up: n = 8 q + r must not end in a ninth commit over r rows
    if wave_rows > 0 and b >= wave_rows:
        b = (b // wave_rows) * wave_rows
    return b


def predict_using_bc_with_0approx(
    y_proba: Matrix,
    binary_metric_func: Union[Callable, List[Callable]],
    k: int,
    metric_aggregation: str = "mean",  # "mean" or "sum"
    normalize_conf_matrix: bool = True,
    metric_kwargs: Optional[Dict[str, Any]] = None,
    maximize: bool = True,
    tolerance: float = 1e-6,
    init_y_pred: Union[str, Matrix] = "top",  # "random", "top", "greedy", Matrix
    max_iters: int = 100,
    shuffle_order: bool = True,
    skip_tn: bool = False,
    return_meta: bool = False,
    seed: Optional[int] = None,
    verbose: bool = False,
    **kwargs,
) -> Union[Matrix, Tuple[Matrix, Dict[str, Any]]]:
    """Block coordinate ascent/descent on the 0-th order approximation of the Expected Test Utility
    for a metric that decomposes into per-label binary metrics.  Arguments, defaults, return value
    and ``meta`` keys follow the reference (xcolumns/block_coordinate.py:296-499).

    Extra keywords (absorbed by the reference's ``**kwargs``): ``mode`` ("auto" | "exact" |
    "batched"), ``batch_size`` (rows per commit in batched mode), ``distributed`` (bool: y_proba is
    this rank's row shard of a torch.distributed job), ``y_pred_format`` ("same" | "indices": return
    the compact (n, k) int32 label ids instead of materialising a dense matrix)."""
    mode = kwargs.pop("mode", None)
    batch_size = kwargs.pop("batch_size", None)
    distributed = kwargs.pop("distributed", False)
    y_pred_format = kwargs.pop("y_pred_format", "same")

    log_info(f"Starting optimization of ETU metric using block coordinate "
             f"{'ascent (maximization)' if maximize else 'descent (minimization)'} algorithm ...", verbose)
    meta: Dict[str, Any] = {"utilities": [], "iters": 0, "time": time()}

    _check_k(k)
    if not isinstance(y_proba, (np.ndarray, torch.Tensor, csr_matrix)):
        raise ValueError("y_proba must be either np.ndarray, torch.Tensor, or csr_matrix")
    if metric_aggregation not in ("mean", "sum"):
        raise ValueError(f"Unsupported utility aggregation function: {metric_aggregation}, must be either 'mean' or 'sum'")
    if k < 0:
        raise ValueError("k must be >= 0")
    if k == 0 and isinstance(y_proba, csr_matrix):
        raise NotImplementedError("xcolumns_b200: BCA without a budget (k=0) is implemented for dense inputs only")
    try:
        metric_id, beta, eps = M.resolve_binary_metric(binary_metric_func, metric_kwargs)
    except M.UnsupportedMetricError:
        # any other callable, or a list of m callables (block_coordinate.py:54-129): evaluated on the device through
        # torch tensors instead of inside the fused kernels (xcolumns_b200/generic_metric.py)
        from .generic_metric import bca_generic
        return bca_generic(y_proba, binary_metric_func, k, metric_aggregation, normalize_conf_matrix, metric_kwargs,
                           maximize, tolerance, init_y_pred, max_iters, shuffle_order, skip_tn, return_meta, seed, verbose,
                           mode, batch_size, y_pred_format, _initial_pred, _finish_pred)
    mix = M.resolve_mix(binary_metric_func)
    if metric_id in M.TN_METRICS and skip_tn:
        log_warning("skip_tn=True with a metric that uses true negatives: tn is the constant -1 like in the reference")

    n, m = y_proba.shape
    if k > m:
        raise ValueError(f"k={k} is larger than the number of labels m={m}")
    greedy = isinstance(init_y_pred, str) and init_y_pred == "greedy"
    batchable = metric_id in M.AFFINE_GAIN_METRICS or metric_id in M.RECORD_GAIN_METRICS
    mode = _resolve_mode(mode, n, greedy or k == 0 or not batchable or (metric_id in M.TN_METRICS and skip_tn))

    device = dev.pick_device(y_proba)
    comm = make_comm(distributed, device)
    if k == 0:
        if mode != "exact" or comm.world > 1:
            raise NotImplementedError("xcolumns_b200: k=0 runs in the sequential mode on one GPU (mode='exact')")
        return _bca_k0_dense(y_proba, binary_metric_func, metric_id, beta, eps, metric_aggregation,
                             normalize_conf_matrix, maximize, tolerance, init_y_pred, max_iters, shuffle_order,
                             skip_tn, return_meta, seed, verbose, meta, device)
    is_csr = isinstance(y_proba, csr_matrix)
    # host inputs with a large dense result: clear the result on host threads while the GPU works
    prefill = None if (is_csr or y_pred_format != "same") else dev.DenseOutputPrefill.start(y_proba, n, m)
    data = dev.csr_to_device(y_proba, device) if is_csr else dev.dense_to_device(y_proba, device)

    n_div = n if normalize_conf_matrix else 1            # block_coordinate.py:403-405
    n_div_global = comm.n_global(n) if normalize_conf_matrix else 1
    n_order = n if normalize_conf_matrix else 1          # order = arange(n) after the overwrite (:414)
    params = _metric_params(metric_id, beta, eps, maximize, skip_tn, n_div_global, comm.n_global(n), mix)
    # the utility / stopping test ignores metric_kwargs (:63); precision@k keeps its k (it is bound, not a kwarg)
    util_beta = beta if metric_id == M.XC_METRIC_PREC_AT_K else 1.0
    util_params = _metric_params(metric_id, util_beta, 1e-9, maximize, skip_tn, n_div_global, comm.n_global(n), mix)

    timing = os.environ.get("XCOLUMNS_B200_TIMING") == "1"

    def _mark(name):
        if timing:  # phase wall-clock (synchronising; diagnostics only)
            torch.cuda.synchronize(device)
            meta.setdefault("timings", {})[name] = time() - meta["time"]

    _mark("h2d")
    sess = BcaSession(data, k, params, util_params, metric_aggregation, comm)
    sess.pred = _initial_pred(y_proba, data, init_y_pred, k, seed, device)
    meta["mode"] = mode
    meta["h2d_bytes"] = data.h2d_bytes
    _mark("init")

    if mode == "exact":
        if comm.world > 1:
            raise NotImplementedError("sequential-exact BCA does not shard (replicas only); use mode='batched'")
        rng = np.random.default_rng(seed)                # :413
        order = np.arange(n_order)                       # :414
        new_u = None
        for j in range(1, max_iters + 1):
            log_info(f"  Starting iteration {j}/{max_iters} ...", verbose)
            if shuffle_order:
                rng.shuffle(order)
            if greedy:
                sess.state.zero_()
            elif new_u is None:
                sess.recompute(XC_SUM_ORDERED)
            # (from the 2nd sweep on the state recomputed after the previous sweep is reused:
            #  the reference recomputes the identical sums again at :430)
            old_u = _host_utility(binary_metric_func, metric_aggregation, sess.state, n_div) if (
                new_u is None or greedy) else new_u
            order_dev = torch.from_numpy(order.astype(np.int32)).to(device)
            sess.sweep_exact(order_dev, greedy)
            sess.recompute(XC_SUM_ORDERED)
            new_u = _host_utility(binary_metric_func, metric_aggregation, sess.state, n_div)
            greedy = False
            meta["iters"] = j
            meta["utilities"].append(new_u)
            log_info(f"    Iteration {j}/{max_iters} finished, expected metric value: {old_u} -> {new_u}", verbose)
            if (maximize and new_u - old_u < tolerance) or (not maximize and new_u - old_u > tolerance):
                log_info(f"  Stopping because improvement of expected metric value is smaller than {tolerance}", verbose)
                break
    else:
        if greedy:
            raise NotImplementedError("init_y_pred='greedy' needs the sequential mode (mode='exact')")
        if not batchable:
            raise NotImplementedError(
                "xcolumns_b200 batched mode: Jaccard / G-mean / H-mean are not fused inside the mixed utilities; "
                "use mode='exact'")
        if metric_id in M.TN_METRICS and skip_tn:
            raise NotImplementedError("batched mode evaluates tn-based metrics with the real tn: pass skip_tn=False")
        if batch_size:
            batch = int(batch_size)
        elif metric_id == M.XC_METRIC_GMEAN:
            # sqrt couples the rows of a batch strongly (a never-predicted label looks equally attractive to every
            # row of the batch): measured on 6000 x 2000, n/8 and n/32 overshoot, ~100 rows per commit agree with
            # the sequential reference to 1e-6.  The damping below shrinks further if a sweep still regresses.
            batch = max(16, min(96, n_order // 64))
        else:
            batch = default_batch_rows(n_order, sess.wave_rows())
        n_batches = comm.max_int((n_order + batch - 1) // batch)
        batch_max = comm.max_int(batch)      # ragged shards: ranks may differ by a row, decisions must not
        min_batch = max(1, 16 // comm.world)  # every rank commits `batch` rows at once: the floor is per job, not per rank
        base_seed = (0x9E3779B97F4A7C15 * (1 + (0 if seed is None else int(seed))) + 7919 * comm.rank) & (2**64 - 1)
        sess._n_order = n_order
        meta["batch_size"] = batch
        # util_all[0]: utility of the initial prediction, util_all[j]: after sweep j (float64, device).  Sweep j+1 is
        # enqueued before sweep j's utility is read on the host, so the GPU queue never drains at the stopping
        # test; if sweep j turns out to be the last one, the snapshot taken before the speculative sweep is restored.
        util_all = torch.zeros(max_iters + 2, dtype=torch.float64, device=device)
        util_host = torch.zeros((max_iters + 2, 2), dtype=torch.float64).pin_memory()
        sess.recompute(XC_SUM_FAST)
        sess.utility_into(util_all)
        events, saved = {}, {}
        attempt = 0

        def enqueue(j):
            saved[j] = sess.run_sweep(j, base_seed + 0x632BE59BD9B4E019 * j + 0x9FB21C651E98DF25 * attempt, shuffle_order,
                                      batch, n_batches, j % 16 == 0, util_all[j:])
            util_host[j].copy_(util_all[j - 1:j + 1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(device))
            events[j] = ev

        enqueue(1)
        j = 1
        while j <= max_iters:
            log_info(f"  Starting iteration {j}/{max_iters} ...", verbose)
            if j + 1 <= max_iters and j + 1 not in events:
                enqueue(j + 1)
            events[j].synchronize()
            old_u, new_u = (float(v) for v in util_host[j])
            regressed = (new_u < old_u - 1e-12) if maximize else (new_u > old_u + 1e-12)
            if regressed:
                # Block-Jacobi overshoot (the rows of a batch all reacted to the same frozen state): the
                # sequential sweep this mode stands in for cannot lose utility.  Roll the sweep (and the
                # speculative one after it) back ...
                sess.restore(saved[j])
                events.clear()
                saved.clear()
                sess.recompute(XC_SUM_FAST)
                sess.utility_into(util_all[j - 1:])
                # (the decision is taken on numbers every rank agrees on: the utilities come from the
                #  replicated state, batch_max is all-reduced)
                if batch_max > min_batch and not batch_size:
                    # ... and repeat it with 4x more commits
                    batch = max(min_batch, batch // 4)
                    batch_max = max(min_batch, batch_max // 4)
                    n_batches = comm.max_int((n_order + batch - 1) // batch)
                    meta["batch_size"] = batch
                    attempt += 1
                    log_info(f"    Iteration {j} lost utility ({old_u} -> {new_u}); repeating with batches of {batch} rows", verbose)
                    enqueue(j)
                    continue
                # ... or, with a caller-chosen / already minimal batch, keep the prediction the sweep started
                # from: the result is never worse than what the sweep began with
                log_info(f"  Stopping: iteration {j} lost utility ({old_u} -> {new_u}) and cannot be repeated "
                         f"with smaller batches", verbose)
                break
            meta["iters"] = j
            meta["utilities"].append(new_u)
            log_info(f"    Iteration {j}/{max_iters} finished, expected metric value: {old_u} -> {new_u}", verbose)
            if (maximize and new_u - old_u < tolerance) or (not maximize and new_u - old_u > tolerance):
                log_info(f"  Stopping because improvement of expected metric value is smaller than {tolerance}", verbose)
                if j + 1 in saved:
                    sess.restore(saved[j + 1])    # undo the speculative sweep
                break
            saved.pop(j, None)
            j += 1
        sess.join()

    meta["launches"] = sess.ctx.launches()
    meta["lag"] = sess.lag if (mode == "batched" and sess.pipe) else 0
    meta["commit"] = "peer-memory" if sess.peer is not None else ("all-reduce" if comm.world > 1 else "local")
    sess.close()
    _mark("sweeps")
    y_pred = _finish_pred(y_proba, sess.pred, m, y_pred_format, prefill)
    _mark("output")
    if return_meta:
        meta["time"] = time() - meta["time"]
        return y_pred, meta
    return y_pred


def _bca_k0_dense(y_proba, binary_metric_func, metric_id, beta, eps, aggregation, normalize_conf_matrix, maximize,
                  tolerance, init_y_pred, max_iters, shuffle_order, skip_tn, return_meta, seed, verbose, meta, device):
    """k = 0: no budget, a label is predicted whenever its gain is >= 0 (block_coordinate.py:199-200).
    Labels are independent, so one thread per label walks the instance order (csrc/bca_exact.cu)."""
    ctx = dev.ctx_for(device)
    data = dev.dense_to_device(y_proba, device)
    n, m = data.n, data.m
    n_div = n if normalize_conf_matrix else 1
    n_order = n if normalize_conf_matrix else 1
    params = _metric_params(metric_id, beta, eps, maximize, skip_tn, n_div, n, M.resolve_mix(binary_metric_func))
    tdt = data.torch_dtype
    if isinstance(init_y_pred, str) and init_y_pred == "top":
        pred = (data.t[:, :m] >= 0).to(tdt)                       # predict_top_k(k=0): gains >= th = 0
    elif isinstance(init_y_pred, str) and init_y_pred in ("random", "greedy"):
        pred = torch.zeros((n, m), dtype=tdt, device=device)      # random_at_k with k = 0 draws nothing
    elif isinstance(init_y_pred, (np.ndarray, torch.Tensor)):
        if tuple(init_y_pred.shape) != (n, m):
            raise ValueError(f"init_y_pred must have shape (n, m) = ({n}, {m}), but has shape {init_y_pred.shape}")
        pred = dev.dense_to_device(init_y_pred, device, tdt, pad=False).t.clone()
    else:
        raise ValueError("init_y_pred must be a dense matrix or one of 'random', 'greedy', 'top'")
    pred = pred.contiguous()
    greedy = isinstance(init_y_pred, str) and init_y_pred == "greedy"
    state = torch.zeros((4, m), dtype=torch.float64, device=device)
    sp = lambda: dev.stream_ptr(device)

    def recompute():
        ctx.call("xc_confmat_dense", dev.ptr(data.t), data.ld, dev.ptr(pred), m, data.code, n, m, 0, XC_SUM_ORDERED, 0,
                 C.c_void_p(state[0].data_ptr()), C.c_void_p(state[1].data_ptr()), C.c_void_p(state[2].data_ptr()), sp())
        if skip_tn:
            state[3].fill_(-1.0)
        else:
            state[3] = -state[0] - state[1] - state[2] + n

    rng = np.random.default_rng(seed)
    order = np.arange(n_order)
    new_u = None
    for j in range(1, max_iters + 1):
        if shuffle_order:
            rng.shuffle(order)
        if greedy:
            state.zero_()
        elif new_u is None:
            recompute()
        old_u = _host_utility(binary_metric_func, aggregation, state, n_div) if (new_u is None or greedy) else new_u
        order_dev = torch.from_numpy(order.astype(np.int32)).to(device)
        ctx.call("xc_bca_exact_sweep_dense_k0", dev.ptr(data.t), data.code, n, m, data.ld, dev.ptr(order_dev),
                 int(order_dev.numel()), C.byref(params), int(greedy), dev.ptr(pred), m,
                 C.c_void_p(state[0].data_ptr()), C.c_void_p(state[1].data_ptr()), C.c_void_p(state[2].data_ptr()),
                 C.c_void_p(state[3].data_ptr()), sp())
        recompute()
        new_u = _host_utility(binary_metric_func, aggregation, state, n_div)
        greedy = False
        meta["iters"] = j
        meta["utilities"].append(new_u)
        log_info(f"    Iteration {j}/{max_iters} finished, expected metric value: {old_u} -> {new_u}", verbose)
        if (maximize and new_u - old_u < tolerance) or (not maximize and new_u - old_u > tolerance):
            break
    meta["mode"] = "exact"
    if isinstance(y_proba, torch.Tensor):
        y_pred = pred.to(device=y_proba.device, dtype=y_proba.dtype)
    else:
        y_pred = pred.cpu().numpy().astype(np.asarray(y_proba).dtype, copy=False)
    if return_meta:
        meta["time"] = time() - meta["time"]
        return y_pred, meta
    return y_pred


def _finish_pred(y_proba, pred: torch.Tensor, m: int, y_pred_format: str, prefill=None):
    if y_pred_format == "indices":
        return pred if isinstance(y_proba, torch.Tensor) and y_proba.is_cuda else pred.cpu().numpy()
    if isinstance(y_proba, csr_matrix):
        return dev.compact_to_csr_like(y_proba, pred, reference_padding=False)
    return dev.compact_to_dense_like(y_proba, pred, m, prefill=prefill)


# ------------------------------------------------------------------------------------------
# coverage (block_coordinate.py:600-701; CSR semantics for both layouts)
# ------------------------------------------------------------------------------------------

class CoverageSession:
    """State of one coverage-BCA run on one device (block_coordinate.py:539-701, CSR semantics for both layouts):
    Ef[j] = prod_i (1 - yhat_ij eta_ij) in float64, the compact prediction, the batch factors dEf.  With a
    torch.distributed communicator every rank holds a row shard: Ef is replicated, the per-batch factors and the
    recomputed state are all-reduced with a product."""

    def __init__(self, data, k: int, alpha: float, comm: Optional[Comm] = None):
        self.data = data
        self.is_csr = isinstance(data, dev.CsrDev)
        self.device = (data.data if self.is_csr else data.t).device
        self.ctx = dev.ctx_for(self.device)
        self.k, self.alpha = k, float(alpha)
        self.comm = comm or Comm(None)
        self.n, self.m = data.n, data.m
        f64 = dict(dtype=torch.float64, device=self.device)
        self.Ef = torch.empty(self.m, **f64)
        self.dEf = torch.ones(self.m, **f64)
        self.util_buf = torch.zeros(4, **f64)
        self.pred: Optional[torch.Tensor] = None
        self._csr_view = data if self.is_csr else None

    def _s(self):
        return dev.stream_ptr(self.device)

    def csr_view(self):
        """stored entries of every row: the rows themselves, or (sequential mode on dense rows only) the
        non-zeros of the dense rows -- what csr_matrix(dense) holds"""
        if self._csr_view is None:
            self._csr_view = _dense_as_csr(self.data)
        return self._csr_view

    def state(self, order: int) -> None:
        """Ef from the current prediction (numba_csr_functions.py:325-382 as called at :665 / :676)"""
        d, k = self.data, self.k
        if self.is_csr or order == XC_SUM_ORDERED:
            v = self.csr_view()
            self.ctx.call("xc_cov_state_csr", dev.ptr(v.data), v.code, dev.ptr(v.indices), dev.ptr(v.indptr), self.n,
                          self.m, dev.ptr(self.pred), k, order, dev.ptr(self.Ef), self._s())
        else:
            self.ctx.call("xc_cov_state_dense", dev.ptr(d.t), d.code, d.n, d.m, d.ld, dev.ptr(self.pred), k,
                          dev.ptr(self.Ef), self._s())
        self.comm.allreduce_prod_(self.Ef)

    def _precision_part(self) -> torch.Tensor:
        """sum_j tp_j / n / k on the device (:593-595)"""
        d, k = self.data, self.k
        tp = torch.zeros(self.m, dtype=torch.float64, device=self.device)
        fp, fn = torch.zeros_like(tp), torch.zeros_like(tp)
        if self.is_csr:
            self.ctx.call("xc_confmat_csr_compact", dev.ptr(d.data), dev.ptr(d.indices), dev.ptr(d.indptr), d.code,
                          dev.ptr(self.pred), k, self.n, self.m, XC_SUM_FAST, dev.ptr(tp), dev.ptr(fp), dev.ptr(fn), self._s())
        else:
            self.ctx.call("xc_confmat_dense_compact", dev.ptr(d.t), d.code, d.ld, dev.ptr(self.pred), k, d.n, d.m,
                          XC_SUM_FAST, None, dev.ptr(tp), dev.ptr(fp), dev.ptr(fn), self._s())
        self.comm.allreduce_sum_(tp)
        return (tp / self.comm.n_global(self.n) / k).sum()

    def utility_host(self) -> float:
        """the reference's expression on the host (numpy's pairwise mean: bit-equal in the sequential mode)"""
        cov = 1 - float(self.Ef.cpu().numpy().mean())       # :592
        if self.alpha < 1:
            cov = self.alpha * cov + (1 - self.alpha) * float(self._precision_part())
        return cov

    def utility_device(self, slot: int) -> None:
        self.ctx.call("xc_cov_utility", dev.ptr(self.Ef), self.m, C.c_void_p(self.util_buf[slot:].data_ptr()), self._s())
        if self.alpha < 1:
            self.util_buf[slot] = self.alpha * self.util_buf[slot] + (1 - self.alpha) * self._precision_part()

    def sweep_exact(self, order_dev: torch.Tensor, greedy: bool) -> None:
        v = self.csr_view()
        self.ctx.call("xc_cov_exact_sweep_csr", dev.ptr(v.data), v.code, dev.ptr(v.indices), dev.ptr(v.indptr), self.n,
                      self.m, dev.ptr(order_dev), int(order_dev.numel()), self.k, C.c_double(self.alpha), int(greedy),
                      dev.ptr(self.pred), dev.ptr(self.Ef), self._s())

    def sweep_batched(self, order_dev: torch.Tensor, batch: int, n_batches: Optional[int] = None) -> None:
        """one block-Jacobi sweep: a single process issues it through ONE C call; sharded rows all-reduce the
        batch factors (product) between the batch kernel and the fold"""
        d, k, n_loc = self.data, self.k, int(order_dev.numel())
        al = C.c_double(self.alpha)
        if self.comm.world == 1:
            if self.is_csr:
                self.ctx.call("xc_cov_sweep_csr", dev.ptr(d.data), d.code, dev.ptr(d.indices), dev.ptr(d.indptr), self.m,
                              dev.ptr(order_dev), n_loc, int(batch), k, al, dev.ptr(self.Ef), dev.ptr(self.pred),
                              dev.ptr(self.dEf), self._s())
            else:
                self.ctx.call("xc_cov_sweep_dense", dev.ptr(d.t), d.code, d.m, d.ld, dev.ptr(order_dev), n_loc,
                              int(batch), k, al, dev.ptr(self.Ef), dev.ptr(self.pred), dev.ptr(self.dEf), self._s())
            return
        nb = n_batches if n_batches is not None else (n_loc + batch - 1) // batch
        for b in range(nb):
            lo = min(b * batch, n_loc)
            hi = min(lo + batch, n_loc)
            if hi > lo:
                rows = C.c_void_p(order_dev.data_ptr() + 4 * lo)
                if self.is_csr:
                    self.ctx.call("xc_cov_batch_csr", dev.ptr(d.data), d.code, dev.ptr(d.indices), dev.ptr(d.indptr),
                                  rows, hi - lo, k, al, dev.ptr(self.Ef), dev.ptr(self.pred), dev.ptr(self.dEf), self._s())
                else:
                    self.ctx.call("xc_cov_batch_dense", dev.ptr(d.t), d.code, d.m, d.ld, rows, hi - lo, k, al,
                                  dev.ptr(self.Ef), dev.ptr(self.pred), dev.ptr(self.dEf), self._s())
            self.comm.allreduce_prod_(self.dEf)
            self.ctx.call("xc_cov_fold", dev.ptr(self.Ef), dev.ptr(self.dEf), self.m, self._s())


def coverage_batch_rows(n: int) -> int:
    """rows one rank commits together in the batched coverage sweep: coverage couples the rows of a batch much
    more strongly than the F-measures (every row of a batch rushes to the same uncovered label): n/32 keeps the
    block-Jacobi fixed point within 1e-4 of the sequential one, n/8 does not (measured, tests/test_gpu_parity.py)"""
    return max(1, min(4096, n // 32))


def predict_optimizing_coverage_using_bc(
    y_proba: Matrix,
    k: int,
    alpha: float = 1,
    tolerance: float = 1e-6,
    init_y_pred: Union[str, Matrix] = "top",
    max_iters: int = 100,
    shuffle_order: bool = True,
    return_meta: bool = False,
    seed: Optional[int] = None,
    verbose: bool = False,
    **kwargs,
) -> Union[Matrix, Tuple[Matrix, Dict[str, Any]]]:
    """Block coordinate ascent for coverage@k (optionally mixed with precision@k through alpha).
    State: Ef[j] = prod_i (1 - yhat_ij * eta_ij), the probability that label j is never covered.
    Same arguments as the reference (block_coordinate.py:600-701); ``mode`` / ``batch_size`` /
    ``distributed`` / ``y_pred_format`` as in :func:`predict_using_bc_with_0approx`."""
    mode = kwargs.pop("mode", None)
    batch_size = kwargs.pop("batch_size", None)
    distributed = kwargs.pop("distributed", False)
    y_pred_format = kwargs.pop("y_pred_format", "same")
    log_info(f"Starting optimization of ETU coverage@{k} metric using block coordinate ascent algorithm ...", verbose)
    if not isinstance(k, int) or k <= 0:
        raise ValueError("k must be an integer > 0")
    _check_k(k)
    if not isinstance(y_proba, (np.ndarray, torch.Tensor, csr_matrix)):
        raise ValueError("y_proba must be either np.ndarray or csr_matrix")
    n, m = y_proba.shape
    meta: Dict[str, Any] = {"utilities": [], "iters": 0, "time": time()}
    if seed is not None:
        np.random.seed(seed)  # global side effect kept from the reference (:632-633)
    greedy = isinstance(init_y_pred, str) and init_y_pred == "greedy"
    mode = _resolve_mode(mode, n, greedy)
    device = dev.pick_device(y_proba)
    comm = make_comm(distributed, device)
    if mode == "exact" and comm.world > 1:
        raise NotImplementedError("sequential-exact coverage BCA does not shard (replicas only); use mode='batched'")
    is_csr = isinstance(y_proba, csr_matrix)
    data = dev.csr_to_device(y_proba, device) if is_csr else dev.dense_to_device(y_proba, device)
    sess = CoverageSession(data, k, alpha, comm)
    sess.pred = _initial_pred(y_proba, data, init_y_pred if not greedy else "random", k, seed, device)
    meta["mode"] = mode
    if mode == "exact":
        rng = np.random.default_rng(seed)
        order = np.arange(n)
        new_cov = None
        for j in range(1, max_iters + 1):
            log_info(f"  Starting iteration {j}/{max_iters} ...", verbose)
            if shuffle_order:
                rng.shuffle(order)
            if greedy:
                sess.Ef.fill_(1.0)
                old_cov = sess.utility_host()
            elif new_cov is None:
                sess.state(XC_SUM_ORDERED)
                old_cov = sess.utility_host()
            else:
                old_cov = new_cov
            order_dev = torch.from_numpy(order.astype(np.int32)).to(device)
            sess.sweep_exact(order_dev, greedy)
            sess.state(XC_SUM_ORDERED)
            new_cov = sess.utility_host()
            greedy = False
            meta["iters"] = j
            meta["utilities"].append(new_cov)
            log_info(f"    Iteration {j}/{max_iters} finished, expected coverage: {old_cov} -> {new_cov}", verbose)
            if new_cov <= old_cov + tolerance:              # :690
                log_info(f"  Stopping because improvement of expected coverage is smaller than {tolerance}", verbose)
                break
    else:
        batch = int(batch_size) if batch_size else coverage_batch_rows(n)
        n_batches = comm.max_int((n + batch - 1) // batch)
        batch_max = comm.max_int(batch)
        min_batch = max(1, 16 // comm.world)   # every rank commits `batch` rows at once: the floor is per job, not per rank
        base_seed = (0x9E3779B97F4A7C15 * (1 + (0 if seed is None else int(seed))) + 7919 * comm.rank) & (2**64 - 1)
        order_dev = torch.arange(n, dtype=torch.int32, device=device)
        sess.state(XC_SUM_FAST)
        sess.utility_device(0)
        meta["batch_size"] = batch
        j, attempt = 1, 0
        while j <= max_iters:
            log_info(f"  Starting iteration {j}/{max_iters} ...", verbose)
            saved = sess.pred.clone()
            if shuffle_order:
                sess.ctx.call("xc_permutation", n, C.c_uint64((base_seed + 0x632BE59BD9B4E019 * j +
                                                               0x9FB21C651E98DF25 * attempt) & (2**64 - 1)),
                              dev.ptr(order_dev), sess._s())
            sess.sweep_batched(order_dev, batch, n_batches)
            sess.state(XC_SUM_FAST)                         # :676 (the products are recomputed, like the reference)
            sess.utility_device(1)
            old_cov, new_cov = (float(v) for v in sess.util_buf[:2].cpu())
            if new_cov < old_cov - 1e-12:
                # block-Jacobi overshoot: the rows of a batch all rushed to the same uncovered labels, which the
                # sequential sweep this mode stands in for cannot do (it never loses coverage).  Roll the sweep back
                # and repeat it with 4x more commits; with a caller-chosen or already minimal batch keep the
                # prediction the sweep started from.
                sess.pred = saved
                sess.state(XC_SUM_FAST)
                sess.utility_device(0)
                if batch_max > min_batch and not batch_size:
                    batch, batch_max = max(min_batch, batch // 4), max(min_batch, batch_max // 4)
                    n_batches = comm.max_int((n + batch - 1) // batch)
                    meta["batch_size"] = batch
                    attempt += 1
                    log_info(f"    Iteration {j} lost coverage ({old_cov} -> {new_cov}); repeating with batches of {batch} rows", verbose)
                    continue
                break
            sess.util_buf[0] = sess.util_buf[1]
            meta["iters"] = j
            meta["utilities"].append(new_cov)
            log_info(f"    Iteration {j}/{max_iters} finished, expected coverage: {old_cov} -> {new_cov}", verbose)
            if new_cov <= old_cov + tolerance:              # :690
                log_info(f"  Stopping because improvement of expected coverage is smaller than {tolerance}", verbose)
                break
            j += 1
    y_pred = _finish_pred(y_proba, sess.pred, m, y_pred_format)
    if return_meta:
        meta["time"] = time() - meta["time"]
        return y_pred, meta
    return y_pred


def _dense_as_csr(d: dev.DenseDev) -> dev.CsrDev:
    """Dense rows seen as CSR rows that store every non-zero label (what csr_matrix(dense) holds); used only by
    the SEQUENTIAL coverage kernels (small inputs), which are defined on stored entries.  The batched mode works
    on the dense rows directly (xc_cov_batch_dense / xc_cov_state_dense)."""
    x = d.t[:, :d.m]
    nz = x != 0
    counts = nz.sum(1)
    indptr = torch.zeros(d.n + 1, dtype=torch.int64, device=x.device)
    indptr[1:] = counts.cumsum(0)
    indices = nz.nonzero()[:, 1].to(torch.int32).contiguous()
    return dev.CsrDev(x[nz].contiguous(), indices, indptr, d.n, d.m, d.code, 0)


# ------------------------------------------------------------------------------------------
# wrappers (block_coordinate.py:709-801)
# ------------------------------------------------------------------------------------------

def make_bc_wrapper(binary_metric_func: Callable, metric_name: str, maximize: bool = True,
                    metric_aggregation: str = "mean", skip_tn: bool = False, warn_k_eq_0: bool = False):
    """Factory of ``predict_optimizing_<metric>_using_bc(y_proba, k, **kwargs)`` wrappers around
    :func:`predict_using_bc_with_0approx`; the wrapper exposes the forwarded keyword arguments in
    its ``__signature__`` so callers can filter kwargs by name."""

    def predict_optimizing_metric_using_bc(y_proba: Matrix, k: int, **kwargs):
        if warn_k_eq_0 and k == 0:
            log_warning(f"Warning: k=0 results in degenerated solution for {metric_name}!")
        return predict_using_bc_with_0approx(y_proba, binary_metric_func, k, metric_aggregation=metric_aggregation,
                                             maximize=maximize, skip_tn=skip_tn, **kwargs)

    predict_optimizing_metric_using_bc.__doc__ = (
        f"Predict for a given test set optimizing {metric_name} with block coordinate "
        f"{'ascent' if maximize else 'descent'}; equivalent to predict_using_bc_with_0approx(y_proba, "
        f"{binary_metric_func.__name__}, k, metric_aggregation={metric_aggregation!r}, maximize={maximize}, "
        f"skip_tn={skip_tn}, ...).")
    return add_kwargs_to_signature(predict_optimizing_metric_using_bc, predict_using_bc_with_0approx,
                                   skip=["metric_func", "metric_aggregation", "maximize", "skip_tn"])


predict_optimizing_macro_precision_using_bc = make_bc_wrapper(
    M.binary_precision_on_conf_matrix, "macro-averaged precision", skip_tn=True, warn_k_eq_0=True)
predict_optimizing_macro_recall_using_bc = make_bc_wrapper(
    M.binary_recall_on_conf_matrix, "macro-averaged recall", skip_tn=True, warn_k_eq_0=True)
predict_optimizing_macro_f1_score_using_bc = make_bc_wrapper(
    M.binary_f1_score_on_conf_matrix, "macro-averaged F1 score", skip_tn=True)
predict_optimizing_macro_jaccard_score_using_bc = make_bc_wrapper(
    M.binary_jaccard_score_on_conf_matrix, "macro-averaged Jaccard score", skip_tn=True)
predict_optimizing_macro_balanced_accuracy_using_bc = make_bc_wrapper(
    M.binary_balanced_accuracy_on_conf_matrix, "macro-averaged balanced accuracy")
predict_optimizing_macro_hmean_using_bc = make_bc_wrapper(M.binary_hmean_on_conf_matrix, "macro-averaged H-mean")
predict_optimizing_macro_gmean_using_bc = make_bc_wrapper(M.binary_gmean_on_conf_matrix, "macro-averaged G-mean")


# ------------------------------------------------------------------------------------------
# instance precision and the mixed instance-precision / macro-metric utilities
# (block_coordinate.py:804-1045)
# ------------------------------------------------------------------------------------------

def predict_optimizing_instance_precision_using_bc(y_proba: Matrix, k: int, tolerance: float = 1e-6,
                                                   init_y_pred: Union[str, Matrix] = "random", max_iters: int = 100,
                                                   shuffle_order: bool = True, verbose: bool = False,
                                                   return_meta: bool = False, **kwargs):
    """BCA with instance precision@k (sum over labels of tp / k) as the target
    (xcolumns/block_coordinate.py:804-838; same defaults, note init_y_pred="random")."""
    return predict_using_bc_with_0approx(y_proba, binary_metric_func=M.PrecisionAtK(k), k=k, metric_aggregation="sum",
                                         tolerance=tolerance, init_y_pred=init_y_pred, max_iters=max_iters,
                                         shuffle_order=shuffle_order, verbose=verbose, return_meta=return_meta,
                                         **kwargs)


def make_mixed_bc_wrapper(binary_metric_func: Callable, metric_name: str):
    """``predict_optimizing_mixed_instance_precision_and_<metric>_using_bc(y_proba, k, alpha=1, **kwargs)``:
    BCA on  sum_j [(1 - alpha) * tp_j / k + alpha * metric_j / m]  (xcolumns/block_coordinate.py:848-1045;
    like there: aggregation "sum", skip_tn=True -- for the tn-based metrics that reproduces the reference's
    constant tn = -1, which only the sequential mode evaluates)."""

    def predict_optimizing_mixed_metric_using_bc(y_proba: Matrix, k: int, alpha: float = 1, **kwargs):
        n, m = y_proba.shape
        return predict_using_bc_with_0approx(
            y_proba, binary_metric_func=M.MixedInstancePrecisionMetric(binary_metric_func, alpha, k, m), k=k,
            metric_aggregation="sum", skip_tn=True, **kwargs)

    predict_optimizing_mixed_metric_using_bc.__doc__ = (
        f"BCA with a weighted average of instance precision@k and {metric_name} as the target; "
        f"see predict_using_bc_with_0approx.")
    return predict_optimizing_mixed_metric_using_bc


predict_optimizing_mixed_instance_precision_and_macro_precision_using_bc = make_mixed_bc_wrapper(
    M.binary_precision_on_conf_matrix, "macro-averaged precision")
predict_optimizing_mixed_instance_precision_and_macro_recall_using_bc = make_mixed_bc_wrapper(
    M.binary_recall_on_conf_matrix, "macro-averaged recall")
predict_optimizing_mixed_instance_precision_and_macro_f1_score_using_bc = make_mixed_bc_wrapper(
    M.binary_f1_score_on_conf_matrix, "macro-averaged F1 score")
predict_optimizing_mixed_instance_precision_and_macro_balanced_accuracy_using_bc = make_mixed_bc_wrapper(
    M.binary_balanced_accuracy_on_conf_matrix, "macro-averaged balanced accuracy")
predict_optimizing_mixed_instance_precision_and_macro_jaccard_score_using_bc = make_mixed_bc_wrapper(
    M.binary_jaccard_score_on_conf_matrix, "macro-averaged Jaccard score")
predict_optimizing_mixed_instance_precision_and_macro_gmean_using_bc = make_mixed_bc_wrapper(
    M.binary_gmean_on_conf_matrix, "macro-averaged G-mean")
predict_optimizing_mixed_instance_precision_and_macro_hmean_using_bc = make_mixed_bc_wrapper(
    M.binary_hmean_on_conf_matrix, "macro-averaged H-mean")
