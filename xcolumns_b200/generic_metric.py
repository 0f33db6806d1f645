"""BCA for metric callables the fused kernels do not know (xcolumns/block_coordinate.py:54-129: ``binary_metric_func``
may be ANY callable of (tp, fp, fn, tn) -- or a list of m callables, one per label).

A user-supplied Python callable cannot run inside a CUDA kernel.  It can, however, run ON the device: the
reference's contract is that the callable is plain arithmetic on arrays, so it is handed float64 torch tensors
that live on the GPU and every operation it performs executes there.  Two drivers, both without any host-side
arithmetic on the data:

``exact``    the reference's sequential sweep, one instance at a time, with the reference's operation order
             (block_coordinate.py:132-209) on m-length device vectors; ties go to the lowest label id.
``batched``  block-Jacobi: the rows of a batch see the same frozen state, the callable is evaluated once per batch
             on (rows x labels) tensors (it must broadcast, which arithmetic callables do).

A list of m callables is grouped by callable object: every distinct callable is evaluated once on the slice of its
labels (the reference calls each of the m callables on scalars, block_coordinate.py:100-121).

This path is orders of magnitude slower than the fused kernels (dozens of elementwise launches per step instead of
one streaming pass); it exists so that every ``binary_metric_func`` the reference accepts works here too.
"""
from __future__ import annotations

import ctypes as C
from time import time
from typing import Any, Callable, Dict, List, Optional, Sequence, Union

import numpy as np
import torch
from scipy.sparse import csr_matrix

from . import _device as dev
from .utils import log_info


class _Metric:
    """binary_metric_func (callable or list of m callables) on float64 device tensors whose LAST axis is labels."""

    def __init__(self, func: Union[Callable, Sequence[Callable]], m: int, kwargs: Optional[Dict[str, Any]], device):
        self.kwargs = dict(kwargs or {})
        self.single = callable(func)
        if self.single:
            self.func = func
        else:
            funcs = list(func)
            if len(funcs) != m:
                raise ValueError(f"binary_metric_func must be a callable or a list of {m} callables, got {len(funcs)}")
            groups: Dict[int, List[int]] = {}
            self.funcs = {}
            for j, f in enumerate(funcs):
                if not callable(f):
                    raise ValueError("binary_metric_func must be a callable or a list of callables")
                groups.setdefault(id(f), []).append(j)
                self.funcs[id(f)] = f
            self.groups = [(self.funcs[key], torch.tensor(idx, dtype=torch.int64, device=device))
                           for key, idx in groups.items()]

    def __call__(self, tp, fp, fn, tn, with_kwargs: bool = True):
        kw = self.kwargs if with_kwargs else {}
        if self.single:
            out = self.func(tp, fp, fn, tn, **kw)
        else:
            out = torch.empty_like(tp)
            for f, idx in self.groups:
                out[..., idx] = f(tp[..., idx], fp[..., idx], fn[..., idx], tn[..., idx], **kw)
        if not isinstance(out, torch.Tensor):
            raise ValueError(f"binary_metric_func must return a tensor of the shape of its inputs, but returned {type(out)}")
        if out.shape != tp.shape:
            raise ValueError(f"binary_metric_func must return a tensor of shape {tuple(tp.shape)}, but returned {tuple(out.shape)}")
        return out


def _utility(metric: _Metric, aggregation: str, state: torch.Tensor, n_div) -> float:
    # block_coordinate.py:54-90 (metric_kwargs are NOT forwarded there, :63)
    vals = metric(state[0] / n_div, state[1] / n_div, state[2] / n_div, state[3] / n_div, with_kwargs=False)
    return float(vals.sum() if aggregation == "sum" else vals.mean())


def _topk_lowest_index(neg_gains: torch.Tensor, k: int) -> torch.Tensor:
    """the k smallest entries of the last axis, ties to the lowest index, returned ascending by index"""
    order = torch.sort(neg_gains, dim=-1, stable=True).indices[..., :k]
    return torch.sort(order, dim=-1).values


def _state_from_pred(eta: torch.Tensor, pred_mask: torch.Tensor, skip_tn: bool, n: int) -> torch.Tensor:
    """confusion sums of a 0/1 prediction (confusion_matrix.py:364-399 with dtype float64), rows in order"""
    e64 = eta.double()
    one_m = (1 - eta).double()          # (1 - eta) in the data dtype, like numpy's weak-scalar promotion
    p = pred_mask
    tp = torch.where(p, e64, torch.zeros_like(e64)).sum(0)
    fp = torch.where(p, one_m, torch.zeros_like(e64)).sum(0)
    fn = torch.where(p, torch.zeros_like(e64), e64).sum(0)
    tn = torch.full_like(tp, -1.0) if skip_tn else -tp - fp - fn + n
    return torch.stack([tp, fp, fn, tn])


def bca_generic(y_proba, binary_metric_func, k: int, metric_aggregation: str, normalize_conf_matrix: bool,
                metric_kwargs, maximize: bool, tolerance: float, init_y_pred, max_iters: int, shuffle_order: bool,
                skip_tn: bool, return_meta: bool, seed, verbose: bool, mode: Optional[str], batch_size: Optional[int],
                y_pred_format: str, initial_pred_fn, finish_pred_fn):
    """predict_using_bc_with_0approx for an arbitrary callable; dense inputs (numpy / torch) here, CSR rows in
    ``_bca_generic_csr``."""
    if isinstance(y_proba, csr_matrix):
        return _bca_generic_csr(y_proba, binary_metric_func, k, metric_aggregation, normalize_conf_matrix, metric_kwargs,
                                maximize, tolerance, init_y_pred, max_iters, shuffle_order, skip_tn, return_meta, seed,
                                verbose, mode, batch_size, y_pred_format, initial_pred_fn, finish_pred_fn)
    if k <= 0:
        raise NotImplementedError("xcolumns_b200: arbitrary metric callables need a budget k > 0")
    meta: Dict[str, Any] = {"utilities": [], "iters": 0, "time": time()}
    device = dev.pick_device(y_proba)
    data = dev.dense_to_device(y_proba, device, pad=False)
    eta = data.t[:, :data.m]
    n, m = eta.shape
    metric = _Metric(binary_metric_func, m, metric_kwargs, device)
    n_div = n if normalize_conf_matrix else 1
    n_order = n if normalize_conf_matrix else 1          # block_coordinate.py:403-414
    pred_idx = initial_pred_fn(y_proba, data, init_y_pred, k, seed, device).long()
    greedy = isinstance(init_y_pred, str) and init_y_pred == "greedy"
    pred = torch.zeros((n, m), dtype=torch.bool, device=device)
    pred.scatter_(1, pred_idx.clamp_min(0), True)
    mode = mode or "auto"
    if mode == "auto":
        mode = "exact" if (n <= 4096 or greedy) else "batched"
    if greedy and mode != "exact":
        raise NotImplementedError("init_y_pred='greedy' needs the sequential mode (mode='exact')")
    meta["mode"] = mode
    sgn = -1.0 if maximize else 1.0                       # the k smallest of -gain (:187-198)
    rng = np.random.default_rng(seed)
    order = np.arange(n_order)
    state = None
    new_u = None
    for j in range(1, max_iters + 1):
        log_info(f"  Starting iteration {j}/{max_iters} ...", verbose)
        if shuffle_order:
            rng.shuffle(order)
        if greedy:
            state = torch.zeros((4, m), dtype=torch.float64, device=device)
            if skip_tn:
                state[3].fill_(-1.0)
        elif new_u is None:
            state = _state_from_pred(eta, pred, skip_tn, n)
        old_u = _utility(metric, metric_aggregation, state, n_div) if (new_u is None or greedy) else new_u
        tp, fp, fn, tn = state[0], state[1], state[2], state[3]
        if mode == "exact":
            for i in order.tolist():
                e = eta[i]
                om = 1 - e
                if not greedy:                                # :158-163
                    yp = pred[i].to(e.dtype)
                    tp -= yp * e
                    fp -= yp * om
                    fn -= (1 - yp) * e
                    if not skip_tn:
                        tn -= (1 - yp) * om
                pos_tp, pos_fp, neg_fn = tp + e, fp + om, fn + e   # :166-172
                neg_tn = tn if skip_tn else tn + om
                gains = metric(pos_tp / n_div, pos_fp / n_div, fn / n_div, tn / n_div) - metric(
                    tp / n_div, fp / n_div, neg_fn / n_div, neg_tn / n_div)   # :174-185
                sel = _topk_lowest_index(sgn * gains, k)
                pred[i] = False
                pred[i, sel] = True
                yp = pred[i].to(e.dtype)                      # :204-209
                tp += yp * e
                fp += yp * om
                fn += (1 - yp) * e
                if not skip_tn:
                    tn += (1 - yp) * om
        else:
            b = int(batch_size) if batch_size else max(1, n_order // 8)
            b = max(1, min(b, (256 << 20) // (8 * m)))        # rows x labels float64 temporaries: <= 256 MB each
            ord_dev = torch.from_numpy(order).to(device)
            for lo in range(0, n_order, b):
                rows = ord_dev[lo:lo + b]
                e = eta[rows].double()
                om = (1 - eta[rows]).double()
                p = pred[rows]
                zero = torch.zeros_like(e)
                tp0 = tp - torch.where(p, e, zero)            # every row sees the frozen state minus its own share
                fp0 = fp - torch.where(p, om, zero)
                fn0 = fn - torch.where(p, zero, e)
                tn0 = tn.expand_as(e) if skip_tn else tn - torch.where(p, zero, om)
                gains = metric((tp0 + e) / n_div, (fp0 + om) / n_div, fn0 / n_div, tn0 / n_div) - metric(
                    tp0 / n_div, fp0 / n_div, (fn0 + e) / n_div, (tn0 if skip_tn else tn0 + om) / n_div)
                sel = _topk_lowest_index(sgn * gains, k)
                newp = torch.zeros_like(p)
                newp.scatter_(1, sel, True)
                d_tp = (torch.where(newp, e, zero) - torch.where(p, e, zero)).sum(0)
                d_fp = (torch.where(newp, om, zero) - torch.where(p, om, zero)).sum(0)
                tp += d_tp
                fp += d_fp
                fn -= d_tp
                if not skip_tn:
                    tn -= d_fp
                pred[rows] = newp
        state = _state_from_pred(eta, pred, skip_tn, n)      # :465-467
        new_u = _utility(metric, metric_aggregation, state, n_div)
        greedy = False
        meta["iters"] = j
        meta["utilities"].append(new_u)
        log_info(f"    Iteration {j}/{max_iters} finished, expected metric value: {old_u} -> {new_u}", verbose)
        if (maximize and new_u - old_u < tolerance) or (not maximize and new_u - old_u > tolerance):
            log_info(f"  Stopping because improvement of expected metric value is smaller than {tolerance}", verbose)
            break
    idx = torch.nonzero(pred)[:, 1].reshape(n, k).to(torch.int32)
    y_pred = finish_pred_fn(y_proba, idx, m, y_pred_format, None)
    if return_meta:
        meta["time"] = time() - meta["time"]
        return y_pred, meta
    return y_pred


# ------------------------------------------------------------------------------------------
# CSR rows (block_coordinate.py:212-293): only the STORED labels of a row are candidates; labels the row does not
# store can still sit in an initial prediction and then count as false positives until the row is visited
# ------------------------------------------------------------------------------------------

def csr_state_from_pred(data: torch.Tensor, indices: torch.Tensor, row_of: torch.Tensor, pred: torch.Tensor, m: int,
                        skip_tn: bool, n: int) -> torch.Tensor:
    """confusion sums [tp, fp, fn, tn] (float64) of a compact prediction ([n, k] label ids, -1 = unused) against CSR
    rows (confusion_matrix.py:386-399 over numba_csr_functions.py:116-258): fp counts (1 - eta) for stored labels and
    1 for predicted labels the row does not store."""
    dev_ = data.device
    in_pred = torch.zeros(data.numel(), dtype=torch.bool, device=dev_)
    step = 1 << 22
    for s0 in range(0, data.numel(), step):
        sl = slice(s0, s0 + step)
        in_pred[sl] = (pred[row_of[sl]] == indices[sl, None]).any(1)
    z = lambda: torch.zeros(m, dtype=torch.float64, device=dev_)
    tp = z().index_add_(0, indices[in_pred], data[in_pred].double())
    fp = z().index_add_(0, indices[in_pred], (1 - data[in_pred]).double())
    fn = z().index_add_(0, indices[~in_pred], data[~in_pred].double())
    predicted = torch.bincount(pred[pred >= 0], minlength=m)
    stored = torch.bincount(indices[in_pred], minlength=m)
    fp += (predicted - stored).double()
    tn = torch.full_like(tp, -1.0) if skip_tn else -tp - fp - fn + n
    return torch.stack([tp, fp, fn, tn])


def bca_generic_csr_core(data: torch.Tensor, indices: torch.Tensor, indptr: np.ndarray, n: int, m: int,
                         pred: torch.Tensor, metric: "_Metric", k: int, metric_aggregation: str,
                         normalize_conf_matrix: bool, maximize: bool, tolerance: float, greedy: bool, max_iters: int,
                         shuffle_order: bool, skip_tn: bool, seed, verbose: bool, mode: str, batch_size: Optional[int],
                         meta: Dict[str, Any], state_fn: Optional[Callable] = None) -> torch.Tensor:
    """The sweeps of predict_using_bc_with_0approx on CSR rows for a callable, on tensors of ANY device (the CPU tests
    drive it with CPU tensors).  ``data`` [nnz] float, ``indices`` [nnz] int64 ascending inside a row, ``indptr`` host
    int64, ``pred`` [n, k] int64 label ids with -1 for unused slots.  ``state_fn(pred) -> [4, m]`` recomputes the
    confusion sums at the sweep boundaries (default: ``csr_state_from_pred``; on the GPU the ordered C kernel, so that
    the sums are added in the reference's row order instead of the order atomics happen to land in).  Returns the final
    compact prediction."""
    dev_ = data.device
    if state_fn is None:
        state_fn = lambda pr: csr_state_from_pred(data, indices, row_of, pr, m, skip_tn, n)
    n_true = n
    n_div = n if normalize_conf_matrix else 1           # utilities (block_coordinate.py:403-405)
    n_order = n if normalize_conf_matrix else 1          # :403-414
    n_step = n_true                                      # the step functions keep their default normalisation (:446-459)
    lens_host = np.diff(indptr)
    row_of = torch.repeat_interleave(torch.arange(n, device=dev_), torch.from_numpy(lens_host).to(dev_))
    # valid slots first, ascending
    pred = torch.where(pred < 0, torch.full_like(pred, m), pred).sort(dim=1).values
    cnt_host = (pred < m).sum(1).cpu().numpy()
    pred = torch.where(pred == m, torch.full_like(pred, -1), pred)
    sgn = -1.0 if maximize else 1.0
    rng = np.random.default_rng(seed)
    order = np.arange(n_order)
    state, new_u = None, None
    for j in range(1, max_iters + 1):
        log_info(f"  Starting iteration {j}/{max_iters} ...", verbose)
        if shuffle_order:
            rng.shuffle(order)
        if greedy:
            state = torch.zeros((4, m), dtype=torch.float64, device=dev_)
        elif new_u is None:
            state = state_fn(pred)
        old_u = _utility(metric, metric_aggregation, state, n_div) if (new_u is None or greedy) else new_u
        tp, fp, fn, tn = state[0], state[1], state[2], state[3]
        if mode == "exact":
            for i in order.tolist():
                s0, e0 = int(indptr[i]), int(indptr[i + 1])
                t_idx, t = indices[s0:e0], data[s0:e0]
                om = 1 - t

                def own_share(p):
                    """(tp, fp, fn) contributions of prediction p of this row, as the reference forms them
                    (numba_csr_functions.py:380-452): products in the data dtype, widened when added"""
                    eq = p[:, None] == t_idx[None, :]
                    t_at_p = (eq.to(t.dtype) * t[None, :]).sum(1)          # eta of a predicted label, 0 if not stored
                    p_at_t = eq.any(0)
                    return t_at_p.double(), (1 - t_at_p).double(), (t * (1 - p_at_t.to(t.dtype))).double()

                if not greedy:
                    p = pred[i, :cnt_host[i]]
                    tp_d, fp_d, fn_d = own_share(p)
                    tp.index_add_(0, p, -tp_d)
                    fp.index_add_(0, p, -fp_d)
                    fn.index_add_(0, t_idx, -fn_d)
                    if not skip_tn:
                        tn -= 1
                        tn.index_add_(0, p, tp_d)
                        tn.index_add_(0, p, fp_d)
                        tn.index_add_(0, t_idx, fn_d)
                neg_tp, neg_fp, pos_fn = tp[t_idx], fp[t_idx], fn[t_idx]
                pos_tp = (neg_tp + t) / n_step
                pos_fp = (neg_fp + om) / n_step
                neg_fn = (pos_fn + t) / n_step
                neg_tp, neg_fp, pos_fn = neg_tp / n_step, neg_fp / n_step, pos_fn / n_step
                pos_tn = tn[t_idx]
                neg_tn = pos_tn
                if not skip_tn:
                    neg_tn = (pos_tn + om) / n_step
                    pos_tn = pos_tn / n_step
                gains = metric(pos_tp, pos_fp, pos_fn, pos_tn) - metric(neg_tp, neg_fp, neg_fn, neg_tn)
                new_p = t_idx[_topk_lowest_index(sgn * gains, k)] if e0 - s0 > k else t_idx
                cnt_host[i] = new_p.numel()
                pred[i] = -1
                pred[i, :cnt_host[i]] = new_p
                tp_d, fp_d, fn_d = own_share(new_p)
                tp.index_add_(0, new_p, tp_d)
                fp.index_add_(0, new_p, fp_d)
                fn.index_add_(0, t_idx, fn_d)
                if not skip_tn:
                    tn += 1
                    tn.index_add_(0, new_p, -tp_d)
                    tn.index_add_(0, new_p, -fp_d)
                    tn.index_add_(0, t_idx, -fn_d)
        else:
            b = int(batch_size) if batch_size else max(1, n_order // 8)
            ar = None
            for lo in range(0, n_order, b):
                rows_h = order[lo:lo + b]
                lens_h = lens_host[rows_h]
                L = int(lens_h.max()) if rows_h.size else 0
                if L == 0:
                    continue
                # rows x L padded tiles of float64 temporaries: keep each below ~256 MB by splitting the batch
                sub = max(1, min(rows_h.size, (256 << 20) // (8 * L * max(1, k))))
                for q in range(0, rows_h.size, sub):
                    rh = rows_h[q:q + sub]
                    rows = torch.from_numpy(rh).to(dev_)
                    lens = torch.from_numpy(lens_host[rh]).to(dev_)
                    starts = torch.from_numpy(indptr[rh]).to(dev_)
                    Lq = int(lens_host[rh].max())
                    if Lq == 0:
                        continue
                    ar = torch.arange(Lq, device=dev_)
                    valid = ar[None, :] < lens[:, None]
                    pos = torch.where(valid, starts[:, None] + ar[None, :], torch.zeros_like(valid, dtype=torch.int64))
                    tix = torch.where(valid, indices[pos], torch.zeros_like(pos))
                    tq = torch.where(valid, data[pos], torch.zeros_like(data[pos]))
                    e = tq.double()
                    om = (1 - tq).double()
                    p = pred[rows]
                    eq = (tix[:, :, None] == p[:, None, :]) & valid[:, :, None]
                    p_at_t = eq.any(2)
                    zero = torch.zeros_like(e)
                    tp0 = tp[tix] - torch.where(p_at_t, e, zero)     # frozen state minus the row's own share
                    fp0 = fp[tix] - torch.where(p_at_t, om, zero)
                    fn0 = fn[tix] - torch.where(p_at_t, zero, e)
                    if skip_tn:
                        pos_tn = neg_tn = tn[tix]
                    else:
                        tn0 = tn[tix] - torch.where(p_at_t, zero, om)
                        pos_tn, neg_tn = tn0 / n_step, (tn0 + om) / n_step
                    gains = metric((tp0 + e) / n_step, (fp0 + om) / n_step, fn0 / n_step, pos_tn) - metric(
                        tp0 / n_step, fp0 / n_step, (fn0 + e) / n_step, neg_tn)
                    key = torch.where(valid, sgn * gains, torch.full_like(gains, float("inf")))
                    kk = min(k, Lq)
                    selpos = torch.sort(key, dim=1, stable=True).indices[:, :kk]
                    keep = torch.arange(kk, device=dev_)[None, :] < torch.clamp(lens, max=k)[:, None]
                    new_in = torch.zeros_like(valid)
                    new_in.scatter_(1, selpos, keep)
                    newp = torch.where(keep, tix.gather(1, selpos), torch.full_like(selpos, m)).sort(dim=1).values
                    newp = torch.where(newp == m, torch.full_like(newp, -1), newp)
                    if kk < k:
                        newp = torch.cat([newp, torch.full((newp.shape[0], k - kk), -1, dtype=newp.dtype, device=dev_)], 1)
                    chg = new_in.double() - p_at_t.double()
                    d_tp = torch.zeros(m, dtype=torch.float64, device=dev_).index_add_(0, tix.reshape(-1), (chg * e).reshape(-1))
                    d_fp = torch.zeros(m, dtype=torch.float64, device=dev_).index_add_(0, tix.reshape(-1), (chg * om).reshape(-1))
                    unstored = (p >= 0) & ~eq.any(1)               # predicted labels the row does not store
                    d_fp.index_add_(0, p[unstored], -torch.ones(int(unstored.sum()), dtype=torch.float64, device=dev_))
                    tp += d_tp
                    fp += d_fp
                    fn -= d_tp
                    if not skip_tn:
                        tn -= d_fp
                    pred[rows] = newp
            cnt_host = None   # not maintained in batched mode
        state = state_fn(pred)   # :465-467
        new_u = _utility(metric, metric_aggregation, state, n_div)
        greedy = False
        meta["iters"] = j
        meta["utilities"].append(new_u)
        log_info(f"    Iteration {j}/{max_iters} finished, expected metric value: {old_u} -> {new_u}", verbose)
        if (maximize and new_u - old_u < tolerance) or (not maximize and new_u - old_u > tolerance):
            log_info(f"  Stopping because improvement of expected metric value is smaller than {tolerance}", verbose)
            break
    return pred


def _bca_generic_csr(y_proba: csr_matrix, binary_metric_func, k: int, metric_aggregation: str,
                     normalize_conf_matrix: bool, metric_kwargs, maximize: bool, tolerance: float, init_y_pred,
                     max_iters: int, shuffle_order: bool, skip_tn: bool, return_meta: bool, seed, verbose: bool,
                     mode: Optional[str], batch_size: Optional[int], y_pred_format: str, initial_pred_fn,
                     finish_pred_fn):
    if not callable(binary_metric_func):
        raise NotImplementedError(
            "xcolumns_b200: a LIST of metric callables needs dense inputs -- the reference's own CSR step indexes the "
            "list by the position of a stored entry, not by its label (block_coordinate.py:110-127)")
    if k <= 0:
        raise NotImplementedError("xcolumns_b200: arbitrary metric callables need a budget k > 0")
    meta: Dict[str, Any] = {"utilities": [], "iters": 0, "time": time()}
    device = dev.pick_device(y_proba)
    c = dev.csr_to_device(y_proba, device)
    n, m = c.n, c.m
    metric = _Metric(binary_metric_func, m, metric_kwargs, device)
    greedy = isinstance(init_y_pred, str) and init_y_pred == "greedy"
    pred = initial_pred_fn(y_proba, c, init_y_pred, k, seed, device).long()
    mode = mode or "auto"
    if mode == "auto":
        mode = "exact" if (n <= 4096 or greedy) else "batched"
    if greedy and mode != "exact":
        raise NotImplementedError("init_y_pred='greedy' needs the sequential mode (mode='exact')")
    meta["mode"] = mode
    ctx = dev.ctx_for(device)

    def state_fn(pr: torch.Tensor) -> torch.Tensor:
        """confusion sums of the compact prediction with one running float64 sum per label in row order
        (XC_SUM_ORDERED = 1: what numba's row loop does, numba_csr_functions.py:144-258)"""
        st = torch.zeros((4, m), dtype=torch.float64, device=device)
        p32 = pr.to(torch.int32).contiguous()
        ctx.call("xc_confmat_csr_compact", dev.ptr(c.data), dev.ptr(c.indices), dev.ptr(c.indptr), c.code, dev.ptr(p32),
                 k, n, m, 1, C.c_void_p(st[0].data_ptr()), C.c_void_p(st[1].data_ptr()), C.c_void_p(st[2].data_ptr()),
                 dev.stream_ptr(device))
        if skip_tn:
            st[3].fill_(-1.0)
        else:
            st[3] = -st[0] - st[1] - st[2] + n
        return st

    pred = bca_generic_csr_core(c.data, c.indices.long(), c.indptr.cpu().numpy().astype(np.int64), n, m, pred, metric, k,
                                metric_aggregation, normalize_conf_matrix, maximize, tolerance, greedy, max_iters,
                                shuffle_order, skip_tn, seed, verbose, mode, batch_size, meta, state_fn=state_fn)
    y_pred = finish_pred_fn(y_proba, pred.to(torch.int32), m, y_pred_format, None)
    if return_meta:
        meta["time"] = time() - meta["time"]
        return y_pred, meta
    return y_pred
