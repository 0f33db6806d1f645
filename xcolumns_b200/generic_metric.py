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
    """predict_using_bc_with_0approx for an arbitrary callable; dense inputs (numpy / torch)."""
    if isinstance(y_proba, csr_matrix):
        raise NotImplementedError(
            "xcolumns_b200: arbitrary metric callables are supported for dense inputs (numpy / torch); CSR inputs "
            "need one of the built-in binary metrics")
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
