"""Label-wise confusion matrix on the GPU (drop-in for xcolumns/confusion_matrix.py:16-399).

tp = sum y*yhat, fp = sum (1-y)*yhat, fn = sum y*(1-yhat) along an axis, in ONE fused pass over
both matrices (the reference makes three passes with n x m temporaries); tn is derived.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch
from scipy.sparse import csr_matrix

from . import _device as dev
from ._lib import XC_SUM_FAST, XC_SUM_ORDERED
from .types import DType, Matrix


class ConfusionMatrix:
    """Per-label (or per-instance) counts or rates of true positives, false positives, false
    negatives and true negatives.  Unpacks as ``tp, fp, fn, tn`` so it can be splatted into the
    metric functions; supports element-wise + - * / // between matrices / with scalars."""

    __slots__ = ("tp", "fp", "fn", "tn")

    def __init__(self, tp, fp, fn, tn):
        self.tp, self.fp, self.fn, self.tn = tp, fp, fn, tn

    def __iter__(self):
        return iter((self.tp, self.fp, self.fn, self.tn))

    def _zip(self, other, op):
        if isinstance(other, ConfusionMatrix):
            return ConfusionMatrix(*(op(x, y) for x, y in zip(self, other)))
        return ConfusionMatrix(*(op(x, other) for x in self))

    def _izip(self, other, op):
        self.tp, self.fp, self.fn, self.tn = tuple(self._zip(other, op))
        return self

    def __eq__(self, other):
        if not isinstance(other, ConfusionMatrix):
            return False
        return all(bool(np.all(np.asarray(x == y))) if not isinstance(x, torch.Tensor) else bool((x == y).all())
                   for x, y in zip(self, other))

    def __add__(self, o): return self._zip(o, lambda x, y: x + y)
    def __sub__(self, o): return self._zip(o, lambda x, y: x - y)
    def __mul__(self, o): return self._zip(o, lambda x, y: x * y)
    def __truediv__(self, o): return self._zip(o, lambda x, y: x / y)
    def __floordiv__(self, o): return self._zip(o, lambda x, y: x // y)
    def __iadd__(self, o): return self._izip(o, lambda x, y: x + y)
    def __isub__(self, o): return self._izip(o, lambda x, y: x - y)
    def __imul__(self, o): return self._izip(o, lambda x, y: x * y)
    def __itruediv__(self, o): return self._izip(o, lambda x, y: x / y)
    def __ifloordiv__(self, o): return self._izip(o, lambda x, y: x // y)

    def normalize(self) -> "ConfusionMatrix":
        """Rates instead of counts: every entry divided by tp + fp + fn + tn."""
        total = self.tp + self.fp + self.fn + self.tn
        return self / total

    def __repr__(self):
        return f"ConfusionMatrix(tp={self.tp!r}, fp={self.fp!r}, fn={self.fn!r}, tn={self.tn!r})"


def _sum_order(order: Optional[str]) -> int:
    order = order or os.environ.get("XCOLUMNS_B200_SUM_ORDER", "fast")
    if order not in ("fast", "ordered"):
        raise ValueError("order must be 'fast' or 'ordered'")
    return XC_SUM_ORDERED if order == "ordered" else XC_SUM_FAST


def _numpy_dtype_of(x):
    return x.dtype if not isinstance(x, torch.Tensor) else np.dtype(str(x.dtype).replace("torch.", ""))


def confusion_sums_device(y_true, y_pred, axis: int, order: int, acc_f32: bool, device):
    """float64 device vectors (tp, fp, fn) of two same-kind matrices."""
    ctx = dev.ctx_for(device)
    n, m = y_true.shape
    length = m if axis == 0 else n
    tp, fp, fn = (torch.empty(length, dtype=torch.float64, device=device) for _ in range(3))
    if isinstance(y_true, csr_matrix):
        if axis != 0:
            raise NotImplementedError("xcolumns_b200: CSR confusion matrix supports axis=0")
        cdt = np.result_type(y_true.dtype, y_pred.dtype, np.float32)
        cdt = np.float32 if cdt == np.float32 else np.float64
        t = dev.csr_to_device(y_true, device, cdt)
        p = dev.csr_to_device(y_pred, device, cdt)
        ctx.call("xc_confmat_csr", dev.ptr(t.data), dev.ptr(t.indices), dev.ptr(t.indptr), dev.ptr(p.data),
                 dev.ptr(p.indices), dev.ptr(p.indptr), t.code, n, m, order, int(acc_f32), dev.ptr(tp), dev.ptr(fp),
                 dev.ptr(fn), dev.stream_ptr(device))
    else:
        tdt = np.result_type(_numpy_dtype_of(y_true), _numpy_dtype_of(y_pred), np.float32)
        tdt = torch.float32 if tdt == np.float32 else torch.float64
        t = dev.dense_to_device(_to_float(y_true, tdt), device, tdt, pad=False)
        p = dev.dense_to_device(_to_float(y_pred, tdt), device, tdt, pad=False)
        ctx.call("xc_confmat_dense", dev.ptr(t.t), t.ld, dev.ptr(p.t), p.ld, t.code, n, m, axis, order, int(acc_f32),
                 dev.ptr(tp), dev.ptr(fp), dev.ptr(fn), dev.stream_ptr(device))
    return tp, fp, fn


def _to_float(x, tdt):
    if isinstance(x, torch.Tensor):
        return x if x.dtype == tdt else x.to(tdt)
    a = np.asarray(x)
    want = np.float32 if tdt == torch.float32 else np.float64
    return a if a.dtype == want else a.astype(want)


def calculate_confusion_matrix(
    y_true: Matrix,
    y_pred: Matrix,
    normalize: bool = False,
    skip_tn: bool = False,
    axis: Optional[int] = 0,
    dtype: Optional[DType] = None,
    order: Optional[str] = None,
) -> ConfusionMatrix:
    """Confusion matrix of true vs predicted labels along an axis
    (xcolumns/confusion_matrix.py:364-399).

    ``order`` (extension, default "fast" or $XCOLUMNS_B200_SUM_ORDER): "ordered" reproduces the
    reference's per-label summation order bit for bit, "fast" splits the rows over the grid."""
    dense = isinstance(y_true, (np.ndarray, torch.Tensor)) and isinstance(y_pred, (np.ndarray, torch.Tensor))
    sparse = isinstance(y_true, csr_matrix) and isinstance(y_pred, csr_matrix)
    if not (dense or sparse):
        raise ValueError("y_true and y_pred must be both np.ndarray, both torch.Tensor, or csr_matrix")
    if tuple(y_true.shape) != tuple(y_pred.shape):
        raise ValueError("y_true and y_pred must have the same shape")
    if axis not in (0, 1):
        raise ValueError("axis must be 0 or 1")
    n, m = y_true.shape
    device = dev.pick_device(y_true, y_pred)

    # dtype of the reference's result: np.sum(..., dtype=dtype) -> dtype, or the product's dtype
    if dtype is not None:
        out_dt = dtype
    elif sparse:
        out_dt = y_true.dtype
    elif isinstance(y_true, torch.Tensor):
        out_dt = torch.promote_types(y_true.dtype, y_pred.dtype if isinstance(y_pred, torch.Tensor) else y_true.dtype)
    else:
        out_dt = np.result_type(y_true.dtype, _numpy_dtype_of(y_pred))
    sum_order = _sum_order(order)
    f32_out = out_dt in (np.float32, torch.float32) or (not isinstance(out_dt, torch.dtype) and np.dtype(out_dt) == np.float32)
    tp, fp, fn = confusion_sums_device(y_true, y_pred, axis, sum_order, f32_out and sum_order == XC_SUM_ORDERED, device)

    as_torch = isinstance(y_true, torch.Tensor)
    if as_torch:
        tdt = out_dt if isinstance(out_dt, torch.dtype) else torch.from_numpy(np.empty(0, dtype=out_dt)).dtype
        tp, fp, fn = (v.to(device=y_true.device, dtype=tdt) for v in (tp, fp, fn))
    else:
        ndt = np.dtype(str(out_dt).replace("torch.", "")) if isinstance(out_dt, torch.dtype) else np.dtype(out_dt)
        host = torch.stack([tp, fp, fn]).cpu().numpy()
        tp, fp, fn = (host[i].astype(ndt) for i in range(3))
    if normalize:
        tp, fp, fn = tp / n, fp / n, fn / n
    if skip_tn:
        tn = tp.clone() if as_torch else tp.copy()
        tn[:] = -1
    else:
        tn = -tp - fp - fn + (1.0 if normalize else (n if axis == 0 else m))
    return ConfusionMatrix(tp, fp, fn, tn)


def calculate_tp(y_true, y_pred, normalize=False, axis=0, dtype=None):
    return calculate_confusion_matrix(y_true, y_pred, normalize=normalize, skip_tn=True, axis=axis, dtype=dtype).tp


def calculate_fp(y_true, y_pred, normalize=False, axis=0, dtype=None):
    return calculate_confusion_matrix(y_true, y_pred, normalize=normalize, skip_tn=True, axis=axis, dtype=dtype).fp


def calculate_fn(y_true, y_pred, normalize=False, axis=0, dtype=None):
    return calculate_confusion_matrix(y_true, y_pred, normalize=normalize, skip_tn=True, axis=axis, dtype=dtype).fn
