"""predict_weighted_per_instance / predict_top_k on the GPU
(drop-in for xcolumns/weighted_prediction.py:91-220).

gains = a * eta + b per row, then the k largest (k > 0) or all gains >= th (k == 0).  The CUDA
kernel streams every row once with 128-bit loads and keeps a warp-level top-k list; ties go to
the lowest label id.  Results come back in the caller's container: numpy -> numpy, torch ->
torch on the same device, csr_matrix -> csr_matrix with k slots per row.
"""
from __future__ import annotations

import ctypes as C
from time import time
from typing import Optional, Tuple, Union

import numpy as np
import torch
from scipy.sparse import csr_matrix

from . import _device as dev
from ._lib import XC_F32, XC_F64
from .types import DenseMatrix, DType, Matrix

_MAX_K = 32


def _check_k(k, limit: Optional[int] = _MAX_K):
    """k is an int; the warp-list kernels (BCA, Frank-Wolfe, CSR top-k) hold k <= 32 labels, the dense weighted
    top-k takes any k <= m (block-level radix select beyond 32)."""
    if not isinstance(k, int) or isinstance(k, bool):
        raise ValueError("k must be an integer")
    if limit is not None and k > limit:
        raise NotImplementedError(f"xcolumns_b200 kernels support k <= {_MAX_K}, got k={k}")


def topk_dense_device(d: dev.DenseDev, k: int, a: Optional[torch.Tensor], b: Optional[torch.Tensor],
                      g_code: int, want_vals: bool = False, rows: Optional[torch.Tensor] = None):
    """Device-level entry: compact [n, k] label ids (+ gains) of a DenseDev."""
    device = d.t.device
    ctx = dev.ctx_for(device)
    n_rows = d.n if rows is None else int(rows.numel())
    out_idx = torch.empty((n_rows, k), dtype=torch.int32, device=device)
    out_val = torch.empty((n_rows, k), dtype=torch.float32 if g_code == XC_F32 else torch.float64,
                          device=device) if want_vals else None
    ctx.call("xc_topk_dense", dev.ptr(d.t), d.code, n_rows, d.m, d.ld, dev.ptr(rows), dev.ptr(a), dev.ptr(b),
             g_code, k, dev.ptr(out_idx), dev.ptr(out_val), dev.stream_ptr(device))
    return out_idx, out_val


def topk_csr_device(c: dev.CsrDev, k: int, a: Optional[torch.Tensor], b: Optional[torch.Tensor],
                    want_vals: bool = False):
    device = c.data.device
    ctx = dev.ctx_for(device)
    out_idx = torch.empty((c.n, k), dtype=torch.int32, device=device)
    out_val = torch.empty((c.n, k), dtype=c.data.dtype, device=device) if want_vals else None
    ctx.call("xc_topk_csr", dev.ptr(c.data), c.code, dev.ptr(c.indices), dev.ptr(c.indptr), c.n, dev.ptr(a),
             dev.ptr(b), k, dev.ptr(out_idx), dev.ptr(out_val), dev.stream_ptr(device))
    return out_idx, out_val


def _gain_dtype(y_dtype: torch.dtype, a, b) -> torch.dtype:
    """numpy/torch type promotion of eta * a + b (weighted_prediction.py:37-41)."""
    out = y_dtype
    for v in (a, b):
        if v is None:
            continue
        vd = v.dtype if isinstance(v, torch.Tensor) else torch.from_numpy(np.empty(0, dtype=np.asarray(v).dtype)).dtype
        if vd == torch.float64:
            out = torch.float64
    return out


def predict_weighted_per_instance(
    y_proba: Matrix,
    k: int,
    th: float = 0.0,
    a: Optional[DenseMatrix] = None,
    b: Optional[DenseMatrix] = None,
    dtype: Optional[DType] = None,
    keep_scores: bool = False,
    return_meta: bool = False,
    return_weights: bool = False,
) -> Union[Matrix, Tuple[Matrix, dict]]:
    """Weighted per-instance prediction: gains g = a * eta_i + b, top-k (k > 0) or g >= th (k = 0).

    Same arguments, validation and return contract as the reference
    (xcolumns/weighted_prediction.py:91-188)."""
    if not isinstance(y_proba, (np.ndarray, torch.Tensor, csr_matrix)):
        raise ValueError("y_proba must be either np.ndarray, torch.Tensor, or csr_matrix")
    if len(y_proba.shape) == 1:
        y_proba = y_proba.reshape(1, -1)
    elif len(y_proba.shape) > 2:
        raise ValueError("y_proba must be 1d or 2d")
    _check_k(k, _MAX_K if isinstance(y_proba, csr_matrix) else None)
    n, m = y_proba.shape
    for name, v in (("a", a), ("b", b)):
        if v is not None:
            if not isinstance(v, (np.ndarray, torch.Tensor)):
                raise ValueError(f"{name} must be np.ndarray or torch.Tensor")
            if tuple(v.shape) != (m,):
                raise ValueError(f"{name} must be of shape (y_proba[1],)")
    if return_meta:
        meta = {"iters": 1, "time": time()}

    device = dev.pick_device(y_proba)
    if isinstance(y_proba, csr_matrix):
        c = dev.csr_to_device(y_proba, device)
        # the reference casts a and b to the data dtype (weighted_prediction.py:72-75)
        ad = dev.vec_to_device(a, device, c.data.dtype, m, "a")
        bd = dev.vec_to_device(b, device, c.data.dtype, m, "b")
        if k > 0:
            idx, vals = topk_csr_device(c, k, ad, bd, want_vals=keep_scores)
            y_pred = dev.compact_to_csr_like(y_proba, idx, out_dtype=dtype, vals=vals)
        else:
            y_pred = _threshold_csr(y_proba, c, ad, bd, th, dtype)
    else:
        if k > m:
            raise ValueError(f"k={k} is larger than the number of labels m={m}")
        prefill = dev.DenseOutputPrefill.start(y_proba, n, m, dtype) if k > 0 else None
        d = dev.dense_to_device(y_proba, device)
        gdt = _gain_dtype(d.torch_dtype, a, b)
        g_code = XC_F32 if gdt == torch.float32 else XC_F64
        ad = dev.vec_to_device(a, device, gdt, m, "a")
        bd = dev.vec_to_device(b, device, gdt, m, "b")
        if k > 0:
            idx, vals = topk_dense_device(d, k, ad, bd, g_code, want_vals=keep_scores)
            y_pred = dev.compact_to_dense_like(y_proba, idx, m, out_dtype=dtype, vals=vals, prefill=prefill)
        else:
            y_pred = _threshold_dense(y_proba, d, ad, bd, g_code, th, dtype)

    if return_meta:
        meta["time"] = time() - meta["time"]
        if return_weights:
            meta["a"] = a
            meta["b"] = b
        return y_pred, meta
    return y_pred


def _threshold_dense(like, d: dev.DenseDev, ad, bd, g_code, th, dtype):
    device = d.t.device
    ctx = dev.ctx_for(device)
    if not np.isscalar(th) and getattr(th, "ndim", 0) > 0:
        # a vector of per-label thresholds (weighted_prediction.py:118): same gains (separate IEEE multiply and
        # add in the gain dtype), compared label-wise; k = 0 is not a hot path, elementwise torch on the device
        gdt = torch.float32 if g_code == XC_F32 else torch.float64
        thv = dev.vec_to_device(th, device, gdt, d.m, "th")
        g = d.t[:, :d.m].to(gdt)
        if ad is not None:
            g = g * ad
        if bd is not None:
            g = g + bd
        out = (g >= thv).to(d.torch_dtype)
        if isinstance(like, torch.Tensor):
            return out.to(device=like.device, dtype=like.dtype if dtype is None else dtype)
        res = out.cpu().numpy()
        return res if dtype is None else res.astype(dtype)
    out = torch.empty((d.n, d.m), dtype=d.torch_dtype, device=device)
    ctx.call("xc_threshold_dense", dev.ptr(d.t), d.code, d.n, d.m, d.ld, dev.ptr(ad), dev.ptr(bd), g_code,
             C.c_double(float(th)), dev.ptr(out), d.m, dev.stream_ptr(device))
    if isinstance(like, torch.Tensor):
        return out.to(device=like.device, dtype=like.dtype if dtype is None else dtype)
    res = out.cpu().numpy()
    return res if dtype is None else res.astype(dtype)


def _threshold_csr(like: csr_matrix, c: dev.CsrDev, ad, bd, th, dtype):
    """k == 0 on CSR rows (numba_csr_functions.py:516-517, :631-653): keep stored labels whose
    gain >= th.  Elementwise on the nnz entries, compaction with torch on the device."""
    idx = c.indices.long()
    g = c.data
    if ad is not None:
        g = g * ad[idx]
    if bd is not None:
        g = g + bd[idx]
    if not np.isscalar(th) and getattr(th, "ndim", 0) > 0:
        th = dev.vec_to_device(th, c.data.device, c.data.dtype, c.m, "th")[idx]
    keep = g >= th
    row_of = torch.repeat_interleave(torch.arange(c.n, device=c.data.device), c.indptr[1:] - c.indptr[:-1])
    counts = torch.zeros(c.n, dtype=torch.int64, device=c.data.device).index_add_(0, row_of[keep], torch.ones_like(row_of[keep]))
    indptr = torch.cat([torch.zeros(1, dtype=torch.int64, device=c.data.device), counts.cumsum(0)])
    new_idx = c.indices[keep].cpu().numpy().astype(like.indices.dtype)
    data = np.ones(new_idx.shape[0], dtype=like.data.dtype)
    return csr_matrix((data, new_idx, indptr.cpu().numpy().astype(like.indptr.dtype)), shape=like.shape, dtype=dtype)


def predict_top_k(
    y_proba: Matrix,
    k: int,
    dtype: Optional[DType] = None,
    keep_scores: bool = False,
    return_meta: bool = False,
) -> Union[Matrix, Tuple[Matrix, dict]]:
    """Top-k labels per row -- optimal for precision@k / nDCG@k
    (xcolumns/weighted_prediction.py:196-220)."""
    return predict_weighted_per_instance(y_proba, k=k, dtype=dtype, keep_scores=keep_scores, return_meta=return_meta)


# ------------------------------------------------------------------------------------------
# closed-form weighted strategies (xcolumns/weighted_prediction.py:223-560): per-label weights from
# priors / propensities, then the weighted top-k kernel
# ------------------------------------------------------------------------------------------

def _check_vec(v, m: int, name: str):
    if v.shape[0] != m:
        raise ValueError(f"{name} must be of shape (y_proba[1],)")


def predict_optimizing_macro_recall(y_proba: Matrix, k: int, priors: DenseMatrix, epsilon: float = 1e-6,
                                    keep_scores: bool = False, dtype: Optional[DType] = None, return_meta: bool = False,
                                    return_weights: bool = False):
    """a = 1 / (priors + epsilon) (xcolumns/weighted_prediction.py:223-264)."""
    _check_vec(priors, y_proba.shape[1], "priors")
    return predict_weighted_per_instance(y_proba, k=k, a=1.0 / (priors + epsilon), dtype=dtype, keep_scores=keep_scores,
                                         return_meta=return_meta, return_weights=return_weights)


def predict_optimizing_macro_balanced_accuracy(y_proba: Matrix, k: int, priors: DenseMatrix, epsilon: float = 1e-6,
                                               dtype: Optional[DType] = None, return_meta: bool = False):
    """Optimal strategy for macro balanced accuracy (xcolumns/weighted_prediction.py:267-368):
    gains = eta / pi - (1 - eta) / (1 - pi) with pi = priors + epsilon.  The gain is affine in eta,
    g = eta * (1/pi + 1/(1-pi)) - 1/(1-pi); it is evaluated in that form (one multiply-add per element
    instead of two divisions), which differs from the reference's expression by <= 1 ulp of the gain."""
    _check_vec(priors, y_proba.shape[1], "priors")
    if not isinstance(y_proba, (np.ndarray, torch.Tensor, csr_matrix)):
        raise ValueError("y_proba must be either np.ndarray, torch.Tensor, or csr_matrix")
    if return_meta:
        meta = {"iters": 1, "time": time()}
    pri = priors + epsilon
    a = 1.0 / pri + 1.0 / (1 - pri)
    b = -1.0 / (1 - pri)
    if isinstance(y_proba, csr_matrix) and y_proba.dtype != np.float64:
        # the reference forms the CSR gains in float64 (float32 data / float64 marginals,
        # numba_csr_functions.py:679); the CSR top-k kernel works in the data dtype, so widen the data
        y64 = csr_matrix((y_proba.data.astype(np.float64), y_proba.indices, y_proba.indptr), shape=y_proba.shape)
        y_pred = predict_weighted_per_instance(y64, k=k, th=0.0, a=np.asarray(a, dtype=np.float64),
                                               b=np.asarray(b, dtype=np.float64),
                                               dtype=y_proba.dtype if dtype is None else dtype)
    else:
        y_pred = predict_weighted_per_instance(y_proba, k=k, th=0.0, a=a, b=b, dtype=dtype)
    if return_meta:
        meta["time"] = time() - meta["time"]
        return y_pred, meta
    return y_pred


def predict_log_weighted_per_instance(y_proba: Matrix, k: int, priors: DenseMatrix, epsilon: float = 1e-6,
                                      keep_scores: bool = False, dtype: Optional[DType] = None,
                                      return_meta: bool = False, return_weights: bool = False):
    """a = -log(priors + epsilon) (xcolumns/weighted_prediction.py:371-416)."""
    _check_vec(priors, y_proba.shape[1], "priors")
    pri = priors + epsilon
    weights = -torch.log(pri) if isinstance(pri, torch.Tensor) else -np.log(pri)
    return predict_weighted_per_instance(y_proba, k=k, a=weights, keep_scores=keep_scores, dtype=dtype,
                                         return_meta=return_meta, return_weights=return_weights)


def predict_power_law_weighted_per_instance(y_proba: Matrix, k: int, priors: DenseMatrix, beta: float,
                                            epsilon: float = 1e-6, keep_scores: bool = False,
                                            dtype: Optional[DType] = None, return_meta: bool = False,
                                            return_weights: bool = False):
    """a = (priors + epsilon) ** -beta (xcolumns/weighted_prediction.py:419-465)."""
    _check_vec(priors, y_proba.shape[1], "priors")
    return predict_weighted_per_instance(y_proba, k=k, a=(priors + epsilon) ** -beta, keep_scores=keep_scores,
                                         dtype=dtype, return_meta=return_meta, return_weights=return_weights)


def predict_optimizing_instance_precision(y_proba: Matrix, k: int, keep_scores: bool = False,
                                          dtype: Optional[DType] = None, return_meta: bool = False):
    """Plain top-k: optimal for precision@k and nDCG@k (xcolumns/weighted_prediction.py:468-497)."""
    if k <= 0:
        raise ValueError("k must be > 0")
    return predict_top_k(y_proba, k=k, keep_scores=keep_scores, dtype=dtype, return_meta=return_meta)


def predict_optimizing_instance_propensity_scored_precision(y_proba: Matrix, k: int,
                                                            inverse_propensities: Optional[DenseMatrix] = None,
                                                            propensities: Optional[DenseMatrix] = None,
                                                            keep_scores: bool = False, dtype: Optional[DType] = None,
                                                            return_meta: bool = False, return_weights: bool = False):
    """a = inverse propensities (xcolumns/weighted_prediction.py:500-560; like there, zero entries of
    `propensities` are set to 1 IN PLACE before inverting)."""
    m = y_proba.shape[1]
    if inverse_propensities is not None:
        _check_vec(inverse_propensities, m, "inverse_propensities")
    elif propensities is not None:
        _check_vec(propensities, m, "propensities")
        propensities[propensities == 0] = 1.0
        inverse_propensities = 1.0 / propensities
    else:
        raise ValueError("either inverse_propensities or propensities must be provided")
    return predict_weighted_per_instance(y_proba, k=k, a=inverse_propensities, keep_scores=keep_scores, dtype=dtype,
                                         return_meta=return_meta, return_weights=return_weights)
