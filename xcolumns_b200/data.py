"""On-disk formats -> the matrices the hot path consumes (SURVEY.md section 8f rank 4;
reference: experiments/utils.py:112-228).

Same file formats and the same resulting ``csr_matrix`` (float32 data, sorted indices) as the
reference's loaders; the text files are parsed with whole-file numpy conversions instead of
per-token Python loops, and the dense -> top-k sparsifier (``load_npy_full_pred``) runs on the GPU
through the weighted top-k kernel (csrc/topk.cu) instead of two n x m ``np.partition`` passes.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Union

import numpy as np
import torch
from scipy.sparse import csr_matrix, load_npz, save_npz

from . import _device as dev
from ._lib import XC_F32, XC_F64
from .weighted_prediction import topk_dense_device


def _csr(data, indices, indptr, shape=None, sort: bool = False) -> csr_matrix:
    mat = csr_matrix((np.asarray(data, dtype=np.float32), np.asarray(indices, dtype=np.int32),
                      np.asarray(indptr, dtype=np.int32)), shape=shape, dtype=np.float32)
    if sort:
        mat.sort_indices()
    return mat


def _needs_sort(indices: np.ndarray, indptr: np.ndarray) -> bool:
    """True if some row's indices are not ascending (reference: `requires_sort`)."""
    if indices.size < 2:
        return False
    desc = indices[1:] < indices[:-1]
    row_start = np.zeros(indices.size, dtype=bool)
    starts = indptr[1:-1]
    row_start[starts[starts < indices.size]] = True
    return bool((desc & ~row_start[1:]).any())


def load_txt_labels(path: str, header: bool = True, labels_delimiter: str = ",",
                    labels_features_delimiter: Optional[str] = " ", labels_map: Optional[dict] = None) -> csr_matrix:
    """Sparse 0/1 label matrix from the XMC-repository text format: an optional ``n_ins n_ftr n_lbl``
    header, then per line ``l1,l2,... f:v f:v`` (experiments/utils.py:112-160)."""
    with open(path) as f:
        lines = f.read().split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    if header:
        lines = lines[1:]
    if labels_features_delimiter is not None:
        lines = [ln.split(labels_features_delimiter, 1)[0] for ln in lines]
    counts = np.fromiter((0 if ln == "" else ln.count(labels_delimiter) + 1 for ln in lines), dtype=np.int64,
                         count=len(lines))
    joined = labels_delimiter.join(ln for ln in lines if ln != "")
    if labels_map is not None:
        indices = np.array([labels_map[t.strip()] for t in joined.split(labels_delimiter)] if joined else [], dtype=np.int64)
    else:
        indices = np.array(joined.split(labels_delimiter), dtype=np.int64) if joined else np.zeros(0, dtype=np.int64)
    indptr = np.concatenate([[0], np.cumsum(counts)])
    return _csr(np.ones(indices.size, dtype=np.float32), indices, indptr, sort=_needs_sort(indices, indptr))


def load_txt_sparse_pred(path: str) -> csr_matrix:
    """Sparse score matrix from the libsvm-like format ``label:value label:value ...`` per line
    (experiments/utils.py:163-187)."""
    with open(path) as f:
        text = f.read()
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    counts = np.fromiter((ln.count(":") for ln in lines), dtype=np.int64, count=len(lines))
    flat = np.array(text.replace(":", " ").split(), dtype=np.float64)
    indices = flat[0::2].astype(np.int64)
    data = flat[1::2].astype(np.float32)
    indptr = np.concatenate([[0], np.cumsum(counts)])
    return _csr(data, indices, indptr, sort=_needs_sort(indices, indptr))


def load_npy_sparse_pred(path: str) -> csr_matrix:
    """LightXML-style pair ``<path>-labels.npy`` / ``<path>-scores.npy`` of shape [n, top]
    (experiments/utils.py:190-196)."""
    indices = np.load(path + "-labels.npy", allow_pickle=True)
    data = np.load(path + "-scores.npy", allow_pickle=True)
    indptr = np.arange(0, indices.shape[0] + 1, 1, dtype=np.int32) * indices.shape[1]
    return _csr(np.asarray(data).flatten(), np.asarray(indices).flatten(), indptr, sort=True)


def sparsify_top_k(dense: Union[np.ndarray, torch.Tensor], keep_top_k: int) -> csr_matrix:
    """The ``keep_top_k`` largest scores of every row as a CSR matrix with sorted indices, selected on the
    GPU (ties -> lowest label id).  Values keep the input's precision, the result is float32 like the
    reference's."""
    if not isinstance(keep_top_k, int) or keep_top_k <= 0:
        raise ValueError("keep_top_k must be a positive integer")
    device = dev.pick_device(dense)
    d = dev.dense_to_device(dense, device)
    if keep_top_k > d.m:
        raise ValueError(f"keep_top_k={keep_top_k} is larger than the number of columns {d.m}")
    idx, vals = topk_dense_device(d, keep_top_k, None, None, XC_F32 if d.code == XC_F32 else XC_F64, want_vals=True)
    n = d.n
    indptr = np.arange(0, n + 1, dtype=np.int32) * keep_top_k
    return _csr(vals.cpu().numpy().reshape(-1), idx.cpu().numpy().reshape(-1), indptr, shape=(n, d.m))


def load_npy_full_pred(path: str, keep_top_k: int = 0, **kwargs) -> csr_matrix:
    """Dense ``[n, m]`` score matrix from a .npy file, sparsified to its top ``keep_top_k`` per row
    (experiments/utils.py:199-211; there: two full np.partition passes on the host)."""
    return sparsify_top_k(np.load(path, allow_pickle=True), keep_top_k)


def load_cache_npz_file(path: str, load_func: Callable, recreate: bool = False, **load_func_args):
    """``<path>.npz`` if it exists, otherwise ``load_func(path, ...)`` cached there
    (experiments/utils.py:214-228)."""
    if not os.path.exists(path + ".npz") or recreate:
        data = load_func(path, **load_func_args)
        save_npz(path + ".npz", data)
    else:
        data = load_npz(path + ".npz")
    return data
