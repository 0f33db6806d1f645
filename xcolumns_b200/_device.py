"""Host <-> device adapters of the boundary: numpy / torch / scipy CSR in, device buffers for the
C ABI, and results handed back in the caller's container type and dtype
(SURVEY.md section 8b "Return-type rule").  torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
from scipy.sparse import csr_matrix

from . import _lib
from ._lib import XC_F32, XC_F64, XColumnsB200Error

_NP2CODE = {np.dtype(np.float32): XC_F32, np.dtype(np.float64): XC_F64}
_T2CODE = {torch.float32: XC_F32, torch.float64: XC_F64}
_CODE2T = {XC_F32: torch.float32, XC_F64: torch.float64}
_T2NP = {torch.float32: np.float32, torch.float64: np.float64}
_STAGED_MIN_BYTES = 64 << 20


def host_threads() -> int:
    """Host threads one process may use for the host-side helpers (clearing / filling the dense result, staging
    copies): all cores of a single process, an equal share when several ranks of one box (torchrun sets
    LOCAL_WORLD_SIZE) run the same helpers at the same time -- eight ranks with 32 threads each only fight over
    the memory bus of the one host."""
    cores = os.cpu_count() or 1
    try:
        local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    except ValueError:
        local_world = 1
    return max(2, min(32, cores // local_world))


def pick_device(*objs) -> torch.device:
    """Device of the first CUDA tensor among objs, else cuda:LOCAL_RANK / current device."""
    for o in objs:
        if isinstance(o, torch.Tensor) and o.is_cuda:
            return o.device
    if not torch.cuda.is_available():
        raise XColumnsB200Error(
            "xcolumns_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")
    if "XCOLUMNS_B200_DEVICE" in os.environ:
        return torch.device(os.environ["XCOLUMNS_B200_DEVICE"])
    return torch.device("cuda", torch.cuda.current_device())


def ctx_for(device: torch.device) -> _lib.Context:
    return _lib.context(device.index if device.index is not None else torch.cuda.current_device())


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device: torch.device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def float_code(dtype) -> int:
    if isinstance(dtype, torch.dtype):
        if dtype in _T2CODE:
            return _T2CODE[dtype]
    else:
        d = np.dtype(dtype)
        if d in _NP2CODE:
            return _NP2CODE[d]
    raise ValueError(f"xcolumns_b200 supports float32 / float64 probability matrices, got {dtype}")


@dataclass
class DenseDev:
    """Row-major [n, ld] device matrix, first m columns valid."""
    t: torch.Tensor          # 2-D tensor of shape [n, ld]
    n: int
    m: int
    ld: int
    code: int                # XC_F32 / XC_F64
    h2d_bytes: int = 0

    @property
    def torch_dtype(self):
        return _CODE2T[self.code]


@dataclass
class CsrDev:
    data: torch.Tensor       # float32 / float64 [nnz]
    indices: torch.Tensor    # int32 [nnz], ascending inside a row
    indptr: torch.Tensor     # int64 [n + 1]
    n: int
    m: int
    code: int
    h2d_bytes: int = 0


def _as_float_numpy(a: np.ndarray, want=None) -> np.ndarray:
    if want is not None:
        return np.ascontiguousarray(a, dtype=want)
    if a.dtype in _NP2CODE:
        return np.ascontiguousarray(a)
    if a.dtype == np.float16:
        return np.ascontiguousarray(a, dtype=np.float32)
    raise ValueError(f"unsupported dtype {a.dtype}: pass float32 or float64")


def dense_to_device(x, device: torch.device, dtype=None, pad: bool = True) -> DenseDev:
    """numpy / torch (any device) -> DenseDev.  Host inputs are uploaded with the leading
    dimension padded to a multiple of 16 bytes so rows can be read with 128-bit loads; CUDA
    tensors are used in place (zero copy) when contiguous."""
    if isinstance(x, torch.Tensor):
        if x.dim() != 2:
            raise ValueError("expected a 2-D matrix")
        tdt = dtype if dtype is not None else x.dtype
        if tdt not in _T2CODE:
            raise ValueError(f"unsupported dtype {x.dtype}: pass float32 or float64")
        n, m = x.shape
        if x.is_cuda:
            t = x.to(device=device, dtype=tdt)
            if t.stride(1) != 1 or (n > 1 and t.stride(0) < m):
                t = t.contiguous()
            ld = t.stride(0) if n > 1 else max(m, t.stride(0))
            return DenseDev(t, n, m, int(ld), _T2CODE[tdt], 0)
        x = x.detach().to(dtype=tdt)
        src = x.contiguous()
    else:
        a = np.asarray(x)
        if a.ndim != 2:
            raise ValueError("expected a 2-D matrix")
        a = _as_float_numpy(a, None if dtype is None else _T2NP[dtype] if isinstance(dtype, torch.dtype) else dtype)
        src = torch.from_numpy(a)
    n, m = src.shape
    code = _T2CODE[src.dtype]
    vec = 4 if code == XC_F32 else 2
    ld = ((m + vec - 1) // vec) * vec if pad else m
    nbytes = n * m * src.element_size()
    if nbytes >= _STAGED_MIN_BYTES and not src.is_pinned():
        # an ordinary (pageable) numpy array -- what a caller of the reference passes: threaded copies into two
        # pinned staging buffers overlapped with the DMA (csrc/host.cu: xc_h2d_staged)
        t = torch.empty((n, ld), dtype=src.dtype, device=device)
        if ld != m:
            t[:, m:].zero_()
        es = src.element_size()
        ctx_for(device).call("xc_h2d_staged", C.c_void_p(t.data_ptr()), ld * es, C.c_void_p(src.data_ptr()), m * es,
                             m * es, n, min(16, host_threads()), stream_ptr(device))
    elif ld == m:
        t = src.to(device, non_blocking=True)
    else:
        t = torch.zeros((n, ld), dtype=src.dtype, device=device)
        t[:, :m].copy_(src, non_blocking=True)
    return DenseDev(t, n, m, ld, code, nbytes)


def csr_to_device(x: csr_matrix, device: torch.device, dtype=None) -> CsrDev:
    if not isinstance(x, csr_matrix):
        raise ValueError("expected scipy.sparse.csr_matrix")
    if not x.has_sorted_indices:
        x = x.sorted_indices()  # the reference requires sorted indices (numba_csr_functions.py:121)
    n, m = x.shape
    data = _as_float_numpy(x.data, dtype)
    indices = np.ascontiguousarray(x.indices, dtype=np.int32)
    indptr = np.ascontiguousarray(x.indptr, dtype=np.int64)
    nbytes = data.nbytes + indices.nbytes + indptr.nbytes
    return CsrDev(torch.from_numpy(data).to(device, non_blocking=True),
                  torch.from_numpy(indices).to(device, non_blocking=True),
                  torch.from_numpy(indptr).to(device, non_blocking=True),
                  n, m, _NP2CODE[data.dtype], nbytes)


def vec_to_device(v, device: torch.device, torch_dtype, m: int, name: str) -> Optional[torch.Tensor]:
    if v is None:
        return None
    if isinstance(v, torch.Tensor):
        t = v.detach().to(device=device, dtype=torch_dtype).contiguous()
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(v), dtype=_T2NP[torch_dtype])).to(device)
    if t.shape != (m,):
        raise ValueError(f"{name} must be of shape (y_proba[1],)")
    return t


# ------------------------------------------------------------------------------------------
# results back to the caller's container
# ------------------------------------------------------------------------------------------

class DenseOutputPrefill:
    """The dense host result of a call with host inputs, cleared on background host threads while the
    scores are uploaded and the GPU works (csrc/host.cu: xc_zero_host); at the end only the n*k
    selected entries are scattered.  At AmazonCat-13K shape the 16 GB zero-fill is as long as the
    upload itself (0.3 s each, measured), so overlapping them nearly halves the end-to-end time."""

    MIN_BYTES = 64 << 20

    def __init__(self, n: int, m: int, np_dtype):
        self.out = np.empty((n, m), dtype=np_dtype)
        lib = _lib.load()

        def clear(arr=self.out):   # the closure keeps the array alive even if the caller bails out early
            lib.xc_zero_host(arr.ctypes.data, arr.nbytes, host_threads())   # ctypes drops the GIL for the duration of the call

        self._thread = threading.Thread(target=clear, daemon=True)
        self._thread.start()

    @staticmethod
    def start(like, n: int, m: int, out_dtype=None) -> Optional["DenseOutputPrefill"]:
        """Prefill for numpy / CPU-tensor inputs whose dense result is large; None otherwise."""
        if isinstance(like, torch.Tensor):
            if like.is_cuda:
                return None
            dt = _T2NP.get(like.dtype if out_dtype is None else out_dtype)
        elif isinstance(like, np.ndarray):
            dt = np.dtype(like.dtype if out_dtype is None else out_dtype)
            dt = dt.type if dt in _NP2CODE else None
        else:
            return None
        if dt is None or n * m * np.dtype(dt).itemsize < DenseOutputPrefill.MIN_BYTES:
            return None
        return DenseOutputPrefill(n, m, dt)

    def finish(self, idx: np.ndarray, vals: Optional[np.ndarray]) -> np.ndarray:
        self._thread.join()
        _scatter_host(self.out, idx, vals, zero=False)
        return self.out


def compact_to_dense_like(like, pred_idx: torch.Tensor, m: int, out_dtype=None, vals: Optional[torch.Tensor] = None,
                          prefill: Optional[DenseOutputPrefill] = None):
    """[n, k] label ids (device) -> dense 0/1 (or gain-valued) matrix of `like`'s container type."""
    n, k = pred_idx.shape
    if prefill is not None:
        out = prefill.finish(pred_idx.cpu().numpy(), None if vals is None else vals.cpu().numpy())
        if isinstance(like, torch.Tensor):
            return torch.from_numpy(out).to(like.dtype if out_dtype is None else out_dtype)
        return out
    if isinstance(like, torch.Tensor):
        dt = like.dtype if out_dtype is None else out_dtype
        if like.is_cuda:
            out = torch.zeros((n, m), dtype=dt if dt in _T2CODE else torch.float32, device=like.device)
            ctx = ctx_for(like.device)
            ctx.call("xc_scatter_pred_dense", ptr(pred_idx), ptr(vals), 0 if vals is None else _T2CODE[vals.dtype],
                     k, n, ptr(out), _T2CODE[out.dtype], m, stream_ptr(like.device))
            return out if out.dtype == dt else out.to(dt)
        idx = pred_idx.cpu().numpy()
        v = None if vals is None else vals.cpu().numpy()
        out = np.empty((n, m), dtype=_T2NP.get(dt, np.float32))
        _scatter_host(out, idx, v)
        return torch.from_numpy(out).to(dt)
    dt = np.dtype(like.dtype if out_dtype is None else out_dtype)
    idx = pred_idx.cpu().numpy()
    v = None if vals is None else vals.cpu().numpy()
    out = np.empty((n, m), dtype=dt)
    _scatter_host(out, idx, v)
    return out


def _scatter_host(out: np.ndarray, idx: np.ndarray, vals: Optional[np.ndarray], zero: bool = True):
    """zero-fill (unless already done) + scatter on host threads (csrc/host.cu)"""
    n, k = idx.shape
    lib = _lib.load()
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    v = None
    if vals is not None:
        v = np.ascontiguousarray(vals, dtype=np.float64 if vals.dtype == np.float64 else np.float32)
    if out.dtype in _NP2CODE and out.flags.c_contiguous:
        fn = lib.xc_fill_pred_dense_host if zero else lib.xc_scatter_pred_dense_host
        rc = fn(out.ctypes.data, _NP2CODE[out.dtype], n, out.shape[1], out.shape[1], idx.ctypes.data,
                None if v is None else v.ctypes.data, 0 if v is None else _NP2CODE[v.dtype], k, host_threads())
        if rc != 0:
            raise XColumnsB200Error(f"xc_fill_pred_dense_host failed ({rc})")
        return
    if zero:
        out[...] = 0
    rows = np.repeat(np.arange(n), k)
    flat = idx.reshape(-1)
    ok = flat >= 0
    out[rows[ok], flat[ok]] = 1 if vals is None else vals.reshape(-1)[ok]


def compact_to_csr_like(like: csr_matrix, pred_idx: torch.Tensor, out_dtype=None, vals: Optional[torch.Tensor] = None,
                        reference_padding: bool = True) -> csr_matrix:
    """[n, k] label ids -> csr_matrix with k slots per row (indptr = arange(n+1)*k), like
    numba_csr_functions.py:598-629: unused slots become the reference's (index 0, value 1) filler
    when reference_padding is set, else rows are compacted."""
    n, k = pred_idx.shape
    idx = pred_idx.cpu().numpy()
    ddt = np.dtype(like.dtype if out_dtype is None else out_dtype)
    data = np.ones((n, k), dtype=like.data.dtype) if vals is None else vals.cpu().numpy().astype(like.data.dtype)
    if reference_padding:
        pad = idx < 0
        idx = np.where(pad, 0, idx)
        data = np.where(pad, 1, data).astype(data.dtype)
        indptr = np.arange(n + 1, dtype=like.indptr.dtype) * k
        return csr_matrix((data.reshape(-1), idx.reshape(-1).astype(like.indices.dtype), indptr),
                          shape=like.shape, dtype=ddt)
    ok = idx >= 0
    indptr = np.concatenate([[0], np.cumsum(ok.sum(1))]).astype(like.indptr.dtype)
    return csr_matrix((data[ok], idx[ok].astype(like.indices.dtype), indptr), shape=like.shape, dtype=ddt)


def dense_pred_to_compact(y_pred, k: int, device: torch.device) -> torch.Tensor:
    """0/1 prediction matrix (numpy / torch / csr) with exactly k ones per row -> [n, k] int32."""
    if isinstance(y_pred, csr_matrix):
        y = y_pred if y_pred.has_sorted_indices else y_pred.sorted_indices()
        n = y.shape[0]
        cnt = np.diff(y.indptr)
        if (cnt > k).any():
            raise ValueError("init_y_pred has rows with more than k predicted labels")
        idx = np.full((n, k), -1, dtype=np.int32)
        cols = np.arange(y.nnz) - np.repeat(y.indptr[:-1], cnt)
        idx[np.repeat(np.arange(n), cnt), cols] = y.indices
        return torch.from_numpy(idx).to(device)
    if isinstance(y_pred, torch.Tensor):
        nz = (y_pred != 0)
        if not bool((nz.sum(1) == k).all()):
            raise ValueError("init_y_pred must have exactly k predicted labels per row")
        cols = nz.nonzero()[:, 1].reshape(-1, k)
        return cols.to(device=device, dtype=torch.int32)
    a = np.asarray(y_pred)
    r, c = np.nonzero(a)
    if not (np.bincount(r, minlength=a.shape[0]) == k).all():
        raise ValueError("init_y_pred must have exactly k predicted labels per row")
    return torch.from_numpy(c.reshape(-1, k).astype(np.int32)).to(device)
