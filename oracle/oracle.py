"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Python drivers around oracle/xc_oracle.c: a CPU restatement of the reference's
prediction-optimisation path (mwydmuch/xCOLUMNs 0.0.3, citations relative to
/root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module; the product package never does.

Parity status: PINNED against outputs of the live reference (tests/golden/make_golden.py
-> tests/golden/*.npz, replayed by tests/test_oracle_golden.py).

Metrics are named by strings: "precision", "recall", "fbeta", "f1", "jaccard",
"balanced_accuracy", "gmean", "hmean" (+ ``beta``/``epsilon`` keyword arguments).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time
from typing import Optional

import numpy as np
from scipy.sparse import csr_matrix

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libxc_oracle.so")

METRIC_IDS = {"precision": 0, "recall": 1, "fbeta": 2, "f1": 2, "jaccard": 3,
              "balanced_accuracy": 4, "gmean": 5, "hmean": 6, "precision_at_k": 7}
USES_TN = {4, 5, 6}


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc via oracle/Makefile)."""
    src = [os.path.join(_HERE, f) for f in ("xc_oracle.c", "xc_oracle_impl.h")]
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "_build/libxc_oracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def _p(a, t=C.c_void_p):
    return None if a is None else a.ctypes.data_as(t)


def _sfx(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise ValueError(f"oracle supports float32/float64 data, got {dtype}")


def metric_params(metric: str, beta: float = 1.0, epsilon: float = 1e-9):
    mid = METRIC_IDS[metric]
    if metric == "f1":
        beta = 1.0
    if metric == "precision_at_k":      # tp / k: k travels in the c1 slot (pass beta=k)
        return mid, float(beta), 0.0, float(epsilon)
    return mid, float(1 + beta ** 2), float(beta ** 2), float(epsilon)


class _Mix:
    """Scope of the mixed utility (1 - alpha) * tp / k + alpha * metric / m
    (block_coordinate.py:848-1045): process-wide switch of the C restatement."""

    def __init__(self, mix):
        self.mix = mix

    def __enter__(self):
        if self.mix is not None:
            a, k, m = self.mix
            lib().orc_set_mix(C.c_int(1), C.c_double(a), C.c_double(k), C.c_double(m))

    def __exit__(self, *exc):
        if self.mix is not None:
            lib().orc_set_mix(C.c_int(0), C.c_double(1.0), C.c_double(1.0), C.c_double(1.0))
        return False


# ------------------------------------------------------------------------------------------
# weighted top-k  (weighted_prediction.py:25-88, 91-220)
# ------------------------------------------------------------------------------------------

def topk_indices_dense(eta: np.ndarray, k: int, a=None, b=None):
    """Compact result: (n, k) int32 label ids ascending + gains of those labels."""
    eta = np.ascontiguousarray(eta)
    n, m = eta.shape
    idx = np.empty((n, k), dtype=np.int32)
    gdt = np.result_type(eta.dtype, *(x.dtype for x in (a, b) if x is not None))
    if gdt == eta.dtype:
        a_ = None if a is None else np.ascontiguousarray(a, dtype=eta.dtype)
        b_ = None if b is None else np.ascontiguousarray(b, dtype=eta.dtype)
        val = np.empty((n, k), dtype=eta.dtype)
        rc = getattr(lib(), "orc_topk_dense_" + _sfx(eta.dtype))(
            _p(eta), C.c_int64(n), C.c_int64(m), C.c_int64(m), _p(a_), _p(b_), C.c_int(k),
            _p(idx), _p(val))
    else:  # float32 eta with float64 weights -> float64 gains
        a_ = None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        b_ = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
        val = np.empty((n, k), dtype=np.float64)
        rc = getattr(lib(), "orc_topk_dense_wd_" + _sfx(eta.dtype))(
            _p(eta), C.c_int64(n), C.c_int64(m), C.c_int64(m), _p(a_), _p(b_), C.c_int(k),
            _p(idx), _p(val))
    if rc:
        raise RuntimeError(f"orc_topk_dense failed: {rc}")
    return idx, val


def topk_indices_csr(y: csr_matrix, k: int, a=None, b=None, keep_scores=False):
    n, m = y.shape
    dt = y.data.dtype
    a_ = None if a is None else np.ascontiguousarray(a, dtype=dt)
    b_ = None if b is None else np.ascontiguousarray(b, dtype=dt)
    idx = np.empty((n, k), dtype=np.int32)
    val = np.empty((n, k), dtype=dt)
    indptr = y.indptr.astype(np.int64)
    indices = np.ascontiguousarray(y.indices, dtype=np.int32)
    rc = getattr(lib(), "orc_topk_csr_" + _sfx(dt))(
        _p(y.data), _p(indices), _p(indptr), C.c_int64(n), _p(a_), _p(b_), C.c_int(k),
        C.c_int(int(keep_scores)), _p(idx), _p(val))
    if rc:
        raise RuntimeError(f"orc_topk_csr failed: {rc}")
    return idx, val


def predict_weighted_per_instance(y_proba, k: int, th: float = 0.0, a=None, b=None,
                                  dtype=None, keep_scores: bool = False):
    """Same container/dtype/shape contract as the reference (weighted_prediction.py:91-188)."""
    if isinstance(y_proba, csr_matrix):
        n, m = y_proba.shape
        if k <= 0:
            # numba_csr_functions.py:631-653 -> :564-570 -> :516-517: the STORED labels of a row whose gain
            # (data * a[idx] + b[idx], numpy promotion) is >= th, appended row by row; data = ones in the data dtype
            g = y_proba.data
            if a is not None:
                g = g * np.asarray(a)[y_proba.indices]
            if b is not None:
                g = g + np.asarray(b)[y_proba.indices]
            thv = np.asarray(th)[y_proba.indices] if np.ndim(th) > 0 else th
            keep = g >= thv
            rows = np.repeat(np.arange(n), np.diff(y_proba.indptr))
            counts = np.bincount(rows[keep], minlength=n)
            indptr = np.concatenate([[0], np.cumsum(counts)]).astype(y_proba.indptr.dtype)
            return csr_matrix((np.ones(int(keep.sum()), dtype=y_proba.data.dtype), y_proba.indices[keep], indptr),
                              shape=(n, m), dtype=dtype)
        idx, val = topk_indices_csr(y_proba, k, a, b, keep_scores)
        indptr = (np.arange(n + 1, dtype=y_proba.indptr.dtype) * k)
        return csr_matrix((val.reshape(-1), idx.reshape(-1).astype(y_proba.indices.dtype), indptr),
                          shape=(n, m), dtype=dtype)
    y_proba = np.asarray(y_proba)
    n, m = y_proba.shape
    out = np.zeros((n, m), dtype=y_proba.dtype if dtype is None else dtype)
    if k > 0:
        idx, val = topk_indices_dense(y_proba, k, a, b)
        out[np.arange(n)[:, None], idx] = val if keep_scores else 1
    else:
        g = y_proba
        if a is not None:
            g = g * a
        if b is not None:
            g = g + b
        out[g >= th] = 1
    return out


def predict_top_k(y_proba, k, dtype=None, keep_scores=False):
    return predict_weighted_per_instance(y_proba, k, dtype=dtype, keep_scores=keep_scores)


# ------------------------------------------------------------------------------------------
# confusion matrix (confusion_matrix.py:160-399)
# ------------------------------------------------------------------------------------------

def calculate_confusion_matrix(y_true, y_pred, normalize=False, skip_tn=False, axis=0, dtype=None):
    """Returns (tp, fp, fn, tn) as numpy vectors in `dtype` (None -> y_true.dtype)."""
    if axis != 0:
        # axis=1 sums run along the contiguous axis where numpy adds pairwise: the oracle
        # states the value (float64 row sums rounded to the output dtype), not the bits.
        return _confmat_axis1(y_true, y_pred, normalize, skip_tn, dtype)
    n, m = y_true.shape
    out_dt = np.dtype(y_true.dtype if dtype is None else dtype)
    tp, fp, fn = (np.empty(m, dtype=np.float64) for _ in range(3))
    if isinstance(y_true, csr_matrix):
        dt = np.result_type(y_true.dtype, np.float32)
        yt, yp = y_true.astype(dt), y_pred.astype(dt)
        acc_is_T = int(out_dt == dt)
        if not acc_is_T and out_dt != np.float64:
            raise NotImplementedError("oracle: csr accumulate dtype")
        rc = getattr(lib(), "orc_confmat_csr_" + _sfx(dt))(
            _p(yt.data), _p(yt.indices.astype(np.int32)), _p(yt.indptr.astype(np.int64)),
            _p(yp.data), _p(yp.indices.astype(np.int32)), _p(yp.indptr.astype(np.int64)),
            C.c_int64(n), C.c_int64(m), C.c_int(acc_is_T), _p(tp), _p(fp), _p(fn))
    else:
        dt = np.result_type(y_true.dtype, y_pred.dtype, np.float32)
        yt = np.ascontiguousarray(y_true, dtype=dt)
        yp = np.ascontiguousarray(y_pred, dtype=dt)
        acc_is_T = int(out_dt == dt)
        if not acc_is_T and out_dt != np.float64:
            raise NotImplementedError("oracle: dense accumulate dtype")
        rc = getattr(lib(), "orc_confmat_dense_" + _sfx(dt))(
            _p(yt), C.c_int64(m), _p(yp), C.c_int64(m), C.c_int64(n), C.c_int64(m),
            C.c_int(acc_is_T), _p(tp), _p(fp), _p(fn))
    if rc:
        raise RuntimeError(f"orc_confmat failed: {rc}")
    tp, fp, fn = tp.astype(out_dt), fp.astype(out_dt), fn.astype(out_dt)
    if normalize:  # confusion_matrix.py:265-266
        tp, fp, fn = tp / n, fp / n, fn / n
    if skip_tn:    # :391-393
        tn = tp.copy()
        tn[:] = -1
    else:          # :397
        tn = -tp - fp - fn + (1.0 if normalize else n)
    return tp, fp, fn, tn


def _confmat_axis1(y_true, y_pred, normalize, skip_tn, dtype):
    n, m = y_true.shape
    yt = np.asarray(y_true.todense()) if isinstance(y_true, csr_matrix) else np.asarray(y_true)
    yp = np.asarray(y_pred.todense()) if isinstance(y_pred, csr_matrix) else np.asarray(y_pred)
    out_dt = np.dtype(yt.dtype if dtype is None else dtype)
    yt64, yp64 = yt.astype(np.float64), yp.astype(np.float64)
    tp = (yt64 * yp64).sum(1).astype(out_dt)
    fp = ((1 - yt64) * yp64).sum(1).astype(out_dt)
    fn = (yt64 * (1 - yp64)).sum(1).astype(out_dt)
    if normalize:
        tp, fp, fn = tp / n, fp / n, fn / n
    if skip_tn:
        tn = tp.copy()
        tn[:] = -1
    else:
        tn = -tp - fp - fn + (1.0 if normalize else m)
    return tp, fp, fn, tn


# ------------------------------------------------------------------------------------------
# BCA, 0-th order ETU approximation (block_coordinate.py:296-499)
# ------------------------------------------------------------------------------------------

def _utility(mid, c1, b2, eps, tp, fp, fn, tn, div, aggregation):
    """block_coordinate.py:54-90 on (E*/n): per-label metric vector in C, numpy .mean()/.sum()
    (pairwise) exactly like the reference."""
    m = tp.shape[0]
    vec = np.empty(m, dtype=np.float64)
    lib().orc_binary_metric_vec(C.c_int(mid), _p(tp), _p(fp), _p(fn), _p(tn), C.c_int64(m),
                                C.c_double(div), C.c_double(c1), C.c_double(b2), C.c_double(eps),
                                _p(vec))
    return vec.mean() if aggregation == "mean" else vec.sum()


def _compact_to_dense_pred(idx, n, m):
    pred = np.zeros((n, m), dtype=np.uint8)
    rows = np.repeat(np.arange(n), idx.shape[1])
    flat = idx.reshape(-1)
    ok = flat >= 0
    pred[rows[ok], flat[ok]] = 1
    return pred


def predict_using_bc_with_0approx(y_proba, metric: str, k: int, metric_aggregation="mean",
                                  normalize_conf_matrix=True, beta=1.0, epsilon=1e-9,
                                  maximize=True, tolerance=1e-6, init_y_pred="top", max_iters=100,
                                  shuffle_order=True, skip_tn=False, seed=None,
                                  return_state=False, mix=None):
    """Sequential BCA; mix=(alpha, k, m) selects the mixed instance-precision utility.  See _bca."""
    with _Mix(mix):
        return _bca(y_proba, metric, k, metric_aggregation, normalize_conf_matrix, beta, epsilon, maximize,
                    tolerance, init_y_pred, max_iters, shuffle_order, skip_tn, seed, return_state)


def _bca(y_proba, metric, k, metric_aggregation, normalize_conf_matrix, beta, epsilon, maximize, tolerance,
         init_y_pred, max_iters, shuffle_order, skip_tn, seed, return_state):
    """Sequential BCA.  Returns (pred, meta): pred is a dense uint8 0/1 matrix for dense input
    and an (n, k) int32 id matrix (ascending, -1 padded) for CSR input; meta carries
    "utilities" and "iters" like the reference's return_meta (:385, :479-480)."""
    mid, c1, b2, eps = metric_params(metric, beta, epsilon)
    # Reference quirk kept on purpose: _calculate_utility (block_coordinate.py:54-63, called at
    # :438/:469) does NOT forward metric_kwargs, so the reported utilities and the stopping test
    # use the metric's default beta=1, epsilon=1e-9 while the gains (:174-185) use the kwargs.
    _, u_c1, u_b2, u_eps = metric_params(metric, beta if metric == "precision_at_k" else 1.0)
    is_csr = isinstance(y_proba, csr_matrix)
    n, m = y_proba.shape
    n_div = n if normalize_conf_matrix else 1            # :403-405
    n_order = n_div if not normalize_conf_matrix else n  # order = arange(n) AFTER the overwrite
    greedy = isinstance(init_y_pred, str) and init_y_pred == "greedy"
    dt = y_proba.dtype
    sfx = _sfx(dt)
    t0 = time.time()

    # ---- initial prediction (:28-51, :410)
    if is_csr:
        if k <= 0:
            raise NotImplementedError("oracle: CSR BCA needs k > 0")
        data = np.ascontiguousarray(y_proba.data)
        indices = np.ascontiguousarray(y_proba.indices, dtype=np.int32)
        indptr = y_proba.indptr.astype(np.int64)
        if isinstance(init_y_pred, str) and init_y_pred == "top":
            pidx, _ = topk_indices_csr(y_proba, k)
            nnz = np.diff(indptr)
            plen = np.minimum(nnz, k).astype(np.int32)
            for i in np.nonzero(nnz < k)[0]:
                pidx[i, nnz[i]:] = -1
        elif isinstance(init_y_pred, np.ndarray):
            pidx = np.ascontiguousarray(init_y_pred, dtype=np.int32).copy()
            plen = (pidx >= 0).sum(1).astype(np.int32)
        else:
            raise NotImplementedError("oracle: init_y_pred for CSR must be 'top' or an id matrix")
    else:
        eta = np.ascontiguousarray(y_proba)
        if isinstance(init_y_pred, str) and init_y_pred == "top":
            pidx, _ = topk_indices_dense(eta, k)
            pred = _compact_to_dense_pred(pidx, n, m)
        elif isinstance(init_y_pred, str) and init_y_pred in ("random", "greedy"):
            # utils.py:104-116 random_at_k_np
            pred = np.zeros((n, m), dtype=np.uint8)
            rng0 = np.random.default_rng(seed)
            labels_range = np.arange(m)
            for i in range(n):
                pred[i, rng0.choice(labels_range, k, replace=False, shuffle=False)] = 1
        else:
            pred = np.ascontiguousarray(np.asarray(init_y_pred) != 0, dtype=np.uint8).copy()

    def recompute():
        tp, fp, fn = (np.empty(m, dtype=np.float64) for _ in range(3))
        if is_csr:
            # confusion_matrix.py:174-228 with y_true = y_proba, y_pred = ones at pidx
            valid = pidx >= 0
            p_ptr = np.concatenate([[0], np.cumsum(valid.sum(1))]).astype(np.int64)
            p_idx = np.ascontiguousarray(pidx[valid], dtype=np.int32)
            p_data = np.ones(p_idx.shape[0], dtype=dt)
            rc = getattr(lib(), "orc_confmat_csr_" + sfx)(
                _p(data), _p(indices), _p(indptr), _p(p_data), _p(p_idx), _p(p_ptr),
                C.c_int64(n), C.c_int64(m), C.c_int(0), _p(tp), _p(fp), _p(fn))
        else:
            predT = pred.astype(dt)
            rc = getattr(lib(), "orc_confmat_dense_" + sfx)(
                _p(eta), C.c_int64(m), _p(predT), C.c_int64(m), C.c_int64(n), C.c_int64(m),
                C.c_int(0), _p(tp), _p(fp), _p(fn))
        if rc:
            raise RuntimeError("oracle confmat failed")
        if skip_tn:
            tn = np.full(m, -1.0)
        else:
            tn = -tp - fp - fn + n   # y_true.shape[0], not the overwritten n
        return tp, fp, fn, tn

    rng = np.random.default_rng(seed)     # :413
    order = np.arange(n_order)            # :414
    meta = {"utilities": [], "iters": 0}
    tp = fp = fn = tn = None
    for j in range(1, max_iters + 1):
        if shuffle_order:
            rng.shuffle(order)            # :419 (cumulative, in place)
        if greedy:
            tp, fp, fn, tn = (np.zeros(m) for _ in range(4))
        else:
            tp, fp, fn, tn = recompute()
        old_u = _utility(mid, u_c1, u_b2, u_eps, tp, fp, fn, tn, n_div, metric_aggregation)
        order64 = np.ascontiguousarray(order, dtype=np.int64)
        if is_csr:
            rc = getattr(lib(), "orc_bca_csr_sweep_" + sfx)(
                _p(data), _p(indices), _p(indptr), C.c_int64(n), C.c_int64(m), _p(pidx),
                _p(plen), _p(order64), C.c_int64(order64.size), C.c_int(k), C.c_int(mid),
                C.c_double(c1), C.c_double(b2), C.c_double(eps), C.c_double(n_div),
                C.c_int(int(maximize)), C.c_int(int(greedy)), _p(tp), _p(fp), _p(fn),
                None if skip_tn else _p(tn), C.c_double(-1.0))
        else:
            rc = getattr(lib(), "orc_bca_dense_sweep_" + sfx)(
                _p(eta), C.c_int64(n), C.c_int64(m), C.c_int64(m), _p(pred), _p(order64),
                C.c_int64(order64.size), C.c_int(k), C.c_int(mid), C.c_double(c1),
                C.c_double(b2), C.c_double(eps), C.c_double(n_div), C.c_int(int(skip_tn)),
                C.c_int(int(maximize)), C.c_int(int(greedy)), _p(tp), _p(fp), _p(fn), _p(tn))
        if rc:
            raise RuntimeError("oracle sweep failed")
        running = (tp.copy(), fp.copy(), fn.copy(), tn.copy())
        tp, fp, fn, tn = recompute()
        new_u = _utility(mid, u_c1, u_b2, u_eps, tp, fp, fn, tn, n_div, metric_aggregation)
        greedy = False
        meta["iters"] = j
        meta["utilities"].append(float(new_u))
        if (maximize and new_u - old_u < tolerance) or (not maximize and new_u - old_u > tolerance):
            break
    meta["time"] = time.time() - t0
    if return_state:
        meta["state"] = (tp, fp, fn, tn)
        meta["running_state"] = running
    return (pidx if is_csr else pred), meta


# ------------------------------------------------------------------------------------------
# BCA for coverage (block_coordinate.py:539-701, CSR semantics -- the dense variant of the
# reference is not a valid oracle; dense inputs are converted to CSR)
# ------------------------------------------------------------------------------------------

def predict_optimizing_coverage_using_bc(y_proba, k: int, alpha: float = 1, tolerance=1e-6,
                                         init_y_pred="top", max_iters=100, shuffle_order=True,
                                         seed=None):
    if not isinstance(y_proba, csr_matrix):
        y_proba = csr_matrix(np.asarray(y_proba))
        y_proba.sort_indices()
    n, m = y_proba.shape
    dt = y_proba.dtype
    sfx = _sfx(dt)
    data = np.ascontiguousarray(y_proba.data)
    indices = np.ascontiguousarray(y_proba.indices, dtype=np.int32)
    indptr = y_proba.indptr.astype(np.int64)
    if (np.diff(indptr) <= k).any():
        raise NotImplementedError("oracle: coverage rows need more than k stored labels")
    greedy = isinstance(init_y_pred, str) and init_y_pred == "greedy"
    if isinstance(init_y_pred, str) and init_y_pred == "top":
        pidx, _ = topk_indices_csr(y_proba, k)
    else:
        pidx = np.ascontiguousarray(init_y_pred, dtype=np.int32).copy()
    plen = np.full(n, k, dtype=np.int32)
    Ef = np.empty(m, dtype=np.float64)
    t0 = time.time()

    def state():
        getattr(lib(), "orc_cov_state_csr_" + sfx)(
            _p(data), _p(indices), _p(indptr), C.c_int64(n), C.c_int64(m), _p(pidx), _p(plen),
            C.c_int(k), _p(Ef))

    def utility():
        cov = 1 - Ef.mean()                # :592
        if alpha < 1:                      # :593-595 (value, float64)
            sel = np.zeros(m)
            rows = np.repeat(np.arange(n), k)
            sub = np.asarray(y_proba[rows, pidx.reshape(-1)]).reshape(-1).astype(np.float64)
            np.add.at(sel, pidx.reshape(-1), sub)
            cov = alpha * cov + (1 - alpha) * (sel / n / k).sum()
        return cov

    rng = np.random.default_rng(seed)
    order = np.arange(n)
    meta = {"utilities": [], "iters": 0}
    for j in range(1, max_iters + 1):
        if shuffle_order:
            rng.shuffle(order)
        if greedy:
            Ef[:] = 1.0
        else:
            state()
        old = utility()
        order64 = np.ascontiguousarray(order, dtype=np.int64)
        rc = getattr(lib(), "orc_cov_csr_sweep_" + sfx)(
            _p(data), _p(indices), _p(indptr), C.c_int64(n), C.c_int64(m), _p(pidx), _p(order64),
            C.c_int64(n), C.c_int(k), C.c_double(float(alpha)), C.c_int(int(greedy)), _p(Ef))
        if rc:
            raise RuntimeError("oracle coverage sweep failed")
        state()
        new = utility()
        greedy = False
        meta["iters"] = j
        meta["utilities"].append(float(new))
        if new <= old + tolerance:          # :690
            break
    meta["time"] = time.time() - t0
    meta["Ef"] = Ef.copy()
    return pidx, meta


# ------------------------------------------------------------------------------------------
# Frank-Wolfe (frank_wolfe.py:407-690) for macro-averaged built-in metrics
# ------------------------------------------------------------------------------------------

def macro_metric_and_grad(metric: str, tp, fp, fn, tn, beta=1.0, epsilon=1e-9):
    """Value and d/d(tp,fp,fn,tn) of mean_j binary_metric -- closed forms that the reference
    obtains by automatic differentiation (frank_wolfe.py:368-376)."""
    tp, fp, fn, tn = (np.asarray(x, dtype=np.float64) for x in (tp, fp, fn, tn))
    m = tp.shape[0]
    z = np.zeros(m)
    b2 = beta ** 2 if metric != "f1" else 1.0
    e = epsilon
    if metric in ("fbeta", "f1"):
        c = 1 + b2
        D = b2 * (tp + fp) + tp + fn + e
        val = c * tp / D
        g = (c * (D - tp * c) / D ** 2, -c * tp * b2 / D ** 2, -c * tp / D ** 2, z)
    elif metric == "precision":
        D = tp + fp + e
        val = tp / D
        g = ((fp + e) / D ** 2, -tp / D ** 2, z, z)
    elif metric == "recall":
        D = tp + fn + e
        val = tp / D
        g = ((fn + e) / D ** 2, z, -tp / D ** 2, z)
    elif metric == "jaccard":
        D = tp + fp + fn + e
        val = tp / D
        g = ((fp + fn + e) / D ** 2, -tp / D ** 2, -tp / D ** 2, z)
    else:
        Dp, Dn = tp + fn + e, tn + fp + e
        tpr, tnr = tp / Dp, tn / Dn
        d_tpr = ((fn + e) / Dp ** 2, z, -tp / Dp ** 2, z)              # wrt tp, fp, fn, tn
        d_tnr = (z, -tn / Dn ** 2, z, (fp + e) / Dn ** 2)
        if metric == "balanced_accuracy":
            val = (tpr + tnr) / 2
            f_p, f_n = np.full(m, 0.5), np.full(m, 0.5)
        elif metric == "gmean":
            val = np.sqrt(tpr * tnr)
            with np.errstate(divide="ignore", invalid="ignore"):
                f_p, f_n = 0.5 * tnr / val, 0.5 * tpr / val
        elif metric == "hmean":
            val = 2 * tpr * tnr / (tpr + tnr)
            f_p, f_n = 2 * tnr ** 2 / (tpr + tnr) ** 2, 2 * tpr ** 2 / (tpr + tnr) ** 2
        else:
            raise ValueError(metric)
        g = tuple(f_p * a + f_n * b for a, b in zip(d_tpr, d_tnr))
    return float(val.mean()), tuple(x / m for x in g)


def find_classifier_using_fw(y_true, y_proba, metric: str, k: int, max_iters=100,
                             init_classifier="top", maximize=True, normalize_conf_matrix=True,
                             beta=1.0, epsilon=1e-9, tolerance=1e-6, search_for_best_alpha=True,
                             alpha_tolerance=0.001, alpha_uniform_search_step=0.0001,
                             skip_tn=False, seed=None, mix=None, alpha_search_algo="uniform", micro=False,
                             recall_precision_alpha=None):
    """Returns (a, b, p, meta) with the truncation rules of frank_wolfe.py:644-670.
    alpha_search_algo="ternary": utils.py:187-201 with eps = alpha_tolerance (:627).
    mix=(alpha, k, m): objective sum_j [(1 - alpha) tp_j / k + alpha metric_j / m] (:838-915)."""
    mid, c1, b2, eps = metric_params(metric, beta, epsilon)

    def macro_metric_and_grad_mix(metric, tp, fp, fn, tn, beta=1.0, epsilon=1e-9):
        if recall_precision_alpha is not None:   # frank_wolfe.py:917-938: sum_j (1 - al) recall_j + al precision_j
            al = recall_precision_alpha
            vr, gr = macro_metric_and_grad("recall", tp, fp, fn, tn, epsilon=epsilon)
            vp, gp = macro_metric_and_grad("precision", tp, fp, fn, tn, epsilon=epsilon)
            return m * ((1 - al) * vr + al * vp), tuple(m * ((1 - al) * x + al * y) for x, y in zip(gr, gp))
        if micro:   # metrics.py:68-100: the metric of the four sums; every label gets the same gradient
            sums = [np.array([np.sum(x)], dtype=np.float64) for x in (tp, fp, fn, tn)]
            v, g = macro_metric_and_grad(metric, *sums, beta=beta, epsilon=epsilon)
            return v, tuple(np.full(m, float(x[0])) for x in g)
        v, (gtp, gfp, gfn, gtn) = macro_metric_and_grad(metric, tp, fp, fn, tn, beta=beta, epsilon=epsilon)
        if mix is None:
            return v, (gtp, gfp, gfn, gtn)
        al, kk, _ = mix
        return al * v + (1 - al) * float(np.sum(tp)) / kk, (al * gtp + (1 - al) / kk, al * gfp, al * gfn, al * gtn)

    n, m = y_proba.shape
    rng = np.random.default_rng(seed)
    A = np.zeros((max_iters + 1, m), dtype=np.float32)   # :503-505
    B = np.zeros((max_iters + 1, m), dtype=np.float32)
    P = np.ones(max_iters + 1, dtype=np.float32)
    if isinstance(init_classifier, str) and init_classifier == "top":
        A[0] = 1.0
        B[0] = -0.5
    elif isinstance(init_classifier, str) and init_classifier == "random":
        A[0] = rng.random(m)
        B[0] = rng.random(m) - 0.5
    else:
        A[0], B[0] = init_classifier

    def conf(a, b):
        pred = predict_weighted_per_instance(y_proba, k, th=0.0, a=a, b=b)
        return [np.asarray(x, dtype=np.float64) for x in calculate_confusion_matrix(
            y_true, pred, normalize=normalize_conf_matrix, skip_tn=skip_tn)]

    def value(c):
        return macro_metric_and_grad_mix(metric, *c, beta=beta, epsilon=epsilon)[0]

    Cm = conf(A[0], B[0])
    u0 = value(Cm)
    meta = {"alphas": [], "classifiers_utilities": [u0], "utilities": [u0]}
    alphas = np.arange(0 + alpha_uniform_search_step, 1, alpha_uniform_search_step)
    t0 = time.time()
    it = 0
    new_u = u0
    for i in range(1, max_iters + 1):
        it = i
        old_u, (gtp, gfp, gfn, gtn) = macro_metric_and_grad_mix(metric, *Cm, beta=beta, epsilon=epsilon)
        A[i] = gtp - gfp - gfn + gtn     # :595-596 (float32 store)
        B[i] = gfp - gtn
        if not maximize:
            A[i] *= -1
            B[i] *= -1
        Ci = conf(A[i], B[i])
        u_i = value(Ci)
        if search_for_best_alpha:
            ba, bv = C.c_double(), C.c_double()

            def f_at(al):
                """metric of (1 - al) C + al C_i (frank_wolfe.py:393-398): a one-point 'grid'"""
                if micro or recall_precision_alpha is not None:
                    return value([(1 - al) * x + al * y for x, y in zip(Cm, Ci)])
                one = np.array([al], dtype=np.float64)
                oa, ov = C.c_double(), C.c_double()
                with _Mix(mix):
                    lib().orc_fw_alpha_search(
                        C.c_int(mid), *[_p(x) for x in Cm], *[_p(x) for x in Ci], C.c_int64(m),
                        _p(one), C.c_int64(1), C.c_double(c1), C.c_double(b2), C.c_double(eps),
                        C.byref(oa), C.byref(ov), C.c_int(1))
                return ov.value

            if alpha_search_algo == "ternary":           # utils.py:187-201, verbatim
                low, high = 0, 1
                while high - low > alpha_tolerance:
                    mid1 = low + (high - low) / 3
                    mid2 = high - (high - low) / 3
                    if f_at(mid1) < f_at(mid2):
                        high = mid2
                    else:
                        low = mid1
                ba.value = (low + high) / 2
            elif recall_precision_alpha is not None:      # utils.py:174-184, point by point
                best, best_val = 0.0, f_at(0.0)
                for al_ in alphas:
                    sc = f_at(al_)
                    if sc > best_val:
                        best, best_val = al_, sc
                ba.value = best
            elif micro:                                   # utils.py:174-184 on the scalar objective
                grid = np.concatenate([[0.0], alphas])
                S, Si = [float(np.sum(x)) for x in Cm], [float(np.sum(x)) for x in Ci]
                T = [np.ascontiguousarray((1 - grid) * s_ + grid * si_) for s_, si_ in zip(S, Si)]
                vals = np.empty(grid.size, dtype=np.float64)
                lib().orc_binary_metric_vec(C.c_int(mid), *[_p(x) for x in T], C.c_int64(grid.size), C.c_double(1.0),
                                            C.c_double(c1), C.c_double(b2), C.c_double(eps), _p(vals))
                ba.value = float(grid[int(np.argmax(vals))])   # argmax = first maximum = first strict improvement
            else:
                with _Mix(mix):
                    lib().orc_fw_alpha_search(
                        C.c_int(mid), *[_p(x) for x in Cm], *[_p(x) for x in Ci], C.c_int64(m),
                        _p(alphas), C.c_int64(alphas.size), C.c_double(c1), C.c_double(b2),
                        C.c_double(eps), C.byref(ba), C.byref(bv), C.c_int(0))
            alpha = ba.value
        else:
            alpha = 2 / (i + 1)
        Cm = [(1 - alpha) * x + alpha * y for x, y in zip(Cm, Ci)]
        new_u = value(Cm)
        if alpha < alpha_tolerance or (maximize and new_u - old_u < tolerance) or (
                not maximize and old_u - new_u < tolerance):
            A, B, P = A[:i], B[:i], P[:i]
            break
        meta["alphas"].append(alpha)
        meta["classifiers_utilities"].append(u_i)
        meta["utilities"].append(new_u)
        P[:i] *= 1 - alpha
        P[i] = alpha
    meta["iters"] = it
    meta["time"] = time.time() - t0
    meta["final_utility"] = new_u
    return A, B, P, meta


# ------------------------------------------------------------------------------------------
# online / greedy steps on CSR rows (block_coordinate.py:212-293 with greedy=True, only_pred=True, then
# confusion_matrix.py:421-432 -> numba_csr_functions.py:421-452), driven like
# experiments/omma_wrappers_online_methods.py:223-266.  numpy restatement, one python step per instance.
# ------------------------------------------------------------------------------------------

def _metric_vec(mid, c1, b2, eps, tp, fp, fn, tn):
    out = np.empty(tp.shape[0], dtype=np.float64)
    lib().orc_binary_metric_vec(C.c_int(mid), _p(np.ascontiguousarray(tp)), _p(np.ascontiguousarray(fp)),
                                _p(np.ascontiguousarray(fn)), _p(np.ascontiguousarray(tn)), C.c_int64(tp.shape[0]),
                                C.c_double(1.0), C.c_double(c1), C.c_double(b2), C.c_double(eps), _p(out))
    return out


def online_greedy_csr(y_proba: csr_matrix, y_true: csr_matrix, k: int, metric: str, skip_tn: bool = False,
                      init=(1e-6, 1e-6, 1e-6, 1e-6), maximize: bool = True):
    """Returns (pred_idx [n, k] ascending label ids, state [4, m]).  Ties between equal gains go to the lowest label
    id (the reference's argpartition order is unpinned)."""
    n, m = y_proba.shape
    mid, c1, b2, eps = metric_params(metric)
    tp, fp, fn, tn = (np.full(m, float(v), dtype=np.float64) for v in init)
    pred = np.full((n, k), -1, dtype=np.int32)
    dt = y_proba.data.dtype
    one = dt.type(1)
    for i in range(n):
        s, e = y_proba.indptr[i], y_proba.indptr[i + 1]
        t_data, t_idx = y_proba.data[s:e], y_proba.indices[s:e]
        om = (one - t_data).astype(dt)                      # :252 (1 - t_data) in the data dtype
        neg_tp, neg_fp, pos_fn = tp[t_idx], fp[t_idx], fn[t_idx]
        pos_tpp = (neg_tp + t_data) / n
        pos_fpp = (neg_fp + om) / n
        neg_fnn = (pos_fn + t_data) / n
        neg_tp, neg_fp, pos_fn = neg_tp / n, neg_fp / n, pos_fn / n
        pos_tn = tn[t_idx]
        neg_tnn = pos_tn
        if not skip_tn:
            neg_tnn = (pos_tn + om) / n
            pos_tn = pos_tn / n
        gains = _metric_vec(mid, c1, b2, eps, pos_tpp, pos_fpp, pos_fn, pos_tn) - _metric_vec(
            mid, c1, b2, eps, neg_tp, neg_fp, neg_fnn, neg_tnn)
        if not maximize:
            gains = -gains
        if t_idx.size > k:                                  # numba_csr_functions.py:456-466
            sel = np.sort(np.lexsort((t_idx, -gains))[:k])
            p_idx = t_idx[sel]
        else:
            p_idx = t_idx.copy()
        pred[i, :p_idx.size] = p_idx
        # update with the true row (prediction entries are ones of the data dtype)
        us, ue = y_true.indptr[i], y_true.indptr[i + 1]
        u_data, u_idx = y_true.data[us:ue].astype(dt), y_true.indices[us:ue]
        true_of = dict(zip(u_idx.tolist(), u_data.tolist()))
        in_pred = set(p_idx.tolist())
        if not skip_tn:
            tn += 1
        for j in p_idx.tolist():
            if j in true_of:
                v = dt.type(true_of[j])
                tpd = dt.type(one * v)
                ftd = dt.type(float(one) * (1.0 - float(v)))
                tp[j] += tpd
                fp[j] += ftd
                if not skip_tn:
                    tn[j] -= tpd
                    tn[j] -= ftd
            else:
                fp[j] += float(one)
                if not skip_tn:
                    tn[j] -= float(one)
        for j, v in zip(u_idx.tolist(), u_data.tolist()):
            v = dt.type(v)
            fnd = dt.type(float(v) * (1.0 - 1.0)) if j in in_pred else v
            fn[j] += fnd
            if not skip_tn:
                tn[j] -= fnd
    return pred, np.stack([tp, fp, fn, tn])
