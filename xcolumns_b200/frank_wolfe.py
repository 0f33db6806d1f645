"""Frank-Wolfe search for a randomized weighted classifier on the GPU
(drop-in for xcolumns/frank_wolfe.py:294-360, :407-690 and the macro wrappers :698-830).

Per iteration the reference does: autograd gradient of the metric -> linear classifier (a, b) ->
weighted top-k over all rows -> confusion matrix vs y_true -> 10^4-point line search.  Here the
n x m work is ONE fused streaming kernel (top-k + tp/count accumulation), the gradient is a
closed form, and the whole alpha grid is evaluated by one kernel; only four scalars per
iteration travel to the host (for the stopping rules and ``meta``).
"""
from __future__ import annotations

import ctypes as C
import os
from time import time
from typing import Any, Callable, Dict, Optional, Tuple, Union

import numpy as np
import torch
from scipy.sparse import csr_matrix

from . import _device as dev
from . import metrics as M
from ._lib import MetricParams
from .distributed import make_comm
from .types import DefaultDataDType, DenseMatrix, DType, Matrix
from .utils import add_kwargs_to_signature, log_info, log_warning
from .weighted_prediction import _check_k


class RandomizedWeightedClassifier:
    """A set of weighted classifiers (rows of ``a`` and ``b``), one of which is drawn per instance
    with probabilities ``p`` (xcolumns/frank_wolfe.py:294-360)."""

    def __init__(self, k: int, a: DenseMatrix, b: DenseMatrix, p: DenseMatrix):
        if not isinstance(k, int):
            raise ValueError("k must be an integer")
        if not all(isinstance(v, (np.ndarray, torch.Tensor)) for v in (a, b, p)):
            raise ValueError("a, b, and p must be ndarray")
        if a.shape != b.shape or a.shape[0] != p.shape[0]:
            raise ValueError("a, b must have the same shape and the number of rows must be equal to the number of rows of p")
        self.k, self.a, self.b, self.p = k, a, b, p

    def predict(self, y_proba: Matrix, dtype: Optional[DType] = None, seed: Optional[int] = None) -> Matrix:
        if y_proba.shape[1] != self.a.shape[1]:
            raise ValueError(f"This classifier support the input matrix with {self.a.shape[1]} columns (labels), got {y_proba.shape[1]}")
        return predict_using_randomized_weighted_classifier(y_proba, self.k, self.a, self.b, self.p, dtype=dtype, seed=seed)


def _draw_classifiers(P: np.ndarray, n: int, seed) -> np.ndarray:
    """One classifier index per row, the same stream as the reference's per-row
    ``rng.choice(arange(len(p)), p=p)`` (frank_wolfe.py:70, :199): Generator.choice with probabilities draws ONE
    uniform double per call and returns cdf.searchsorted(u, side="right") on the normalised cumulative sum, so n
    calls consume exactly the n doubles of one rng.random(n).  Vectorised: seconds -> milliseconds at n = 307 k."""
    rng = np.random.default_rng(seed)
    p = np.asarray(P, dtype=np.float64)
    # the checks Generator.choice applies to p (non-negative, sums to 1 within sqrt(eps))
    if p.ndim != 1 or (p < 0).any() or not np.isfinite(p).all():
        raise ValueError("probabilities are not non-negative")
    atol = np.sqrt(np.finfo(np.float64).eps)
    if isinstance(P, np.ndarray) and np.issubdtype(P.dtype, np.floating):
        atol = max(atol, np.sqrt(np.finfo(P.dtype).eps))
    if abs(float(np.sum(p)) - 1.0) > atol:
        raise ValueError("probabilities do not sum to 1")
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return cdf.searchsorted(rng.random(n), side="right").astype(np.int64)


def predict_using_randomized_weighted_classifier(y_proba: Matrix, k: int, classifiers_a, classifiers_b,
                                                 classifiers_proba, dtype=None, seed=None) -> Matrix:
    """Prediction of a randomized weighted classifier (xcolumns/frank_wolfe.py:175-291): every row
    draws one classifier c_i ~ p with numpy's Generator (same stream as the reference), rows are
    grouped by classifier and each group runs through the weighted top-k kernel."""
    if not isinstance(y_proba, (np.ndarray, torch.Tensor, csr_matrix)):
        raise ValueError("y_proba must be either np.ndarray, torch.Tensor, or csr_matrix")
    _check_k(k)
    if k < 0:
        raise ValueError("k must be >= 0")
    to_np = lambda v: v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    A, B, P = to_np(classifiers_a), to_np(classifiers_b), to_np(classifiers_proba)
    n, m = y_proba.shape
    if A.shape[1] != m or B.shape[1] != m:
        raise ValueError("classifiers_a, classifier_b, and classifiers_proba must have the same number of columns as y_proba")
    if A.shape[0] != B.shape[0] or A.shape[0] != P.shape[0]:
        raise ValueError("classifiers_a, classifier_b, and classifiers_proba must have the same number of rows")
    choice = _draw_classifiers(P, n, seed)
    device = dev.pick_device(y_proba)
    if k == 0 and isinstance(y_proba, csr_matrix):
        # no budget on CSR rows: the STORED labels whose gain under the row's classifier is >= 0
        # (frank_wolfe.py:152-170 -> numba_csr_functions.py:516-517); elementwise on the nnz entries on the device,
        # gains formed like numpy does there (data dtype x classifier dtype, separate multiply and add)
        c = dev.csr_to_device(y_proba, device)
        gdt = torch.promote_types(c.data.dtype, torch.from_numpy(A[:0]).dtype)
        Ad = torch.from_numpy(np.ascontiguousarray(A)).to(device=device, dtype=gdt)
        Bd = torch.from_numpy(np.ascontiguousarray(B)).to(device=device, dtype=gdt)
        counts = c.indptr[1:] - c.indptr[:-1]
        cls_of = torch.repeat_interleave(torch.from_numpy(choice).to(device), counts)
        row_of = torch.repeat_interleave(torch.arange(n, device=device), counts)
        idx = c.indices.long()
        keep = (c.data.to(gdt) * Ad[cls_of, idx] + Bd[cls_of, idx]) >= 0
        kept = torch.zeros(n, dtype=torch.int64, device=device).index_add_(0, row_of[keep], torch.ones_like(row_of[keep]))
        indptr = torch.cat([torch.zeros(1, dtype=torch.int64, device=device), kept.cumsum(0)])
        new_idx = c.indices[keep].cpu().numpy().astype(y_proba.indices.dtype)
        out_dt = y_proba.data.dtype if dtype is None else dtype
        return csr_matrix((np.ones(new_idx.shape[0], dtype=out_dt), new_idx,
                           indptr.cpu().numpy().astype(y_proba.indptr.dtype)), shape=y_proba.shape)
    if k == 0:
        # no budget: every label with a non-negative gain under the row's classifier (frank_wolfe.py:80-105
        # with k = 0 -> weighted_prediction.py:55-56); rows are grouped by classifier like below
        from .weighted_prediction import _threshold_dense
        d = dev.dense_to_device(y_proba, device)
        gdt = torch.promote_types(d.torch_dtype, torch.from_numpy(A[:0]).dtype)
        g_code = 0 if gdt == torch.float32 else 1
        out = torch.zeros((n, m), dtype=d.torch_dtype, device=device)
        for c in np.unique(choice):
            rows = torch.from_numpy(np.nonzero(choice == c)[0]).to(device)
            sub = dev.DenseDev(d.t[rows][:, :m].contiguous(), int(rows.numel()), m, m, d.code, 0)
            a = dev.vec_to_device(A[c], device, gdt, m, "a")
            b = dev.vec_to_device(B[c], device, gdt, m, "b")
            out[rows] = _threshold_dense(out, sub, a, b, g_code, 0.0, None)
        if isinstance(y_proba, torch.Tensor):
            return out.to(device=y_proba.device, dtype=y_proba.dtype if dtype is None else dtype)
        res = out.cpu().numpy()
        return res if dtype is None else res.astype(dtype)
    pred = torch.full((n, k), -1, dtype=torch.int32, device=device)
    if isinstance(y_proba, csr_matrix):
        from .weighted_prediction import topk_csr_device
        for c in np.unique(choice):
            rows = np.nonzero(choice == c)[0]
            sub = dev.csr_to_device(y_proba[rows], device)
            a = dev.vec_to_device(A[c], device, sub.data.dtype, m, "a")
            b = dev.vec_to_device(B[c], device, sub.data.dtype, m, "b")
            pred[torch.from_numpy(rows).to(device)] = topk_csr_device(sub, k, a, b)[0]
        return dev.compact_to_csr_like(y_proba, pred, out_dtype=dtype)
    from .weighted_prediction import topk_dense_device
    d = dev.dense_to_device(y_proba, device)
    gdt = torch.promote_types(d.torch_dtype, torch.from_numpy(A[:0]).dtype)
    g_code = 0 if gdt == torch.float32 else 1
    for c in np.unique(choice):
        rows = torch.from_numpy(np.nonzero(choice == c)[0].astype(np.int32)).to(device)
        a = dev.vec_to_device(A[c], device, gdt, m, "a")
        b = dev.vec_to_device(B[c], device, gdt, m, "b")
        pred[rows.long()] = topk_dense_device(d, k, a, b, g_code, rows=rows)[0]
    return dev.compact_to_dense_like(y_proba, pred, m, out_dtype=dtype)


def find_classifier_using_fw(
    y_true: Matrix,
    y_proba: Matrix,
    metric_func: Callable,
    k: int,
    max_iters: int = 100,
    init_classifier: Union[str, Tuple[DenseMatrix, DenseMatrix]] = "top",
    maximize: bool = True,
    normalize_conf_matrix: bool = True,
    metric_kwargs: Optional[Dict[str, Any]] = None,
    tolerance: float = 1e-6,
    search_for_best_alpha: bool = True,
    alpha_search_algo: str = "uniform",
    alpha_tolerance: float = 0.001,
    alpha_uniform_search_step: float = 0.0001,
    skip_tn: bool = False,
    seed: Optional[int] = None,
    verbose: bool = False,
    return_meta: bool = False,
    **kwargs,
) -> Union[RandomizedWeightedClassifier, Tuple[RandomizedWeightedClassifier, Dict[str, Any]]]:
    """Frank-Wolfe over the confusion-matrix polytope; same arguments / defaults / return value / ``meta`` keys as
    xcolumns/frank_wolfe.py:407-690.  Built-in macro / micro / mixed objectives run the fused per-label kernels; any
    other differentiable callable of (tp, fp, fn, tn) keeps the streaming kernels and takes its gradient and line
    search through torch on the device (generic_fw.py).
    ``distributed=True``: y_true / y_proba are this rank's row shard; the per-iterate confusion sums
    are all-reduced and every rank ends with the same classifier."""
    distributed = kwargs.pop("distributed", False)
    log_info("Starting searching for optimal randomized classifier using Frank-Wolfe algorithm ...", verbose)
    if type(y_true) != type(y_proba) and isinstance(y_true, (np.ndarray, torch.Tensor, csr_matrix)):
        raise ValueError(
            f"y_true and y_proba have unsupported combination of types {type(y_true)} and {type(y_proba)}, should be both np.ndarray, both torch.Tensor, or both csr_matrix")
    if tuple(y_true.shape) != tuple(y_proba.shape):
        raise ValueError(f"y_true and y_proba must have the same shape, got {y_true.shape} and {y_proba.shape}")
    _check_k(k)
    if k < 0:
        raise ValueError("k must be >= 0")
    if alpha_search_algo not in ("uniform", "ternary") and search_for_best_alpha:
        raise ValueError(f"Unknown search algorithm {alpha_search_algo}")
    ternary_eps = float(alpha_tolerance) if (search_for_best_alpha and alpha_search_algo == "ternary") else 0.0
    if ternary_eps < 0 or (search_for_best_alpha and alpha_search_algo == "ternary" and not ternary_eps > 0):
        raise ValueError("alpha_search_algo='ternary' needs alpha_tolerance > 0 (it is the search's epsilon)")
    try:
        metric_id, beta, eps = M.resolve_macro_metric(metric_func, metric_kwargs)
        generic = False
    except M.UnsupportedMetricError:
        # any other differentiable callable of (tp, fp, fn, tn): same streaming kernels, gradient and line search
        # through torch on the device (generic_fw.py; frank_wolfe.py:368-376 differentiates with autograd)
        if not callable(metric_func):
            raise
        metric_id, beta, eps, generic = 0, 1.0, 1e-9, True
    n, m = y_proba.shape
    device = dev.pick_device(y_proba, y_true)
    comm = make_comm(distributed, device)
    ctx = dev.ctx_for(device)
    sp = lambda: dev.stream_ptr(device)
    n_global = comm.n_global(n)

    # ---- classifiers (frank_wolfe.py:500-540), kept on the host as float32 like the reference
    rng = np.random.default_rng(seed)
    A = np.zeros((max_iters + 1, m), dtype=DefaultDataDType)
    B = np.zeros((max_iters + 1, m), dtype=DefaultDataDType)
    P = np.ones(max_iters + 1, dtype=DefaultDataDType)
    if isinstance(init_classifier, str) and init_classifier == "top":
        A[0] = 1.0
        B[0] = -0.5
    elif isinstance(init_classifier, str) and init_classifier == "random":
        A[0] = rng.random(m)
        B[0] = rng.random(m) - 0.5
        if comm.world > 1:   # one classifier for the whole job: rank 0's draw (seed=None draws differ per rank)
            ab = torch.from_numpy(np.stack([A[0], B[0]])).to(device)
            torch.distributed.broadcast(ab, src=torch.distributed.get_global_rank(comm.group, 0), group=comm.group)
            A[0], B[0] = ab.cpu().numpy()
    elif isinstance(init_classifier, str) and init_classifier == "prior":
        freq = np.asarray(y_true.sum(axis=0), dtype=DefaultDataDType).flatten() if not isinstance(
            y_true, torch.Tensor) else y_true.sum(0).float().cpu().numpy()
        if comm.world > 1:   # label frequencies of ALL rows, not of this rank's shard
            fq = torch.from_numpy(np.ascontiguousarray(freq, dtype=np.float64)).to(device)
            comm.allreduce_sum_(fq)
            freq = fq.cpu().numpy().astype(DefaultDataDType)
        A[0] = 1.0 / ((freq + 0.1) / n_global)
        B[0] = 0
    elif (isinstance(init_classifier, (tuple, list)) and len(init_classifier) == 2
          and all(isinstance(v, (np.ndarray, torch.Tensor)) and tuple(v.shape) == (m,) for v in init_classifier)):
        A[0], B[0] = (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v for v in init_classifier)
    else:
        raise ValueError(
            "Unsupported type of init_classifier, it should be in ['random', 'top'], or a tuple of two np.ndarray or torch.Tensor of shape (y_true.shape[1], )")

    # ---- inputs on the device
    is_csr = isinstance(y_proba, csr_matrix)
    if is_csr:
        pd_ = dev.csr_to_device(y_proba, device)
        td_ = dev.csr_to_device(y_true, device, np.dtype(np.float32 if pd_.code == 0 else np.float64))
        wdt = pd_.data.dtype
    else:
        pd_ = dev.dense_to_device(y_proba, device)
        same = y_true is y_proba
        td_ = pd_ if same else dev.dense_to_device(
            y_true if not isinstance(y_true, torch.Tensor) else y_true, device, pd_.torch_dtype)
        wdt = pd_.torch_dtype
    f64 = dict(dtype=torch.float64, device=device)
    colsum = torch.empty(m, **f64)
    if is_csr:
        ctx.call("xc_colsum_csr", dev.ptr(td_.data), td_.code, dev.ptr(td_.indices), int(td_.data.numel()), m,
                 dev.ptr(colsum), sp())
    else:
        ctx.call("xc_colsum_dense", dev.ptr(td_.t), td_.code, n, m, td_.ld, dev.ptr(colsum), sp())
    comm.allreduce_sum_(colsum)

    if generic:
        from . import generic_fw
        conf = generic_fw.device_conf(ctx=ctx, comm=comm, device=device, pd_=pd_, td_=td_, is_csr=is_csr, colsum=colsum,
                                      n=n, m=m, n_global=n_global, k=k, normalize_conf_matrix=normalize_conf_matrix,
                                      skip_tn=skip_tn)
        A_used, B_used, P, meta = generic_fw.run(
            conf=conf, device=device, m=m, A=A, B=B, P=P, max_iters=max_iters, maximize=maximize,
            metric_func=metric_func, metric_kwargs=metric_kwargs, tolerance=tolerance,
            search_for_best_alpha=search_for_best_alpha, alpha_search_algo=alpha_search_algo,
            alpha_tolerance=alpha_tolerance, alpha_uniform_search_step=alpha_uniform_search_step, verbose=verbose)
        if isinstance(y_true, torch.Tensor):
            A, B = (v.to(device=y_proba.device, dtype=y_proba.dtype, copy=True) for v in (A_used, B_used))
            P = torch.tensor(P, dtype=y_proba.dtype, device=y_proba.device)
        else:
            A, B = A_used.cpu().numpy(), B_used.cpu().numpy()
        log_info(f"  Final utility of the randomized classifier: {meta.pop('final_utility')}, "
                 f"number of sub-classifiers: {len(A)}", verbose)
        clf = RandomizedWeightedClassifier(k, A, B, P)
        if return_meta:
            meta["time"] = time() - meta["time"]
            meta["launches"] = ctx.launches()
            return clf, meta
        return clf

    mix = M.resolve_mix(metric_func)
    mix_alpha, mix_k, mix_m = mix if mix is not None else (1.0, 1.0, 1.0)
    mix_code = 2 if M.is_micro_metric(metric_func) else int(mix is not None)
    if isinstance(metric_func, M.MixedMacroRecallPrecisionMetric):
        mix_code, mix_alpha, mix_m = 3, float(metric_func.alpha), float(metric_func.m)
    params = MetricParams(metric=metric_id, maximize=int(bool(maximize)), skip_tn=int(bool(skip_tn)),
                          mix=mix_code, c1=float(1 + beta**2), beta2=float(beta**2), eps=float(eps),
                          n_div=1.0, n_rows=float(n_global), mix_alpha=mix_alpha, mix_k=mix_k, mix_m=mix_m)
    Cm = torch.empty(4 * m, **f64)       # running confusion vectors [tp, fp, fn, tn]
    Ci = torch.empty(4 * m, **f64)       # confusion vectors of the newest classifier
    raw = torch.empty((2, m), **f64)     # tp_raw, cnt of one iterate
    scal_all = torch.zeros((max_iters + 2, 8), **f64)   # per iteration: [old_u, u_i, alpha, best_val, new_u, a, b]
    host_all = torch.zeros((max_iters + 2, 8), dtype=torch.float64).pin_memory()
    events = {}
    alphas = np.arange(0 + alpha_uniform_search_step, 1, alpha_uniform_search_step)  # utils.py:179
    alphas_dev = torch.from_numpy(alphas).to(device)
    vals_dev = torch.empty((int(ctx.lib.xc_fw_alpha_scratch_bytes(m, int(alphas.size))) + 7) // 8, **f64)
    debug = bool(os.environ.get("XCOLUMNS_B200_FW_DEBUG"))   # diagnostics: meta["alpha_search_ctl"] (syncs!)
    debug_ctl = os.environ.get("XCOLUMNS_B200_FW_DEBUG") == "2"   # also read the search control block per step
    dbg_ctl = []
    ctl_off = int(ctx.lib.xc_fw_alpha_ctl_offset(m, int(alphas.size)))

    # -------- two C calls per iteration (begin: streaming pass, finish: everything per label); the classifier
    # matrices live on the device, rows padded so that every row starts 16-byte aligned
    ldc = (m + 3) // 4 * 4
    A_dev = torch.zeros((max_iters + 1, ldc), dtype=torch.float32, device=device)
    B_dev = torch.zeros((max_iters + 1, ldc), dtype=torch.float32, device=device)
    A_dev[0, :m].copy_(torch.from_numpy(A[0]))
    B_dev[0, :m].copy_(torch.from_numpy(B[0]))
    code = pd_.code
    ab64 = torch.empty(2 * (m + 1), **f64) if (code == 1 and not is_csr) else None
    rowp = lambda t, i: C.c_void_p(t.data_ptr() + 4 * ldc * i)
    n_alphas = int(alphas.size)

    def step(i, first):
        """enqueue iteration i (no host sync); its scalars land in scal_all[i] / host_all[i].  The
        finish call also writes classifier row i + 1 (gradient at the new running matrix) and
        the next iteration's "old utility" into scal_all[i + 1][0]."""
        sc = C.c_void_p(scal_all[i].data_ptr())
        has_next = i + 1 <= max_iters
        if is_csr:
            # the CSR kernel takes the classifier in the data dtype (weighted_prediction.py:72-75)
            aw, bw = A_dev[i, :m].to(wdt), B_dev[i, :m].to(wdt)
            ctx.call("xc_fw_iterate_csr", dev.ptr(pd_.data), code, dev.ptr(pd_.indices), dev.ptr(pd_.indptr), n, m,
                     dev.ptr(td_.data), dev.ptr(td_.indices), dev.ptr(td_.indptr), dev.ptr(aw), dev.ptr(bw), k,
                     C.c_void_p(raw[0].data_ptr()), C.c_void_p(raw[1].data_ptr()), None, sp())
        else:
            ctx.call("xc_fw_step_begin", dev.ptr(pd_.t), code, n, m, pd_.ld, dev.ptr(td_.t), td_.ld,
                     rowp(A_dev, i), rowp(B_dev, i), dev.ptr(ab64), k, dev.ptr(raw), 0 if first else 1, sp())
        comm.allreduce_sum_(raw)
        ctx.call("xc_fw_step_finish", C.byref(params), 1 if first else 0, dev.ptr(raw), dev.ptr(colsum), m,
                 C.c_double(float(n_global)), int(bool(normalize_conf_matrix)), int(bool(skip_tn)), dev.ptr(Cm),
                 dev.ptr(Ci), dev.ptr(alphas_dev) if search_for_best_alpha else None, n_alphas,
                 C.c_double(2 / (i + 1)), dev.ptr(vals_dev), sc,
                 rowp(A_dev, i + 1) if has_next else None, rowp(B_dev, i + 1) if has_next else None,
                 C.c_void_p(scal_all[i + 1].data_ptr()), 1, C.c_double(ternary_eps), sp())
        host_all[i].copy_(scal_all[i], non_blocking=True)
        if debug_ctl:
            w = vals_dev.view(torch.int32)[ctl_off // 4: ctl_off // 4 + 32].cpu().tolist()
            dbg_ctl.append((w[16], w[0], w[2]))   # stage-1 candidates, after refinement, label slices
        ev = torch.cuda.Event(enable_timing=debug)
        ev.record(torch.cuda.current_stream(device))
        events[i] = ev

    step(0, True)
    utility_i = float(scal_all[0, 0].item())
    meta: Dict[str, Any] = {"alphas": [], "classifiers_utilities": [utility_i], "utilities": [utility_i], "time": time()}
    log_info(f"    Metric value of the first (sub)classifier 0: {utility_i}", verbose)

    new_utility = utility_i
    i = 0
    n_used = max_iters + 1
    next_step = 1
    depth = max(1, int(os.environ.get("XCOLUMNS_B200_FW_SPECULATE", "3")))
    for i in range(1, max_iters + 1):
        log_info(f"  Starting iteration {i}/{max_iters} ...", verbose)
        # The stopping rules need iteration i's scalars on the host, which would drain the GPU
        # queue once per iteration (measured: 0.5 ms of idle GPU per 0.65 ms iteration).  So up to
        # `depth` iterations are enqueued speculatively BEFORE iteration i's scalars are read; if i
        # turns out to be the last one, the speculative classifier rows are simply truncated like
        # the reference truncates its arrays (frank_wolfe.py:659-661).  A depth above 1 matters
        # when rows are sharded: every iteration ends in an exchange all ranks must have enqueued,
        # so one rank's host hiccup would otherwise stall all GPUs (8 GPUs, depth 1: 1.04 ms per
        # iteration instead of 0.42).
        while next_step <= min(max_iters, i + depth):
            step(next_step, False)
            next_step += 1
        events[i].synchronize()
        host = host_all[i].numpy()
        old_utility, utility_i, alpha, new_utility = float(host[0]), float(host[1]), float(host[2]), float(host[4])
        log_info(f"    Iteration {i}/{max_iters} finished, alpha: {alpha}, metric: {old_utility} -> {new_utility}", verbose)
        if alpha < alpha_tolerance or (maximize and new_utility - old_utility < tolerance) or (
                not maximize and old_utility - new_utility < tolerance):
            n_used = i                                   # :659-661
            break
        meta["alphas"].append(alpha)
        meta["classifiers_utilities"].append(utility_i)
        meta["utilities"].append(new_utility)
        P[:i] *= 1 - alpha
        P[i] = alpha
    P = P[:n_used]
    if isinstance(y_true, torch.Tensor):
        # tensors in -> tensors out on the caller's device; dense classifiers never leave the GPU
        A, B = (v.to(device=y_proba.device, dtype=y_proba.dtype, copy=True) for v in (A_dev[:n_used, :m], B_dev[:n_used, :m]))
        P = torch.tensor(P, dtype=y_proba.dtype, device=y_proba.device)
    else:
        # one pinned staging buffer, one async copy, one sync (pageable copies cost ~1 ms each here)
        stage = torch.empty((2, n_used, m), dtype=torch.float32).pin_memory()
        stage[0].copy_(A_dev[:n_used, :m], non_blocking=True)
        stage[1].copy_(B_dev[:n_used, :m], non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
        A, B = stage[0].numpy(), stage[1].numpy()
    log_info(f"  Final utility of the randomized classifier: {new_utility}, number of sub-classifiers: {len(A)}", verbose)

    clf = RandomizedWeightedClassifier(k, A, B, P)
    if return_meta:
        meta["time"] = time() - meta["time"]
        meta["iters"] = i
        meta["launches"] = ctx.launches()
        if debug:
            meta["alpha_search_ctl"] = dbg_ctl   # per enqueued iteration: (stage-1 candidates, refined, slices)
            torch.cuda.synchronize(device)
            ks = sorted(events)
            meta["step_ms"] = [events[a_].elapsed_time(events[b_]) for a_, b_ in zip(ks[:-1], ks[1:])]
        return clf, meta
    return clf


def make_frank_wolfe_wrapper(metric_func: Callable, metric_name: str, maximize: bool = True, skip_tn: bool = False,
                             warn_k_eq_0: bool = False):
    """Factory of ``find_classifier_optimizing_<metric>_using_fw(y_true, y_proba, k, **kwargs)``
    (xcolumns/frank_wolfe.py:698-748)."""

    def find_classifier_for_metric_using_fw(y_true: Matrix, y_proba: Matrix, k: int, **kwargs):
        if warn_k_eq_0 and k == 0:
            log_warning(f"Warning: k=0 results in degenerated solution for {metric_name}!")
        return find_classifier_using_fw(y_true, y_proba, metric_func, k, maximize=maximize, skip_tn=skip_tn, **kwargs)

    find_classifier_for_metric_using_fw.__doc__ = (
        f"Find a randomized classifier that {'maximizes' if maximize else 'minimizes'} {metric_name} with the "
        f"Frank-Wolfe algorithm; see find_classifier_using_fw.")
    return add_kwargs_to_signature(find_classifier_for_metric_using_fw, find_classifier_using_fw,
                                   skip=["metric_func", "maximize", "skip_tn"])


find_classifier_optimizing_macro_precision_using_fw = make_frank_wolfe_wrapper(
    M.macro_precision_on_conf_matrix, "macro-averaged precision", skip_tn=True, warn_k_eq_0=True)
find_classifier_optimizing_macro_recall_using_fw = make_frank_wolfe_wrapper(
    M.macro_recall_on_conf_matrix, "macro-averaged recall", skip_tn=True, warn_k_eq_0=True)
find_classifier_optimizing_macro_f1_score_using_fw = make_frank_wolfe_wrapper(
    M.macro_f1_score_on_conf_matrix, "macro-averaged F1 score", skip_tn=True)
find_classifier_optimizing_macro_jaccard_score_using_fw = make_frank_wolfe_wrapper(
    M.macro_jaccard_score_on_conf_matrix, "macro-averaged Jaccard score", skip_tn=True)
find_classifier_optimizing_macro_balanced_accuracy_using_fw = make_frank_wolfe_wrapper(
    M.macro_balanced_accuracy_on_conf_matrix, "macro-averaged balanced accuracy")
find_classifier_optimizing_macro_hmean_using_fw = make_frank_wolfe_wrapper(M.macro_hmean_on_conf_matrix, "macro-averaged H-mean")
find_classifier_optimizing_macro_gmean_using_fw = make_frank_wolfe_wrapper(M.macro_gmean_on_conf_matrix, "macro-averaged G-mean")
# micro-averaged objectives (xcolumns/frank_wolfe.py:758-832)
find_classifier_optimizing_micro_precision_using_fw = make_frank_wolfe_wrapper(
    M.micro_precision_on_conf_matrix, "micro-averaged precision", skip_tn=True, warn_k_eq_0=True)
find_classifier_optimizing_micro_recall_using_fw = make_frank_wolfe_wrapper(
    M.micro_recall_on_conf_matrix, "micro-averaged recall", skip_tn=True, warn_k_eq_0=True)
find_classifier_optimizing_micro_f1_score_using_fw = make_frank_wolfe_wrapper(
    M.micro_f1_score_on_conf_matrix, "micro-averaged F1 score", skip_tn=True)
find_classifier_optimizing_micro_jaccard_score_using_fw = make_frank_wolfe_wrapper(
    M.micro_jaccard_score_on_conf_matrix, "micro-averaged Jaccard score", skip_tn=True)
find_classifier_optimizing_micro_balanced_accuracy_using_fw = make_frank_wolfe_wrapper(
    M.micro_balanced_accuracy_on_conf_matrix, "micro-averaged balanced accuracy")
find_classifier_optimizing_micro_hmean_using_fw = make_frank_wolfe_wrapper(M.micro_hmean_on_conf_matrix, "micro-averaged H-mean")
find_classifier_optimizing_micro_gmean_using_fw = make_frank_wolfe_wrapper(M.micro_gmean_on_conf_matrix, "micro-averaged G-mean")


# ------------------------------------------------------------------------------------------
# mixed instance-precision / macro-metric objectives (frank_wolfe.py:838-915)
# ------------------------------------------------------------------------------------------

def make_mixed_frank_wolfe_wrapper(binary_metric_func: Callable, metric_name: str):
    """``find_classifier_optimizing_mixed_instance_precision_and_<metric>_using_fw(y_true, y_proba, k,
    alpha=1, **kwargs)``: Frank-Wolfe on  sum_j [(1 - alpha) * tp_j / k + alpha * metric_j / m].  The line
    search of these objectives runs the exhaustive float64 grid (the two-stage screen covers the pure
    c*tp/D metrics only)."""

    def find_classifier_optimizing_mixed_metric_using_fw(y_true: Matrix, y_proba: Matrix, k: int, alpha: float = 1,
                                                         **kwargs):
        n, m = y_true.shape
        return find_classifier_using_fw(y_true, y_proba, M.MixedInstancePrecisionMacroMetric(binary_metric_func, alpha, k, m),
                                        k, **kwargs)

    find_classifier_optimizing_mixed_metric_using_fw.__doc__ = (
        f"Find a randomized classifier maximizing a weighted average of instance precision@k and {metric_name} "
        f"with the Frank-Wolfe algorithm; see find_classifier_using_fw.")
    return find_classifier_optimizing_mixed_metric_using_fw


find_classifier_optimizing_mixed_instance_precision_and_macro_precision_using_fw = make_mixed_frank_wolfe_wrapper(
    M.binary_precision_on_conf_matrix, "macro-averaged precision")
find_classifier_optimizing_mixed_instance_precision_and_macro_f1_score_using_fw = make_mixed_frank_wolfe_wrapper(
    M.binary_f1_score_on_conf_matrix, "macro-averaged F1 score")
find_classifier_optimizing_mixed_instance_precision_and_macro_recall_using_fw = make_mixed_frank_wolfe_wrapper(
    M.binary_recall_on_conf_matrix, "macro-averaged recall")


def find_classifier_optimizing_mixed_macro_recall_and_macro_precision_using_fw(y_true: Matrix, y_proba: Matrix, k: int,
                                                                               alpha: float = 1, **kwargs):
    """Find a randomized classifier maximizing  sum_j [(1 - alpha) * recall_j + alpha * precision_j]  with the
    Frank-Wolfe algorithm (xcolumns/frank_wolfe.py:917-938); see find_classifier_using_fw."""
    n, m = y_true.shape
    return find_classifier_using_fw(y_true, y_proba, M.MixedMacroRecallPrecisionMetric(alpha, m), k, **kwargs)
