"""Frank-Wolfe for objectives the fused kernels do not know (xcolumns/frank_wolfe.py:368-376, :589-637:
``metric_func`` may be ANY differentiable callable of the confusion vectors (tp, fp, fn, tn)).

The n x m work of an iteration -- weighted top-k of every row fused with the accumulation of tp / count -- is the same
streaming kernel the built-in objectives use (``xc_fw_step_begin`` / ``xc_fw_iterate_csr``).  What a user-supplied
Python callable cannot have is the fused per-label finish (closed-form gradient, screened line search); here that part
runs ON the device through torch, like the reference's own torch flavour does (frank_wolfe.py:18-40):

* the callable is handed float64 CUDA tensors, its gradient comes from ``torch.autograd.grad`` (the reference:
  ``autograd.grad`` on float32 / float64 numpy vectors, the same derivative up to rounding);
* the uniform line search (utils.py:174-184) evaluates the callable on the whole alpha grid with ``torch.vmap`` in
  chunks (one batched evaluation instead of 10^4 Python calls); callables vmap cannot trace (data-dependent control
  flow, ``.item()``) are evaluated point by point.  ``torch.argmax`` returns the FIRST maximum, which is what the
  reference's strict ``score > best_val`` scan keeps;
* nothing but the scalars of the stopping rules travels to the host.

The callable must be written with arithmetic operators / tensor methods / torch functions (numpy ufuncs cannot run on
CUDA tensors) -- the reference's own metrics are.
"""
from __future__ import annotations

import ctypes as C
from time import time
from typing import Any, Callable, Dict, Optional

import numpy as np
import torch

from . import _device as dev
from .utils import log_info

_ALPHA_CHUNK_BYTES = 256 << 20   # memory of one vmapped line-search chunk (4 stacked float64 vectors per alpha)


class _Objective:
    """metric_func(tp, fp, fn, tn, **metric_kwargs) on views of a stacked [4, m] float64 device tensor."""

    def __init__(self, func: Callable, kwargs: Optional[Dict[str, Any]]):
        self.func = func
        self.kwargs = dict(kwargs or {})
        self.vmap_ok = True

    def __call__(self, c4: torch.Tensor):
        out = self.func(c4[0], c4[1], c4[2], c4[3], **self.kwargs)
        if not isinstance(out, torch.Tensor):
            raise ValueError(
                f"metric_func must return a scalar tensor computed from its arguments (it is evaluated on float64 CUDA "
                f"tensors), but returned {type(out)}")
        if out.numel() != 1:
            raise ValueError(f"metric_func must return a scalar, but returned a tensor of shape {tuple(out.shape)}")
        return out.reshape(())

    def value(self, c4: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            return self(c4)

    def value_and_grad(self, c4: torch.Tensor):
        """(value, d value / d [tp, fp, fn, tn]) -- frank_wolfe.py:18-40 (materialize_grads, allow_unused)."""
        parts = [c4[r].detach().clone().requires_grad_(True) for r in range(4)]
        val = self.func(*parts, **self.kwargs)
        if not isinstance(val, torch.Tensor) or val.numel() != 1:
            raise ValueError("metric_func must return a scalar tensor computed from its arguments")
        if not val.requires_grad:   # a constant objective: zero gradient
            return val.detach().reshape(()), [torch.zeros_like(p) for p in parts]
        grads = torch.autograd.grad(val.reshape(()), parts, allow_unused=True, materialize_grads=True)
        return val.detach().reshape(()), grads

    def on_grid(self, cm: torch.Tensor, ci: torch.Tensor, alphas: torch.Tensor) -> torch.Tensor:
        """values of metric((1 - alpha) C + alpha C_i) for every alpha of the grid (frank_wolfe.py:393-398)"""
        m4 = cm.numel()
        chunk = max(1, int(_ALPHA_CHUNK_BYTES // (8 * m4)))
        out = torch.empty(alphas.numel(), dtype=torch.float64, device=cm.device)
        with torch.no_grad():
            for s in range(0, alphas.numel(), chunk):
                al = alphas[s:s + chunk]
                if self.vmap_ok:
                    try:
                        comb = (1 - al)[:, None, None] * cm[None] + al[:, None, None] * ci[None]
                        out[s:s + chunk] = torch.vmap(self)(comb).to(torch.float64)
                        continue
                    except ValueError:
                        raise
                    except Exception:   # not traceable by vmap: point by point from here on
                        self.vmap_ok = False
                for q in range(al.numel()):
                    a_ = al[q]
                    out[s + q] = self((1 - a_) * cm + a_ * ci)
        return out


def device_conf(*, ctx, comm, device, pd_, td_, is_csr: bool, colsum: torch.Tensor, n: int, m: int, n_global: int,
                k: int, normalize_conf_matrix: bool, skip_tn: bool):
    """conf(i, A_dev, B_dev) -> confusion vectors [tp, fp, fn, tn] ([4, m] float64 on the device) of classifier row i
    over all rows of all ranks: the fused streaming kernel + ``xc_fw_make_conf`` (frank_wolfe.py:601-606)."""
    sp = lambda: dev.stream_ptr(device)
    f64 = dict(dtype=torch.float64, device=device)
    code = pd_.code
    raw = torch.empty((2, m), **f64)
    ab64 = torch.empty(2 * (m + 1), **f64) if (code == 1 and not is_csr) else None
    wdt = pd_.data.dtype if is_csr else pd_.torch_dtype

    def conf(i: int, A_dev: torch.Tensor, B_dev: torch.Tensor) -> torch.Tensor:
        if is_csr:
            aw, bw = A_dev[i, :m].to(wdt), B_dev[i, :m].to(wdt)
            ctx.call("xc_fw_iterate_csr", dev.ptr(pd_.data), code, dev.ptr(pd_.indices), dev.ptr(pd_.indptr), n, m,
                     dev.ptr(td_.data), dev.ptr(td_.indices), dev.ptr(td_.indptr), dev.ptr(aw), dev.ptr(bw), k,
                     C.c_void_p(raw[0].data_ptr()), C.c_void_p(raw[1].data_ptr()), None, sp())
        else:
            row = lambda t: C.c_void_p(t.data_ptr() + 4 * t.stride(0) * i)
            ctx.call("xc_fw_step_begin", dev.ptr(pd_.t), code, n, m, pd_.ld, dev.ptr(td_.t), td_.ld,
                     row(A_dev), row(B_dev), dev.ptr(ab64), k, dev.ptr(raw), 0, sp())
        comm.allreduce_sum_(raw)
        c4 = torch.empty(4 * m, **f64)
        ctx.call("xc_fw_make_conf", C.c_void_p(raw[0].data_ptr()), C.c_void_p(raw[1].data_ptr()), dev.ptr(colsum), m,
                 C.c_double(float(n_global)), int(bool(normalize_conf_matrix)), int(bool(skip_tn)), dev.ptr(c4), sp())
        return c4.view(4, m)

    return conf


def run(*, conf: Callable, device, m: int, A: np.ndarray, B: np.ndarray, P: np.ndarray, max_iters: int, maximize: bool,
        metric_func: Callable, metric_kwargs, tolerance: float, search_for_best_alpha: bool, alpha_search_algo: str,
        alpha_tolerance: float, alpha_uniform_search_step: float, verbose: bool):
    """The loop of frank_wolfe.py:565-670 for an arbitrary objective.  ``conf(i, A_dev, B_dev)`` returns the confusion
    vectors of classifier row i ([4, m] float64 on ``device``).  A / B / P are the host classifier arrays with row 0
    initialised; returns (A_dev [n_used, m] float32, B_dev, P [n_used], meta)."""
    obj = _Objective(metric_func, metric_kwargs)
    ldc = (m + 3) // 4 * 4     # classifier rows start 16-byte aligned (128-bit loads in the streaming kernel)
    A_dev = torch.zeros((max_iters + 1, ldc), dtype=torch.float32, device=device)
    B_dev = torch.zeros((max_iters + 1, ldc), dtype=torch.float32, device=device)
    A_dev[0, :m].copy_(torch.from_numpy(np.ascontiguousarray(A[0])))
    B_dev[0, :m].copy_(torch.from_numpy(np.ascontiguousarray(B[0])))
    alphas = np.arange(0 + alpha_uniform_search_step, 1, alpha_uniform_search_step)   # utils.py:179
    grid = torch.from_numpy(np.concatenate([[0.0], alphas])).to(device)              # best starts at low = 0
    conf_ = conf
    conf = lambda i: conf_(i, A_dev, B_dev)

    def best_alpha(cm, ci) -> float:
        if alpha_search_algo == "uniform":
            vals = obj.on_grid(cm, ci, grid)
            return float(grid[int(torch.argmax(vals).item())].item())
        low, high = 0, 1                                   # utils.py:187-201
        f = lambda a_: float(obj.value((1 - a_) * cm + a_ * ci).item())
        while high - low > alpha_tolerance:
            mid1 = low + (high - low) / 3
            mid2 = high - (high - low) / 3
            if f(mid1) < f(mid2):
                high = mid2
            else:
                low = mid1
        return (low + high) / 2

    cm = conf(0)
    utility_i = float(obj.value(cm).item())
    meta: Dict[str, Any] = {"alphas": [], "classifiers_utilities": [utility_i], "utilities": [utility_i], "time": time()}
    log_info(f"    Metric value of the first (sub)classifier 0: {utility_i}", verbose)
    new_utility = utility_i
    n_used = max_iters + 1
    i = 0
    for i in range(1, max_iters + 1):
        log_info(f"  Starting iteration {i}/{max_iters} ...", verbose)
        old_u, (gtp, gfp, gfn, gtn) = obj.value_and_grad(cm)
        a_i = gtp - gfp - gfn + gtn                        # frank_wolfe.py:595-596
        b_i = gfp - gtn
        if not maximize:
            a_i, b_i = -a_i, -b_i
        A_dev[i, :m].copy_(a_i)                            # float32 store, like the reference's classifier arrays
        B_dev[i, :m].copy_(b_i)
        ci = conf(i)
        utility_i = float(obj.value(ci).item())
        log_info(f"    Metric value of new (sub)classifier {i}: {utility_i}", verbose)
        alpha = best_alpha(cm, ci) if search_for_best_alpha else 2 / (i + 1)
        cm = (1 - alpha) * cm + alpha * ci
        old_utility, new_utility = float(old_u.item()), float(obj.value(cm).item())
        log_info(f"    Iteration {i}/{max_iters} finished, alpha: {alpha}, metric: {old_utility} -> {new_utility}", verbose)
        if alpha < alpha_tolerance or (maximize and new_utility - old_utility < tolerance) or (
                not maximize and old_utility - new_utility < tolerance):
            n_used = i                                     # :659-661
            break
        meta["alphas"].append(alpha)
        meta["classifiers_utilities"].append(utility_i)
        meta["utilities"].append(new_utility)
        P[:i] *= 1 - alpha
        P[i] = alpha
    meta["iters"] = i
    meta["final_utility"] = new_utility
    return A_dev[:n_used, :m], B_dev[:n_used, :m], P[:n_used], meta
