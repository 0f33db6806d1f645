"""Online / greedy prediction with a device-resident confusion matrix.

The reference's online experiments (experiments/omma_wrappers_online_methods.py:192-270) call, per arriving
instance, ``_bc_with_0approx_step_dense(..., greedy=True, only_pred=True)`` -- predict the k labels with the
largest marginal gain against the confusion matrix observed so far -- and then
``_update_unnormalized_confusion_matrix(C, y_true[i], y_pred[i])`` (xcolumns/block_coordinate.py:132-209,
xcolumns/confusion_matrix.py:402-435).  Here a whole micro-batch of instances runs through ONE kernel
launch (the sequential-exact cluster kernel, csrc/bca_exact.cu) with the state kept in registers / HBM
between calls; the arithmetic is the reference's (float64 state, float32 products), so predictions and
state are bit-equal to the reference's step functions driven with a float64 confusion matrix.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Callable, Dict, Optional, Sequence, Union

import numpy as np
import torch
from scipy.sparse import csr_matrix

from . import _device as dev
from . import metrics as M
from .block_coordinate import _metric_params
from .confusion_matrix import ConfusionMatrix
from .types import Matrix
from .weighted_prediction import _check_k


class OnlineGreedy:
    """Greedy online maximisation of a label-wise utility at k.

    Args mirror the reference's ``OnlineGreedy`` (m, k, binary_utility_func, skip_tn, etu_variant,
    initial_confusion_matrix) plus ``metric_kwargs`` / ``maximize`` of the step function and the device.
    ``etu_variant=True`` advances the state with the probabilities instead of the true labels."""

    def __init__(self, m: int, k: int, binary_utility_func: Callable, skip_tn: bool = False, etu_variant: bool = False,
                 initial_confusion_matrix: Sequence[float] = (1e-6, 1e-6, 1e-6, 1e-6),
                 metric_kwargs: Optional[Dict[str, Any]] = None, maximize: bool = True,
                 device: Optional[Union[str, torch.device]] = None):
        _check_k(k)
        if k <= 0 or k > m:
            raise ValueError("k must be in 1..m")
        self.m, self.k, self.skip_tn, self.etu_variant, self.maximize = m, k, bool(skip_tn), bool(etu_variant), maximize
        self.metric = binary_utility_func
        self._resolved = M.resolve_binary_metric(binary_utility_func, metric_kwargs)
        self._mix = M.resolve_mix(binary_utility_func)
        self.device = torch.device(device) if device is not None else dev.pick_device()
        self.ctx = dev.ctx_for(self.device)
        self.state = torch.empty((4, m), dtype=torch.float64, device=self.device)
        for i, v in enumerate(initial_confusion_matrix):
            self.state[i].fill_(float(v))
        self.n = float(sum(initial_confusion_matrix))   # like the reference: instances seen (+ the regulariser)
        self._steps = 0                                  # instances processed through the CSR path (lazy tn)
        self._tn_last: Optional[torch.Tensor] = None

    @property
    def C(self) -> ConfusionMatrix:
        """The current confusion matrix (host copy, float64)."""
        tp, fp, fn, tn = self.state.cpu().numpy()
        return ConfusionMatrix(tp, fp, fn, tn)

    def predict_update(self, y_proba: Matrix, y_true: Optional[Matrix] = None, n_div: Optional[int] = None,
                       y_pred_format: str = "same") -> Matrix:
        """Predict the rows of ``y_proba`` one after another, advancing the state after each row with the
        matching row of ``y_true`` (or of ``y_proba`` itself for the ETU variant).  ``n_div``: the divisor of
        the step's normalisation (block_coordinate.py:149-151: the number of rows of the matrix the reference
        step is handed); defaults to ``y_proba.shape[0]``."""
        if isinstance(y_proba, csr_matrix):
            return self._predict_update_csr(y_proba, y_true, n_div, y_pred_format)
        if not isinstance(y_proba, (np.ndarray, torch.Tensor)):
            raise ValueError("y_proba must be np.ndarray, torch.Tensor or csr_matrix")
        if y_proba.shape[1] != self.m:
            raise ValueError(f"y_proba must have {self.m} columns")
        if not self.etu_variant:
            if y_true is None:
                raise ValueError("y_true is required unless etu_variant=True")
            if tuple(y_true.shape) != tuple(y_proba.shape):
                raise ValueError("y_true and y_proba must have the same shape")
        n = y_proba.shape[0]
        d = dev.dense_to_device(y_proba, self.device)
        t = None if self.etu_variant else dev.dense_to_device(y_true, self.device, d.torch_dtype)
        metric_id, beta, eps = self._resolved
        p = _metric_params(metric_id, beta, eps, self.maximize, self.skip_tn, float(n if n_div is None else n_div),
                           mix=self._mix)
        pred = torch.empty((n, self.k), dtype=torch.int32, device=self.device)
        sp = lambda i: C.c_void_p(self.state[i].data_ptr())
        self.ctx.call("xc_bca_online_dense", dev.ptr(d.t), d.code, n, self.m, d.ld, None if t is None else dev.ptr(t.t),
                      0 if t is None else t.ld, self.k, C.byref(p), dev.ptr(pred), sp(0), sp(1), sp(2), sp(3),
                      dev.stream_ptr(self.device))
        self.n += n
        if y_pred_format == "indices":
            return pred if isinstance(y_proba, torch.Tensor) and y_proba.is_cuda else pred.cpu().numpy()
        return dev.compact_to_dense_like(y_proba, pred, self.m)

    def _predict_update_csr(self, y_proba: csr_matrix, y_true: Optional[csr_matrix], n_div: Optional[int],
                            y_pred_format: str):
        """CSR rows (block_coordinate.py:212-293 with greedy=True, only_pred=True, then
        confusion_matrix.py:421-432, as driven by experiments/omma_wrappers_online_methods.py:223-266): candidates are
        the row's stored labels; rows with fewer than k of them keep them all.  Returns a csr_matrix of ones (same
        dtypes as y_proba) or, with y_pred_format="indices", the (n, k) label ids (-1 = unused slot)."""
        if y_proba.shape[1] != self.m:
            raise ValueError(f"y_proba must have {self.m} columns")
        if not self.etu_variant:
            if y_true is None:
                raise ValueError("y_true is required unless etu_variant=True")
            if not isinstance(y_true, csr_matrix) or tuple(y_true.shape) != tuple(y_proba.shape):
                raise ValueError("y_true must be a csr_matrix of y_proba's shape")
        n = y_proba.shape[0]
        c = dev.csr_to_device(y_proba, self.device)
        t = c if self.etu_variant else dev.csr_to_device(y_true, self.device, np.dtype(np.float32 if c.code == 0 else np.float64))
        metric_id, beta, eps = self._resolved
        p = _metric_params(metric_id, beta, eps, self.maximize, self.skip_tn, float(n if n_div is None else n_div),
                           mix=self._mix)
        if self._tn_last is None:
            self._tn_last = torch.zeros(self.m, dtype=torch.int32, device=self.device)
        pred = torch.empty((n, self.k), dtype=torch.int32, device=self.device)
        sp = lambda i: C.c_void_p(self.state[i].data_ptr())
        self.ctx.call("xc_bca_online_csr", dev.ptr(c.data), c.code, dev.ptr(c.indices), dev.ptr(c.indptr), dev.ptr(t.data),
                      dev.ptr(t.indices), dev.ptr(t.indptr), n, self.m, self.k, C.byref(p), dev.ptr(pred), sp(0), sp(1),
                      sp(2), sp(3), dev.ptr(self._tn_last), int(self._steps), dev.stream_ptr(self.device))
        self._steps += n
        self.n += n
        if y_pred_format == "indices":
            return pred.cpu().numpy()
        return dev.compact_to_csr_like(y_proba, pred, reference_padding=False)
