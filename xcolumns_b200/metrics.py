"""Binary / macro metrics on confusion vectors -- the formulas the hot path evaluates
(xcolumns/metrics.py:585-944) -- plus the resolver that maps a metric callable onto the fused
kernels' metric id.

The functions work on numpy arrays, torch tensors and scalars and keep the reference's operation
order, so they are usable exactly like the reference's (``binary_f1_score_on_conf_matrix(tp, fp,
fn, tn, epsilon=...)``).  Inside the CUDA path they are never called per element: the kernels
carry the same expressions (csrc/xc_common.cuh: xc_binary_metric).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional, Tuple

from .utils import add_kwargs_to_signature

XC_METRIC_PRECISION, XC_METRIC_RECALL, XC_METRIC_FBETA, XC_METRIC_JACCARD = 0, 1, 2, 3
XC_METRIC_BALANCED_ACC, XC_METRIC_GMEAN, XC_METRIC_HMEAN, XC_METRIC_PREC_AT_K = 4, 5, 6, 7
AFFINE_GAIN_METRICS = (XC_METRIC_PRECISION, XC_METRIC_RECALL, XC_METRIC_FBETA, XC_METRIC_BALANCED_ACC,
                       XC_METRIC_PREC_AT_K)
TN_METRICS = (XC_METRIC_BALANCED_ACC, XC_METRIC_GMEAN, XC_METRIC_HMEAN)
# gain not affine in eta but a closed form of eta and four per-label numbers (dense batched mode)
RECORD_GAIN_METRICS = (XC_METRIC_JACCARD, XC_METRIC_GMEAN, XC_METRIC_HMEAN)


def binary_precision_on_conf_matrix(tp, fp, fn, tn, epsilon: float = 1e-9):
    return tp / (tp + fp + epsilon)


def binary_recall_on_conf_matrix(tp, fp, fn, tn, epsilon: float = 1e-9):
    return tp / (tp + fn + epsilon)


def binary_fbeta_score_on_conf_matrix(tp, fp, fn, tn, beta: float = 1.0, epsilon: float = 1e-9):
    return (1 + beta**2) * tp / ((beta**2 * (tp + fp)) + tp + fn + epsilon)


def binary_f1_score_on_conf_matrix(tp, fp, fn, tn, epsilon: float = 1e-9):
    return binary_fbeta_score_on_conf_matrix(tp, fp, fn, tn, beta=1.0, epsilon=epsilon)


def binary_jaccard_score_on_conf_matrix(tp, fp, fn, tn, epsilon: float = 1e-9):
    return tp / (tp + fp + fn + epsilon)


def _rates(tp, fp, fn, tn, epsilon):
    return tp / (tp + fn + epsilon), tn / (tn + fp + epsilon)


def binary_balanced_accuracy_on_conf_matrix(tp, fp, fn, tn, epsilon: float = 1e-9):
    tpr, tnr = _rates(tp, fp, fn, tn, epsilon)
    return (tpr + tnr) / 2


def binary_gmean_on_conf_matrix(tp, fp, fn, tn, epsilon: float = 1e-9):
    tpr, tnr = _rates(tp, fp, fn, tn, epsilon)
    return (tpr * tnr) ** 0.5


def binary_hmean_on_conf_matrix(tp, fp, fn, tn, epsilon: float = 1e-9):
    tpr, tnr = _rates(tp, fp, fn, tn, epsilon)
    return (2 * tpr * tnr) / (tpr + tnr)


def binary_precision_at_k_on_conf_matrix(tp, fp, fn, tn, k: int):
    """tp / k (xcolumns/metrics.py:497-513)."""
    return tp / k


class PrecisionAtK:
    """``binary_precision_at_k_on_conf_matrix`` with k bound -- the callable
    predict_optimizing_instance_precision_using_bc builds at block_coordinate.py:822-823, as an
    object the resolver can recognise."""

    def __init__(self, k: int):
        self.k = k
        self.__name__ = "instance_precision_with_specific_k"

    def __call__(self, tp, fp, fn, tn):
        return binary_precision_at_k_on_conf_matrix(tp, fp, fn, tn, self.k)


class MixedInstancePrecisionMetric:
    """The mixed utilities of block_coordinate.py:848-1045:
    ``(1 - alpha) * binary_precision_at_k(tp, k) + alpha * binary_metric(tp, fp, fn, tn, epsilon) / m``
    (evaluated per label and summed by the caller)."""

    def __init__(self, binary_metric: Callable, alpha: float, k: int, m: int):
        self.binary_metric, self.alpha, self.k, self.m = binary_metric, alpha, k, m
        self.__name__ = "mixed_utility_fn"

    def __call__(self, tp, fp, fn, tn, epsilon: float = 1e-9):
        return (1 - self.alpha) * binary_precision_at_k_on_conf_matrix(tp, fp, fn, tn, self.k) + self.alpha * \
            self.binary_metric(tp, fp, fn, tn, epsilon=epsilon) / self.m


class MixedInstancePrecisionMacroMetric:
    """Frank-Wolfe objective of frank_wolfe.py:838-915: the sum over labels of the mixed utility."""

    def __init__(self, binary_metric: Callable, alpha: float, k: int, m: int):
        self.per_label = MixedInstancePrecisionMetric(binary_metric, alpha, k, m)
        self.__name__ = "mixed_metric_fn"

    def __call__(self, tp, fp, fn, tn, epsilon: float = 1e-9):
        return self.per_label(tp, fp, fn, tn, epsilon=epsilon).sum()


class MixedMacroRecallPrecisionMetric:
    """Frank-Wolfe objective of frank_wolfe.py:917-938:
    ``((1 - alpha) * binary_recall + alpha * binary_precision).sum()``."""

    def __init__(self, alpha: float, m: int):
        self.alpha, self.m = alpha, m
        self.__name__ = "mixed_metric_fn"

    def __call__(self, tp, fp, fn, tn, epsilon: float = 1e-9):
        return ((1 - self.alpha) * binary_recall_on_conf_matrix(tp, fp, fn, tn, epsilon=epsilon) + self.alpha *
                binary_precision_on_conf_matrix(tp, fp, fn, tn, epsilon=epsilon)).sum()


_BINARY_IDS = {
    "binary_precision_on_conf_matrix": XC_METRIC_PRECISION,
    "binary_recall_on_conf_matrix": XC_METRIC_RECALL,
    "binary_fbeta_score_on_conf_matrix": XC_METRIC_FBETA,
    "binary_f1_score_on_conf_matrix": XC_METRIC_FBETA,
    "binary_jaccard_score_on_conf_matrix": XC_METRIC_JACCARD,
    "binary_balanced_accuracy_on_conf_matrix": XC_METRIC_BALANCED_ACC,
    "binary_gmean_on_conf_matrix": XC_METRIC_GMEAN,
    "binary_hmean_on_conf_matrix": XC_METRIC_HMEAN,
}
_OWN_MODULES = ("xcolumns_b200.metrics", "xcolumns.metrics")


def make_macro_metric_on_conf_matrix(binary_metric: Callable, name: str) -> Callable:
    """Macro average of a binary metric (xcolumns/metrics.py:38-67)."""

    def macro_metric_on_conf_matrix(tp, fp, fn, tn, **kwargs):
        return binary_metric(tp, fp, fn, tn, **kwargs).mean()

    macro_metric_on_conf_matrix.__doc__ = f"Macro-averaged {name} on the confusion vectors (tp, fp, fn, tn)."
    macro_metric_on_conf_matrix._xc_binary_metric = binary_metric
    return add_kwargs_to_signature(macro_metric_on_conf_matrix, binary_metric)


macro_precision_on_conf_matrix = make_macro_metric_on_conf_matrix(binary_precision_on_conf_matrix, "precision")
macro_recall_on_conf_matrix = make_macro_metric_on_conf_matrix(binary_recall_on_conf_matrix, "recall")
macro_fbeta_score_on_conf_matrix = make_macro_metric_on_conf_matrix(binary_fbeta_score_on_conf_matrix, "F-beta score")
macro_f1_score_on_conf_matrix = make_macro_metric_on_conf_matrix(binary_fbeta_score_on_conf_matrix, "F1 score")
macro_jaccard_score_on_conf_matrix = make_macro_metric_on_conf_matrix(binary_jaccard_score_on_conf_matrix, "Jaccard score")
macro_balanced_accuracy_on_conf_matrix = make_macro_metric_on_conf_matrix(
    binary_balanced_accuracy_on_conf_matrix, "balanced accuracy")
macro_gmean_on_conf_matrix = make_macro_metric_on_conf_matrix(binary_gmean_on_conf_matrix, "G-mean")
macro_hmean_on_conf_matrix = make_macro_metric_on_conf_matrix(binary_hmean_on_conf_matrix, "H-mean")


def make_micro_metric_on_conf_matrix(binary_metric: Callable, name: str) -> Callable:
    """Micro average of a binary metric: the metric of the summed confusion entries
    (xcolumns/metrics.py:68-100)."""

    def micro_metric_on_conf_matrix(tp, fp, fn, tn, **kwargs):
        return binary_metric(tp.sum(), fp.sum(), fn.sum(), tn.sum(), **kwargs)

    micro_metric_on_conf_matrix.__doc__ = f"Micro-averaged {name} on the confusion vectors (tp, fp, fn, tn)."
    micro_metric_on_conf_matrix._xc_binary_metric = binary_metric
    micro_metric_on_conf_matrix._xc_micro = True
    return add_kwargs_to_signature(micro_metric_on_conf_matrix, binary_metric)


micro_precision_on_conf_matrix = make_micro_metric_on_conf_matrix(binary_precision_on_conf_matrix, "precision")
micro_recall_on_conf_matrix = make_micro_metric_on_conf_matrix(binary_recall_on_conf_matrix, "recall")
micro_fbeta_score_on_conf_matrix = make_micro_metric_on_conf_matrix(binary_fbeta_score_on_conf_matrix, "F-beta score")
micro_f1_score_on_conf_matrix = make_micro_metric_on_conf_matrix(binary_fbeta_score_on_conf_matrix, "F1 score")
micro_jaccard_score_on_conf_matrix = make_micro_metric_on_conf_matrix(binary_jaccard_score_on_conf_matrix, "Jaccard score")
micro_balanced_accuracy_on_conf_matrix = make_micro_metric_on_conf_matrix(
    binary_balanced_accuracy_on_conf_matrix, "balanced accuracy")
micro_gmean_on_conf_matrix = make_micro_metric_on_conf_matrix(binary_gmean_on_conf_matrix, "G-mean")
micro_hmean_on_conf_matrix = make_micro_metric_on_conf_matrix(binary_hmean_on_conf_matrix, "H-mean")


def is_micro_metric(func: Callable) -> bool:
    """True for our micro-averaged metrics and for the reference's own factory closures."""
    return bool(getattr(func, "_xc_micro", False)) or getattr(func, "__name__", "") == "micro_metric_on_conf_matrix"


def coverage_on_conf_matrix(tp, fp, fn, tn):
    """Fraction of labels with at least one true positive (xcolumns/metrics.py:972-990)."""
    return (tp > 0).mean()


class UnsupportedMetricError(NotImplementedError):
    pass


def resolve_binary_metric(func: Callable, metric_kwargs: Optional[Dict[str, Any]] = None) -> Tuple[int, float, float]:
    """(metric id, beta, epsilon) of a built-in binary metric callable -- ours or the reference's
    own (matched by module + name).  Other callables cannot run inside the fused kernels."""
    kw = dict(metric_kwargs or {})
    if isinstance(func, PrecisionAtK):
        if kw:
            raise ValueError(f"unknown metric_kwargs for precision@k: {sorted(kw)}")
        return XC_METRIC_PREC_AT_K, float(func.k), 1e-9   # the "beta" slot carries k (see metric_c1)
    if isinstance(func, MixedInstancePrecisionMetric):
        return resolve_binary_metric(func.binary_metric, kw)
    name = getattr(func, "__name__", None)
    mod = getattr(func, "__module__", None)
    if name in _BINARY_IDS and mod in _OWN_MODULES:
        beta = 1.0
        if name == "binary_fbeta_score_on_conf_matrix":
            beta = float(kw.pop("beta", 1.0))
        eps = float(kw.pop("epsilon", 1e-9))
        if kw:
            raise ValueError(f"unknown metric_kwargs for {name}: {sorted(kw)}")
        return _BINARY_IDS[name], beta, eps
    raise UnsupportedMetricError(
        f"binary_metric_func={func!r} is not one of the built-in binary metrics "
        f"({', '.join(sorted(_BINARY_IDS))}); arbitrary Python callables cannot be evaluated inside "
        f"the CUDA sweep and xcolumns_b200 has no CPU fallback")


def resolve_mix(func: Callable) -> Optional[Tuple[float, float, float]]:
    """(alpha, k, m) when func is one of the mixed instance-precision utilities, else None."""
    if isinstance(func, MixedInstancePrecisionMacroMetric):
        func = func.per_label
    if isinstance(func, MixedInstancePrecisionMetric):
        return float(func.alpha), float(func.k), float(func.m)
    return None


def metric_c1_beta2(metric_id: int, beta: float) -> Tuple[float, float]:
    """(c1, beta2) of xc_metric_params: 1 + beta**2 and beta**2 exactly like python computes them;
    precision@k passes k through c1."""
    if metric_id == XC_METRIC_PREC_AT_K:
        return float(beta), 0.0
    return float(1 + beta**2), float(beta**2)


def resolve_macro_metric(func: Callable, metric_kwargs: Optional[Dict[str, Any]] = None) -> Tuple[int, float, float]:
    """Same for a macro-averaged metric on the confusion matrix (Frank-Wolfe objective)."""
    if isinstance(func, MixedInstancePrecisionMacroMetric):
        return resolve_binary_metric(func.per_label.binary_metric, metric_kwargs)
    if isinstance(func, MixedMacroRecallPrecisionMetric):
        return resolve_binary_metric(binary_precision_on_conf_matrix, metric_kwargs)
    inner = getattr(func, "_xc_binary_metric", None)
    if inner is None and getattr(func, "__name__", "") in ("macro_metric_on_conf_matrix",
                                                            "micro_metric_on_conf_matrix") and func.__closure__:
        # the reference's factory closure (xcolumns/metrics.py:51-58) captures `binary_metric`
        for cell in func.__closure__:
            try:
                v = cell.cell_contents
            except ValueError:
                continue
            if callable(v) and getattr(v, "__name__", None) in _BINARY_IDS:
                inner = v
    if inner is None:
        raise UnsupportedMetricError(
            f"metric_func={func!r} is not a built-in macro- or micro-averaged metric; xcolumns_b200 fuses "
            f"precision / recall / F-beta / Jaccard / balanced accuracy / G-mean / H-mean")
    return resolve_binary_metric(inner, metric_kwargs)
