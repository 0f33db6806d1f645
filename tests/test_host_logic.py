"""Host-side logic of the boundary that needs no GPU: metric resolution (ours, the reference's own callables
by module + name, the mixed / micro / precision@k objects), the wrappers' rewritten signatures (the
reference's experiment scripts filter kwargs by them: experiments/utils.py:16-26, xcolumns/utils.py:209-230),
mode / batch-size rules, and the metric formulas against literal restatements of xcolumns/metrics.py."""
import inspect

import numpy as np
import pytest


def test_metric_formulas_match_reference_expressions():
    from xcolumns_b200 import metrics as M
    rng = np.random.default_rng(0)
    tp, fp, fn, tn = (rng.random(50) for _ in range(4))
    e = 1e-9
    assert np.array_equal(M.binary_precision_on_conf_matrix(tp, fp, fn, tn), tp / (tp + fp + e))       # :613
    assert np.array_equal(M.binary_recall_on_conf_matrix(tp, fp, fn, tn), tp / (tp + fn + e))          # :652
    assert np.array_equal(M.binary_fbeta_score_on_conf_matrix(tp, fp, fn, tn, beta=2.0),
                          (1 + 2.0**2) * tp / ((2.0**2 * (tp + fp)) + tp + fn + e))                        # :703
    assert np.array_equal(M.binary_jaccard_score_on_conf_matrix(tp, fp, fn, tn), tp / (tp + fp + fn + e))  # :797
    tpr, tnr = tp / (tp + fn + e), tn / (tn + fp + e)
    assert np.array_equal(M.binary_balanced_accuracy_on_conf_matrix(tp, fp, fn, tn), (tpr + tnr) / 2)  # :843-845
    assert np.allclose(M.binary_gmean_on_conf_matrix(tp, fp, fn, tn), (tpr * tnr) ** 0.5, rtol=1e-15)  # :890-892
    assert np.array_equal(M.binary_hmean_on_conf_matrix(tp, fp, fn, tn), 2 * tpr * tnr / (tpr + tnr))  # :939-941
    assert M.macro_f1_score_on_conf_matrix(tp, fp, fn, tn) == M.binary_f1_score_on_conf_matrix(tp, fp, fn, tn).mean()
    assert M.micro_recall_on_conf_matrix(tp, fp, fn, tn) == M.binary_recall_on_conf_matrix(
        tp.sum(), fp.sum(), fn.sum(), tn.sum())
    assert np.array_equal(M.binary_precision_at_k_on_conf_matrix(tp, fp, fn, tn, 5), tp / 5)            # :513
    mixed = M.MixedInstancePrecisionMetric(M.binary_f1_score_on_conf_matrix, 0.3, 5, 50)
    assert np.array_equal(mixed(tp, fp, fn, tn), (1 - 0.3) * (tp / 5) + 0.3 * M.binary_f1_score_on_conf_matrix(
        tp, fp, fn, tn) / 50)                                                                             # bc:922-925


def test_metric_resolution():
    from xcolumns_b200 import metrics as M
    assert M.resolve_binary_metric(M.binary_f1_score_on_conf_matrix) == (M.XC_METRIC_FBETA, 1.0, 1e-9)
    assert M.resolve_binary_metric(M.binary_fbeta_score_on_conf_matrix, {"beta": 2.0, "epsilon": 1e-6}) == (
        M.XC_METRIC_FBETA, 2.0, 1e-6)
    assert M.resolve_binary_metric(M.PrecisionAtK(7))[0] == M.XC_METRIC_PREC_AT_K
    assert M.metric_c1_beta2(M.XC_METRIC_PREC_AT_K, 7.0) == (7.0, 0.0) and M.metric_c1_beta2(M.XC_METRIC_FBETA, 2.0) == (5.0, 4.0)
    mixed = M.MixedInstancePrecisionMetric(M.binary_recall_on_conf_matrix, 0.25, 3, 40)
    assert M.resolve_binary_metric(mixed)[0] == M.XC_METRIC_RECALL and M.resolve_mix(mixed) == (0.25, 3.0, 40.0)
    assert M.resolve_mix(M.binary_recall_on_conf_matrix) is None
    assert M.resolve_macro_metric(M.macro_jaccard_score_on_conf_matrix)[0] == M.XC_METRIC_JACCARD
    assert M.is_micro_metric(M.micro_f1_score_on_conf_matrix) and not M.is_micro_metric(M.macro_f1_score_on_conf_matrix)
    with pytest.raises(ValueError):
        M.resolve_binary_metric(M.binary_recall_on_conf_matrix, {"beta": 2.0})       # unknown kwarg for this metric
    with pytest.raises(NotImplementedError):
        M.resolve_binary_metric(lambda tp, fp, fn, tn: tp)
    with pytest.raises(NotImplementedError):
        M.resolve_macro_metric(lambda tp, fp, fn, tn: tp.sum())

    # the reference's own callables are recognised by module + name / by their factory closure
    def binary_recall_on_conf_matrix(tp, fp, fn, tn, epsilon=1e-9):
        return tp / (tp + fn + epsilon)

    binary_recall_on_conf_matrix.__module__ = "xcolumns.metrics"
    assert M.resolve_binary_metric(binary_recall_on_conf_matrix)[0] == M.XC_METRIC_RECALL

    def make(binary_metric):
        def macro_metric_on_conf_matrix(tp, fp, fn, tn, **kw):
            return binary_metric(tp, fp, fn, tn, **kw).mean()
        return macro_metric_on_conf_matrix

    assert M.resolve_macro_metric(make(binary_recall_on_conf_matrix))[0] == M.XC_METRIC_RECALL


def test_wrappers_expose_the_forwarded_keywords():
    import xcolumns_b200 as xb
    sig = inspect.signature(xb.predict_optimizing_macro_f1_score_using_bc).parameters
    assert list(sig)[:2] == ["y_proba", "k"]
    for name in ("tolerance", "init_y_pred", "max_iters", "shuffle_order", "return_meta", "seed", "verbose"):
        assert name in sig, name
    for name in ("binary_metric_func", "metric_aggregation", "maximize", "skip_tn"):
        assert name not in sig, name                       # bound by the wrapper (block_coordinate.py:733-748)
    assert sig["tolerance"].default == 1e-6 and sig["max_iters"].default == 100 and sig["init_y_pred"].default == "top"
    fsig = inspect.signature(xb.find_classifier_optimizing_macro_f1_score_using_fw).parameters
    assert list(fsig)[:3] == ["y_true", "y_proba", "k"]
    assert "alpha_search_algo" in fsig and "metric_func" not in fsig and fsig["alpha_uniform_search_step"].default == 1e-4
    bsig = inspect.signature(xb.predict_using_bc_with_0approx).parameters
    assert [bsig[n].default for n in ("metric_aggregation", "normalize_conf_matrix", "maximize", "tolerance", "init_y_pred",
                                      "max_iters", "shuffle_order", "skip_tn", "return_meta")] == [
        "mean", True, True, 1e-6, "top", 100, True, False, False]                      # block_coordinate.py:296-313
    csig = inspect.signature(xb.predict_optimizing_coverage_using_bc).parameters
    assert csig["alpha"].default == 1 and csig["max_iters"].default == 100                # :600-611
    isig = inspect.signature(xb.predict_optimizing_instance_precision_using_bc).parameters
    assert isig["init_y_pred"].default == "random"                                        # :804-813


def test_mode_and_batch_rules():
    from xcolumns_b200.block_coordinate import _resolve_mode, default_batch_rows
    assert _resolve_mode(None, 100, False) == "exact" and _resolve_mode(None, 100000, False) == "batched"
    assert _resolve_mode(None, 100000, True) == "exact" and _resolve_mode("batched", 10, False) == "batched"
    with pytest.raises(ValueError):
        _resolve_mode("fast", 10, False)
    assert default_batch_rows(307000, 7104) % 7104 == 0 and 307000 // 10 < default_batch_rows(307000, 7104) <= 307000 // 8
    assert 1 <= default_batch_rows(10, 7104) <= 10 and default_batch_rows(10000, 0) == 1250


def test_confusion_matrix_container():
    from xcolumns_b200 import ConfusionMatrix
    a = ConfusionMatrix(np.array([1.0, 2.0]), np.array([0.5, 0.5]), np.array([2.0, 1.0]), np.array([6.5, 6.5]))
    tp, fp, fn, tn = a                                        # iterable -> 4-tuple (confusion_matrix.py:16-152)
    assert tp[1] == 2.0 and tn[0] == 6.5
    b = a + a
    assert (b.tp == 2 * a.tp).all() and ((a * 2).fn == b.fn).all() and ((b / 2).fp == a.fp).all()
    n = a.normalize()
    assert np.allclose(n.tp + n.fp + n.fn + n.tn, 1.0)


def test_vectorised_classifier_draw_is_the_reference_stream():
    """predict_using_randomized_weighted_classifier draws one classifier per row; the vectorised draw consumes the
    same Generator stream as the reference's per-row rng.choice(arange(len(p)), p=p) (frank_wolfe.py:70)"""
    from xcolumns_b200.frank_wolfe import _draw_classifiers
    for seed in (0, 3):
        for dt in (np.float32, np.float64):
            p = np.random.default_rng(5).random(21).astype(dt)
            p /= p.sum()
            rng = np.random.default_rng(seed)
            ref = np.array([rng.choice(np.arange(p.shape[0]), p=p) for _ in range(4000)])
            assert (_draw_classifiers(p, 4000, seed) == ref).all()
    with pytest.raises(ValueError):
        _draw_classifiers(np.array([0.5, 0.6]), 10, 0)


def test_batch_and_lag_policy(monkeypatch):
    """host-side scheduling rules of the batched sweeps (xcolumns_b200/block_coordinate.py)"""
    from xcolumns_b200 import _device as dev
    from xcolumns_b200.block_coordinate import _default_lag, coverage_batch_rows, default_batch_rows
    wave = 7104
    # n/8 rounded UP (no ninth commit over a handful of rows), whole waves once a batch exceeds one
    assert default_batch_rows(38375, wave) == 4797 and -(-38375 // 4797) == 8
    assert default_batch_rows(307000, wave) == 35520 and default_batch_rows(307000, wave) % wave == 0
    assert default_batch_rows(5, 0) == 1 and default_batch_rows(0, wave) == 1
    monkeypatch.delenv("XCOLUMNS_B200_LAG", raising=False)
    # two batch kernels in flight fill the GPU when a batch is >= 4 waves; smaller batches get a third
    assert _default_lag(35520, wave) == 1 and _default_lag(14208, wave) == 2 and _default_lag(4797, wave) == 2
    monkeypatch.setenv("XCOLUMNS_B200_LAG", "0")
    assert _default_lag(4797, wave) == 0
    monkeypatch.setenv("XCOLUMNS_B200_LAG", "9")
    assert _default_lag(4797, wave) == 3
    assert coverage_batch_rows(153000) == 4096 and coverage_batch_rows(3000) == 93 and coverage_batch_rows(10) == 1
    # host threads of the result helpers: an equal share per local rank
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    monkeypatch.setattr("os.cpu_count", lambda: 32)
    assert dev.host_threads() == 4
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "1")
    assert dev.host_threads() == 32
