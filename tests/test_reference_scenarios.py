"""The reference's own tests (tests/test_block_coordinate.py, tests/test_frank_wolfe.py,
tests/test_weighted_prediction.py of mwydmuch/xCOLUMNs) replayed against xcolumns_b200 on the reference's
test-data recipe (tests/golden/make_reference_fixture.py: the conftest.py fixture, first 4 000 validation /
test rows): same assertions -- container / dtype / row-sum rules, agreement of the confusion matrices across
float64 / float32 / CSR inputs within 3 counts, BCA and Frank-Wolfe no worse than top-k and within 0.02 of
the closed-form optimum for macro recall -- plus the scores the live reference reaches on this fixture."""
import numpy as np
import pytest
from scipy.sparse import csr_matrix

pytestmark = pytest.mark.gpu

# what the live reference computes on the fixture (CPU run in the build container)
REF_TOPK, REF_OPT, REF_BCA, REF_FW = 0.2706989645730515, 0.3184660011959975, 0.3191484336887993, 0.3184660011959975


@pytest.fixture(scope="module")
def data(golden):
    g = golden("reference_fixture")
    return {k: v.astype(np.float64) for k, v in g.items()}


@pytest.fixture(scope="module")
def xb():
    import xcolumns_b200
    return xcolumns_b200


def _variants(args):
    """float64, float32 and CSR copies of the matrix arguments (tests/conftest.py:119-170)"""
    yield "float64", [a.astype(np.float64) if isinstance(a, np.ndarray) and a.ndim == 2 else a for a in args]
    yield "float32", [a.astype(np.float32) if isinstance(a, np.ndarray) and a.ndim == 2 else a for a in args]
    yield "csr", [csr_matrix(a) if isinstance(a, np.ndarray) and a.ndim == 2 else a for a in args]


def _max_diff(C1, C2):
    return max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).max() for a, b in zip(C1, C2))


def _check_pred(y_pred, y_proba, k):
    assert type(y_pred) == type(y_proba)
    assert y_pred.dtype == y_proba.dtype
    assert (np.asarray(y_pred.sum(axis=1)).ravel() == k).all()


def test_block_coordinate_macro_recall(xb, data):
    from xcolumns_b200.metrics import macro_recall_on_conf_matrix
    y_val, y_test, y_proba_test = data["y_val"], data["y_test"], data["y_proba_test"]
    k = 3
    top_C = xb.calculate_confusion_matrix(y_test, xb.predict_top_k(y_proba_test, k), normalize=False, skip_tn=False)
    conf = {}
    for name, (yt, yp) in _variants((y_test, y_proba_test)):
        y_pred = xb.predict_optimizing_macro_recall_using_bc(yp, k, seed=2024)
        _check_pred(y_pred, yp, k)
        conf[name] = xb.calculate_confusion_matrix(yt, y_pred, normalize=False, skip_tn=False)
    assert _max_diff(conf["float64"], conf["float32"]) <= 3 and _max_diff(conf["float64"], conf["csr"]) <= 3
    opt = xb.predict_optimizing_macro_recall(y_proba_test, k, priors=y_val.mean(axis=0))
    opt_C = xb.calculate_confusion_matrix(y_test, opt, normalize=False, skip_tn=False)
    top_s, bc_s, opt_s = (float(macro_recall_on_conf_matrix(*c)) for c in (top_C, conf["float64"], opt_C))
    assert bc_s >= top_s and abs(opt_s - bc_s) < 0.02                     # tests/test_block_coordinate.py:95-96
    assert abs(top_s - REF_TOPK) < 1e-12 and abs(opt_s - REF_OPT) < 1e-12 and abs(bc_s - REF_BCA) < 1e-12


def test_frank_wolfe_macro_recall(xb, data):
    from xcolumns_b200.metrics import macro_recall_on_conf_matrix
    y_val, y_proba_val, y_test, y_proba_test = data["y_val"], data["y_proba_val"], data["y_test"], data["y_proba_test"]
    k = 3
    rs = np.random.RandomState(0)
    init_a, init_b = rs.rand(y_proba_val.shape[1]), rs.rand(y_proba_val.shape[1])
    top_C = xb.calculate_confusion_matrix(y_test, xb.predict_top_k(y_proba_test, k), normalize=False, skip_tn=False)
    conf = {}
    for name, (yv, pv, yt, pt) in _variants((y_val, y_proba_val, y_test, y_proba_test)):
        clf, meta = xb.find_classifier_using_fw(yv, pv, macro_recall_on_conf_matrix, k, return_meta=True, seed=2024,
                                                init_classifier=(init_a, init_b))
        y_pred = clf.predict(pt, seed=2024)
        _check_pred(y_pred, pt, k)
        conf[name] = xb.calculate_confusion_matrix(yt, y_pred, normalize=False, skip_tn=False)
    assert _max_diff(conf["float64"], conf["float32"]) <= 3 and _max_diff(conf["float64"], conf["csr"]) <= 3
    opt = xb.predict_optimizing_macro_recall(y_proba_test, k, priors=y_val.mean(axis=0))
    opt_C = xb.calculate_confusion_matrix(y_test, opt, normalize=False, skip_tn=False)
    top_s, fw_s, opt_s = (float(macro_recall_on_conf_matrix(*c)) for c in (top_C, conf["float64"], opt_C))
    assert fw_s >= top_s and abs(opt_s - fw_s) < 0.02                      # tests/test_frank_wolfe.py:102-103
    assert abs(fw_s - REF_FW) < 1e-3


def test_weighted_prediction_containers(xb, data):
    """tests/test_weighted_prediction.py: predict_weighted_per_instance across containers with a and b"""
    y_test, y_proba_test = data["y_test"], data["y_proba_test"]
    k = 3
    rs = np.random.RandomState(1)
    a, b = rs.rand(y_proba_test.shape[1]), rs.rand(y_proba_test.shape[1]) * 0.01
    conf = {}
    for name, (yt, yp) in _variants((y_test, y_proba_test)):
        y_pred = xb.predict_weighted_per_instance(yp, k, a=a, b=b)
        _check_pred(y_pred, yp, k)
        conf[name] = xb.calculate_confusion_matrix(yt, y_pred, normalize=False, skip_tn=False)
    assert _max_diff(conf["float64"], conf["float32"]) <= 3 and _max_diff(conf["float64"], conf["csr"]) <= 3
