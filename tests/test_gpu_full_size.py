"""BASELINE.json's full-size configurations, checked through size-independent properties (the CPU
oracle cannot run these shapes in test time): exactly k distinct ascending labels per row, sweep
utilities that never decrease, agreement of the reported utility with an independent recomputation
from the returned prediction, no regression against the top-k start, idempotence of a converged
solution, and sampled rows against torch.topk."""
import numpy as np
import pytest
import torch
from scipy.sparse import csr_matrix

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xb():
    import xcolumns_b200
    return xcolumns_b200


def _macro(metric, y_dev, pred_idx, n, m, eps=1e-9):
    """macro metric of compact predictions against y = eta, float64 on the device (independent of the library)"""
    tp = torch.zeros(m, dtype=torch.float64, device=y_dev.device)
    cnt = torch.zeros(m, dtype=torch.float64, device=y_dev.device)
    rows = torch.arange(n, device=y_dev.device).repeat_interleave(pred_idx.shape[1])
    flat = pred_idx.reshape(-1).long()
    tp.index_add_(0, flat, y_dev[rows, flat].double())
    cnt.index_add_(0, flat, torch.ones_like(flat, dtype=torch.float64))
    col = y_dev.sum(0, dtype=torch.float64)
    tp, fp, fn = tp / n, (cnt - tp) / n, (col - tp) / n
    if metric == "f1":
        return float((2 * tp / (2 * tp + fp + fn + eps)).mean())
    return float((tp / (tp + fn + eps)).mean())


def test_c3_amazoncat_shape_dense_bca(xb):
    from xcolumns_b200.synth import dense_probs_device
    n, m, k = 307000, 13000, 5
    dev_ = torch.device("cuda", 0)
    eta = dense_probs_device(n, m, seed=1003, device=dev_)
    pred, meta = xb.predict_optimizing_macro_f1_score_using_bc(eta, k, seed=0, mode="batched", return_meta=True,
                                                               y_pred_format="indices", max_iters=8)
    assert pred.shape == (n, k) and pred.dtype == torch.int32
    assert bool((pred[:, 1:] > pred[:, :-1]).all()) and int(pred.min()) >= 0 and int(pred.max()) < m
    u = np.array(meta["utilities"])
    assert (np.diff(u) > -1e-9).all(), u
    assert abs(_macro("f1", eta, pred, n, m) - u[-1]) < 1e-9
    top = torch.topk(eta[:2048], k, dim=1).indices.sort(dim=1).values.int()
    top_lib = xb.predict_top_k(eta[:2048], k)
    assert bool((torch.nonzero(top_lib)[:, 1].reshape(-1, k).int() == top).all())
    top_all = torch.cat([torch.topk(eta[s:s + 32768], k, dim=1).indices.sort(dim=1).values.int()
                         for s in range(0, n, 32768)])
    assert u[-1] > _macro("f1", eta, top_all, n, m) + 1e-3            # BCA beats its top-k start
    # a converged solution is a fixed point up to the batched mode's tolerance: restarting from it with a
    # compact (n, k) warm start cannot lose utility
    init = torch.zeros((4096, m), dtype=torch.float32, device=dev_)
    init.scatter_(1, pred[:4096].long(), 1.0)
    pred2, meta2 = xb.predict_optimizing_macro_f1_score_using_bc(eta[:4096], k, seed=1, mode="batched",
                                                                 init_y_pred=init, return_meta=True,
                                                                 y_pred_format="indices", max_iters=3)
    assert (np.diff(np.array(meta2["utilities"])) > -1e-9).all()


def test_c4_amazon670k_shape_csr_bca(xb):
    from xcolumns_b200.synth import csr_probs_device
    n, m, nnz, k = 153000, 670000, 100, 5
    data, idx, ptr = csr_probs_device(n, m, nnz, 1004, torch.device("cuda", 0))
    y = csr_matrix((data.cpu().numpy(), idx.cpu().numpy(), ptr.cpu().numpy().astype(np.int32)), shape=(n, m))
    assert y.has_sorted_indices
    pred, meta = xb.predict_optimizing_macro_recall_using_bc(y, k, seed=0, mode="batched", return_meta=True, max_iters=6)
    assert isinstance(pred, csr_matrix) and pred.shape == y.shape and pred.dtype == y.dtype
    assert (np.diff(pred.indptr) == k).all() and pred.has_sorted_indices
    # every predicted label is one of the row's stored labels
    stored = y.copy()
    stored.data[:] = 1
    assert pred.multiply(stored).nnz == pred.nnz
    u = np.array(meta["utilities"])
    assert (np.diff(u) > -1e-9).all(), u
    cov, cmeta = xb.predict_optimizing_coverage_using_bc(y, k, seed=0, mode="batched", return_meta=True, max_iters=5)
    assert (np.diff(cov.indptr) == k).all() and cov.has_sorted_indices
    cu = np.array(cmeta["utilities"])
    assert (np.diff(cu) > -1e-9).all() and 0.0 < cu[-1] <= 1.0, cu
    # coverage of the result recomputed on the host: 1 - mean_j prod_i (1 - yhat_ij eta_ij)
    sel = cov.multiply(y).tocoo()
    logf = np.zeros(m)
    np.add.at(logf, sel.col, np.log1p(-sel.data.astype(np.float64)))
    assert abs((1.0 - np.exp(logf).mean()) - cu[-1]) < 1e-9


def test_c5_wiki10_shape_frank_wolfe(xb):
    from xcolumns_b200 import metrics as M
    from xcolumns_b200.synth import dense_probs_device
    n, m, k = 14000, 31000, 5
    eta = dense_probs_device(n, m, seed=1005, device=torch.device("cuda", 0))
    clf, meta = xb.find_classifier_using_fw(eta, eta, M.macro_f1_score_on_conf_matrix, k, max_iters=20,
                                            tolerance=-np.inf, alpha_tolerance=0.0, skip_tn=True, seed=0,
                                            return_meta=True)
    assert meta["iters"] == 20 and clf.a.shape == (21, m) and clf.b.shape == (21, m)
    p = clf.p.double().cpu().numpy()
    assert abs(p.sum() - 1.0) < 1e-5 and (p >= 0).all()
    u = np.array(meta["utilities"])
    assert (np.diff(u) > -1e-12).all(), u          # the line search includes alpha = 0: never worse
    assert all(0.0 <= a <= 1.0 for a in meta["alphas"])
    # the first classifier is the plain top-k; its reported utility matches an independent recomputation
    top = torch.topk(eta, k, dim=1).indices.sort(dim=1).values.int()
    assert abs(_macro("f1", eta, top, n, m) - u[0]) < 1e-9
    # prediction with the randomized classifier: exactly k labels per row
    yp = clf.predict(eta[:3000], seed=0)
    assert bool(((yp != 0).sum(1) == k).all())
