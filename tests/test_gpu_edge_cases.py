"""Edge cases of the boundary against the CPU oracle: degenerate shapes (one row, m == k, k = 1, k = 32),
unaligned / strided / non-contiguous inputs, float64, CSR rows that are empty or shorter than k, and the
reference's error behaviour (SURVEY.md section 8b "Error convention")."""
import numpy as np
import pytest
import torch
from scipy.sparse import csr_matrix

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xb():
    import xcolumns_b200
    return xcolumns_b200


def _idx(pred, k):
    r, c = np.nonzero(pred)
    return c.reshape(pred.shape[0], k).astype(np.int32)


@pytest.mark.parametrize("n,m,k", [(1, 5, 5), (1, 7, 1), (3, 33, 32), (2, 1000, 1), (5, 6, 3), (64, 31, 7)])
def test_topk_degenerate_shapes(xb, oracle, n, m, k):
    from xcolumns_b200.synth import dense_probs
    for dt in (np.float32, np.float64):
        eta = dense_probs(n, m, seed=100 + n + m).astype(dt)
        got = xb.predict_top_k(eta, k)
        assert got.shape == (n, m) and got.dtype == dt and (got.sum(1) == k).all()
        assert (_idx(got, k) == oracle.topk_indices_dense(eta, k)[0]).all()


def test_dense_layouts_agree(xb, oracle):
    """odd leading dimension, column slice, transposed view, CUDA / CPU tensors: same selection as numpy"""
    from xcolumns_b200.synth import dense_probs
    big = dense_probs(40, 203, seed=5)
    ref = oracle.topk_indices_dense(np.ascontiguousarray(big[:, 1:200]), 4)[0]
    for x in (big[:, 1:200],                                   # numpy view, unaligned rows
              torch.from_numpy(big)[:, 1:200],                 # CPU tensor view
              torch.from_numpy(big).cuda()[:, 1:200],          # CUDA tensor view (stride 203, offset 1)
              torch.from_numpy(np.ascontiguousarray(big[:, 1:200].T)).cuda().T):   # column-major
        got = xb.predict_top_k(x, 4)
        assert type(got) is type(x) and tuple(got.shape) == (40, 199)
        g = got.cpu().numpy() if isinstance(got, torch.Tensor) else got
        assert (_idx(g, 4) == ref).all()
    a = (0.5 + np.random.default_rng(0).random(199)).astype(np.float32)
    x = torch.from_numpy(big).cuda()[:, 1:200]
    got = xb.predict_weighted_per_instance(x, 3, a=torch.from_numpy(a).cuda(), b=torch.from_numpy(-a / 7).cuda())
    ref = oracle.topk_indices_dense(np.ascontiguousarray(big[:, 1:200]), 3, a, -a / 7)[0]
    assert (_idx(got.cpu().numpy(), 3) == ref).all()


def test_csr_short_and_empty_rows(xb, oracle):
    """rows with nnz < k (incl. nnz = 0): top-k keeps what is there with the reference's (0, 1) padding;
    BCA keeps all stored labels of such rows"""
    rng = np.random.default_rng(3)
    m, k = 50, 4
    rows = [rng.choice(m, size=s, replace=False) for s in (0, 1, 3, 4, 9, 0, 12, 2)]
    indptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int32)
    indices = np.concatenate([np.sort(r) for r in rows]).astype(np.int32)
    data = (0.05 + 0.9 * rng.random(indices.size)).astype(np.float32)
    y = csr_matrix((data, indices, indptr), shape=(len(rows), m))
    got = xb.predict_top_k(y, k)
    ref = oracle.predict_weighted_per_instance(y, k)
    assert (got.indptr == ref.indptr).all() and (got.indices == ref.indices).all() and (got.data == ref.data).all()
    tp, fp, fn, tn = xb.calculate_confusion_matrix(y, got, dtype=np.float64)
    otp, ofp, ofn, otn = oracle.calculate_confusion_matrix(y, ref, dtype=np.float64)
    assert np.allclose(tp, otp, atol=1e-12) and np.allclose(fp, ofp, atol=1e-12) and np.allclose(fn, ofn, atol=1e-12)
    for mode in ("exact", "batched"):
        pred = xb.predict_optimizing_macro_recall_using_bc(y, k, seed=0, mode=mode, y_pred_format="indices")
        for i, r in enumerate(rows):
            sel = pred[i][pred[i] >= 0]
            assert len(sel) == min(k, len(r)) and set(sel) <= set(r.tolist())


def test_single_row_and_tiny_bca(xb, oracle):
    from xcolumns_b200.synth import dense_probs
    for n, m, k in ((1, 9, 3), (2, 5, 5), (7, 12, 1)):
        eta = dense_probs(n, m, seed=n * 10 + m)
        pred, meta = xb.predict_optimizing_macro_f1_score_using_bc(eta, k, seed=0, mode="exact", return_meta=True)
        opred, ometa = oracle.predict_using_bc_with_0approx(eta, "f1", k, seed=0, skip_tn=True)
        assert (pred.astype(np.uint8) == opred).all() and meta["utilities"] == ometa["utilities"]
        predb = xb.predict_optimizing_macro_f1_score_using_bc(eta, k, seed=0, mode="batched")
        assert (predb.sum(1) == k).all()


def test_confusion_matrix_axis_dtype_and_compact(xb, oracle):
    from xcolumns_b200.synth import dense_probs
    eta = dense_probs(33, 77, seed=12)
    lab = (np.random.default_rng(1).random(eta.shape) < eta).astype(np.float32)
    pred = xb.predict_top_k(eta, 6)
    for kw in (dict(), dict(normalize=True), dict(axis=1, dtype=np.float64), dict(skip_tn=True, dtype=np.float64),
               dict(axis=1, normalize=True, dtype=np.float64)):
        c = xb.calculate_confusion_matrix(lab, pred, **kw)
        o = oracle.calculate_confusion_matrix(lab, pred, **kw)
        for a, b in zip(c, o):
            assert np.asarray(a).dtype == np.asarray(b).dtype and np.allclose(a, b, rtol=0, atol=1e-6), kw
    # integer labels with dtype=None accumulate in the input dtype like numpy (confusion_matrix.py:166)
    c = xb.calculate_confusion_matrix(lab.astype(np.int64), pred.astype(np.int64))
    assert np.asarray(c.tp).dtype == np.int64 and (np.asarray(c.tp) == (lab * pred).sum(0)).all()
    t = xb.calculate_confusion_matrix(torch.from_numpy(lab).cuda(), torch.from_numpy(pred).cuda(), dtype=torch.float64)
    assert isinstance(t.tp, torch.Tensor) and t.tp.is_cuda
    assert np.allclose(t.tp.cpu().numpy(), (lab * pred).sum(0))


def test_error_convention(xb):
    from xcolumns_b200 import metrics as M
    eta = np.random.default_rng(0).random((6, 10)).astype(np.float32)
    with pytest.raises(ValueError):
        xb.predict_top_k(eta, 2.0)                               # k must be an int
    with pytest.raises(ValueError):
        xb.predict_top_k(eta, 11)                                # k > m
    with pytest.raises(ValueError):
        xb.predict_weighted_per_instance(eta, 2, a=np.ones(9, dtype=np.float32))
    with pytest.raises(ValueError):
        xb.predict_top_k([[0.1, 0.2]], 1)                        # unsupported container
    with pytest.raises(ValueError):
        xb.calculate_confusion_matrix(eta, csr_matrix(eta))     # mixed containers
    with pytest.raises(ValueError):
        xb.calculate_confusion_matrix(eta, eta[:, :5])           # shape mismatch
    with pytest.raises(ValueError):
        xb.predict_using_bc_with_0approx(eta, M.binary_f1_score_on_conf_matrix, 2, metric_aggregation="median")
    with pytest.raises(ValueError):
        xb.find_classifier_using_fw(eta, eta[:, :5], M.macro_f1_score_on_conf_matrix, 2)
    with pytest.raises(ValueError):
        xb.find_classifier_using_fw(eta, eta, M.macro_f1_score_on_conf_matrix, 2, init_classifier="nope")
    with pytest.raises(ValueError):
        xb.find_classifier_using_fw(eta, eta, M.macro_f1_score_on_conf_matrix, 2, alpha_search_algo="golden")
    with pytest.raises(NotImplementedError):   # a LIST of callables on CSR rows: loud, no fallback (the reference's own
        # CSR step indexes such a list by stored position, block_coordinate.py:110-127)
        xb.predict_using_bc_with_0approx(csr_matrix(eta), [lambda tp, fp, fn, tn: tp] * 10, 2)
    with pytest.raises(ValueError):             # a list of callables must have one entry per label
        xb.predict_using_bc_with_0approx(eta, [M.binary_f1_score_on_conf_matrix] * 3, 2)
