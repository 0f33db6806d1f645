"""The CPU oracle (oracle/) replayed against outputs of the LIVE reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  Bit-exact on label ids,
float64 state and per-sweep utilities; Frank-Wolfe within 1e-6 (the reference differentiates
with autograd and keeps float32 confusion vectors; contract tolerance is 1e-4)."""
import numpy as np
import pytest
from scipy.sparse import csr_matrix


def _idx(pred, k):
    n = pred.shape[0]
    r, c = np.nonzero(pred)
    return c.reshape(n, k).astype(np.int32)


def test_topk_dense(golden, oracle):
    g = golden("topk_dense")
    eta = g["eta"]
    assert (oracle.topk_indices_dense(eta, 5)[0] == g["top5"]).all()
    assert (oracle.topk_indices_dense(eta, 1)[0] == g["top1"]).all()
    assert (oracle.topk_indices_dense(eta, 5, g["a32"], g["b32"])[0] == g["w5_ab32"]).all()
    assert (oracle.topk_indices_dense(eta, 3, g["a64"])[0] == g["w3_a64"]).all()
    assert (oracle.topk_indices_dense(eta.astype(np.float64), 5, g["a64"], g["b32"].astype(np.float64))[0]
            == g["w5_f64"]).all()
    ks = oracle.predict_weighted_per_instance(eta, 4, a=g["a32"], b=g["b32"], keep_scores=True)
    assert ks.dtype == g["w4_scores"].dtype and (ks == g["w4_scores"]).all()
    th = oracle.predict_weighted_per_instance(eta, 0, th=0.3, a=g["a32"], b=g["b32"])
    assert (th == g["th0"]).all()


def test_topk_csr(golden, oracle):
    g = golden("topk_csr")
    y = csr_matrix((g["data"], g["indices"], g["indptr"]), shape=tuple(g["shape"]))
    for name, kw, k in (("top5", {}, 5), ("top8", {}, 8), ("w5", {"a": g["a"], "b": g["b"]}, 5),
                        ("w5s", {"a": g["a"], "b": g["b"], "keep_scores": True}, 5)):
        r = oracle.predict_weighted_per_instance(y, k, **kw)
        assert (r.indices == g[name + "_indices"]).all(), name
        assert (r.indptr == g[name + "_indptr"]).all(), name
        assert r.data.dtype == g[name + "_data"].dtype and (r.data == g[name + "_data"]).all(), name


def test_confmat(golden, oracle):
    g = golden("confmat")
    eta, lab = g["eta"], g["lab"]
    n, m = eta.shape
    pred = np.zeros_like(eta)
    pred[np.arange(n)[:, None], g["pred"]] = 1
    for name, yt, kw in (("probs_f64", eta, dict(dtype=np.float64, skip_tn=True)),
                         ("probs_none", eta, dict()),
                         ("lab_norm", lab, dict(normalize=True)),
                         ("lab_f64_norm_skip", lab, dict(normalize=True, skip_tn=True, dtype=np.float64))):
        c = np.stack([np.asarray(v, dtype=np.float64) for v in oracle.calculate_confusion_matrix(yt, pred, **kw)])
        assert (c == g[name]).all(), name
    c = np.stack(oracle.calculate_confusion_matrix(lab, pred, axis=1, dtype=np.float64))
    assert np.allclose(c, g["lab_axis1"], rtol=0, atol=1e-9)
    y = csr_matrix((g["c_data"], g["c_indices"], g["c_indptr"]), shape=tuple(g["c_shape"]))
    nn = y.shape[0]
    p = csr_matrix((np.ones(nn * 4, dtype=np.float32), g["c_pred"].reshape(-1), np.arange(nn + 1) * 4), shape=y.shape)
    for name, kw in (("csr_f64", dict(dtype=np.float64, skip_tn=True)), ("csr_none", dict()),
                     ("csr_norm", dict(normalize=True, dtype=np.float64))):
        c = np.stack([np.asarray(v, dtype=np.float64) for v in oracle.calculate_confusion_matrix(y, p, **kw)])
        assert (c == g[name]).all(), name


BCA_DENSE = [
    ("f1", "f1", 5, dict(seed=0, skip_tn=True)),
    ("f1_f64", "f1", 5, dict(seed=0, skip_tn=True)),
    ("recall", "recall", 5, dict(seed=3, skip_tn=True)),
    ("precision", "precision", 3, dict(seed=4, skip_tn=True)),
    ("jaccard", "jaccard", 5, dict(seed=5, skip_tn=True)),
    ("fbeta2", "fbeta", 5, dict(seed=6, skip_tn=True, beta=2.0, epsilon=1e-6)),
    ("balacc", "balanced_accuracy", 5, dict(seed=7, skip_tn=False)),
    ("gmean", "gmean", 5, dict(seed=8, skip_tn=False)),
    ("hmean", "hmean", 5, dict(seed=9, skip_tn=False)),
    ("f1_noshuffle", "f1", 5, dict(seed=0, skip_tn=True, shuffle_order=False)),
    ("f1_random", "f1", 5, dict(seed=10, skip_tn=True, init_y_pred="random")),
    ("f1_greedy", "f1", 5, dict(seed=11, skip_tn=True, init_y_pred="greedy")),
    ("f1_sum", "f1", 5, dict(seed=12, skip_tn=True, metric_aggregation="sum", tolerance=1e-4)),
]


@pytest.mark.parametrize("name,metric,k,kw", BCA_DENSE, ids=[c[0] for c in BCA_DENSE])
def test_bca_dense(golden, oracle, name, metric, k, kw):
    g = golden("bca_dense")
    eta = g["eta"].astype(np.float64) if name.endswith("f64") else g["eta"]
    pred, meta = oracle.predict_using_bc_with_0approx(eta, metric, k, **kw)
    assert (_idx(pred, k) == g[name + "_pred"]).all()
    assert meta["iters"] == len(g[name + "_util"])
    assert (np.array(meta["utilities"]) == g[name + "_util"]).all()


def test_bca_dense_k0(golden, oracle):
    g = golden("bca_dense")
    eta = g["eta"]
    init = np.zeros(eta.shape, dtype=np.uint8)
    init[np.arange(eta.shape[0])[:, None], oracle.topk_indices_dense(eta, 3)[0]] = 1
    pred, meta = oracle.predict_using_bc_with_0approx(eta, "f1", 0, seed=13, skip_tn=True, init_y_pred=init)
    assert (pred == g["f1_k0_pred"]).all()
    assert (np.array(meta["utilities"]) == g["f1_k0_util"]).all()


BCA_CSR = [("f1", "f1", 5, 0, True), ("recall", "recall", 5, 1, True), ("jaccard", "jaccard", 3, 2, True),
           ("f1_f64", "f1", 5, 0, True), ("balacc", "balanced_accuracy", 5, 3, False), ("hmean", "hmean", 5, 4, False)]


@pytest.mark.parametrize("name,metric,k,seed,skip_tn", BCA_CSR, ids=[c[0] for c in BCA_CSR])
def test_bca_csr(golden, oracle, name, metric, k, seed, skip_tn):
    g = golden("bca_csr")
    data = g["data"].astype(np.float64) if name.endswith("f64") else g["data"]
    y = csr_matrix((data, g["indices"], g["indptr"]), shape=tuple(g["shape"]))
    pidx, meta = oracle.predict_using_bc_with_0approx(y, metric, k, seed=seed, skip_tn=skip_tn)
    assert (pidx == g[name + "_pred"]).all()
    assert (np.array(meta["utilities"]) == g[name + "_util"]).all()


@pytest.mark.parametrize("name,kw", [("cov", dict(seed=0)), ("cov_a07", dict(seed=1, alpha=0.7))])
def test_coverage_csr(golden, oracle, name, kw):
    g = golden("coverage_csr")
    y = csr_matrix((g["data"], g["indices"], g["indptr"]), shape=tuple(g["shape"]))
    pidx, meta = oracle.predict_optimizing_coverage_using_bc(y, 5, **kw)
    assert (pidx == g[name + "_pred"]).all()
    if name == "cov":
        assert (np.array(meta["utilities"]) == g[name + "_util"]).all()
    else:  # the precision@k mix is summed in float32 by the reference
        assert np.allclose(meta["utilities"], g[name + "_util"], rtol=0, atol=1e-6)


FW = [("f1_proba", "f1", "eta", dict(max_iters=10, skip_tn=True)),
      ("f1_lab", "f1", "lab", dict(max_iters=10, skip_tn=True, epsilon=1e-4)),
      ("recall_lab", "recall", "lab", dict(max_iters=6, skip_tn=True, epsilon=1e-4)),
      ("balacc_lab", "balanced_accuracy", "lab", dict(max_iters=6, epsilon=1e-4))]


@pytest.mark.parametrize("name,metric,yt,kw", FW, ids=[c[0] for c in FW])
def test_fw_dense(golden, oracle, name, metric, yt, kw):
    g = golden("fw_dense")
    A, B, P, meta = oracle.find_classifier_using_fw(g[yt], g["eta"], metric, 5, seed=0, **kw)
    assert meta["iters"] == int(g[name + "_iters"])
    assert A.shape == g[name + "_a"].shape
    assert np.allclose(meta["utilities"], g[name + "_util"], rtol=0, atol=1e-5)
    assert np.allclose(meta["alphas"], g[name + "_alphas"], rtol=0, atol=2e-3)
    assert np.allclose(P, g[name + "_p"], rtol=0, atol=2e-3)


MIXED = [("mix_f1", "f1", 5, 0.3, dict(seed=0)), ("mix_prec", "precision", 3, 0.5, dict(seed=1)),
         ("mix_recall", "recall", 5, 0.8, dict(seed=2)), ("mix_jaccard", "jaccard", 5, 0.5, dict(seed=3)),
         ("mix_balacc", "balanced_accuracy", 5, 0.5, dict(seed=4)),
         ("mix_f1_eps", "f1", 5, 0.6, dict(seed=5, epsilon=1e-5))]


@pytest.mark.parametrize("name,metric,k,alpha,kw", MIXED, ids=[c[0] for c in MIXED])
def test_bca_mixed_dense(golden, oracle, name, metric, k, alpha, kw):
    """mixed instance-precision / macro-metric utilities (block_coordinate.py:848-1045), bit-exact"""
    g = golden("mixed")
    eta = g["eta"]
    pred, meta = oracle.predict_using_bc_with_0approx(eta, metric, k, metric_aggregation="sum", skip_tn=True,
                                                      mix=(alpha, k, eta.shape[1]), **kw)
    assert (_idx(pred, k) == g[name + "_pred"]).all()
    assert (np.array(meta["utilities"]) == g[name + "_util"]).all()


def test_bca_instance_precision_and_csr_mixed(golden, oracle):
    g = golden("mixed")
    eta = g["eta"]
    pred, meta = oracle.predict_using_bc_with_0approx(eta, "precision_at_k", 5, metric_aggregation="sum", beta=5,
                                                      init_y_pred="random", seed=6)
    assert (_idx(pred, 5) == g["inst_prec_pred"]).all()
    assert (np.array(meta["utilities"]) == g["inst_prec_util"]).all()
    y = csr_matrix((g["data"], g["indices"], g["indptr"]), shape=tuple(g["shape"]))
    pred, meta = oracle.predict_using_bc_with_0approx(y, "f1", 5, metric_aggregation="sum", skip_tn=True, seed=0,
                                                      mix=(0.4, 5, y.shape[1]))
    assert (pred == g["csr_mix_f1_pred"]).all()
    assert (np.array(meta["utilities"]) == g["csr_mix_f1_util"]).all()


@pytest.mark.parametrize("name,metric,alpha", [("fw_mix_f1", "f1", 0.5), ("fw_mix_prec", "precision", 0.7)])
def test_fw_mixed(golden, oracle, name, metric, alpha):
    g = golden("mixed")
    eta = g["eta_fw"]
    a, b, p, meta = oracle.find_classifier_using_fw(eta, eta, metric, 5, max_iters=8, skip_tn=True, seed=0,
                                                    mix=(alpha, 5, eta.shape[1]))
    assert len(meta["utilities"]) == len(g[name + "_util"])
    assert np.allclose(meta["utilities"], g[name + "_util"], rtol=0, atol=1e-5)
    assert a.shape == g[name + "_a"].shape


@pytest.mark.parametrize("name,metric,skip_tn", [("tern_f1", "f1", True), ("tern_balacc", "balanced_accuracy", False)])
def test_fw_ternary_search(golden, oracle, name, metric, skip_tn):
    """alpha_search_algo="ternary" (utils.py:187-201, which narrows towards the SMALLER probe): the
    restatement follows it verbatim, so it stops where the live reference stops"""
    g = golden("extra")
    eta = g["eta"]
    a, b, p, meta = oracle.find_classifier_using_fw(eta, eta, metric, 5, max_iters=8, skip_tn=skip_tn, seed=0,
                                                    alpha_search_algo="ternary")
    assert len(meta["utilities"]) == len(g[name + "_util"]) and a.shape == g[name + "_a"].shape
    assert np.allclose(meta["utilities"], g[name + "_util"], rtol=0, atol=1e-5)


@pytest.mark.parametrize("name,metric,skip_tn", [("micro_f1", "f1", True), ("micro_balacc", "balanced_accuracy", False)])
def test_fw_micro(golden, oracle, name, metric, skip_tn):
    """micro-averaged objectives (frank_wolfe.py:758-832): metric of the summed confusion entries"""
    g = golden("extra")
    eta = g["eta"]
    a, b, p, meta = oracle.find_classifier_using_fw(eta, eta, metric, 5, max_iters=5, skip_tn=skip_tn, seed=0,
                                                    init_classifier="random", micro=True)
    assert np.allclose(meta["alphas"], g[name + "_alphas"], rtol=0, atol=1e-9)
    assert np.allclose(meta["utilities"], g[name + "_util"], rtol=0, atol=1e-6)
    assert a.shape == g[name + "_a"].shape and np.allclose(p, g[name + "_p"], atol=1e-6)


def test_fw_no_budget(golden, oracle):
    """k = 0: every label with eta * a + b >= 0 is predicted (frank_wolfe.py:601 with th = 0)"""
    g = golden("extra")
    eta = g["eta"]
    a, b, p, meta = oracle.find_classifier_using_fw(eta, eta, "f1", 0, max_iters=6, skip_tn=True, seed=0)
    assert np.allclose(meta["alphas"], g["fw_k0_alphas"], rtol=0, atol=1e-9)
    assert np.allclose(meta["utilities"], g["fw_k0_util"], rtol=0, atol=1e-5)


def test_fw_mixed_macro_recall_and_precision(golden, oracle):
    """frank_wolfe.py:917-938: sum_j [(1 - alpha) recall_j + alpha precision_j]"""
    g = golden("extra")
    eta = g["eta"]
    a, b, p, meta = oracle.find_classifier_using_fw(eta, eta, "precision", 5, max_iters=4, skip_tn=True, seed=0,
                                                    alpha_uniform_search_step=0.002, recall_precision_alpha=0.4)
    assert np.allclose(meta["alphas"], g["fw_rp_alphas"], rtol=0, atol=1e-9)
    assert np.allclose(meta["utilities"], g["fw_rp_util"], rtol=1e-6, atol=0)


OPT_BCA = [("opt_bca_nonorm", dict(normalize_conf_matrix=False, seed=0, skip_tn=True)),
           ("opt_bca_min", dict(maximize=False, seed=1, skip_tn=True, max_iters=3)),
           ("opt_bca_noshuffle_tn", dict(shuffle_order=False, seed=2, skip_tn=False))]
OPT_FW = [("opt_fw_nonorm", dict(normalize_conf_matrix=False, skip_tn=True, max_iters=4)),
          ("opt_fw_fixed", dict(search_for_best_alpha=False, skip_tn=True, max_iters=5)),
          ("opt_fw_tuple", dict(init_classifier="tuple", skip_tn=True, max_iters=4)),
          ("opt_fw_random", dict(init_classifier="random", skip_tn=True, max_iters=4, seed=5)),
          ("opt_fw_coarse", dict(alpha_uniform_search_step=0.01, skip_tn=True, max_iters=4))]


@pytest.mark.parametrize("name,kw", OPT_BCA, ids=[c[0] for c in OPT_BCA])
def test_bca_driver_options(golden, oracle, name, kw):
    """un-normalised confusion matrix (only instance 0 is visited, block_coordinate.py:403-414), minimisation,
    fixed order with true negatives: bit-equal to the live reference"""
    g = golden("extra")
    pred, meta = oracle.predict_using_bc_with_0approx(g["opt_eta"], "f1", 4, **kw)
    assert (_idx(pred, 4) == g[name + "_pred"]).all()
    assert (np.array(meta["utilities"]) == g[name + "_util"]).all()


@pytest.mark.parametrize("name,kw", OPT_FW, ids=[c[0] for c in OPT_FW])
def test_fw_driver_options(golden, oracle, name, kw):
    g = golden("extra")
    kw = dict(kw)
    if kw.get("init_classifier") == "tuple":
        kw["init_classifier"] = (g["opt_a0"], g["opt_b0"])
    kw.setdefault("seed", 0)
    a, b, p, meta = oracle.find_classifier_using_fw(g["eta"], g["eta"], "f1", 5, **kw)
    assert tuple(a.shape) == tuple(g[name + "_ashape"])
    assert np.allclose(meta["alphas"], g[name + "_alphas"], rtol=0, atol=1e-9)
    assert np.allclose(meta["utilities"], g[name + "_util"], rtol=0, atol=1e-5)
    assert np.allclose(p, g[name + "_p"], atol=1e-6)


@pytest.mark.parametrize("name,metric,skip_tn,etu", [("f1", "f1", True, False), ("f1_etu", "f1", True, True),
                                                       ("balacc", "balanced_accuracy", False, False),
                                                       ("gmean_etu", "gmean", False, True)])
def test_online_greedy_csr_golden(oracle, golden, name, metric, skip_tn, etu):
    """online / greedy steps on CSR rows: oracle == live reference, label ids and float64 state bit for bit"""
    from scipy.sparse import csr_matrix
    g = golden("online_csr")
    n, m, k = (int(v) for v in g["shape"])
    y = csr_matrix((g["y_data"], g["y_indices"], g["y_indptr"]), shape=(n, m))
    t = csr_matrix((g["t_data"], g["t_indices"], g["t_indptr"]), shape=(n, m))
    pred, state = oracle.online_greedy_csr(y, y if etu else t, k, metric, skip_tn=skip_tn)
    assert (pred == g[name + "_pred"]).all()
    assert np.array_equal(state, g[name + "_state"])


def _fw_generic_csr(g):
    from scipy.sparse import csr_matrix
    shape = tuple(g["shape"])
    y = csr_matrix((g["y_data"], g["y_indices"], g["y_indptr"]), shape=shape)
    yt = csr_matrix((g["yt_data"], g["yt_indices"], g["yt_indptr"]), shape=shape)
    return y, yt


def test_fw_no_budget_csr(golden, oracle):
    """k = 0 on CSR rows: the STORED labels with data * a + b >= 0 are predicted (numba_csr_functions.py:516-517,
    :631-653), in Frank-Wolfe (frank_wolfe.py:601 with th = 0) -- pinned to the live reference"""
    g = golden("fw_generic")
    y, yt = _fw_generic_csr(g)
    a, b, p, meta = oracle.find_classifier_using_fw(yt, y, "f1", 0, max_iters=8, skip_tn=True, seed=0)
    assert np.allclose(meta["alphas"], g["k0_f1_alphas"], rtol=0, atol=1e-9)
    assert np.allclose(meta["utilities"], g["k0_f1_util"], rtol=0, atol=1e-5)
    assert a.shape == g["k0_f1_a"].shape
    a, b, p, meta = oracle.find_classifier_using_fw(yt, y, "balanced_accuracy", 0, max_iters=5, seed=0)
    assert np.allclose(meta["alphas"], g["k0_balacc_alphas"], rtol=0, atol=1e-9)
    assert np.allclose(meta["utilities"], g["k0_balacc_util"], rtol=0, atol=1e-5)


def test_threshold_csr_rows_of_randomized_prediction(golden, oracle):
    """randomized-classifier prediction without a budget on CSR rows (frank_wolfe.py:130-172): per classifier the
    oracle's thresholded rows, rows drawn with numpy's Generator like the reference draws them"""
    g = golden("fw_generic")
    y, _ = _fw_generic_csr(g)
    n = y.shape[0]
    rng = np.random.default_rng(11)
    choice = np.array([rng.choice(np.arange(3), p=g["rnd_p"]) for _ in range(n)])
    per_cls = [oracle.predict_weighted_per_instance(y, 0, th=0.0, a=g["rnd_a"][c], b=g["rnd_b"][c]) for c in range(3)]
    idx, ptr = [], [0]
    for i in range(n):
        r = per_cls[choice[i]]
        idx.extend(r.indices[r.indptr[i]:r.indptr[i + 1]])
        ptr.append(len(idx))
    assert (np.array(ptr) == g["rnd_pred_indptr"]).all()
    assert (np.array(idx) == g["rnd_pred_indices"]).all()
