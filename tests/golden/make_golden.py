"""Generate golden vectors by running the LIVE reference (mwydmuch/xCOLUMNs 0.0.3).

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the reference unmodified, with two environment shims that do not touch its
arithmetic:
  * ``np.product = np.prod``        (numpy >= 2 removed the alias used at block_coordinate.py:663)
  * a stub ``autograd`` package     (absent here; grad() is served by torch float64 autograd)
and stores inputs + outputs as small .npz files next to this script.  tests/test_oracle_golden.py
replays them against oracle/ (bit-exact where the path is deterministic), and the -m gpu tests
replay them against the CUDA path.
"""
import os
import sys
import tempfile

import numpy as np
from scipy.sparse import csr_matrix

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

REF = os.environ.get("XC_REFERENCE", "/root/reference")


def _install_shims():
    np.product = np.prod
    stub = tempfile.mkdtemp(prefix="autograd_stub_")
    os.makedirs(os.path.join(stub, "autograd"))
    with open(os.path.join(stub, "autograd", "__init__.py"), "w") as f:
        f.write(
            "import numpy as _np\n"
            "import torch as _torch\n"
            "def grad(fun, argnum=0):\n"
            "    argnums = list(argnum) if isinstance(argnum, (list, tuple)) else [argnum]\n"
            "    def g(*args):\n"
            "        ts = [_torch.tensor(_np.asarray(a, dtype=_np.float64), requires_grad=(i in argnums)) for i, a in enumerate(args)]\n"
            "        val = fun(*ts)\n"
            "        gs = _torch.autograd.grad(val, [ts[i] for i in argnums], allow_unused=True, materialize_grads=True)\n"
            "        return tuple(x.detach().numpy() for x in gs)\n"
            "    return g\n")
    with open(os.path.join(stub, "autograd", "numpy.py"), "w") as f:
        f.write("from numpy import *\nimport numpy as _np\nrandom = _np.random\n")
    sys.path.insert(0, stub)
    sys.path.insert(0, REF)


def pred_to_idx(y_pred, k):
    """(n, k) ascending label ids of a 0/1 prediction matrix with exactly k ones per row."""
    if isinstance(y_pred, csr_matrix):
        n = y_pred.shape[0]
        return np.asarray(y_pred.indices, dtype=np.int32).reshape(n, k).copy()
    n = y_pred.shape[0]
    r, c = np.nonzero(y_pred)
    assert (np.bincount(r, minlength=n) == k).all()
    return c.reshape(n, k).astype(np.int32)


def main():
    _install_shims()
    from xcolumns import block_coordinate as bc
    from xcolumns import confusion_matrix as cm
    from xcolumns import frank_wolfe as fw
    from xcolumns import metrics as mt
    from xcolumns import weighted_prediction as wp

    from xcolumns_b200.synth import csr_probs, dense_probs

    def save(name, **kw):
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **kw)
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")

    rng = np.random.default_rng(7)

    # ---------------- weighted top-k, dense ----------------
    eta = dense_probs(96, 333, seed=11)
    a32 = (0.5 + rng.random(333)).astype(np.float32)
    b32 = (0.01 * rng.standard_normal(333)).astype(np.float32)
    a64 = 1.0 / (eta.mean(0).astype(np.float64) + 1e-6)
    out = {"eta": eta, "a32": a32, "b32": b32, "a64": a64}
    out["top5"] = pred_to_idx(wp.predict_top_k(eta, 5), 5)
    out["top1"] = pred_to_idx(wp.predict_top_k(eta, 1), 1)
    out["w5_ab32"] = pred_to_idx(wp.predict_weighted_per_instance(eta, 5, a=a32, b=b32), 5)
    out["w3_a64"] = pred_to_idx(wp.predict_weighted_per_instance(eta, 3, a=a64), 3)
    ks = wp.predict_weighted_per_instance(eta, 4, a=a32, b=b32, keep_scores=True)
    out["w4_scores"] = ks
    eta64 = eta.astype(np.float64)
    out["w5_f64"] = pred_to_idx(wp.predict_weighted_per_instance(eta64, 5, a=a64, b=b32.astype(np.float64)), 5)
    out["th0"] = wp.predict_weighted_per_instance(eta, 0, th=0.3, a=a32, b=b32)
    save("topk_dense", **out)

    # ---------------- weighted top-k, CSR (ragged rows, some with nnz <= k) ----------------
    ycsr = csr_probs(80, 1500, 24, seed=12, ragged=True)
    ac = (0.5 + rng.random(1500)).astype(np.float64)
    bcf = (0.01 * rng.standard_normal(1500)).astype(np.float64)
    out = {"data": ycsr.data, "indices": ycsr.indices, "indptr": ycsr.indptr, "shape": np.array(ycsr.shape),
           "a": ac, "b": bcf}
    for name, kw, k in (("top5", {}, 5), ("top8", {}, 8), ("w5", {"a": ac, "b": bcf}, 5),
                        ("w5s", {"a": ac, "b": bcf, "keep_scores": True}, 5)):
        r = wp.predict_weighted_per_instance(ycsr, k, **kw)
        out[name + "_data"], out[name + "_indices"], out[name + "_indptr"] = r.data, r.indices, r.indptr
    save("topk_csr", **out)

    # ---------------- confusion matrix ----------------
    eta = dense_probs(200, 150, seed=13)
    pred = wp.predict_top_k(eta, 4)
    lab = (rng.random(eta.shape) < eta).astype(np.float32)
    out = {"eta": eta, "pred": pred_to_idx(pred, 4), "lab": lab}
    for name, yt, kw in (("probs_f64", eta, dict(dtype=np.float64, skip_tn=True)),
                         ("probs_none", eta, dict()),
                         ("lab_norm", lab, dict(normalize=True)),
                         ("lab_f64_norm_skip", lab, dict(normalize=True, skip_tn=True, dtype=np.float64)),
                         ("lab_axis1", lab, dict(axis=1, dtype=np.float64))):
        c = cm.calculate_confusion_matrix(yt, pred, **kw)
        out[name] = np.stack([np.asarray(v, dtype=np.float64) for v in c])
    ycsr = csr_probs(150, 900, 20, seed=14)
    pcsr = wp.predict_top_k(ycsr, 4)
    out.update({"c_data": ycsr.data, "c_indices": ycsr.indices, "c_indptr": ycsr.indptr,
                "c_shape": np.array(ycsr.shape), "c_pred": pred_to_idx(pcsr, 4)})
    for name, kw in (("csr_f64", dict(dtype=np.float64, skip_tn=True)), ("csr_none", dict()),
                     ("csr_norm", dict(normalize=True, dtype=np.float64))):
        c = cm.calculate_confusion_matrix(ycsr, pcsr, **kw)
        out[name] = np.stack([np.asarray(v, dtype=np.float64) for v in c])
    save("confmat", **out)

    # ---------------- BCA dense ----------------
    eta = dense_probs(300, 200, seed=1001)
    out = {"eta": eta}

    def run_bc(name, y, metric, k, **kw):
        yp, meta = bc.predict_using_bc_with_0approx(y, metric, k, return_meta=True, **kw)
        if k > 0:
            out[name + "_pred"] = pred_to_idx(yp, k)
        else:
            out[name + "_pred"] = np.asarray(yp != 0, dtype=np.uint8)
        out[name + "_util"] = np.array(meta["utilities"], dtype=np.float64)
        print(f"  {name}: iters={meta['iters']} util={meta['utilities'][-1]:.6f} time={meta['time']:.2f}s")

    run_bc("f1", eta, mt.binary_f1_score_on_conf_matrix, 5, seed=0, skip_tn=True)
    run_bc("f1_f64", eta.astype(np.float64), mt.binary_f1_score_on_conf_matrix, 5, seed=0, skip_tn=True)
    run_bc("recall", eta, mt.binary_recall_on_conf_matrix, 5, seed=3, skip_tn=True)
    run_bc("precision", eta, mt.binary_precision_on_conf_matrix, 3, seed=4, skip_tn=True)
    run_bc("jaccard", eta, mt.binary_jaccard_score_on_conf_matrix, 5, seed=5, skip_tn=True)
    run_bc("fbeta2", eta, mt.binary_fbeta_score_on_conf_matrix, 5, seed=6, skip_tn=True,
           metric_kwargs={"beta": 2.0, "epsilon": 1e-6})
    run_bc("balacc", eta, mt.binary_balanced_accuracy_on_conf_matrix, 5, seed=7, skip_tn=False)
    run_bc("gmean", eta, mt.binary_gmean_on_conf_matrix, 5, seed=8, skip_tn=False)
    run_bc("hmean", eta, mt.binary_hmean_on_conf_matrix, 5, seed=9, skip_tn=False)
    run_bc("f1_noshuffle", eta, mt.binary_f1_score_on_conf_matrix, 5, seed=0, skip_tn=True, shuffle_order=False)
    run_bc("f1_random", eta, mt.binary_f1_score_on_conf_matrix, 5, seed=10, skip_tn=True, init_y_pred="random")
    run_bc("f1_greedy", eta, mt.binary_f1_score_on_conf_matrix, 5, seed=11, skip_tn=True, init_y_pred="greedy")
    run_bc("f1_sum", eta, mt.binary_f1_score_on_conf_matrix, 5, seed=12, skip_tn=True, metric_aggregation="sum",
           tolerance=1e-4)
    run_bc("f1_k0", eta, mt.binary_f1_score_on_conf_matrix, 0, seed=13, skip_tn=True, init_y_pred=wp.predict_top_k(eta, 3))
    yp, meta = bc.predict_optimizing_macro_f1_score_using_bc(eta, 5, seed=0, return_meta=True)
    assert (pred_to_idx(yp, 5) == out["f1_pred"]).all()
    save("bca_dense", **out)

    # ---------------- BCA CSR ----------------
    ycsr = csr_probs(300, 3000, 40, seed=1004)
    out = {"data": ycsr.data, "indices": ycsr.indices, "indptr": ycsr.indptr, "shape": np.array(ycsr.shape)}
    run_bc("f1", ycsr, mt.binary_f1_score_on_conf_matrix, 5, seed=0, skip_tn=True)
    run_bc("recall", ycsr, mt.binary_recall_on_conf_matrix, 5, seed=1, skip_tn=True)
    run_bc("jaccard", ycsr, mt.binary_jaccard_score_on_conf_matrix, 3, seed=2, skip_tn=True)
    y64 = csr_matrix((ycsr.data.astype(np.float64), ycsr.indices, ycsr.indptr), shape=ycsr.shape)
    run_bc("f1_f64", y64, mt.binary_f1_score_on_conf_matrix, 5, seed=0, skip_tn=True)
    run_bc("balacc", ycsr, mt.binary_balanced_accuracy_on_conf_matrix, 5, seed=3, skip_tn=False)
    run_bc("hmean", ycsr, mt.binary_hmean_on_conf_matrix, 5, seed=4, skip_tn=False)
    save("bca_csr", **out)

    # ---------------- coverage BCA (CSR) ----------------
    ycsr = csr_probs(400, 600, 30, seed=1005)
    out = {"data": ycsr.data, "indices": ycsr.indices, "indptr": ycsr.indptr, "shape": np.array(ycsr.shape)}
    for name, kw in (("cov", dict(seed=0)), ("cov_a07", dict(seed=1, alpha=0.7))):
        yp, meta = bc.predict_optimizing_coverage_using_bc(ycsr, 5, return_meta=True, **kw)
        out[name + "_pred"] = pred_to_idx(yp, 5)
        out[name + "_util"] = np.array(meta["utilities"], dtype=np.float64)
        print(f"  {name}: iters={meta['iters']} util={meta['utilities']}")
    save("coverage_csr", **out)

    # ---------------- Frank-Wolfe ----------------
    eta = dense_probs(400, 300, seed=1005)
    out = {"eta": eta}

    def run_fw(name, yt, yp, metric, **kw):
        clf, meta = fw.find_classifier_using_fw(yt, yp, metric, 5, return_meta=True, **kw)
        out[name + "_a"], out[name + "_b"], out[name + "_p"] = clf.a, clf.b, clf.p
        out[name + "_alphas"] = np.array(meta["alphas"], dtype=np.float64)
        out[name + "_util"] = np.array([float(u) for u in meta["utilities"]], dtype=np.float64)
        out[name + "_cutil"] = np.array([float(u) for u in meta["classifiers_utilities"]], dtype=np.float64)
        out[name + "_iters"] = np.array(meta["iters"])
        print(f"  {name}: iters={meta['iters']} util={out[name + '_util']}")

    run_fw("f1_proba", eta, eta, mt.macro_f1_score_on_conf_matrix, max_iters=10, skip_tn=True, seed=0)
    lab = (np.random.default_rng(5).random(eta.shape) < eta).astype(np.float32)
    out["lab"] = lab
    run_fw("f1_lab", lab, eta, mt.macro_f1_score_on_conf_matrix, max_iters=10, skip_tn=True, seed=0,
           metric_kwargs={"epsilon": 1e-4})
    run_fw("recall_lab", lab, eta, mt.macro_recall_on_conf_matrix, max_iters=6, skip_tn=True, seed=0,
           metric_kwargs={"epsilon": 1e-4})
    run_fw("balacc_lab", lab, eta, mt.macro_balanced_accuracy_on_conf_matrix, max_iters=6, seed=0,
           metric_kwargs={"epsilon": 1e-4})
    save("fw_dense", **out)


def main_mixed():
    """mixed.npz: instance precision and the mixed instance-precision / macro-metric wrappers
    (block_coordinate.py:804-1045, frank_wolfe.py:838-915) on the live reference."""
    _install_shims()
    from xcolumns import block_coordinate as bc
    from xcolumns import frank_wolfe as fw

    from xcolumns_b200.synth import csr_probs, dense_probs

    eta = dense_probs(300, 200, seed=1001)
    ycsr = csr_probs(300, 3000, 40, seed=1004)
    out = {"eta": eta, "data": ycsr.data, "indices": ycsr.indices, "indptr": ycsr.indptr, "shape": np.array(ycsr.shape)}

    def run(name, fn, y, k, **kw):
        yp, meta = fn(y, k, return_meta=True, **kw)
        out[name + "_pred"] = pred_to_idx(yp, k)
        out[name + "_util"] = np.array(meta["utilities"], dtype=np.float64)
        print(f"  {name}: iters={meta['iters']} util={meta['utilities'][-1]:.9f}")

    run("mix_f1", bc.predict_optimizing_mixed_instance_precision_and_macro_f1_score_using_bc, eta, 5, alpha=0.3, seed=0)
    run("mix_prec", bc.predict_optimizing_mixed_instance_precision_and_macro_precision_using_bc, eta, 3, alpha=0.5, seed=1)
    run("mix_recall", bc.predict_optimizing_mixed_instance_precision_and_macro_recall_using_bc, eta, 5, alpha=0.8, seed=2)
    run("mix_jaccard", bc.predict_optimizing_mixed_instance_precision_and_macro_jaccard_score_using_bc, eta, 5, alpha=0.5,
        seed=3)
    run("mix_balacc", bc.predict_optimizing_mixed_instance_precision_and_macro_balanced_accuracy_using_bc, eta, 5,
        alpha=0.5, seed=4)
    run("mix_f1_eps", bc.predict_optimizing_mixed_instance_precision_and_macro_f1_score_using_bc, eta, 5, alpha=0.6,
        seed=5, metric_kwargs={"epsilon": 1e-5})
    run("inst_prec", bc.predict_optimizing_instance_precision_using_bc, eta, 5, seed=6)
    run("csr_mix_f1", bc.predict_optimizing_mixed_instance_precision_and_macro_f1_score_using_bc, ycsr, 5, alpha=0.4,
        seed=0)

    eta2 = dense_probs(400, 300, seed=1005)
    out["eta_fw"] = eta2
    for name, fn, al in (("fw_mix_f1", fw.find_classifier_optimizing_mixed_instance_precision_and_macro_f1_score_using_fw, 0.5),
                         ("fw_mix_prec", fw.find_classifier_optimizing_mixed_instance_precision_and_macro_precision_using_fw, 0.7)):
        clf, meta = fn(eta2, eta2, 5, alpha=al, max_iters=8, skip_tn=True, seed=0, return_meta=True)
        out[name + "_a"], out[name + "_b"], out[name + "_p"] = clf.a, clf.b, clf.p
        out[name + "_alphas"] = np.array(meta["alphas"], dtype=np.float64)
        out[name + "_util"] = np.array([float(u) for u in meta["utilities"]], dtype=np.float64)
        print(f"  {name}: iters={meta['iters']} util={out[name + '_util']}")
    path = os.path.join(HERE, "mixed.npz")
    np.savez_compressed(path, **out)
    print(f"mixed: {os.path.getsize(path) / 1024:.0f} KiB")


def main_extra():
    """extra.npz: Frank-Wolfe with alpha_search_algo="ternary" (utils.py:187-201) on the live reference."""
    _install_shims()
    from xcolumns import frank_wolfe as fw
    from xcolumns import metrics as mt

    from xcolumns_b200.synth import dense_probs

    eta = dense_probs(400, 300, seed=1005)
    out = {"eta": eta}
    for name, metric, kw in (("tern_f1", mt.macro_f1_score_on_conf_matrix, dict(skip_tn=True)),
                             ("tern_balacc", mt.macro_balanced_accuracy_on_conf_matrix, dict())):
        clf, meta = fw.find_classifier_using_fw(eta, eta, metric, 5, max_iters=8, seed=0, alpha_search_algo="ternary",
                                                return_meta=True, **kw)
        out[name + "_a"], out[name + "_b"], out[name + "_p"] = clf.a, clf.b, clf.p
        out[name + "_alphas"] = np.array(meta["alphas"], dtype=np.float64)
        out[name + "_util"] = np.array([float(u) for u in meta["utilities"]], dtype=np.float64)
        print(f"  {name}: iters={meta['iters']} alphas={out[name + '_alphas']} util={out[name + '_util']}")
    # online / greedy steps (block_coordinate.py:132-209 with greedy=True, only_pred=True, then
    # confusion_matrix.py:402-435), driven like experiments/omma_wrappers_online_methods.py:192-270
    from xcolumns import block_coordinate as bc
    from xcolumns import confusion_matrix as cm
    eo = dense_probs(250, 180, seed=33)
    lab = (np.random.default_rng(9).random(eo.shape) < eo).astype(np.float32)
    out.update({"on_eta": eo, "on_lab": lab})
    for name, metric, skip_tn, etu in (("on_f1", mt.binary_f1_score_on_conf_matrix, True, False),
                                       ("on_f1_etu", mt.binary_f1_score_on_conf_matrix, True, True),
                                       ("on_balacc", mt.binary_balanced_accuracy_on_conf_matrix, False, False)):
        m_ = eo.shape[1]
        C_ = cm.ConfusionMatrix(*[np.full(m_, 1e-6, dtype=np.float64) for _ in range(4)])
        yp = np.zeros_like(eo)
        yt = eo if etu else lab
        for i in range(eo.shape[0]):
            bc._bc_with_0approx_step_dense(eo, yp, i, C_.tp, C_.fp, C_.fn, C_.tn, 5, metric, greedy=True,
                                           skip_tn=skip_tn, only_pred=True)
            cm._update_unnormalized_confusion_matrix(C_, yt[i], yp[i], skip_tn=skip_tn)
        out[name + "_pred"] = pred_to_idx(yp, 5)
        out[name + "_state"] = np.stack([C_.tp, C_.fp, C_.fn, C_.tn])
        print(f"  {name}: tp sum {C_.tp.sum():.6f}")
    # Frank-Wolfe without a budget (k = 0: every label with a non-negative gain)
    clf, meta = fw.find_classifier_using_fw(eta, eta, mt.macro_f1_score_on_conf_matrix, 0, max_iters=6, skip_tn=True,
                                            seed=0, return_meta=True)
    out["fw_k0_a"], out["fw_k0_b"], out["fw_k0_p"] = clf.a, clf.b, clf.p
    out["fw_k0_alphas"] = np.array(meta["alphas"], dtype=np.float64)
    out["fw_k0_util"] = np.array([float(u) for u in meta["utilities"]], dtype=np.float64)
    yk0 = clf.predict(eta, seed=3)
    out["fw_k0_pred_rowsums"] = np.asarray(yk0.sum(1), dtype=np.float64)
    out["fw_k0_pred"] = np.asarray(yk0 != 0, dtype=np.uint8)
    print(f"  fw_k0: iters={meta['iters']} alphas={out['fw_k0_alphas']} util={out['fw_k0_util']}")
    # driver options: un-normalised confusion matrix (quirk: only instance 0 is visited, block_coordinate.py
    # :403-414), minimisation, fixed step sizes, explicit / random initial classifiers
    from xcolumns import block_coordinate as bc2
    eo2 = dense_probs(120, 90, seed=44)
    out["opt_eta"] = eo2
    for name, kw in (("opt_bca_nonorm", dict(normalize_conf_matrix=False, seed=0, skip_tn=True)),
                     ("opt_bca_min", dict(maximize=False, seed=1, skip_tn=True, max_iters=3)),
                     ("opt_bca_noshuffle_tn", dict(shuffle_order=False, seed=2, skip_tn=False))):
        yp, meta = bc2.predict_using_bc_with_0approx(eo2, mt.binary_f1_score_on_conf_matrix, 4, return_meta=True, **kw)
        out[name + "_pred"] = pred_to_idx(yp, 4)
        out[name + "_util"] = np.array(meta["utilities"], dtype=np.float64)
        print(f"  {name}: iters={meta['iters']} util={meta['utilities']}")
    rng_ab = np.random.default_rng(11)
    a0 = (0.5 + rng_ab.random(eta.shape[1])).astype(np.float32)
    b0 = (0.1 * rng_ab.standard_normal(eta.shape[1])).astype(np.float32)
    out["opt_a0"], out["opt_b0"] = a0, b0
    for name, kw in (("opt_fw_nonorm", dict(normalize_conf_matrix=False, skip_tn=True, max_iters=4)),
                     ("opt_fw_fixed", dict(search_for_best_alpha=False, skip_tn=True, max_iters=5)),
                     ("opt_fw_tuple", dict(init_classifier=(a0, b0), skip_tn=True, max_iters=4)),
                     ("opt_fw_random", dict(init_classifier="random", skip_tn=True, max_iters=4, seed=5)),
                     ("opt_fw_coarse", dict(alpha_uniform_search_step=0.01, skip_tn=True, max_iters=4))):
        kw.setdefault("seed", 0)
        clf, meta = fw.find_classifier_using_fw(eta, eta, mt.macro_f1_score_on_conf_matrix, 5, return_meta=True, **kw)
        out[name + "_p"] = clf.p
        out[name + "_ashape"] = np.array(clf.a.shape)
        out[name + "_alphas"] = np.array(meta["alphas"], dtype=np.float64)
        out[name + "_util"] = np.array([float(u) for u in meta["utilities"]], dtype=np.float64)
        print(f"  {name}: iters={meta['iters']} alphas={out[name + '_alphas']} util={out[name + '_util']}")
    # mixed macro recall / macro precision (frank_wolfe.py:917-938)
    clf, meta = fw.find_classifier_optimizing_mixed_macro_recall_and_macro_precision_using_fw(
        eta, eta, 5, alpha=0.4, max_iters=4, skip_tn=True, seed=0, alpha_uniform_search_step=0.002, return_meta=True)
    out["fw_rp_a"], out["fw_rp_p"] = clf.a, clf.p
    out["fw_rp_alphas"] = np.array(meta["alphas"], dtype=np.float64)
    out["fw_rp_util"] = np.array([float(u) for u in meta["utilities"]], dtype=np.float64)
    print(f"  fw_rp: iters={meta['iters']} alphas={out['fw_rp_alphas']} util={out['fw_rp_util']}")
    # micro-averaged Frank-Wolfe objectives (frank_wolfe.py:758-832)
    for name, fn, kw in (("micro_f1", fw.find_classifier_optimizing_micro_f1_score_using_fw, {}),
                         ("micro_balacc", fw.find_classifier_optimizing_micro_balanced_accuracy_using_fw, {})):
        clf, meta = fn(eta, eta, 5, max_iters=5, seed=0, init_classifier="random", return_meta=True, **kw)
        out[name + "_a"], out[name + "_b"], out[name + "_p"] = clf.a, clf.b, clf.p
        out[name + "_alphas"] = np.array(meta["alphas"], dtype=np.float64)
        out[name + "_util"] = np.array([float(u) for u in meta["utilities"]], dtype=np.float64)
        print(f"  {name}: iters={meta['iters']} alphas={out[name + '_alphas']} util={out[name + '_util']}")
    # closed-form weighted strategies (weighted_prediction.py:223-560)
    from xcolumns import weighted_prediction as wp
    from xcolumns_b200.synth import csr_probs
    etaw = dense_probs(120, 400, seed=21)
    pri = etaw.mean(0).astype(np.float64)
    ycsr = csr_probs(90, 1200, 30, seed=22, ragged=True)
    pric = np.asarray(ycsr.mean(0)).ravel().astype(np.float64) + 1e-3
    out.update({"w_eta": etaw, "w_pri": pri, "w_data": ycsr.data, "w_indices": ycsr.indices, "w_indptr": ycsr.indptr,
                "w_shape": np.array(ycsr.shape), "w_pric": pric})
    out["w_recall"] = pred_to_idx(wp.predict_optimizing_macro_recall(etaw, 5, pri), 5)
    out["w_balacc"] = pred_to_idx(wp.predict_optimizing_macro_balanced_accuracy(etaw, 5, pri), 5)
    out["w_balacc_k0"] = np.asarray(wp.predict_optimizing_macro_balanced_accuracy(etaw, 0, pri) != 0, dtype=np.uint8)
    out["w_log"] = pred_to_idx(wp.predict_log_weighted_per_instance(etaw, 4, pri), 4)
    out["w_pow"] = pred_to_idx(wp.predict_power_law_weighted_per_instance(etaw, 5, pri, 0.5), 5)
    out["w_psp"] = pred_to_idx(wp.predict_optimizing_instance_propensity_scored_precision(
        etaw, 3, propensities=(0.2 + 0.8 * pri / pri.max()).copy()), 3)
    r = wp.predict_optimizing_macro_balanced_accuracy(ycsr, 5, pric)
    out["w_balacc_csr_data"], out["w_balacc_csr_indices"], out["w_balacc_csr_indptr"] = r.data, r.indices, r.indptr
    r = wp.predict_optimizing_macro_recall(ycsr, 5, pric)
    out["w_recall_csr_data"], out["w_recall_csr_indices"], out["w_recall_csr_indptr"] = r.data, r.indices, r.indptr
    path = os.path.join(HERE, "extra.npz")
    np.savez_compressed(path, **out)
    print(f"extra: {os.path.getsize(path) / 1024:.0f} KiB")


def main_online_csr():
    """online_csr.npz: the online / greedy steps on CSR rows (block_coordinate.py:212-293 with greedy=True,
    only_pred=True, then confusion_matrix.py:421-432), driven like experiments/omma_wrappers_online_methods.py:223-266
    on the live reference."""
    _install_shims()
    from xcolumns import block_coordinate as bc
    from xcolumns import confusion_matrix as cm
    from xcolumns import metrics as mt

    from xcolumns_b200.synth import csr_probs

    n, m, k = 300, 120, 5
    y = csr_probs(n, m, 18, seed=71)
    rng = np.random.default_rng(13)
    # true labels: Bernoulli draws on the stored entries plus a few positives the probability row does not store
    keep = rng.random(y.nnz) < y.data
    rows = np.repeat(np.arange(n), np.diff(y.indptr))
    extra_r, extra_c = rng.integers(0, n, 60), rng.integers(0, m, 60)
    t = csr_matrix((np.ones(keep.sum() + 60, dtype=np.float32),
                    (np.concatenate([rows[keep], extra_r]), np.concatenate([y.indices[keep], extra_c]))), shape=(n, m))
    t.sum_duplicates()
    t.data[:] = 1.0
    t.sort_indices()
    out = {"y_data": y.data, "y_indices": y.indices, "y_indptr": y.indptr, "t_data": t.data, "t_indices": t.indices,
           "t_indptr": t.indptr, "shape": np.array([n, m, k])}
    for name, metric, skip_tn, etu in (("f1", mt.binary_f1_score_on_conf_matrix, True, False),
                                       ("f1_etu", mt.binary_f1_score_on_conf_matrix, True, True),
                                       ("balacc", mt.binary_balanced_accuracy_on_conf_matrix, False, False),
                                       ("gmean_etu", mt.binary_gmean_on_conf_matrix, False, True)):
        C_ = cm.ConfusionMatrix(*[np.full(m, 1e-6, dtype=np.float64) for _ in range(4)])
        yp = csr_matrix((n, m), dtype=np.float32)
        yt = y if etu else t
        for i in range(n):
            bc._bc_with_0approx_step_csr(y, yp, i, C_.tp, C_.fp, C_.fn, C_.tn, k, metric, greedy=True, skip_tn=skip_tn,
                                         only_pred=True)
            cm._update_unnormalized_confusion_matrix(C_, yt[i], yp[i], skip_tn=skip_tn)
        assert (np.diff(yp.indptr) == k).all()
        out[name + "_pred"] = np.asarray(yp.indices[:n * k], dtype=np.int32).reshape(n, k).copy()
        out[name + "_state"] = np.stack([C_.tp, C_.fp, C_.fn, C_.tn])
        print(f"  online csr {name}: tp sum {C_.tp.sum():.6f} tn sum {C_.tn.sum():.6f}")
    np.savez_compressed(os.path.join(HERE, "online_csr.npz"), **out)


def custom_fmeasure_like(tp, fp, fn, tn, gamma=0.3):
    """a metric none of the libraries ships (used as the "arbitrary callable" of the goldens and of the tests)"""
    return (tp + gamma * tp * tp) / (tp + 0.5 * fp + 0.7 * fn + 1e-6)


def custom_with_tn(tp, fp, fn, tn):
    return tp / (tp + fn + 1e-7) - 0.25 * fp / (fp + tn + 1e-7)


def main_callables():
    """callables.npz: predict_using_bc_with_0approx with an arbitrary callable and with a list of m callables
    (block_coordinate.py:54-129) on the live reference."""
    _install_shims()
    from functools import partial

    from xcolumns import block_coordinate as bc

    from xcolumns_b200.synth import dense_probs

    eta = dense_probs(260, 90, seed=52)
    out = {"eta": eta}
    m = eta.shape[1]
    cases = {
        "custom": (custom_fmeasure_like, dict(seed=0, skip_tn=True, metric_kwargs={"gamma": 0.3})),
        "custom_tn_sum": (custom_with_tn, dict(seed=1, skip_tn=False, metric_aggregation="sum")),
        "custom_min": (custom_fmeasure_like, dict(seed=2, skip_tn=True, maximize=False, max_iters=3)),
        # a list of m callables: two distinct functions alternating over the labels
        "list": ([custom_fmeasure_like if j % 2 == 0 else partial(custom_fmeasure_like, gamma=0.0) for j in range(m)],
                 dict(seed=3, skip_tn=True)),
    }
    for name, (func, kw) in cases.items():
        yp, meta = bc.predict_using_bc_with_0approx(eta, func, 4, return_meta=True, **kw)
        out[name + "_pred"] = pred_to_idx(yp, 4)
        out[name + "_util"] = np.array([float(u) for u in meta["utilities"]], dtype=np.float64)
        print(f"  callables {name}: iters={meta['iters']} util={out[name + '_util']}")
    np.savez_compressed(os.path.join(HERE, "callables.npz"), **out)


def custom_tversky(tp, fp, fn, tn, gamma=0.5):
    """a macro objective the library does not ship (Tversky index): plain arithmetic + .mean(), so it runs on numpy
    arrays, autograd boxes and torch tensors alike"""
    return ((1 + gamma) * tp / ((1 + gamma) * tp + gamma * fp + fn + 1e-6)).mean()


def custom_tpr_tnr(tp, fp, fn, tn):
    return (tp / (tp + fn + 1e-6) * tn / (tn + fp + 1e-6)).mean()


def main_fw_generic():
    """fw_generic.npz: (a) Frank-Wolfe and randomized-classifier prediction WITHOUT a budget (k = 0) on CSR inputs
    (frank_wolfe.py:130-172, numba_csr_functions.py:516-517, :631-653); (b) find_classifier_using_fw with objective
    callables that are not built-in metrics (frank_wolfe.py:368-376), dense and CSR -- all on the live reference."""
    _install_shims()
    from xcolumns import frank_wolfe as fw
    from xcolumns import metrics as mt

    from xcolumns_b200.synth import csr_probs, dense_probs

    out = {}
    y = csr_probs(300, 400, 12, seed=3)
    rng = np.random.default_rng(1)
    yt = y.copy()
    yt.data = (rng.random(y.nnz) < y.data).astype(y.data.dtype)
    yt.eliminate_zeros()
    for nm, mat in (("y", y), ("yt", yt)):
        out[nm + "_data"], out[nm + "_indices"], out[nm + "_indptr"] = mat.data, mat.indices, mat.indptr
    out["shape"] = np.array(y.shape)

    def store(name, clf, meta):
        out[name + "_a"], out[name + "_b"], out[name + "_p"] = np.asarray(clf.a), np.asarray(clf.b), np.asarray(clf.p)
        out[name + "_alphas"] = np.array(meta["alphas"], dtype=np.float64)
        out[name + "_util"] = np.array([float(u) for u in meta["utilities"]], dtype=np.float64)
        out[name + "_cutil"] = np.array([float(u) for u in meta["classifiers_utilities"]], dtype=np.float64)
        out[name + "_iters"] = np.array(meta["iters"])
        print(f"  {name}: iters={meta['iters']} alphas={out[name + '_alphas']} util={out[name + '_util']}")

    # (a) k = 0 on CSR rows
    clf, meta = fw.find_classifier_using_fw(yt, y, mt.macro_f1_score_on_conf_matrix, 0, max_iters=8, seed=0,
                                            skip_tn=True, return_meta=True)
    store("k0_f1", clf, meta)
    yp = clf.predict(y, seed=5)
    yp.sort_indices()
    out["k0_f1_pred_indices"], out["k0_f1_pred_indptr"] = yp.indices[:yp.indptr[-1]].copy(), yp.indptr.copy()
    clf, meta = fw.find_classifier_using_fw(yt, y, mt.macro_balanced_accuracy_on_conf_matrix, 0, max_iters=5, seed=0,
                                            return_meta=True)
    store("k0_balacc", clf, meta)
    # randomized prediction without a budget from hand-made classifiers (weights in the data dtype: the reference's
    # numba step cannot unify float32 rows with float64 weights)
    ra = (0.5 + rng.random((3, 400))).astype(np.float32)
    rb = (0.4 * rng.standard_normal((3, 400)) - 0.3).astype(np.float32)
    rp = np.array([0.5, 0.3, 0.2])
    yp = fw.predict_using_randomized_weighted_classifier(y, 0, ra, rb, rp, seed=11)
    yp.sort_indices()
    out["rnd_a"], out["rnd_b"], out["rnd_p"] = ra, rb, rp
    out["rnd_pred_indices"], out["rnd_pred_indptr"] = yp.indices[:yp.indptr[-1]].copy(), yp.indptr.copy()
    print(f"  rnd k0: nnz={yp.indptr[-1]}")

    # (b) objectives that are not built-in metrics
    eta = dense_probs(300, 200, seed=77)
    lab = (np.random.default_rng(6).random(eta.shape) < eta).astype(np.float32)
    out["eta"], out["lab"] = eta, lab
    clf, meta = fw.find_classifier_using_fw(lab, eta, custom_tversky, 4, max_iters=8, seed=0, skip_tn=True,
                                            metric_kwargs={"gamma": 0.7}, return_meta=True)
    store("tversky", clf, meta)
    clf, meta = fw.find_classifier_using_fw(lab, eta, custom_tpr_tnr, 4, max_iters=6, seed=0, return_meta=True)
    store("tpr_tnr", clf, meta)
    clf, meta = fw.find_classifier_using_fw(lab, eta, custom_tversky, 4, max_iters=6, seed=0, skip_tn=True,
                                            alpha_search_algo="ternary", return_meta=True)
    store("tversky_ternary", clf, meta)
    clf, meta = fw.find_classifier_using_fw(lab, eta, custom_tversky, 4, max_iters=5, seed=0, skip_tn=True,
                                            search_for_best_alpha=False, return_meta=True)
    store("tversky_fixed", clf, meta)
    clf, meta = fw.find_classifier_using_fw(yt, y, custom_tversky, 3, max_iters=6, seed=0, skip_tn=True,
                                            return_meta=True)
    store("tversky_csr", clf, meta)
    np.savez_compressed(os.path.join(HERE, "fw_generic.npz"), **out)
    print(f"fw_generic: {os.path.getsize(os.path.join(HERE, 'fw_generic.npz')) / 1024:.0f} KiB")


def main_callables_csr():
    """callables_csr.npz: predict_using_bc_with_0approx on CSR rows with callables that are not built-in metrics
    (block_coordinate.py:212-293 through :93-129) on the live reference."""
    _install_shims()
    from xcolumns import block_coordinate as bc
    from xcolumns import weighted_prediction as wp

    from xcolumns_b200.synth import csr_probs

    y = csr_probs(220, 300, 14, seed=61, ragged=True)
    out = {"data": y.data, "indices": y.indices, "indptr": y.indptr, "shape": np.array(y.shape)}
    cases = {
        "custom": (custom_fmeasure_like, 4, dict(seed=0, skip_tn=True, metric_kwargs={"gamma": 0.3})),
        "custom_tn_sum": (custom_with_tn, 3, dict(seed=1, skip_tn=False, metric_aggregation="sum")),
        "custom_min": (custom_fmeasure_like, 4, dict(seed=2, skip_tn=True, maximize=False, max_iters=3)),
    }
    for name, (func, k, kw) in cases.items():
        yp, meta = bc.predict_using_bc_with_0approx(y, func, k, return_meta=True, **kw)
        yp.sort_indices()
        out[name + "_indices"], out[name + "_indptr"] = yp.indices[:yp.indptr[-1]].copy(), yp.indptr.copy()
        out[name + "_util"] = np.array([float(u) for u in meta["utilities"]], dtype=np.float64)
        print(f"  callables_csr {name}: iters={meta['iters']} util={out[name + '_util']}")
    np.savez_compressed(os.path.join(HERE, "callables_csr.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "callables":
        main_callables()
    elif len(sys.argv) > 1 and sys.argv[1] == "callables_csr":
        main_callables_csr()
    elif len(sys.argv) > 1 and sys.argv[1] == "fw_generic":
        main_fw_generic()
    elif len(sys.argv) > 1 and sys.argv[1] == "online_csr":
        main_online_csr()
    elif len(sys.argv) > 1 and sys.argv[1] == "extra":
        main_extra()
    elif len(sys.argv) > 1 and sys.argv[1] == "mixed":
        main_mixed()
    else:
        main()
