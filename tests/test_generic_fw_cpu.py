"""Host logic of the Frank-Wolfe path for arbitrary objective callables (xcolumns_b200/generic_fw.py), on CPU tensors:
the loop (gradient by torch autograd, vmapped line search, stopping rules, classifier bookkeeping) is driven with the
ORACLE's confusion vectors in place of the CUDA kernels and must reproduce the live reference's golden runs
(tests/golden/make_golden.py fw_generic; frank_wolfe.py:565-670)."""
import numpy as np
import pytest
import torch
from scipy.sparse import csr_matrix

from xcolumns_b200 import generic_fw


def tversky(tp, fp, fn, tn, gamma=0.5):
    return ((1 + gamma) * tp / ((1 + gamma) * tp + gamma * fp + fn + 1e-6)).mean()


def tpr_tnr(tp, fp, fn, tn):
    return (tp / (tp + fn + 1e-6) * tn / (tn + fp + 1e-6)).mean()


def _oracle_conf(orc, y_true, y_proba, k, skip_tn):
    m = y_proba.shape[1]

    def conf(i, A_dev, B_dev):
        a, b = A_dev[i, :m].numpy(), B_dev[i, :m].numpy()
        pred = orc.predict_weighted_per_instance(y_proba, k, th=0.0, a=a, b=b)
        c = orc.calculate_confusion_matrix(y_true, pred, normalize=True, skip_tn=skip_tn, dtype=np.float64)
        return torch.from_numpy(np.stack([np.asarray(v, dtype=np.float64) for v in c]))

    return conf


def _run(orc, y_true, y_proba, func, k, skip_tn, max_iters, metric_kwargs=None, **kw):
    m = y_proba.shape[1]
    A = np.zeros((max_iters + 1, m), dtype=np.float32)
    B = np.zeros((max_iters + 1, m), dtype=np.float32)
    P = np.ones(max_iters + 1, dtype=np.float32)
    A[0], B[0] = 1.0, -0.5
    args = dict(conf=_oracle_conf(orc, y_true, y_proba, k, skip_tn), device=torch.device("cpu"), m=m, A=A, B=B, P=P,
                max_iters=max_iters, maximize=True, metric_func=func, metric_kwargs=metric_kwargs, tolerance=1e-6,
                search_for_best_alpha=True, alpha_search_algo="uniform", alpha_tolerance=0.001,
                alpha_uniform_search_step=0.0001, verbose=False)
    args.update(kw)
    return generic_fw.run(**args)


def _check(g, name, res, alpha_atol=1e-9):
    a, b, p, meta = res
    assert meta["iters"] == int(g[name + "_iters"]), (name, meta["iters"])
    assert np.allclose(meta["alphas"], g[name + "_alphas"], rtol=0, atol=alpha_atol), (name, meta["alphas"])
    assert np.allclose(meta["utilities"], g[name + "_util"], rtol=0, atol=1e-5), (name, meta["utilities"])
    assert np.allclose(meta["classifiers_utilities"], g[name + "_cutil"], rtol=0, atol=1e-5), name
    assert tuple(a.shape) == g[name + "_a"].shape and np.allclose(p, g[name + "_p"], atol=1e-6), name
    # classifier rows: gradients of the objective (float32 in the reference, float64 -> float32 here)
    assert np.allclose(a.numpy(), g[name + "_a"], rtol=2e-3, atol=1e-4), name
    assert np.allclose(b.numpy(), g[name + "_b"], rtol=2e-3, atol=1e-4), name


def test_generic_fw_loop_matches_reference_goldens(golden, oracle):
    g = golden("fw_generic")
    eta, lab = g["eta"], g["lab"]
    _check(g, "tversky", _run(oracle, lab, eta, tversky, 4, True, 8, metric_kwargs={"gamma": 0.7}))
    _check(g, "tpr_tnr", _run(oracle, lab, eta, tpr_tnr, 4, False, 6))
    _check(g, "tversky_ternary", _run(oracle, lab, eta, tversky, 4, True, 6, alpha_search_algo="ternary"))
    _check(g, "tversky_fixed", _run(oracle, lab, eta, tversky, 4, True, 5, search_for_best_alpha=False))
    shape = tuple(g["shape"])
    y = csr_matrix((g["y_data"], g["y_indices"], g["y_indptr"]), shape=shape)
    yt = csr_matrix((g["yt_data"], g["yt_indices"], g["yt_indptr"]), shape=shape)
    _check(g, "tversky_csr", _run(oracle, yt, y, tversky, 3, True, 6))


def test_objective_grid_first_maximum_and_vmap_fallback():
    torch.manual_seed(0)
    cm, ci = torch.rand(4, 50, dtype=torch.float64), torch.rand(4, 50, dtype=torch.float64)
    grid = torch.linspace(0, 1, 101, dtype=torch.float64)
    obj = generic_fw._Objective(tversky, {"gamma": 0.2})
    vals = obj.on_grid(cm, ci, grid)
    assert obj.vmap_ok
    ref = torch.stack([tversky(*((1 - a) * cm + a * ci), gamma=0.2) for a in grid])
    assert torch.allclose(vals, ref, rtol=0, atol=1e-15)

    def branching(tp, fp, fn, tn):   # data-dependent control flow: vmap cannot trace it
        v = (tp / (tp + fp + fn + 1e-6)).mean()
        return v if float(v) > 0.1 else v * 0.5

    obj2 = generic_fw._Objective(branching, None)
    vals2 = obj2.on_grid(cm, ci, grid)
    assert not obj2.vmap_ok
    ref2 = torch.stack([branching(*((1 - a) * cm + a * ci)) for a in grid])
    assert torch.equal(vals2, ref2)
    # a flat objective: the reference's strict `score > best_val` scan keeps alpha = 0 (the first maximum)
    flat = generic_fw._Objective(lambda tp, fp, fn, tn: (tp * 0).sum() + 1.0, None)
    assert int(torch.argmax(flat.on_grid(cm, ci, grid))) == 0
    v, grads = flat.value_and_grad(cm)
    assert float(v) == 1.0 and all(float(x.abs().sum()) == 0 for x in grads)


def test_objective_rejects_non_scalar_and_non_tensor():
    cm = torch.rand(4, 10, dtype=torch.float64)
    with pytest.raises(ValueError):
        generic_fw._Objective(lambda tp, fp, fn, tn: tp, None).value(cm)
    with pytest.raises(ValueError):
        generic_fw._Objective(lambda tp, fp, fn, tn: 1.0, None).value(cm)
