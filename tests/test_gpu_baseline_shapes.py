"""Parity at BASELINE.json's OWN shapes.

C1 (10 000 x 1 000) and C2 (3 800 x 4 000): the sequential-exact GPU mode against the CPU oracle, bit for bit
(label ids and per-sweep utilities) -- the oracle itself is pinned to the live reference by tests/golden.
C3 / C4: the batched (block-Jacobi, pipelined) mode against the bit-pinned sequential-exact GPU mode used as the
at-scale oracle, final utility within a FLAT 1e-4 (north_star), the measured |delta| recorded in
gpurun_out/parity_deltas.jsonl.  XC_TEST_FULL_C3=1 runs the dense comparison on all 307 000 rows."""
import json
import os

import numpy as np
import pytest
import torch
from scipy.sparse import csr_matrix

pytestmark = pytest.mark.gpu

TOL = 1e-4
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def xb():
    import xcolumns_b200
    return xcolumns_b200


def record(name, **kw):
    """append the measured deviation to gpurun_out/parity_deltas.jsonl (scratch; summarised in DESIGN.md)"""
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_deltas.jsonl"), "a") as f:
            f.write(json.dumps({"case": name, **kw}) + "\n")
    except OSError:
        pass
    print(name, kw)


def _idx(pred, k):
    n = pred.shape[0]
    r, c = np.nonzero(pred)
    assert (np.bincount(r, minlength=n) == k).all()
    return c.reshape(n, k).astype(np.int32)


@pytest.mark.parametrize("name,n,m,seed", [("C1", 10000, 1000, 1001), ("C2", 3800, 4000, 1002)])
def test_exact_mode_bit_equal_to_oracle_at_baseline_shape(xb, oracle, name, n, m, seed):
    from xcolumns_b200.synth import dense_probs
    eta = dense_probs(n, m, seed=seed)
    opred, ometa = oracle.predict_using_bc_with_0approx(eta, "f1", 5, seed=0, skip_tn=True)
    pred, meta = xb.predict_optimizing_macro_f1_score_using_bc(eta, 5, seed=0, return_meta=True, mode="exact")
    assert pred.dtype == eta.dtype and pred.shape == eta.shape
    assert (pred.astype(np.uint8) == opred).all(), f"{name}: {int((pred.astype(np.uint8) != opred).any(1).sum())} rows differ"
    assert meta["iters"] == ometa["iters"]
    assert np.array_equal(np.asarray(meta["utilities"]), np.asarray(ometa["utilities"]))
    # the batched mode on the same input: flat 1e-4 on the final utility
    _, bmeta = xb.predict_optimizing_macro_f1_score_using_bc(eta, 5, seed=0, return_meta=True, mode="batched")
    d = bmeta["utilities"][-1] - ometa["utilities"][-1]
    record(f"{name}_f1_batched_vs_oracle", delta=d, sweeps=[bmeta["iters"], ometa["iters"]], lag=bmeta.get("lag"))
    assert abs(d) < TOL
    if name == "C2":   # BASELINE config 2 also names predict_weighted_per_instance
        rng = np.random.default_rng(3)
        a, b = rng.random(m).astype(np.float32), (rng.random(m).astype(np.float32) - 0.5) * 0.1
        w = xb.predict_weighted_per_instance(eta, 5, a=a, b=b)
        assert (_idx(w, 5) == oracle.topk_indices_dense(eta, 5, a, b)[0]).all()
        assert (_idx(xb.predict_top_k(eta, 5), 5) == oracle.topk_indices_dense(eta, 5)[0]).all()


def _batched_vs_exact_dense(xb, n, m, tag):
    from xcolumns_b200.synth import dense_probs_device
    eta = dense_probs_device(n, m, seed=1003, device=torch.device("cuda", 0))
    out = {}
    for mode in ("exact", "batched"):
        _, meta = xb.predict_optimizing_macro_f1_score_using_bc(eta, 5, seed=0, mode=mode, return_meta=True,
                                                                y_pred_format="indices")
        out[mode] = meta
    os.environ["XCOLUMNS_B200_LAG"] = "0"
    try:
        _, meta0 = xb.predict_optimizing_macro_f1_score_using_bc(eta, 5, seed=0, mode="batched", return_meta=True,
                                                                 y_pred_format="indices")
    finally:
        del os.environ["XCOLUMNS_B200_LAG"]
    ue, ub, u0 = out["exact"]["utilities"][-1], out["batched"]["utilities"][-1], meta0["utilities"][-1]
    record(tag, rows=n, labels=m, delta_pipelined=ub - ue, delta_strict=u0 - ue,
           sweeps=[out["exact"]["iters"], out["batched"]["iters"], meta0["iters"]], lag=out["batched"].get("lag"),
           seconds_exact=out["exact"]["time"], seconds_batched=out["batched"]["time"])
    assert abs(ub - ue) < TOL and abs(u0 - ue) < TOL
    assert (np.diff(np.asarray(out["batched"]["utilities"])) > -1e-9).all()


def test_batched_vs_exact_c3_slice(xb):
    """AmazonCat-13K label space, 60 000 rows: the bit-pinned sequential mode is the oracle"""
    _batched_vs_exact_dense(xb, 60000, 13000, "C3_slice_f1_batched_vs_exact")


@pytest.mark.skipif(os.environ.get("XC_TEST_FULL_C3") != "1", reason="full 307 000-row comparison: set XC_TEST_FULL_C3=1")
def test_batched_vs_exact_c3_full(xb):
    _batched_vs_exact_dense(xb, 307000, 13000, "C3_full_f1_batched_vs_exact")


@pytest.mark.parametrize("metric", ["recall", "f1"])
def test_batched_vs_exact_c4_slice(xb, metric):
    """Amazon-670K label space and row shape (100 stored labels per row), 40 000 rows"""
    from xcolumns_b200.synth import csr_probs
    y = csr_probs(40000, 670000, 100, seed=1004)
    fn = xb.predict_optimizing_macro_recall_using_bc if metric == "recall" else xb.predict_optimizing_macro_f1_score_using_bc
    pe, me = fn(y, 5, seed=0, mode="exact", return_meta=True)
    pb, mb = fn(y, 5, seed=0, mode="batched", return_meta=True)
    assert isinstance(pb, csr_matrix) and (np.diff(pb.indptr) == 5).all()
    d = mb["utilities"][-1] - me["utilities"][-1]
    record(f"C4_slice_{metric}_batched_vs_exact", delta=d, sweeps=[me["iters"], mb["iters"]])
    if metric == "recall":    # state-independent gains: the two modes reach the same optimum
        assert abs(d) < TOL
    else:
        # ~6 stored entries per label: the sequential algorithm itself stops in order-dependent fixed points on this
        # input (its own seed-to-seed spread is recorded next to the deviation); the batched mode must not be worse
        _, me1 = fn(y, 5, seed=1, mode="exact", return_meta=True)
        record("C4_slice_f1_exact_seed_spread", spread=abs(me1["utilities"][-1] - me["utilities"][-1]))
        assert d > -TOL


def test_coverage_batched_vs_exact_c4_slice(xb):
    from xcolumns_b200.synth import csr_probs
    y = csr_probs(20000, 670000, 100, seed=1004)
    _, me = xb.predict_optimizing_coverage_using_bc(y, 5, seed=0, mode="exact", return_meta=True)
    _, mb = xb.predict_optimizing_coverage_using_bc(y, 5, seed=0, mode="batched", return_meta=True)
    d = mb["utilities"][-1] - me["utilities"][-1]
    record("C4_slice_coverage_batched_vs_exact", delta=d, sweeps=[me["iters"], mb["iters"]])
    assert abs(d) < TOL


def test_pipelined_sweep_equals_its_serialised_schedule(xb):
    """lag = 1 on two streams against the same dependency order issued on ONE stream: the overlap must not change
    what a batch reads (coefficient sets, delta buffers)."""
    from xcolumns_b200.synth import dense_probs_device
    eta = dense_probs_device(30000, 3000, seed=5, device=torch.device("cuda", 0))
    res = {}
    for serial in ("0", "1"):
        os.environ["XCOLUMNS_B200_PIPE_SERIAL"] = serial
        try:
            pred, meta = xb.predict_optimizing_macro_f1_score_using_bc(eta, 5, seed=0, mode="batched", return_meta=True,
                                                                       y_pred_format="indices", max_iters=4,
                                                                       tolerance=-np.inf)
        finally:
            del os.environ["XCOLUMNS_B200_PIPE_SERIAL"]
        res[serial] = (pred.cpu().numpy(), meta)
    assert res["0"][1]["lag"] >= 1      # (2 at this size: batches of less than four waves)
    same = (res["0"][0] == res["1"][0]).all(1).mean()
    record("pipelined_vs_serialised", identical_rows=float(same),
           du=res["0"][1]["utilities"][-1] - res["1"][1]["utilities"][-1])
    # float64 atomics reorder the sums (1e-16 relative): a float32 coefficient may round differently once in a while
    assert same > 0.9995
    assert abs(res["0"][1]["utilities"][-1] - res["1"][1]["utilities"][-1]) < 1e-7


def test_pageable_input_upload(xb, monkeypatch):
    """pageable numpy input takes the staged upload (xc_h2d_staged) -- same result as the plain copy, padded and
    unpadded leading dimensions"""
    from xcolumns_b200 import _device as dev
    from xcolumns_b200.synth import dense_probs
    monkeypatch.setattr(dev, "_STAGED_MIN_BYTES", 0)
    for m in (1000, 1003):
        eta = dense_probs(3000, m, seed=m, tie_free=False)
        d = dev.dense_to_device(eta, torch.device("cuda", 0))
        torch.cuda.synchronize()
        assert d.ld % 4 == 0 and np.array_equal(d.t[:, :m].cpu().numpy(), eta)
        if d.ld != m:
            assert float(d.t[:, m:].abs().sum()) == 0.0
