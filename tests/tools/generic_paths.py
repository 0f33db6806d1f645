"""GPU probe of the paths that run user-supplied Python callables on the device through torch (generic_metric.py,
generic_fw.py): Frank-Wolfe with an objective that is not a built-in metric at the C5 shape (14 000 x 31 000), BCA
with a callable on dense and CSR rows.  Prints one JSON line per measurement (these paths are correct-but-unfused;
the numbers document what that costs next to the fused kernels)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import xcolumns_b200 as xb
from xcolumns_b200 import metrics as M
from xcolumns_b200.synth import csr_probs, dense_probs, dense_probs_device


def tversky(tp, fp, fn, tn, gamma=0.5):
    return ((1 + gamma) * tp / ((1 + gamma) * tp + gamma * fp + fn + 1e-6)).mean()


def f1_as_a_callable(tp, fp, fn, tn):       # the built-in objective, hidden from the resolver
    return (2 * tp / (2 * tp + fp + fn + 1e-9)).mean()


def binary_f1_as_a_callable(tp, fp, fn, tn):
    return 2 * tp / (2 * tp + fp + fn + 1e-9)


dev_ = torch.device("cuda", 0)
n, m = 14000, 31000
eta = dense_probs_device(n, m, seed=1005, device=dev_)
for name, func in (("tversky", tversky), ("f1 as an opaque callable", f1_as_a_callable),
                   ("f1 built-in (fused path)", M.macro_f1_score_on_conf_matrix)):
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.time()
        clf, meta = xb.find_classifier_using_fw(eta, eta, func, 5, max_iters=10, skip_tn=True, seed=0, return_meta=True)
        torch.cuda.synchronize()
        dt = time.time() - t0
    print(json.dumps({"op": "find_classifier_using_fw", "objective": name, "shape": [n, m], "iters": meta["iters"],
                      "ms_per_iteration": round(1e3 * dt / max(1, meta["iters"]), 3), "utility": meta["utilities"][-1]}))

eta = torch.from_numpy(dense_probs(20000, 2000, seed=9, tie_free=False)).to(dev_)
for name, func, mode in (("opaque f1, block-Jacobi", binary_f1_as_a_callable, "batched"),
                         ("built-in f1, block-Jacobi (fused)", M.binary_f1_score_on_conf_matrix, "batched")):
    torch.cuda.synchronize()
    t0 = time.time()
    _, meta = xb.predict_using_bc_with_0approx(eta, func, 5, seed=0, skip_tn=True, mode=mode, return_meta=True,
                                               y_pred_format="indices")
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(json.dumps({"op": "predict_using_bc_with_0approx dense 20000 x 2000", "metric": name, "sweeps": meta["iters"],
                      "ms_per_sweep": round(1e3 * dt / meta["iters"], 3), "utility": meta["utilities"][-1]}))

y = csr_probs(20000, 100000, 50, seed=10)
for name, func, mode in (("opaque f1, block-Jacobi", binary_f1_as_a_callable, "batched"),
                         ("built-in f1, block-Jacobi (fused)", M.binary_f1_score_on_conf_matrix, "batched")):
    t0 = time.time()
    _, meta = xb.predict_using_bc_with_0approx(y, func, 5, seed=0, skip_tn=True, mode=mode, return_meta=True,
                                               y_pred_format="indices")
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(json.dumps({"op": "predict_using_bc_with_0approx CSR 20000 x 100000, 50 stored labels per row", "metric": name,
                      "sweeps": meta["iters"], "ms_per_sweep_incl_upload": round(1e3 * dt / meta["iters"], 3),
                      "utility": meta["utilities"][-1]}))
ys = csr_probs(600, 5000, 30, seed=11)
t0 = time.time()
_, meta = xb.predict_using_bc_with_0approx(ys, binary_f1_as_a_callable, 5, seed=0, skip_tn=True, mode="exact",
                                           return_meta=True, y_pred_format="indices")
torch.cuda.synchronize()
dt = time.time() - t0
print(json.dumps({"op": "predict_using_bc_with_0approx CSR 600 x 5000 sequential, opaque f1", "sweeps": meta["iters"],
                  "us_per_instance": round(1e6 * dt / meta["iters"] / 600, 1)}))
