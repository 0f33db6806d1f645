"""Helper of tests/test_gpu_variants.py: runs the public API once under whatever kernel-selection
environment variables the parent set and prints a digest (JSON) of the results."""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xcolumns_b200 as xb
from xcolumns_b200 import metrics as M
from xcolumns_b200.synth import csr_probs, dense_probs


def digest(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


out = {}
eta = dense_probs(4000, 2100, seed=7)
pred, meta = xb.predict_optimizing_macro_f1_score_using_bc(eta, 5, seed=0, mode="batched", return_meta=True,
                                                           y_pred_format="indices")
out["bca_batched"] = (digest(pred), meta["utilities"])
y = csr_probs(600, 5000, 40, seed=8, ragged=True)
pred, meta = xb.predict_optimizing_macro_f1_score_using_bc(y, 5, seed=1, mode="exact", return_meta=True)
out["bca_exact_csr"] = (digest(pred.indices), meta["utilities"])
eta2 = dense_probs(900, 700, seed=9)
clf, meta = xb.find_classifier_using_fw(eta2, eta2, M.macro_f1_score_on_conf_matrix, 5, max_iters=6, skip_tn=True,
                                        seed=0, return_meta=True)
out["fw"] = (digest(clf.a), meta["alphas"], meta["utilities"])
print("DIGEST " + json.dumps(out))
