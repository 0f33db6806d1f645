"""GPU probe: Frank-Wolfe at the C5 shape for 20 iterations; prints the per-iteration time and the
line-search control blocks ([candidates, full, slices, tiles]) when XCOLUMNS_B200_FW_DEBUG=1."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xcolumns_b200 import metrics as M
from xcolumns_b200.frank_wolfe import find_classifier_using_fw
from xcolumns_b200.synth import dense_probs_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 14000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 31000
eta = dense_probs_device(n, m, seed=1005, device=torch.device("cuda", 0))
for rep in range(2):
    clf, meta = find_classifier_using_fw(eta, eta, M.macro_f1_score_on_conf_matrix, 5, max_iters=20, tolerance=-np.inf,
                                         alpha_tolerance=0.0, skip_tn=True, seed=0, return_meta=True)
print("ms/iter", 1e3 * meta["time"] / meta["iters"], "iters", meta["iters"])
print("utilities", meta["utilities"][-3:], "alphas", meta["alphas"][:6])
if "alpha_search_ctl" in meta:
    print("ctl", meta["alpha_search_ctl"])
    print("step_ms", [round(v, 3) for v in meta.get("step_ms", [])])
