"""Workload for the round-2 ncu captures of the two latency / issue-bound kernels VERDICT round 1 asked evidence for:

    ncu --set full --clock-control none --import-source on \
        -k regex:"bca_exact_dense_cluster_kernel|fw_alpha_evalfast_kernel" -c 3 -o gpurun_out/r02_targets \
        python scripts/ncu_targets.py

First one sequential-exact dense sweep at C1 (10 000 x 1 000, one launch), then three Frank-Wolfe iterations at C5
(14 000 x 31 000; the float32 screen of the line search runs once per iteration)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xcolumns_b200 as xb
from xcolumns_b200 import metrics as M
from xcolumns_b200.synth import dense_probs, dense_probs_device

dev_ = torch.device("cuda", 0)
eta = torch.from_numpy(dense_probs(10000, 1000, seed=1001, tie_free=False)).to(dev_)
_, meta = xb.predict_optimizing_macro_f1_score_using_bc(eta, 5, seed=0, mode="exact", max_iters=1, return_meta=True,
                                                        y_pred_format="indices")
print("exact sweep", meta["utilities"])
eta = dense_probs_device(14000, 31000, seed=1005, device=dev_)
clf, meta = xb.find_classifier_using_fw(eta, eta, M.macro_f1_score_on_conf_matrix, 5, max_iters=3, skip_tn=True, seed=0,
                                        return_meta=True)
torch.cuda.synchronize()
print("fw", meta["utilities"])
