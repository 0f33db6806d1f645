"""One sequential-exact call at C1 (10 000 x 1 000) through the public API on device-resident scores: the workload of
the per-kernel launch list `ncu --metrics gpu__time_duration.sum --clock-control none --csv` (profiles/r02_launches_exact_c1.csv)
and, without ncu, the wall-clock split between the sweep kernel and everything else of a sweep."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import xcolumns_b200 as xb
from xcolumns_b200.synth import dense_probs

eta = torch.from_numpy(dense_probs(10000, 1000, seed=1001, tie_free=False)).cuda()
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.time()
    _, meta = xb.predict_optimizing_macro_f1_score_using_bc(eta, 5, seed=0, mode="exact", return_meta=True,
                                                            y_pred_format="indices")
    torch.cuda.synchronize()
    dt = time.time() - t0
print(json.dumps({"sweeps": meta["iters"], "ms_per_call": round(1e3 * dt, 2),
                  "us_per_instance": round(1e6 * dt / (10000 * meta["iters"]), 3)}))
