"""GPU probe for the two small configurations of BASELINE.json (C1: 10 000 x 1 000, C2: 3 800 x 4 000):
per-sweep time of both BCA modes through the public API on device-resident scores, the weighted top-k,
and the CPU port of the reference on the same inputs.  Prints one JSON line per measurement."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import xcolumns_b200 as xb
from oracle import oracle as orc
from xcolumns_b200.synth import dense_probs

dev_ = torch.device("cuda", 0)
for name, n, m in (("C1", 10000, 1000), ("C2", 3800, 4000)):
    eta = dense_probs(n, m, seed=1001 if name == "C1" else 1002)
    eta_d = torch.from_numpy(eta).to(dev_)
    t0 = time.time()
    _, ometa = orc.predict_using_bc_with_0approx(eta, "f1", 5, seed=0, skip_tn=True)
    cpu = time.time() - t0
    for mode in ("exact", "batched"):
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.time()
            pred, meta = xb.predict_optimizing_macro_f1_score_using_bc(eta_d, 5, seed=0, mode=mode, return_meta=True,
                                                                       y_pred_format="indices")
            torch.cuda.synchronize()
            dt = time.time() - t0
        print(json.dumps({"config": name, "mode": mode, "sweeps": meta["iters"], "ms_per_call": round(1e3 * dt, 3),
                          "instances_per_s_per_sweep": round(n * meta["iters"] / dt), "utility": meta["utilities"][-1],
                          "cpu_port_utility": ometa["utilities"][-1],
                          "cpu_port_instances_per_s_per_sweep": round(n * ometa["iters"] / cpu)}))
    a = torch.rand(m, device=dev_) + 0.5
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(20):
            xb.predict_weighted_per_instance(eta_d, 5, a=a)
        torch.cuda.synchronize()
        dt = (time.time() - t0) / 20
    print(json.dumps({"config": name, "op": "predict_weighted_per_instance (device tensor in, dense device tensor out)",
                      "us_per_call": round(1e6 * dt, 1), "GB_per_s_read": round(n * m * 4 / dt / 1e9)}))
