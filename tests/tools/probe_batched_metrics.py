import sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import xcolumns_b200 as xb
from xcolumns_b200 import metrics as M
from oracle import oracle as orc
from xcolumns_b200.synth import dense_probs
eta = dense_probs(6000, 2000, seed=1002)
for metric, f in (("gmean", M.binary_gmean_on_conf_matrix), ("hmean", M.binary_hmean_on_conf_matrix), ("jaccard", M.binary_jaccard_score_on_conf_matrix)):
    skip = metric == "jaccard"
    _, om = orc.predict_using_bc_with_0approx(eta, metric, 5, seed=0, skip_tn=skip)
    print(metric, "oracle", om["utilities"])
    for bs in (750, 188, 94, 47, 16):
        _, meta = xb.predict_using_bc_with_0approx(eta, f, 5, seed=0, skip_tn=skip, return_meta=True, mode="batched", batch_size=bs)
        print("  batch", bs, [round(u, 6) for u in meta["utilities"]], "diff %.2e" % (meta["utilities"][-1] - om["utilities"][-1]))

if len(sys.argv) > 1 and sys.argv[1] == "full":
    import time
    import torch
    from xcolumns_b200.synth import dense_probs_device
    big = dense_probs_device(307000, 13000, seed=1003, device=torch.device("cuda", 0))
    for metric, f in (("jaccard", M.binary_jaccard_score_on_conf_matrix), ("hmean", M.binary_hmean_on_conf_matrix),
                      ("gmean", M.binary_gmean_on_conf_matrix)):
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.time()
            _, meta = xb.predict_using_bc_with_0approx(big, f, 5, seed=0, skip_tn=(metric == "jaccard"), return_meta=True,
                                                       mode="batched", max_iters=4, tolerance=-np.inf,
                                                       y_pred_format="indices")
            torch.cuda.synchronize()
        print(f"C3 {metric}: {1e3 * (time.time() - t0) / meta['iters']:.1f} ms per sweep incl. setup, batch {meta['batch_size']},"
              f" utilities {[round(u, 6) for u in meta['utilities']]}")
