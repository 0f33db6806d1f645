import sys; sys.path.insert(0, "/root/repo")
import numpy as np, xcolumns_b200 as xb
from xcolumns_b200.synth import csr_probs
y = csr_probs(16000, 3000, 40, seed=1006)
for bs in (496, 64, 16, 2):
    _, m = xb.predict_optimizing_coverage_using_bc(y, 5, seed=0, mode="batched", return_meta=True, batch_size=bs, max_iters=4)
    print("batch", bs, m["utilities"], m["iters"], flush=True)
_, m = xb.predict_optimizing_coverage_using_bc(y, 5, seed=0, mode="exact", return_meta=True, max_iters=4)
print("exact", m["utilities"])
