"""Multi-GPU check of the peer-memory commit (run under torchrun, one rank per GPU):
the sharded batched BCA through the public API with the peer-memory commit kernel and with the
NCCL all-reduce path must give the same predictions and utilities."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xcolumns_b200 as xb
from xcolumns_b200.synth import dense_probs_device

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
n, m = 40000, 3000
eta = dense_probs_device(n, m, seed=500 + rank, device=device)
out = {}
for mode in ("1", "0"):
    os.environ["XCOLUMNS_B200_P2P"] = mode
    pred, meta = xb.predict_optimizing_macro_f1_score_using_bc(eta, 5, seed=0, mode="batched", distributed=True,
                                                               return_meta=True, y_pred_format="indices", max_iters=6,
                                                               batch_size=4096)
    out[mode] = (pred.clone() if isinstance(pred, torch.Tensor) else torch.as_tensor(pred), meta)
    if rank == 0:
        print(f"P2P={mode}: commit={meta['commit']} iters={meta['iters']} utilities={meta['utilities']}", flush=True)
same_pred = bool((out["1"][0] == out["0"][0]).all())
du = max(abs(a - b) for a, b in zip(out["1"][1]["utilities"], out["0"][1]["utilities"]))
t = torch.tensor([0 if (same_pred and du < 1e-12) else 1], device=device)
dist.all_reduce(t)
if rank == 0:
    print("peer-memory vs all-reduce: same predictions on every rank:", int(t.item()) == 0, "max |du| =", du, flush=True)
assert out["1"][1]["commit"] == "peer-memory" and out["0"][1]["commit"] == "all-reduce"
assert int(t.item()) == 0
dist.destroy_process_group()
