"""Multi-GPU correctness of the sharded paths (run under torchrun, one rank per GPU; tests/test_gpu_multi.py
launches it with two ranks, bench.py's parity_multi_gpu repeats check 1 before timing):

1. sharded batched BCA (peer-memory commits, pipelined) == the single-GPU sequential-exact mode on the concatenated
   matrix within 1e-4, and its reported utility == a recomputation of the gathered prediction on one GPU;
2. the same with NCCL all-reduce commits ($XCOLUMNS_B200_P2P=0) and with the strict batch order (lag 0);
3. stress of the peer-memory protocol: many short sweeps with a device-side delay injected on one rank before
   every other sweep -- the replicated float64 state must stay BIT-identical on all ranks (a rank that read a
   cleared or half-written peer buffer would diverge);
4. sharded coverage BCA == single-GPU coverage BCA (sequential-exact) within 1e-4;
5. Frank-Wolfe with init_classifier='prior' / 'random': every rank ends with the same classifier.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xcolumns_b200 as xb  # noqa: E402
from xcolumns_b200 import _device as dev  # noqa: E402
from xcolumns_b200 import metrics as M  # noqa: E402
from xcolumns_b200._lib import XC_F32, XC_SUM_FAST  # noqa: E402
from xcolumns_b200.block_coordinate import BcaSession, _metric_params  # noqa: E402
from xcolumns_b200.distributed import make_comm, shard_rows  # noqa: E402
from xcolumns_b200.synth import csr_probs, dense_probs_device  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
comm = make_comm(True, device)
TOL = 1e-4
k = 5


def say(*a):
    if rank == 0:
        print(*a, flush=True)


def gather_rows(t, n):
    parts = [torch.empty((shard_rows(n, r, world)[1] - shard_rows(n, r, world)[0],) + tuple(t.shape[1:]), dtype=t.dtype,
                         device=device) for r in range(world)]
    dist.all_gather(parts, t.contiguous())
    return torch.cat(parts)


# ---- 1 + 2: sharded == single GPU ------------------------------------------------------------------------------
n, m = 6000 * world, 2000
full = dense_probs_device(n, m, seed=977, device=device)
lo, hi = shard_rows(n, rank, world)
shard = full[lo:hi].contiguous()
ref_u = None
if rank == 0:
    _, mx = xb.predict_optimizing_macro_f1_score_using_bc(full, k, seed=0, mode="exact", return_meta=True,
                                                          y_pred_format="indices")
    ref_u = mx["utilities"][-1]
for p2p, lag in (("1", "1"), ("1", "0"), ("0", "1")):
    os.environ["XCOLUMNS_B200_P2P"], os.environ["XCOLUMNS_B200_LAG"] = p2p, lag
    pred, meta = xb.predict_optimizing_macro_f1_score_using_bc(shard, k, seed=0, mode="batched", distributed=True,
                                                               return_meta=True, y_pred_format="indices")
    assert meta["commit"] == ("peer-memory" if p2p == "1" else "all-reduce"), meta["commit"]
    allp = gather_rows(pred, n)
    if rank == 0:
        params = _metric_params(M.XC_METRIC_FBETA, 1.0, 1e-9, True, True, n)
        s1 = BcaSession(dev.DenseDev(full, n, m, m, XC_F32, 0), k, params, params, "mean")
        s1.pred = allp
        s1.recompute(XC_SUM_FAST)
        s1.utility_device(0)
        u_re = float(s1.util_buf[0].item())
        u = meta["utilities"][-1]
        say(f"p2p={p2p} lag={lag}: commit={meta['commit']} sweeps={meta['iters']} u={u:.9f} recomputed={u_re:.9f} "
            f"exact={ref_u:.9f} d={u - ref_u:+.2e}")
        assert abs(u - u_re) < 1e-9 and abs(u - ref_u) < TOL
    # every rank reports the same utilities (replicated state)
    ul = torch.tensor(meta["utilities"], dtype=torch.float64, device=device)
    ug = [torch.empty_like(ul) for _ in range(world)]
    dist.all_gather(ug, ul)
    assert all(torch.equal(ug[0], x) for x in ug), "ranks disagree on the utilities"
os.environ["XCOLUMNS_B200_P2P"], os.environ["XCOLUMNS_B200_LAG"] = "1", "1"

# ---- 3: protocol stress with a delayed rank ---------------------------------------------------------------------
ns, ms = 3000, 1500
eta_s = dense_probs_device(ns, ms, seed=31 + rank, device=device)
params = _metric_params(M.XC_METRIC_FBETA, 1.0, 1e-9, True, True, ns * world)
sess = BcaSession(dev.DenseDev(eta_s, ns, ms, ms, XC_F32, 0), k, params, params, "mean", comm)
assert sess.peer is not None and sess.pipe
from xcolumns_b200.weighted_prediction import topk_dense_device  # noqa: E402
sess.pred = topk_dense_device(sess.data, k, None, None, XC_F32)[0]
sess.recompute(XC_SUM_FAST)
util = torch.zeros(64, dtype=torch.float64, device=device)
for it in range(60):
    if (it + rank) % 2 == 0:
        torch.cuda._sleep(int(2e6 * (1 + (it % 3))))        # ~1-3 ms of device-side delay on alternating ranks
    sess.run_sweep(it + 1, 11 + 7919 * it + 13 * rank, True, 100, 30, it % 16 == 15, util[it:])
sess.join()
st = sess.state[:3].clone()
gs = [torch.empty_like(st) for _ in range(world)]
dist.all_gather(gs, st)
assert all(torch.equal(gs[0], x) for x in gs), "replicated state diverged between ranks"
# ... and equals the state recomputed from the predictions
sess.recompute(XC_SUM_FAST)
assert torch.allclose(sess.state[:3], st, rtol=0, atol=1e-9)
sess.close()
say("protocol stress: 60 sweeps x 30 commits with a delayed rank, state bit-identical on all ranks")

# ---- 4: coverage -------------------------------------------------------------------------------------------------
# the same 4 000 x 3 000 problem whatever the number of ranks (the rows every commit gathers, batch x world, stay a
# fixed fraction of n like in the single-GPU mode)
nc, mc = 4000, 3000
y = csr_probs(nc, mc, 40, seed=1006)
lo, hi = shard_rows(nc, rank, world)
# one sharded sweep by hand: the folded Ef must equal the Ef recomputed from the new predictions (absolute: products
# of thousands of factors run into the denormals, where the two orders of multiplication legitimately differ)
from xcolumns_b200.block_coordinate import CoverageSession  # noqa: E402
from xcolumns_b200.weighted_prediction import topk_csr_device  # noqa: E402
cs = CoverageSession(dev.csr_to_device(y[lo:hi], device), k, 1.0, comm)
cs.pred = topk_csr_device(cs.data, k, None, None)[0]
cs.state(XC_SUM_FAST)
cs.utility_device(0)
order_c = torch.randperm(hi - lo, device=device).int()
bsz = max(1, 16 // world)
cs.sweep_batched(order_c, bsz, comm.max_int((hi - lo + bsz - 1) // bsz))
ef_fold = cs.Ef.clone()
cs.state(XC_SUM_FAST)
cs.utility_device(1)
err = float((ef_fold - cs.Ef).abs().max())
say(f"coverage by hand: u0={float(cs.util_buf[0]):.9f} u1={float(cs.util_buf[1]):.9f} max |Ef_fold - Ef_recomputed| = {err:.2e}")
assert err < 1e-12, err
assert float(cs.util_buf[1]) >= float(cs.util_buf[0]) - 1e-9
predc, metac = xb.predict_optimizing_coverage_using_bc(y[lo:hi], k, seed=0, mode="batched", distributed=True,
                                                       return_meta=True, y_pred_format="indices")
say("coverage sharded meta:", {kk: metac[kk] for kk in ("utilities", "iters", "batch_size")})
if rank == 0:
    _, mex = xb.predict_optimizing_coverage_using_bc(y, k, seed=0, mode="exact", return_meta=True)
    say(f"coverage: sharded {metac['utilities'][-1]:.9f} exact {mex['utilities'][-1]:.9f}")
    assert abs(metac["utilities"][-1] - mex["utilities"][-1]) < TOL
allc = gather_rows(torch.as_tensor(predc, device=device), nc).cpu().numpy()
if rank == 0:   # coverage of the gathered prediction recomputed on the host
    rows = np.repeat(np.arange(nc), k)
    sel = np.asarray(y[rows, allc.reshape(-1)]).reshape(-1).astype(np.float64)
    logf = np.zeros(mc)
    np.add.at(logf, allc.reshape(-1), np.log1p(-sel))
    assert abs((1.0 - np.exp(logf).mean()) - metac["utilities"][-1]) < 1e-9

# ---- 5: Frank-Wolfe, data-dependent initial classifiers -----------------------------------------------------------
for init in ("prior", "random"):
    clf, mfw = xb.find_classifier_using_fw(shard, shard, M.macro_f1_score_on_conf_matrix, k, max_iters=4,
                                           init_classifier=init, seed=None if init == "random" else 0, skip_tn=True,
                                           return_meta=True, distributed=True)
    a = torch.as_tensor(np.asarray(clf.a.cpu() if isinstance(clf.a, torch.Tensor) else clf.a), device=device).double()
    ga = [torch.empty_like(a) for _ in range(world)]
    dist.all_gather(ga, a)
    assert all(torch.equal(ga[0], x) for x in ga), f"FW init={init}: ranks ended with different classifiers"
say("frank-wolfe: identical classifiers on all ranks for init = prior / random")

dist.barrier()
say("MULTI_GPU_CHECK_OK")
dist.destroy_process_group()
