#!/usr/bin/env python
"""Benchmark of the BCA macro-F1@5 hot path (BASELINE.json metric: instances/sec per sweep).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rows n] [--labels m]

A "step" is one block-Jacobi BCA sweep (macro-F1@5) over the whole synthetic AmazonCat-13K-shape
probability matrix (n=307000, m=13000, float32, 15.96 GB -- far larger than the 126 MB L2, so no
L2 flush is needed between steps).  With N > 1 (torchrun, one rank per GPU) the SAME 307 000 rows are
sharded over the ranks (BASELINE.json config 3: "sharded over 8xB200", strong scaling, n/N rows per
GPU); the per-batch confusion deltas are exchanged by the peer-memory commit kernel over NVLink.  The
weak-scaling figure (307 000 rows per GPU) is reported next to it under "weak_scaling".

Printed JSON (one line, rank 0):
  value     : instances/sec/sweep with y_proba resident in HBM, CUDA-event timed, max over ranks
  e2e       : same metric through the public Python API with HOST (pinned) buffers: H2D of
              y_proba, top-k init, K sweeps, D2H + host materialisation of the prediction
  roofline  : the dominant kernel (bca_batch_dense_kernel): algorithmic bytes per launch
              (rows * m * 4) / CUDA-event time per launch vs the measured HBM copy peak
  cpu_baseline : the CPU oracle (C port of the reference algorithm, 1 core) on a row subsample
  parity_multi_gpu (N > 1): a small sharded problem against the single-GPU sequential mode
  secondary (N = 1): C5 Frank-Wolfe, C4 CSR macro-recall@5 and coverage@5, each with a CPU-port baseline
`--impl reference` times that same CPU port (the reference is pure Python + numba and cannot
travel to the GPU box) and prints the same line with "impl": "reference".
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOAD = "amazoncat13k-shape dense f32 n=307000 m=13000 k=5 macro-F1 BCA (batched)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=307000, help="rows of the whole job (sharded over the ranks)")
    ap.add_argument("--labels", type=int, default=13000)
    ap.add_argument("--k", type=int, default=5)
    ap.add_argument("--batch", type=int, default=0, help="rows per commit per rank (0 = default)")
    ap.add_argument("--cpu-rows", type=int, default=1500, help="row subsample of the CPU baseline")
    ap.add_argument("--workload", default="bca_dense",
                    choices=["bca_dense", "bca_csr", "bca_csr_recall", "coverage_csr", "fw_dense"],
                    help="bca_dense is the headline (BASELINE.json metric); the others print their own line")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-weak", action="store_true")
    return ap.parse_args()


def kernel_source_sha():
    """hash of the sources the dominant kernel is compiled from (ties profiles/traffic.json to a build)"""
    import hashlib
    h = hashlib.sha256()
    for f in ("bca_batched.cu", "xc_scan.cuh", "xc_common.cuh"):
        with open(os.path.join(ROOT, "xcolumns_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def load_traffic():
    """DRAM bytes (read + write) of ONE launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/traffic.json, written by scripts/ncu_traffic.py).  Returned only if the capture was taken with the
    sources of THIS build; a stale capture reads as null instead of silently describing another kernel."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if t.get("source_sha") == kernel_source_sha():
            return t
        return {"stale": True, "captured_with": t.get("source_sha"), "this_build": kernel_source_sha()}
    except Exception:
        return None


def load_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle port on a bounded sample
# --------------------------------------------------------------------------------------------

def cpu_port_sample(eta_sub: np.ndarray, k: int, sweeps: int):
    """instances/sec/sweep of the sequential reference algorithm (C port, 1 core)."""
    from oracle import oracle as orc
    orc.build()
    t0 = time.time()
    _, meta = orc.predict_using_bc_with_0approx(eta_sub, "f1", k, seed=0, skip_tn=True, max_iters=sweeps,
                                                tolerance=-np.inf)
    dt = time.time() - t0
    return eta_sub.shape[0] * meta["iters"] / dt, dt, meta["iters"]


def run_reference(args):
    """--impl reference: rank 0 times the CPU port, other ranks exit."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from xcolumns_b200.synth import dense_probs
    n_sub = args.cpu_rows
    eta = dense_probs(n_sub, args.labels, seed=1003, tie_free=False)
    cpu_port_sample(eta[:64], args.k, 1)  # warm (page-in, build)
    for _ in range(args.warmup):
        cpu_port_sample(eta, args.k, 1)
    t0 = time.time()
    for _ in range(args.steps):
        cpu_port_sample(eta, args.k, 1)
    total = time.time() - t0
    value = n_sub * args.steps / total
    line = {
        "impl": "reference", "metric": "BCA macro-F1@5 instances/sec per sweep", "value": value,
        "unit": "instances/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows": args.rows, "labels": args.labels, "k": args.k},
        "cpu_baseline": {"value": value, "unit": "instances/s", "cores": 1, "kind": "port",
                         "sample": f"{n_sub} rows x {args.labels} labels of the same distribution, 1 sequential sweep per step "
                                   f"(a sweep's per-instance cost does not depend on n; the chain of instance updates is "
                                   f"sequential, so 1 core is all the reference algorithm can use; the Python reference "
                                   f"itself runs ~1.0k inst/s at this m, BASELINE.md table B)"},
        "e2e": {"value": value, "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------

class ClockSampler(threading.Thread):
    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_sm = index, False, [], set(), None
        self.active = False  # samples are kept only while the timed region runs
        self.ready = threading.Event()   # NVML is initialised (can take seconds on a fresh box)
        self._h = self._nv = None
        self._names = {}

    def _sample(self):
        clk = self._nv.nvmlDeviceGetClockInfo(self._h, self._nv.NVML_CLOCK_SM)
        r = self._nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        self.sm.append(clk)
        for bit, nm in self._names.items():
            if r & bit:
                self.reasons.add(nm)

    def sample_now(self):
        """one sample from the calling thread (the timed region of a short run can end before the sampling
        thread gets its turn: the caller takes one while its queued sweeps are still executing)"""
        if self._nv is not None:
            try:
                self._sample()
            except Exception:
                pass

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
            self._names = {
                pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
            self._nv = pynvml
            self.ready.set()
            while not self.stop_flag:
                if self.active:
                    self._sample()
                time.sleep(0.001)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")
            self.ready.set()

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------

def read_kernel_timing(ctx, cap=8192):
    """(start_ms, end_ms, rows) of every batch-kernel launch recorded since xc_timing_enable"""
    st, en = (C.c_double * cap)(), (C.c_double * cap)()
    rows = (C.c_int64 * cap)()
    cnt = C.c_int(0)
    ctx.call("xc_timing_read", cap, st, en, rows, C.byref(cnt))
    n = min(cnt.value, cap)
    return np.array(st[:n]), np.array(en[:n]), np.array(rows[:n])


def union_ms(st, en):
    """total length of the union of the intervals [st_i, en_i] (two batch kernels are in flight at a time)"""
    if len(st) == 0:
        return 0.0
    o = np.argsort(st)
    tot, cur_s, cur_e = 0.0, st[o[0]], en[o[0]]
    for i in o[1:]:
        if st[i] > cur_e:
            tot += cur_e - cur_s
            cur_s, cur_e = st[i], en[i]
        else:
            cur_e = max(cur_e, en[i])
    return float(tot + cur_e - cur_s)


def timed(fn, steps, device, comm=None, after_enqueue=None):
    """steps calls of fn between two CUDA events, barrier + synchronize on both sides; ms (max over ranks).
    after_enqueue runs on the host once everything (incl. the closing event) is queued, i.e. while the GPU is
    still executing the timed steps."""
    import torch
    import torch.distributed as dist
    if comm is not None:
        comm.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if after_enqueue is not None:
        after_enqueue()
    torch.cuda.synchronize(device)
    if comm is not None:
        comm.barrier()
    ms = e0.elapsed_time(e1)
    if comm is not None and comm.world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


# --------------------------------------------------------------------------------------------
# secondary workloads (BASELINE.json configs 4 and 5)
# --------------------------------------------------------------------------------------------

def bench_fw(args, device, comm, n=14000, m=31000, cpu=True):
    """C5: find_classifier_using_fw macro-F1@5, Wiki10-31K shape, through the public API on a device tensor"""
    import torch
    import torch.distributed as dist
    from xcolumns_b200 import metrics as M
    from xcolumns_b200.frank_wolfe import find_classifier_using_fw
    from xcolumns_b200.synth import dense_probs, dense_probs_device
    peak, _ = load_peak()
    k, world, rank = args.k, comm.world, comm.rank
    eta_t = dense_probs_device(n, m, seed=1005 + rank, device=device)   # n rows per GPU
    state = {}

    def run(iters):
        comm.barrier()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        clf, meta = find_classifier_using_fw(eta_t, eta_t, M.macro_f1_score_on_conf_matrix, k, max_iters=iters,
                                             tolerance=-np.inf, alpha_tolerance=0.0, skip_tn=True, seed=0,
                                             return_meta=True, distributed=(world > 1))
        e1.record()
        torch.cuda.synchronize(device)
        state["meta"] = meta
        return e0.elapsed_time(e1), meta["iters"]

    run(3)
    ms_call, iters = run(20)
    # the reference's own clock: meta["time"] covers the iteration loop (after the initial classifier's pass)
    ms = state["meta"]["time"] * 1e3
    if world > 1:
        t = torch.tensor([ms, ms_call], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_call = float(t[0]), float(t[1])
    ach = n * m * 4 * iters / (ms / 1e3) / 1e9   # per GPU
    out = {"metric": "Frank-Wolfe macro-F1@5 iterations/sec", "value": iters / (ms / 1e3), "unit": "iterations/s",
           "n_gpus": world, "steps": iters, "ms_per_step": ms / iters, "ms_per_call": ms_call,
           "instances_per_s": n * world * iters / (ms / 1e3), "dtype": "f32", "data": "synthetic (y_true := y_proba)",
           "config": {"workload": f"wiki10-31k-shape dense f32 n={n} rows per GPU, m={m} k={k} macro-F1 FW, "
                                  f"{iters} iterations incl. init pass"},
           "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "note": "iteration loop: one pass over y_proba (n*m*4 bytes) per iteration / loop time"},
           "utilities": state["meta"]["utilities"][-2:]}
    del eta_t
    if cpu and rank == 0:
        from oracle import oracle as orc
        orc.build()
        n_sub = 400
        eta = dense_probs(n_sub, m, seed=1005, tie_free=False)
        t0 = time.time()
        r = orc.find_classifier_using_fw(eta, eta, "f1", k, max_iters=3, tolerance=-np.inf, alpha_tolerance=0.0,
                                         skip_tn=True, seed=0)
        dt = time.time() - t0
        it = r[-1]["iters"]
        out["cpu_baseline"] = {"value": it / dt, "unit": "iterations/s", "cores": 1, "kind": "port",
                               "sample": f"{n_sub} rows x {m} labels, {it} iterations in {dt:.1f} s (numpy + C port; the "
                                         f"10^4-point line search does not depend on n, the weighted top-k and confusion "
                                         f"passes grow with n: an upper bound of the full-size CPU rate)"}
    return out


def bench_csr(args, device, what, n=153000, m=670000, nnz=100, cpu=True):
    """C4: Amazon-670K-shape CSR top-100 scores; what = "f1" | "recall" (BCA sessions) | "coverage" """
    import torch
    from xcolumns_b200 import _device as dev
    from xcolumns_b200 import metrics as M
    from xcolumns_b200._lib import XC_F32, XC_SUM_FAST
    from xcolumns_b200.block_coordinate import BcaSession, CoverageSession, _metric_params, coverage_batch_rows
    from xcolumns_b200.synth import csr_probs, csr_probs_device
    from xcolumns_b200.weighted_prediction import topk_csr_device
    peak, _ = load_peak()
    k = args.k
    data_t, idx_t, ptr_t = csr_probs_device(n, m, nnz, 1004, device)
    data = dev.CsrDev(data_t, idx_t, ptr_t, n, m, XC_F32, 0)
    init_pred = topk_csr_device(data, k, None, None)[0]
    sweep_no = [0]
    if what == "coverage":
        sess = CoverageSession(data, k, 1.0)
        batch = args.batch or coverage_batch_rows(n)
        order = torch.empty(n, dtype=torch.int32, device=device)

        def step():
            sweep_no[0] += 1
            sess.ctx.call("xc_permutation", n, C.c_uint64(3 + 7919 * sweep_no[0]), dev.ptr(order), sess._s())
            sess.sweep_batched(order, batch)
            sess.state(XC_SUM_FAST)
            sess.utility_device(1)

        def reset():
            sess.pred = init_pred.clone()
            sess.state(XC_SUM_FAST)
        name = "coverage@5"
    else:
        mid = M.XC_METRIC_FBETA if what == "f1" else M.XC_METRIC_RECALL
        params = _metric_params(mid, 1.0, 1e-9, True, True, n)
        sess = BcaSession(data, k, params, params, "mean")
        batch = args.batch or max(1, n // 8)

        def step():
            sweep_no[0] += 1
            sess.run_sweep(sweep_no[0], 3 + 7919 * sweep_no[0], True, batch, None, sweep_no[0] % 16 == 0, sess.util_buf[1:])

        def reset():
            sess.pred = init_pred.clone()
            sess.recompute(XC_SUM_FAST)
        name = "macro-F1@5" if what == "f1" else "macro-recall@5"
    reset()
    for _ in range(60):   # sweeps of this workload are < 1 ms: settle clocks and caches first
        step()
    reset()
    steps = max(args.steps, 20)
    l0 = sess.ctx.launches()
    ms = timed(step, steps, device)
    bytes_per_step = n * nnz * 8 + (n + 1) * 8
    ach = bytes_per_step * steps / (ms / 1e3) / 1e9
    out = {"metric": f"BCA {name} instances/sec per sweep (CSR)", "value": n * steps / (ms / 1e3), "unit": "instances/s",
           "n_gpus": 1, "steps": steps, "ms_per_step": ms / steps, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"amazon670k-shape CSR f32 n={n} m={m} nnz/row={nnz} k={k} {name} BCA (batched)",
                      "batch_rows": batch},
           "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "note": "algorithmic bytes = nnz*8 + (n+1)*8 per sweep (122 MB: 19 us at the HBM peak); a sweep of "
                                "this shape is bound by kernel-launch latency and L2 gathers of per-label state, see DESIGN.md"},
           "gpu_launches": int(sess.ctx.launches() - l0), "utility": float(sess.util_buf[1].item())}
    if cpu:
        from oracle import oracle as orc
        orc.build()
        n_sub = 20000
        y = csr_probs(n_sub, m, nnz, seed=1004)
        t0 = time.time()
        if what == "coverage":
            _, meta = orc.predict_optimizing_coverage_using_bc(y, k, seed=0, max_iters=3, tolerance=-np.inf)
        else:
            _, meta = orc.predict_using_bc_with_0approx(y, what, k, seed=0, skip_tn=True, max_iters=3, tolerance=-np.inf)
        dt = time.time() - t0
        out["cpu_baseline"] = {"value": n_sub * meta["iters"] / dt, "unit": "instances/s", "cores": 1, "kind": "port",
                               "sample": f"{n_sub} rows x {m} labels x {nnz} nnz/row, {meta['iters']} sequential sweeps in "
                                         f"{dt:.2f} s (C port; the Python reference runs ~10 k inst/s, SURVEY.md 8a-3)"}
    return out


def bench_exact_small(args, device, cpu=True):
    """C1 / C2 (BASELINE.json configs[0] and the small-m case): the bit-exact sequential mode through the public API
    on device-resident scores -- one launch per sweep, latency bound; the CPU port runs the same
    call and must report the same utilities bit for bit."""
    import torch
    from xcolumns_b200 import predict_optimizing_macro_f1_score_using_bc
    from xcolumns_b200.synth import dense_probs
    out = {"metric": "BCA macro-F1@5 instances/sec per sweep (sequential-exact mode)", "unit": "instances/s",
           "dtype": "f64 gains on f32 scores", "data": "synthetic", "configs": {}}
    for name, n, m, seed in (("C1", 10000, 1000, 1001), ("C2", 3800, 4000, 1002)):
        eta = dense_probs(n, m, seed=seed, tie_free=False)
        eta_d = torch.from_numpy(eta).to(device)
        best, meta = None, None
        for _ in range(3):
            torch.cuda.synchronize(device)
            t0 = time.time()
            _, meta = predict_optimizing_macro_f1_score_using_bc(eta_d, args.k, seed=0, mode="exact", return_meta=True,
                                                                 y_pred_format="indices")
            torch.cuda.synchronize(device)
            dt = time.time() - t0
            best = dt if best is None else min(best, dt)
        cfg = {"workload": f"dense f32 n={n} m={m} k={args.k} macro-F1 BCA (sequential, bit-exact)", "sweeps": meta["iters"],
               "value": n * meta["iters"] / best, "us_per_instance": 1e6 * best / (n * meta["iters"]),
               "ms_per_call": 1e3 * best, "utility": meta["utilities"][-1]}
        if cpu:
            from oracle import oracle as orc
            orc.build()
            t0 = time.time()
            _, ometa = orc.predict_using_bc_with_0approx(eta, "f1", args.k, seed=0, skip_tn=True)
            dtc = time.time() - t0
            cfg["cpu_baseline"] = {"value": n * ometa["iters"] / dtc, "unit": "instances/s", "cores": 1, "kind": "port",
                                   "sample": f"the whole call ({ometa['iters']} sweeps in {dtc:.2f} s)"}
            cfg["utilities_bit_equal_to_cpu_port"] = bool(list(meta["utilities"]) == list(ometa["utilities"]))
        out["configs"][name] = cfg
    out["value"] = out["configs"]["C1"]["value"]
    return out


# --------------------------------------------------------------------------------------------
# multi-GPU parity: a small sharded problem against the single-GPU sequential (bit-pinned) mode
# --------------------------------------------------------------------------------------------

def parity_multi_gpu(device, comm, k=5, rows_per_rank=5000, m=2000):
    import torch
    import torch.distributed as dist
    from xcolumns_b200 import predict_optimizing_macro_f1_score_using_bc
    from xcolumns_b200.distributed import shard_rows
    from xcolumns_b200.synth import dense_probs_device
    n = rows_per_rank * comm.world
    full = dense_probs_device(n, m, seed=977, device=device)        # the same matrix on every rank
    lo, hi = shard_rows(n, comm.rank, comm.world)
    pred_sh, meta_sh = predict_optimizing_macro_f1_score_using_bc(full[lo:hi].contiguous(), k, seed=0, mode="batched",
                                                                  distributed=True, return_meta=True,
                                                                  y_pred_format="indices")
    gathered = [torch.empty((shard_rows(n, r, comm.world)[1] - shard_rows(n, r, comm.world)[0], k), dtype=torch.int32,
                            device=device) for r in range(comm.world)]
    dist.all_gather(gathered, pred_sh.contiguous())
    out = None
    if comm.rank == 0:
        pred_all = torch.cat(gathered)
        # utility of the gathered sharded prediction, recomputed on ONE GPU from the whole matrix
        from xcolumns_b200 import _device as dev
        from xcolumns_b200 import metrics as M
        from xcolumns_b200._lib import XC_F32, XC_SUM_FAST
        from xcolumns_b200.block_coordinate import BcaSession, _metric_params
        params = _metric_params(M.XC_METRIC_FBETA, 1.0, 1e-9, True, True, n)
        s1 = BcaSession(dev.DenseDev(full, n, m, m, XC_F32, 0), k, params, params, "mean")
        s1.pred = pred_all.contiguous()
        s1.recompute(XC_SUM_FAST)
        s1.utility_device(0)
        u_re = float(s1.util_buf[0].item())
        del s1
        _, meta_1 = predict_optimizing_macro_f1_score_using_bc(full, k, seed=0, mode="batched", return_meta=True,
                                                               y_pred_format="indices")
        _, meta_x = predict_optimizing_macro_f1_score_using_bc(full, k, seed=0, mode="exact", return_meta=True,
                                                               y_pred_format="indices")
        u_sh, u_1, u_x = meta_sh["utilities"][-1], meta_1["utilities"][-1], meta_x["utilities"][-1]
        out = {"rows": n, "labels": m, "ranks": comm.world, "commit": meta_sh.get("commit"),
               "utility_sharded": u_sh, "utility_sharded_recomputed_on_one_gpu": u_re,
               "utility_single_gpu_batched": u_1, "utility_single_gpu_sequential_exact": u_x,
               "abs_diff_sharded_vs_exact": abs(u_sh - u_x), "abs_diff_sharded_vs_recomputed": abs(u_sh - u_re),
               "sweeps": [meta_sh["iters"], meta_1["iters"], meta_x["iters"]],
               "tolerance": 1e-4,
               "ok": bool(abs(u_sh - u_x) < 1e-4 and abs(u_sh - u_re) < 1e-9 and abs(u_1 - u_x) < 1e-4)}
    comm.barrier()
    del full
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------
# the headline workload
# --------------------------------------------------------------------------------------------

def bench_dense(args, device, comm, n_local, n_global_hint, sampler=None, want_roofline=True, seed_base=1003):
    """Timed batched sweeps over this rank's n_local rows (device resident).  Returns (dict, eta tensor)."""
    import torch
    from xcolumns_b200 import _device as dev
    from xcolumns_b200 import metrics as M
    from xcolumns_b200._lib import XC_F32, XC_SUM_FAST
    from xcolumns_b200.block_coordinate import BcaSession, _metric_params, default_batch_rows
    from xcolumns_b200.synth import dense_probs_device
    from xcolumns_b200.weighted_prediction import topk_dense_device
    m, k = args.labels, args.k
    eta_t = dense_probs_device(n_local, m, seed=seed_base + comm.rank, device=device)
    data = dev.DenseDev(eta_t, n_local, m, m, XC_F32, 0)
    n_global = comm.n_global(n_local)
    params = _metric_params(M.XC_METRIC_FBETA, 1.0, 1e-9, True, True, n_global)
    sess = BcaSession(data, k, params, params, "mean", comm)
    init_pred = topk_dense_device(data, k, None, None, XC_F32)[0]
    batch = args.batch or default_batch_rows(n_local, sess.wave_rows())
    n_batches = comm.max_int((n_local + batch - 1) // batch)
    sweep_no = [0]

    util = torch.zeros(4096, dtype=torch.float64, device=device)

    def one_sweep():   # what the public driver enqueues per sweep (order, snapshot, batches + commits, utility)
        sweep_no[0] += 1
        j = sweep_no[0]
        sess.run_sweep(j, 17 + 1000003 * comm.rank + 7919 * j, True, batch, n_batches, j % 16 == 0, util[j % 4096:])

    def reset():
        sess.pred = init_pred.clone()
        sess.recompute(XC_SUM_FAST)
        sess.utility_device(0)

    reset()
    for _ in range(args.warmup):
        one_sweep()
    reset()
    torch.cuda.synchronize(device)
    if sampler is not None:
        sampler.ready.wait(timeout=30)
        sampler.active = True
    launches0 = sess.ctx.launches()
    if want_roofline:
        sess.ctx.call("xc_timing_enable", 2 if os.environ.get("BENCH_DUMP_TIMELINE") else 1)

    # the queued sweeps are still executing when the host has enqueued them all: one clock sample under load
    ms = timed(one_sweep, args.steps, device, comm, after_enqueue=(sampler.sample_now if sampler is not None else None))
    if sampler is not None:
        sampler.active = False
    launches = sess.ctx.launches() - launches0
    sess.join()
    utilities = [0.0, float(util[sweep_no[0] % 4096].item())]
    res = {"ms": ms, "value": n_global * args.steps / (ms / 1e3), "n_global": n_global, "batch": batch,
           "n_batches": n_batches, "launches": int(launches), "utility": utilities[1], "lag": sess.lag,
           "commit": ("peer-memory" if sess.peer is not None else ("all-reduce" if comm.world > 1 else "local")),
           "pipe": sess.pipe}
    if want_roofline:
        st, en, rows = read_kernel_timing(sess.ctx)
        sess.ctx.call("xc_timing_enable", 0)
        peak, peak_src = load_peak()
        traffic = load_traffic()
        if os.environ.get("BENCH_DUMP_TIMELINE") and len(st):   # diagnostics: last sweep incl. commits (rows = 0), ms
            per_all = len(st) // args.steps
            t0 = float(st[-per_all:].min())
            res["timeline_last_sweep"] = [[round(float(a - t0), 4), round(float(b - t0), 4), int(r)]
                                          for a, b, r in zip(st[-per_all:], en[-per_all:], rows[-per_all:])]
            keep = rows > 0
            st, en, rows = st[keep], en[keep], rows[keep]
        if len(st):
            per = len(st) // args.steps
            busy = [union_ms(st[i * per:(i + 1) * per], en[i * per:(i + 1) * per]) for i in range(args.steps)]
            busy_ms = union_ms(st, en)   # over the whole timed region (consecutive sweeps overlap, too)
            bytes_total = float(rows.sum()) * m * 4
            achieved = bytes_total / (busy_ms / 1e3) / 1e9
            res["roofline"] = {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (traffic or {}).get("dram_bytes_per_launch"),
                "traffic_detail": traffic, "kernel": "bca_batch_dense_kernel<float,1>",
                "bytes_per_launch": bytes_total / len(st), "avg_launch_ms": busy_ms / len(st),
                "launches": int(len(st)),
                "timing": ("CUDA events around every launch on the stream it is launched on; consecutive batches run on two "
                           "streams and overlap (commits are applied one batch late), so the kernel's time is the UNION of "
                           "its [start, end] intervals, divided by the launches" if sess.lag else
                           "CUDA events around every launch on its stream"),
                "mean_event_interval_ms": float(np.mean(en - st)),
                "kernel_share_of_step": busy_ms / ms,
                "kernel_ms_per_sweep": [round(b, 4) for b in busy],
                "peak_source": peak_src, "frac_of_nominal_8tbs": achieved / 8000.0,
                "whole_step": {"achieved": n_local * m * 4 / (ms / args.steps / 1e3) / 1e9,
                               "frac": n_local * m * 4 / (ms / args.steps / 1e3) / 1e9 / peak,
                               "note": "score-matrix bytes of one sweep (this rank's shard) / ms_per_step"}}
    sess.close()
    del sess, init_pred
    return res, eta_t


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from xcolumns_b200 import predict_optimizing_macro_f1_score_using_bc
    from xcolumns_b200.distributed import make_comm, shard_rows

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    comm = make_comm(world > 1, device)

    if args.workload != "bca_dense":
        if args.workload == "fw_dense":
            out = bench_fw(args, device, comm, cpu=not args.no_cpu)
        elif rank == 0:
            what = {"bca_csr": "f1", "bca_csr_recall": "recall", "coverage_csr": "coverage"}[args.workload]
            out = bench_csr(args, device, what, cpu=not args.no_cpu)
        if rank == 0:
            print(json.dumps(out))
        if world > 1:
            dist.destroy_process_group()
        return

    n, m, k = args.rows, args.labels, args.k
    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- N > 1: sharded == single GPU?  (before anything is timed) ------------------------------------------
    parity = None
    if world > 1 and not args.no_parity:
        parity = parity_multi_gpu(device, comm, k)

    # ---- headline: the SAME n rows, sharded over the ranks (strong scaling; N = 1: the whole matrix) ---------
    lo, hi = shard_rows(n, rank, world)
    n_local = hi - lo
    res, eta_t = bench_dense(args, device, comm, n_local, n, sampler)
    sampler.stop_flag = True
    ms = res["ms"]
    line = {
        "metric": "BCA macro-F1@5 instances/sec per sweep", "value": res["value"], "unit": "instances/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows_total": n, "rows_per_gpu": n_local, "labels": m, "k": k, "mode": "batched",
                   "batch_rows_per_gpu": res["batch"], "commits_per_sweep": res["n_batches"], "commit_lag": res["lag"],
                   "l2": f"input {n_local * m * 4 / 1e9:.2f} GB per GPU >> 126 MB L2, no flush needed",
                   "collective": ("none" if world == 1 else
                                  "peer-memory commit kernel (flags + P2P reads of the 3*m float64 deltas over NVLink) per "
                                  "batch, overlapped with the next batch's streaming; NCCL all-reduce of 2*m float64 when "
                                  "the state is recomputed" if res["commit"] == "peer-memory"
                                  else "NCCL all-reduce of 3*m float64 deltas per commit")},
        "roofline": res.get("roofline"), "gpu_launches": res["launches"], "clocks": sampler.summary(),
        "utility_after_timed_sweeps": res["utility"],
    }
    if parity is not None:
        line["parity_multi_gpu"] = parity
    if "timeline_last_sweep" in res:
        line["timeline_last_sweep"] = res["timeline_last_sweep"]

    # ---- end-to-end through the public API with host buffers (rank-local shard) ----------------
    if not args.no_e2e:
        host = torch.empty((n_local, m), dtype=torch.float32, pin_memory=True)
        host.copy_(eta_t)
        del eta_t
        torch.cuda.empty_cache()
        y_np = host.numpy()
        os.environ["XCOLUMNS_B200_TIMING"] = "1" if os.environ.get("BENCH_E2E_PHASES") else "0"

        def e2e_call(src):
            comm.barrier()
            torch.cuda.synchronize(device)
            t0 = time.time()
            pred, meta = predict_optimizing_macro_f1_score_using_bc(
                src, k, seed=0, mode="batched", max_iters=args.steps, tolerance=-np.inf, return_meta=True,
                distributed=(world > 1))
            torch.cuda.synchronize(device)
            dt = time.time() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=device)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            assert pred.shape == (n_local, m)
            return dt, meta

        dt, meta = e2e_call(y_np)
        line["e2e"] = {"value": n * meta["iters"] / dt, "unit": "instances/s", "rows_per_gpu": n_local,
                       "h2d_bytes_per_step": n_local * m * 4 / meta["iters"],
                       "d2h_bytes_per_step": n_local * k * 4 / meta["iters"],
                       "seconds_per_call": dt, "sweeps_per_call": meta["iters"], "phases_s": meta.get("timings"),
                       "what": "predict_optimizing_macro_f1_score_using_bc(numpy pinned host array, this rank's row "
                               "shard) -> dense numpy y_pred; bytes are per GPU"}
        # what reference users pass: an ordinary (pageable) numpy array
        if world == 1 or n_local * m * 4 <= (6 << 30):
            pageable = np.empty_like(y_np)
            np.copyto(pageable, y_np)
            del y_np, host
            dt2, meta2 = e2e_call(pageable)
            line["e2e"]["pageable_input"] = {"value": n * meta2["iters"] / dt2, "seconds_per_call": dt2,
                                             "note": "same call on a pageable numpy array (chunked double-buffered "
                                                     "pinned staging inside dense_to_device)"}
            del pageable
        else:
            del y_np, host
    else:
        del eta_t
    torch.cuda.empty_cache()

    # ---- N > 1: the weak-scaling figure (n rows PER GPU), device resident, no roofline leg ---------------------
    if world > 1 and not args.no_weak:
        wres, weta = bench_dense(args, device, comm, n, n * world, None, want_roofline=False, seed_base=2003)
        del weta
        torch.cuda.empty_cache()
        line["weak_scaling"] = {"value": wres["value"], "unit": "instances/s", "rows_per_gpu": n,
                                "ms_per_step": wres["ms"] / args.steps, "commits_per_sweep": wres["n_batches"]}

    # ---- secondary workloads and the CPU baseline (rank 0, N = 1) ------------------------------------------------
    if world == 1 and not args.no_secondary:
        sec = {}
        for name, fn in (("fw_dense", lambda: bench_fw(args, device, comm, cpu=not args.no_cpu)),
                         ("bca_csr_recall", lambda: bench_csr(args, device, "recall", cpu=not args.no_cpu)),
                         ("coverage_csr", lambda: bench_csr(args, device, "coverage", cpu=not args.no_cpu)),
                         ("bca_csr_f1", lambda: bench_csr(args, device, "f1", cpu=False)),
                         ("bca_exact_small", lambda: bench_exact_small(args, device, cpu=not args.no_cpu))):
            try:
                sec[name] = fn()
            except Exception as e:  # a secondary line must never take the headline down
                sec[name] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()
        line["secondary"] = sec
    if rank == 0 and world == 1 and not args.no_cpu:
        from xcolumns_b200.synth import dense_probs
        n_sub = args.cpu_rows
        eta_sub = dense_probs(n_sub, m, seed=1003, tie_free=False)
        v, dt, iters = cpu_port_sample(eta_sub, k, 3)
        line["cpu_baseline"] = {"value": v, "unit": "instances/s", "cores": 1, "kind": "port",
                                "sample": f"{n_sub} rows x {m} labels, {iters} sequential sweeps in {dt:.1f} s "
                                          f"(C port of the reference algorithm, 1 core: the chain of instance updates is "
                                          f"sequential; the Python reference itself runs ~1.0k inst/s at this m, "
                                          f"BASELINE.md table B -- it is pure Python + numba and cannot travel to this box)"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
