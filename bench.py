#!/usr/bin/env python
"""Benchmark of the BCA macro-F1@5 hot path (BASELINE.json metric: instances/sec per sweep).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rows n] [--labels m]

A "step" is one block-Jacobi BCA sweep (macro-F1@5) over the whole synthetic AmazonCat-13K-shape
probability matrix (n=307000, m=13000, float32, 15.96 GB -- far larger than the 126 MB L2, so no
L2 flush is needed between steps).  With N > 1 (torchrun, one rank per GPU) every rank holds its
own n-row shard (weak scaling) and the per-batch confusion deltas are all-reduced over NCCL.

Printed JSON (one line, rank 0):
  value     : instances/sec/sweep with y_proba resident in HBM, CUDA-event timed, max over ranks
  e2e       : same metric through the public Python API with HOST (pinned) buffers: H2D of
              y_proba, top-k init, K sweeps, D2H + host materialisation of the prediction
  roofline  : the dominant kernel (bca_batch_dense_kernel): algorithmic bytes per launch
              (rows * m * 4) / CUDA-event duration per launch vs the measured HBM copy peak
  cpu_baseline : the CPU oracle (C port of the reference algorithm, 1 core) on a row subsample
`--impl reference` times that same CPU port (the reference is pure Python + numba and cannot
travel to the GPU box) and prints the same line with "impl": "reference".
"""
import argparse
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
    ap.add_argument("--rows", type=int, default=307000)
    ap.add_argument("--labels", type=int, default=13000)
    ap.add_argument("--k", type=int, default=5)
    ap.add_argument("--batch", type=int, default=0, help="rows per commit per rank (0 = default)")
    ap.add_argument("--cpu-rows", type=int, default=1500, help="row subsample of the CPU baseline")
    ap.add_argument("--workload", default="bca_dense", choices=["bca_dense", "bca_csr", "fw_dense"],
                    help="bca_dense is the headline (BASELINE.json metric); the others are secondary lines")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


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
    per_step = []
    for _ in range(args.warmup):
        cpu_port_sample(eta, args.k, 1)
    t0 = time.time()
    for _ in range(args.steps):
        v, dt, _ = cpu_port_sample(eta, args.k, 1)
        per_step.append(dt)
    total = time.time() - t0
    value = n_sub * args.steps / total
    line = {
        "impl": "reference", "metric": "BCA macro-F1@5 instances/sec per sweep", "value": value,
        "unit": "instances/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows": args.rows, "labels": args.labels, "k": args.k},
        "cpu_baseline": {"value": value, "unit": "instances/s", "cores": 1, "kind": "port",
                         "sample": f"{n_sub} rows x {args.labels} labels of the same distribution, 1 sequential sweep per step "
                                   f"(a sweep's per-instance cost does not depend on n)"},
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

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
            self.ready.set()
            while not self.stop_flag:
                clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                if self.active:
                    self.sm.append(clk)
                    for bit, nm in names.items():
                        if r & bit:
                            self.reasons.add(nm)
                time.sleep(0.001)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")
            self.ready.set()

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------

def secondary(args):
    """Secondary workloads (not the headline line): C4-shape CSR BCA and C5-shape dense Frank-Wolfe,
    device-resident inputs, CUDA-event timed."""
    import torch

    from xcolumns_b200 import _device as dev
    from xcolumns_b200 import metrics as M
    from xcolumns_b200._lib import XC_F32, XC_SUM_FAST
    from xcolumns_b200.block_coordinate import BcaSession, _metric_params
    from xcolumns_b200.synth import csr_probs_device, dense_probs_device
    from xcolumns_b200.weighted_prediction import topk_csr_device

    device = torch.device("cuda", 0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6650.0))
    k = args.k
    if args.workload == "bca_csr":
        n, m, nnz = (args.rows if args.rows != 307000 else 153000), 670000, 100
        data_t, idx_t, ptr_t = csr_probs_device(n, m, nnz, 1004, device)
        data = dev.CsrDev(data_t, idx_t, ptr_t, n, m, XC_F32, 0)
        params = _metric_params(M.XC_METRIC_FBETA, 1.0, 1e-9, True, True, n)
        sess = BcaSession(data, k, params, params, "mean")
        init_pred = topk_csr_device(data, k, None, None)[0]
        batch = args.batch or max(1, n // 8)
        sweep_no = [0]

        def step():
            sweep_no[0] += 1
            order = sess.permutation(n, 3 + 7919 * sweep_no[0])
            sess.zero_delta()
            sess.sweep_batched(order, batch)
            sess.finish_sweep(full=(sweep_no[0] % 16 == 0))
            sess.utility_device(1)

        def reset():
            sess.pred = init_pred.clone()
            sess.recompute(XC_SUM_FAST)

        bytes_per_step = n * nnz * 8 + (n + 1) * 8
        unit_count, metric, unit = n, "BCA macro-F1@5 instances/sec per sweep (CSR)", "instances/s"
        wl = f"amazon670k-shape CSR f32 n={n} m={m} nnz/row={nnz} k={k} macro-F1 BCA (batched)"
    else:
        import torch.distributed as dist
        from xcolumns_b200.frank_wolfe import find_classifier_using_fw
        world = int(os.environ.get("WORLD_SIZE", "1"))
        rank = int(os.environ.get("RANK", "0"))
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local_rank)
        device = torch.device("cuda", local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=device)
        n, m = (args.rows if args.rows != 307000 else 14000), (args.labels if args.labels != 13000 else 31000)
        eta_t = dense_probs_device(n, m, seed=1005 + rank, device=device)   # weak scaling: n rows per GPU
        state = {}

        # FW runs through the public API on the device tensor (zero copy); per-iteration time from
        # CUDA events around the whole call divided by the iterations it performed
        def run(iters):
            if world > 1:
                dist.barrier()
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

        run(args.warmup)
        ms_call, iters = run(args.steps)
        # the reference's own clock: meta["time"] covers the iteration loop (after the initial
        # classifier's pass), so one-time setup (column sums, pinned buffers) is not smeared over
        # the iterations; the whole-call figure is reported next to it
        ms = state["meta"]["time"] * 1e3
        if world > 1:
            t = torch.tensor([ms, ms_call], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, ms_call = float(t[0]), float(t[1])
        bytes_per_step = n * m * 4
        ach = bytes_per_step * iters / (ms / 1e3) / 1e9   # per GPU
        line = {"metric": "Frank-Wolfe macro-F1@5 iterations/sec", "value": iters / (ms / 1e3), "unit": "iterations/s",
                "n_gpus": world, "steps": iters, "warmup": args.warmup, "ms_per_step": ms / iters,
                "ms_per_call": ms_call, "instances_per_s": n * world * iters / (ms / 1e3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic (y_true := y_proba)",
                "config": {"workload": f"wiki10-31k-shape dense f32 n={n} rows per GPU, m={m} k={k} macro-F1 FW, "
                                       f"{iters} iterations incl. init pass",
                           "collective": "none" if world == 1 else "NCCL all-reduce of 2*m float64 per iterate"},
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "traffic": None, "note": "iteration loop: one pass over the rank's y_proba shard per "
                                                      "iteration / loop time (max over ranks)"},
                "utilities": state["meta"]["utilities"][-3:]}
        if rank == 0:
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return

    reset()
    # the sweeps of this workload are ~0.5 ms: warm up for >= 100 sweeps so that clocks and caches are settled
    for _ in range(max(args.warmup, 100)):
        step()
    reset()
    torch.cuda.synchronize()
    l0 = sess.ctx.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ach = bytes_per_step * args.steps / (ms / 1e3) / 1e9
    line = {"metric": metric, "value": unit_count * args.steps / (ms / 1e3), "unit": unit, "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl, "batch_rows": batch},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                         "note": "algorithmic bytes = nnz*8 + (n+1)*8 per sweep; the sweep is bound by L2 gathers of "
                                 "the per-label coefficients and launch latency, not by HBM"},
            "gpu_launches": int(sess.ctx.launches() - l0), "utility": float(sess.util_buf[1].item())}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload != "bca_dense":
        secondary(args)
        return

    import torch
    import torch.distributed as dist

    from xcolumns_b200 import _device as dev
    from xcolumns_b200 import metrics as M
    from xcolumns_b200 import predict_optimizing_macro_f1_score_using_bc
    from xcolumns_b200._lib import XC_F32, XC_SUM_FAST
    from xcolumns_b200.block_coordinate import BcaSession, _metric_params, default_batch_rows
    from xcolumns_b200.distributed import make_comm
    from xcolumns_b200.synth import dense_probs_device
    from xcolumns_b200.weighted_prediction import topk_dense_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    comm = make_comm(world > 1, device)

    n, m, k = args.rows, args.labels, args.k
    eta_t = dense_probs_device(n, m, seed=1003 + rank, device=device)
    data = dev.DenseDev(eta_t, n, m, m, XC_F32, 0)
    n_global = comm.n_global(n)
    params = _metric_params(M.XC_METRIC_FBETA, 1.0, 1e-9, True, True, n_global)
    sess = BcaSession(data, k, params, params, "mean", comm)
    init_pred = topk_dense_device(data, k, None, None, XC_F32)[0]
    batch = args.batch or default_batch_rows(n, sess.wave_rows())
    n_batches = comm.max_int((n + batch - 1) // batch)
    sweep_no = [0]

    def one_sweep(events=None):
        sweep_no[0] += 1
        order = sess.permutation(n, 17 + 1000003 * rank + 7919 * sweep_no[0])
        sess.zero_delta()
        sess.sweep_batched(order, batch, n_batches, events=events)
        sess.finish_sweep(full=(sweep_no[0] % 16 == 0))   # like the public driver: fold, recompute every 16th sweep
        sess.utility_device(1)

    def reset():
        sess.pred = init_pred.clone()
        sess.recompute(XC_SUM_FAST)
        sess.utility_device(0)

    sampler = ClockSampler(local_rank)
    sampler.start()
    reset()
    for _ in range(args.warmup):
        one_sweep()
    reset()
    torch.cuda.synchronize(device)
    sampler.ready.wait(timeout=30)
    comm.barrier()
    torch.cuda.synchronize(device)
    sampler.active = True
    launches0 = sess.ctx.launches()
    events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_sweep(events)
    e1.record()
    torch.cuda.synchronize(device)
    comm.barrier()
    torch.cuda.synchronize(device)
    sampler.active = False
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    launches = sess.ctx.launches() - launches0
    utilities = sess.util_buf[:2].cpu().tolist()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = n_global * args.steps / (ms / 1e3)

    # per-launch roofline of the dominant kernel
    kern_ms = [a.elapsed_time(b) for a, b, _ in events]
    kern_rows = [r for _, _, r in events]
    full = [(t_, r) for t_, r in zip(kern_ms, kern_rows) if r == batch] or list(zip(kern_ms, kern_rows))
    avg_ms = float(np.mean([t_ for t_, _ in full]))
    bytes_per_launch = full[0][1] * m * 4
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = bytes_per_launch / (avg_ms / 1e3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("bca_batch_dense_kernel")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "bca_batch_dense_kernel<float,1>",
                "bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_ms,
                "kernel_share_of_step": float(np.sum(kern_ms)) / ms,
                "kernel_ms_per_sweep": [round(float(np.sum(kern_ms[i * len(kern_ms) // args.steps:(i + 1) * len(kern_ms) // args.steps])), 4)
                                        for i in range(args.steps)],
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                "frac_of_nominal_8tbs": achieved / 8000.0,
                "whole_step": {"achieved": n * m * 4 / (ms / args.steps / 1e3) / 1e9,
                               "frac": n * m * 4 / (ms / args.steps / 1e3) / 1e9 / peak,
                               "note": "score-matrix bytes of one sweep / ms_per_step (per GPU)"}}

    line = {
        "metric": "BCA macro-F1@5 instances/sec per sweep", "value": value, "unit": "instances/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows_per_gpu": n, "labels": m, "k": k, "mode": "batched",
                   "batch_rows_per_gpu": batch, "commits_per_sweep": n_batches,
                   "l2": "input 15.96 GB per GPU >> 126 MB L2, no flush needed",
                   "collective": ("none" if world == 1 else
                                  "peer-memory commit kernel (flags + P2P reads of the 3*m float64 deltas over NVLink) "
                                  "between batches; NCCL all-reduce of 2*m float64 at sweep boundaries"
                                  if sess.peer is not None else "NCCL all-reduce of 3*m float64 deltas per commit")},
        "roofline": roofline, "gpu_launches": int(launches), "clocks": sampler.summary(),
        "utility_after_timed_sweeps": utilities[1],
    }

    # ---- end-to-end through the public API with host buffers (rank-local shard) ----------------
    sess.close()
    if not args.no_e2e:
        del sess, init_pred
        # every rank holds its input shard pinned plus the dense result: 2 * n * m * 4 bytes of host memory per
        # rank; shrink the e2e shard if the box cannot hold that for all local ranks (and say so)
        n_e2e = n
        try:
            import psutil
            avail = psutil.virtual_memory().available
            fit = int(0.6 * avail / (max(1, world) * 2 * m * 4))
            n_e2e = max(1024, min(n, fit))
        except Exception:
            pass
        if world > 1:
            t = torch.tensor([n_e2e], dtype=torch.int64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            n_e2e = int(t.item())
        host = torch.empty((n_e2e, m), dtype=torch.float32, pin_memory=True)
        host.copy_(eta_t[:n_e2e])
        del eta_t, data
        torch.cuda.empty_cache()
        y_np = host.numpy()
        comm.barrier()
        torch.cuda.synchronize(device)
        os.environ["XCOLUMNS_B200_TIMING"] = "1" if os.environ.get("BENCH_E2E_PHASES") else "0"
        t0 = time.time()
        pred, meta = predict_optimizing_macro_f1_score_using_bc(
            y_np, k, seed=0, mode="batched", max_iters=args.steps, tolerance=-np.inf, return_meta=True,
            distributed=(world > 1), batch_size=min(batch, max(1, n_e2e // 8)) if n_e2e < n else batch)
        torch.cuda.synchronize(device)
        dt = time.time() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        line["e2e"] = {"value": n_e2e * world * meta["iters"] / dt, "unit": "instances/s",
                       "rows_per_gpu": n_e2e,
                       "h2d_bytes_per_step": n_e2e * m * 4 / meta["iters"],
                       "d2h_bytes_per_step": n_e2e * k * 4 / meta["iters"],
                       "seconds_per_call": dt, "sweeps_per_call": meta["iters"], "phases_s": meta.get("timings"),
                       "what": "predict_optimizing_macro_f1_score_using_bc(numpy pinned host array) -> dense numpy y_pred"}
        assert pred.shape == (n_e2e, m)
        del pred

    # ---- CPU baseline on a bounded sample (rank 0, N = 1) -------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        from xcolumns_b200.synth import dense_probs
        n_sub = args.cpu_rows
        eta_sub = dense_probs(n_sub, m, seed=1003, tie_free=False)
        v, dt, iters = cpu_port_sample(eta_sub, k, 3)
        line["cpu_baseline"] = {"value": v, "unit": "instances/s", "cores": 1, "kind": "port",
                                "sample": f"{n_sub} rows x {m} labels, {iters} sequential sweeps in {dt:.1f} s "
                                          f"(C port of the reference algorithm; the Python reference itself runs "
                                          f"~1.0k inst/s at this m, BASELINE.md table B)"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
