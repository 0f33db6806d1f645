"""GPU probe (development aid, not a test): times the Frank-Wolfe iterate kernel with CUDA events for the
rows-per-warp variants ($XCOLUMNS_B200_DENSE_R) and checks the selection against torch.topk.

    python scripts/probe_fw.py [n] [m]
"""
import ctypes as C
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(r, n, m):
    env = dict(os.environ)
    if r:
        env["XCOLUMNS_B200_DENSE_R"] = r
    env["XC_PROBE_CHILD"] = "1"
    out = subprocess.run([sys.executable, __file__, str(n), str(m)], env=env, capture_output=True, text=True)
    print(f"[rows per warp: {r or 'auto'}]", out.stdout.strip(), out.stderr.strip()[-400:])


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 14000
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 31000
    if "XC_PROBE_CHILD" not in os.environ:
        for r in ("", "1", "2", "4"):
            child(r, n, m)
        return
    from xcolumns_b200 import _device as dev
    from xcolumns_b200.synth import dense_probs_device

    device = torch.device("cuda", 0)
    ctx = dev.ctx_for(device)
    eta = dense_probs_device(n, m, seed=1005, device=device)
    g = torch.Generator(device=device).manual_seed(3)
    a = torch.rand(m, device=device, generator=g) + 0.5
    b = torch.rand(m, device=device, generator=g) - 0.5
    raw = torch.zeros((2, m), dtype=torch.float64, device=device)
    pred = torch.empty((n, 5), dtype=torch.int32, device=device)
    sp = dev.stream_ptr(device)

    def call(with_pred):
        ctx.call("xc_fw_iterate_dense", dev.ptr(eta), 0, n, m, eta.stride(0), dev.ptr(eta), eta.stride(0), dev.ptr(a),
                 dev.ptr(b), 5, C.c_void_p(raw[0].data_ptr()), C.c_void_p(raw[1].data_ptr()),
                 dev.ptr(pred) if with_pred else None, sp)

    call(True)
    torch.cuda.synchronize()
    ref = torch.topk(eta * a + b, 5, dim=1).indices.sort(dim=1).values.int()
    ok = bool((ref == pred).all())
    for _ in range(3):
        call(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        call(False)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"iterate {us:.1f} us  {n * m * 4 / us / 1e3:.0f} GB/s  parity_vs_torch={ok} cnt_sum={raw[1].sum().item():.0f}")


if __name__ == "__main__":
    main()
