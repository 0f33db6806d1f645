"""Multi-GPU tests (skipped on a one-GPU box): scripts/multi_gpu_check.py under torchrun with two ranks -- sharded
result == single-GPU result, peer-memory protocol under injected delay, sharded coverage, Frank-Wolfe initial
classifiers -- and the single-process several-GPU use of the C ABI (inputs on a non-current device)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_paths_two_ranks():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "MULTI_GPU_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_inputs_on_a_non_current_device(oracle):
    """one process, two GPUs: tensors on cuda:1 while cuda:0 is current (and the other way round); every entry
    point runs on its context's device and leaves the caller's current device alone"""
    import xcolumns_b200 as xb
    from xcolumns_b200.synth import dense_probs
    eta = dense_probs(600, 300, seed=3)
    want = oracle.topk_indices_dense(eta, 5)[0]
    opred, ometa = oracle.predict_using_bc_with_0approx(eta, "f1", 5, seed=0, skip_tn=True)
    for cur, other in ((0, 1), (1, 0), (0, 1)):
        torch.cuda.set_device(cur)
        t = torch.from_numpy(eta).to(f"cuda:{other}")
        top = xb.predict_top_k(t, 5)
        assert top.device.index == other and torch.cuda.current_device() == cur
        assert (torch.nonzero(top)[:, 1].reshape(-1, 5).cpu().numpy() == want).all()
        pred, meta = xb.predict_optimizing_macro_f1_score_using_bc(t, 5, seed=0, mode="exact", return_meta=True)
        assert pred.device.index == other and torch.cuda.current_device() == cur
        assert (pred.cpu().numpy().astype(np.uint8) == opred).all()
        _, mb = xb.predict_optimizing_macro_f1_score_using_bc(t, 5, seed=0, mode="batched", return_meta=True)
        assert abs(mb["utilities"][-1] - ometa["utilities"][-1]) < 1e-4 and torch.cuda.current_device() == cur
    torch.cuda.set_device(0)
