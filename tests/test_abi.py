"""The C-ABI library loads without a GPU and exports every symbol include/xcolumns_b200.h
declares; the ctypes prototypes in xcolumns_b200/_lib.py agree with the header, argument by
argument (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "xcolumns_b200.h")


def _declarations():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for mobj in re.finditer(r"XC_API\s+([\w\s\*]+?)\s*\b(xc_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        ret, name, args = mobj.group(1).strip(), mobj.group(2), mobj.group(3)
        ret = ret.replace("XC_API", "").strip()
        params = [a.strip() for a in args.replace("\n", " ").split(",")] if args.strip() != "void" else []
        decls[name] = (ret, params)
    return decls


def _ctype_of(param: str):
    if "*" in param:
        return "ptr"
    if param.startswith("int64_t"):
        return C.c_int64
    if param.startswith("uint64_t"):
        return C.c_uint64
    if param.startswith("double"):
        return C.c_double
    if param.startswith("int"):
        return C.c_int
    raise AssertionError(f"unparsed parameter {param!r}")


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as entry
    entry.build()
    from xcolumns_b200 import _lib
    lib = _lib.load()
    decls = _declarations()
    assert len(decls) >= 30
    for name in decls:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert lib.xc_abi_version() == 3
    assert lib.xc_strerror(-1) == b"invalid argument"


def test_ctypes_prototypes_match_header():
    from xcolumns_b200 import _lib
    decls = _declarations()
    for name, argtypes in _lib._SIGNATURES.items():
        assert name in decls, f"{name} bound in _lib.py but not declared"
        ret, params = decls[name]
        assert ret == "int"
        assert params[0].startswith("xc_ctx")
        params = params[1:]
        assert len(params) == len(argtypes), f"{name}: header has {len(params)} args after ctx, binding {len(argtypes)}"
        for p, t in zip(params, argtypes):
            want = _ctype_of(p)
            if want == "ptr":
                assert t in (C.c_void_p, _lib._MP), f"{name}: {p!r} bound as {t}"
                if "xc_metric_params" in p:
                    assert t is _lib._MP, f"{name}: {p!r} must be the params struct pointer"
            else:
                assert t is want, f"{name}: {p!r} bound as {t}"
    bound = set(_lib._SIGNATURES) | {"xc_abi_version", "xc_strerror", "xc_ctx_create", "xc_ctx_destroy",
                                     "xc_last_cuda_error", "xc_launch_count", "xc_sm_count", "xc_fill_pred_dense_host", "xc_scatter_pred_dense_host",
                                     "xc_zero_host", "xc_bca_coef_len", "xc_fw_alpha_scratch_bytes",
                                     "xc_fw_alpha_ctl_offset", "xc_p2p_payload", "xc_p2p_destroy", "xc_bca_delta_stride",
                                     "xc_bca_pipe_buffers", "xc_bca_window_bytes"}
    assert set(decls) == bound, f"unbound: {set(decls) - bound}, undeclared: {bound - set(decls)}"


def test_metric_params_layout():
    from xcolumns_b200 import _lib
    assert C.sizeof(_lib.MetricParams) == 80
    assert _lib.MetricParams.mix.offset == 12 and _lib.MetricParams.c1.offset == 16
    assert _lib.MetricParams.n_div.offset == 40 and _lib.MetricParams.mix_alpha.offset == 56
    assert _lib.MetricParams.mix_m.offset == 72


def test_no_gpu_fails_loudly():
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import xcolumns_b200 as xb
    from xcolumns_b200._lib import XColumnsB200Error
    with pytest.raises(XColumnsB200Error):
        xb.predict_top_k(np.random.rand(4, 8).astype(np.float32), 2)


def test_streaming_kernels_keep_their_occupancy():
    """The HBM-bound scan kernels are tuned for 40 registers (6 CTAs of 256 threads per SM, no local
    memory); a refactor of the shared scan loop that costs registers silently costs ~20 % of the
    headline throughput (measured), so the built library is checked."""
    import re
    import shutil
    import subprocess
    from xcolumns_b200 import _lib
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-res-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    usage = {}
    for m_ in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out):
        usage[m_.group(1)] = (int(m_.group(2)), int(m_.group(3)))
    checked = 0
    for name, (reg, stack) in usage.items():
        if "bca_batch_dense_kernelIfLi1ELb0E" in name or "fw_iterate_dense_kernelIfLi1E11XfMulAddVec" in name:
            assert reg <= 40 and stack == 0, (name, reg, stack)
            checked += 1
        if "bca_batch_dense_kernelIfLi1ELb1E" in name:   # deep-prefetch variant: 4 CTAs + one commit CTA per SM
            assert reg <= 56 and stack == 0, (name, reg, stack)
            checked += 1
        if "bca_commit_kernel" in name or "bca_push_kernel" in name:             # must fit next to six resident streaming CTAs
            assert reg <= 64 and stack == 0, (name, reg, stack)
            checked += 1
    assert checked >= 5, sorted(usage)[:5]


def test_dense_output_prefill_host_side(monkeypatch):
    """host-only half of the boundary: the dense result is cleared on background threads and only scattered
    at the end (csrc/host.cu: xc_zero_host / xc_scatter_pred_dense_host); an abandoned prefill must not crash"""
    import gc
    import time
    import numpy as np
    from xcolumns_b200 import _device as dev
    monkeypatch.setattr(dev.DenseOutputPrefill, "MIN_BYTES", 0)
    like = np.zeros((300, 50), dtype=np.float32)
    p = dev.DenseOutputPrefill.start(like, 300, 50)
    idx = np.tile(np.array([1, 7, 49], dtype=np.int32), (300, 1))
    idx[5] = [-1, 2, 3]                                  # unused slot
    vals = np.full((300, 3), 0.25, dtype=np.float32)
    out = p.finish(idx, vals)
    assert out.shape == (300, 50) and out.dtype == np.float32
    assert (out[0, [1, 7, 49]] == 0.25).all() and out[0].sum() == 0.75 and out[5].sum() == 0.5
    for dt in (np.float64, np.float32):
        q = dev.DenseOutputPrefill.start(like.astype(dt), 300, 50)
        o = q.finish(idx, None)
        assert o.dtype == dt and (o.sum(1)[:5] == 3).all()
    assert dev.DenseOutputPrefill.start(like.astype(np.int32), 300, 50) is None      # unsupported dtype: plain path
    abandoned = dev.DenseOutputPrefill.start(like, 300, 50)
    del abandoned
    gc.collect()
    time.sleep(0.1)


def test_pipe_args_layout():
    from xcolumns_b200 import _lib
    assert C.sizeof(_lib.PipeArgs) == 200   # static_assert'ed on the C side (csrc/bca_batched.cu)
    assert _lib.PipeArgs.seed.offset == 80 and _lib.PipeArgs.order.offset == 96
    assert _lib.PipeArgs.delta.offset == 152 and _lib.PipeArgs.util_tn_rows.offset == 184
