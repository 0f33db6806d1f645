"""ctypes binding of the C ABI (include/xcolumns_b200.h).

The product path has no CPU fallback: if the shared library is missing or no CUDA device is
visible every public call fails loudly with :class:`XColumnsB200Error`.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libxcolumns_b200.so")

XC_F32, XC_F64 = 0, 1
XC_SUM_FAST, XC_SUM_ORDERED = 0, 1


class XColumnsB200Error(RuntimeError):
    pass


class MetricParams(C.Structure):
    """Mirror of xc_metric_params."""
    _fields_ = [("metric", C.c_int32), ("maximize", C.c_int32), ("skip_tn", C.c_int32),
                ("mix", C.c_int32), ("c1", C.c_double), ("beta2", C.c_double),
                ("eps", C.c_double), ("n_div", C.c_double), ("n_rows", C.c_double),
                ("mix_alpha", C.c_double), ("mix_k", C.c_double), ("mix_m", C.c_double)]


_vp, _i32, _i64, _dbl, _int = C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_int
_MP = C.POINTER(MetricParams)

XC_PIPE_FORK, XC_PIPE_SHUFFLE = 1, 2


class PipeArgs(C.Structure):
    """Mirror of xc_bca_pipe_args (one pipelined dense sweep, include/xcolumns_b200.h)."""
    _fields_ = [("params", _MP), ("eta", C.c_void_p), ("dtype", C.c_int32), ("k", C.c_int32),
                ("m", C.c_int64), ("ld", C.c_int64), ("n_rows", C.c_int64),
                ("batch", C.c_int64), ("n_batches", C.c_int64), ("batch0", C.c_int64),
                ("lag", C.c_int32), ("flags", C.c_int32), ("seed", C.c_uint64), ("sweep", C.c_int64),
                ("order", C.c_void_p), ("coef", C.c_void_p), ("pred_idx", C.c_void_p), ("pred_snapshot", C.c_void_p),
                ("tp", C.c_void_p), ("fp", C.c_void_p), ("fn", C.c_void_p), ("delta", C.c_void_p),
                ("util_params", _MP), ("util_out", C.c_void_p), ("agg", C.c_int32), ("reserved", C.c_int32),
                ("util_tn_rows", C.c_double), ("prev_tail_from", C.c_int64)]

# name -> argtypes (after ctx); every function returns int unless listed in _RESTYPES
_SIGNATURES = {
    "xc_permutation": [_i64, C.c_uint64, _vp, _vp],
    "xc_topk_dense": [_vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _int, _int, _vp, _vp, _vp],
    "xc_topk_csr": [_vp, _int, _vp, _vp, _i64, _vp, _vp, _int, _vp, _vp, _vp],
    "xc_threshold_dense": [_vp, _int, _i64, _i64, _i64, _vp, _vp, _int, _dbl, _vp, _i64, _vp],
    "xc_scatter_pred_dense": [_vp, _vp, _int, _int, _i64, _vp, _int, _i64, _vp],
    "xc_confmat_dense": [_vp, _i64, _vp, _i64, _int, _i64, _i64, _int, _int, _int, _vp, _vp, _vp, _vp],
    "xc_confmat_dense_compact": [_vp, _int, _i64, _vp, _int, _i64, _i64, _int, _vp, _vp, _vp, _vp, _vp],
    "xc_confmat_csr": [_vp, _vp, _vp, _vp, _vp, _vp, _int, _i64, _i64, _int, _int, _vp, _vp, _vp, _vp],
    "xc_confmat_csr_compact": [_vp, _vp, _vp, _int, _vp, _int, _i64, _i64, _int, _vp, _vp, _vp, _vp],
    "xc_confmat_csc_ordered": [_vp, _int, _vp, _vp, _vp, _vp, _vp, _int, _i64, _i64, _vp, _vp, _vp, _vp, _vp],
    "xc_p2p_create": [_int, _int, _i64, _vp, _vp],
    "xc_p2p_open": [_vp, _vp],
    "xc_p2p_error": [_vp, _vp],
    "xc_bca_pipe_sweep": [_vp, _vp, _vp],
    "xc_bca_pipe_join": [_vp],
    "xc_h2d_staged": [_vp, _i64, _vp, _i64, _i64, _i64, _int, _vp],
    "xc_timing_enable": [_int],
    "xc_timing_read": [_int, _vp, _vp, _vp, _vp],
    "xc_bca_online_csr": [_vp, _int, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _int, _MP, _vp, _vp, _vp, _vp, _vp, _vp, _i64,
                          _vp],
    "xc_bca_online_dense": [_vp, _int, _i64, _i64, _i64, _vp, _i64, _int, _MP, _vp, _vp, _vp, _vp, _vp, _vp],
    "xc_bca_batch_csr_rec": [_MP, _vp, _int, _vp, _vp, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "xc_bca_rec": [_MP, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp],
    "xc_bca_batch_dense_rec": [_MP, _vp, _int, _i64, _i64, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "xc_bca_sweep_dense": [_MP, _vp, _int, _i64, _i64, _vp, _i64, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                           _vp],
    "xc_bca_sweep_csr": [_MP, _vp, _int, _vp, _vp, _i64, _vp, _i64, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                         _vp, _int, _vp, _vp, _vp, _int, _vp],
    "xc_bca_fold_touched": [_MP, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "xc_colsum_dense": [_vp, _int, _i64, _i64, _i64, _vp, _vp],
    "xc_colsum_csr": [_vp, _int, _vp, _i64, _i64, _vp, _vp],
    "xc_utility": [_MP, _int, _vp, _vp, _vp, _vp, _i64, _vp, _vp],
    "xc_bca_exact_sweep_dense": [_vp, _int, _i64, _i64, _i64, _vp, _i64, _int, _MP, _int, _vp, _vp, _vp, _vp, _vp, _vp],
    "xc_bca_exact_sweep_dense_k0": [_vp, _int, _i64, _i64, _i64, _vp, _i64, _MP, _int, _vp, _i64, _vp, _vp, _vp, _vp, _vp],
    "xc_bca_exact_sweep_csr": [_vp, _int, _vp, _vp, _i64, _i64, _vp, _i64, _int, _MP, _int, _vp, _vp, _vp, _vp, _vp, _vp],
    "xc_cov_exact_sweep_csr": [_vp, _int, _vp, _vp, _i64, _i64, _vp, _i64, _int, _dbl, _int, _vp, _vp, _vp],
    "xc_cov_state_csr": [_vp, _int, _vp, _vp, _i64, _i64, _vp, _int, _int, _vp, _vp],
    "xc_bca_coef": [_MP, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp],
    "xc_bca_wave_rows": [_int, _i64],
    "xc_bca_batch_dense": [_vp, _int, _i64, _i64, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "xc_bca_batch_csr": [_vp, _int, _vp, _vp, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp, _vp, _vp, _vp],
    "xc_cov_batch_csr": [_vp, _int, _vp, _vp, _vp, _i64, _int, _dbl, _vp, _vp, _vp, _vp],
    "xc_cov_batch_dense": [_vp, _int, _i64, _i64, _vp, _i64, _int, _dbl, _vp, _vp, _vp, _vp],
    "xc_cov_fold": [_vp, _vp, _i64, _vp],
    "xc_cov_sweep_csr": [_vp, _int, _vp, _vp, _i64, _vp, _i64, _i64, _int, _dbl, _vp, _vp, _vp, _vp],
    "xc_cov_sweep_dense": [_vp, _int, _i64, _i64, _vp, _i64, _i64, _int, _dbl, _vp, _vp, _vp, _vp],
    "xc_cov_state_dense": [_vp, _int, _i64, _i64, _i64, _vp, _int, _vp, _vp],
    "xc_cov_utility": [_vp, _i64, _vp, _vp],
    "xc_fw_iterate_dense": [_vp, _int, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _int, _vp, _vp, _vp, _vp],
    "xc_fw_iterate_csr": [_vp, _int, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _int, _vp, _vp, _vp, _vp],
    "xc_fw_make_conf": [_vp, _vp, _vp, _i64, _dbl, _int, _int, _vp, _vp],
    "xc_fw_metric_grad": [_MP, _vp, _i64, _vp, _vp, _vp, _vp],
    "xc_fw_alpha_search": [_MP, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp],
    "xc_fw_combine": [_vp, _vp, _i64, _vp, _vp],
    "xc_fw_step_begin": [_vp, _int, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _int, _vp, _int, _vp],
    "xc_fw_step_finish": [_MP, _int, _vp, _vp, _i64, _dbl, _int, _int, _vp, _vp, _vp, _i64, _dbl, _vp, _vp, _vp, _vp,
                          _vp, _int, _dbl, _vp],
    "xc_fw_alpha_ternary": [_MP, _vp, _vp, _i64, _dbl, _vp, _vp],
}

_lib = None
_lib_lock = threading.Lock()
_ctxs = {}


def load():
    """dlopen the library and declare the prototypes (no GPU needed for this step)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise XColumnsB200Error(
                f"{LIB_PATH} not found: the CUDA extension is not built (run "
                f"`python __graft_entry__.py`). xcolumns_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        lib.xc_abi_version.restype = C.c_int
        lib.xc_strerror.restype = C.c_char_p
        lib.xc_strerror.argtypes = [C.c_int]
        lib.xc_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        lib.xc_ctx_destroy.argtypes = [C.c_void_p]
        lib.xc_ctx_destroy.restype = None
        lib.xc_last_cuda_error.argtypes = [C.c_void_p]
        lib.xc_last_cuda_error.restype = C.c_char_p
        lib.xc_launch_count.argtypes = [C.c_void_p]
        lib.xc_launch_count.restype = C.c_int64
        lib.xc_sm_count.argtypes = [C.c_void_p]
        lib.xc_fw_alpha_scratch_bytes.argtypes = [C.c_int64, C.c_int64]
        lib.xc_fw_alpha_scratch_bytes.restype = C.c_int64
        lib.xc_fw_alpha_ctl_offset.argtypes = [C.c_int64, C.c_int64]
        lib.xc_fw_alpha_ctl_offset.restype = C.c_int64
        lib.xc_p2p_payload.argtypes = [C.c_void_p]
        lib.xc_p2p_payload.restype = C.c_void_p
        lib.xc_p2p_destroy.argtypes = [C.c_void_p, C.c_void_p]
        lib.xc_p2p_destroy.restype = None
        lib.xc_bca_delta_stride.argtypes = [C.c_int64]
        lib.xc_bca_delta_stride.restype = C.c_int64
        lib.xc_bca_pipe_buffers.argtypes = [C.c_int]
        lib.xc_bca_pipe_buffers.restype = C.c_int
        lib.xc_bca_window_bytes.argtypes = [C.c_int64, C.c_int, C.c_int]
        lib.xc_bca_window_bytes.restype = C.c_int64
        lib.xc_bca_coef_len.argtypes = [C.c_int64]
        lib.xc_bca_coef_len.restype = C.c_int64
        lib.xc_fill_pred_dense_host.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                                C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.xc_fill_pred_dense_host.restype = C.c_int
        lib.xc_scatter_pred_dense_host.argtypes = lib.xc_fill_pred_dense_host.argtypes
        lib.xc_scatter_pred_dense_host.restype = C.c_int
        lib.xc_zero_host.argtypes = [C.c_void_p, C.c_int64, C.c_int]
        lib.xc_zero_host.restype = C.c_int
        for name, args in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = [C.c_void_p] + args
            fn.restype = C.c_int
        _lib = lib
        return lib


class Context:
    """One xc_ctx per CUDA device (per process)."""

    def __init__(self, device_index: int):
        lib = load()
        h = C.c_void_p()
        rc = lib.xc_ctx_create(device_index, C.byref(h))
        if rc != 0:
            raise XColumnsB200Error(
                f"xc_ctx_create(device={device_index}) failed: {lib.xc_strerror(rc).decode()} -- "
                f"a CUDA device is required (no CPU fallback)")
        self.handle = h
        self.lib = lib
        self.device_index = device_index
        self.sm_count = lib.xc_sm_count(h)

    def launches(self) -> int:
        return int(self.lib.xc_launch_count(self.handle))

    def call(self, name: str, *args):
        rc = getattr(self.lib, name)(self.handle, *args)
        if rc != 0:
            detail = self.lib.xc_strerror(rc).decode()
            if rc == -3:
                detail += ": " + self.lib.xc_last_cuda_error(self.handle).decode()
            raise XColumnsB200Error(f"{name} failed ({rc}): {detail}")


def context(device_index: int) -> Context:
    ctx = _ctxs.get(device_index)
    if ctx is None:
        ctx = Context(device_index)
        _ctxs[device_index] = ctx
    return ctx
