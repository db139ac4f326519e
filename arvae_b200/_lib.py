"""ctypes binding of ``libarvae_b200.so`` (the C ABI in include/arvae_b200.h).

There is no fallback: if the library is missing or a call fails, a
``RuntimeError`` is raised -- the product path never routes around the CUDA
kernels.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ARVAE_LIB_PATH") or os.path.join(_HERE, "csrc", "libarvae_b200.so")  # override: A/B experiments

ALGO_AUTO, ALGO_DENSE, ALGO_SORTED, ALGO_TRIANGLE = 0, 1, 2, 3
MAX_REG_DIMS = 32

_c_i32p = ctypes.POINTER(ctypes.c_int32)
_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_i32 = ctypes.c_int32
_f = ctypes.c_float
_sz = ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/arvae_b200.h declares
SIGNATURES = {
    "arvae_version": (ctypes.c_int, []),
    "arvae_last_error": (ctypes.c_char_p, []),
    "arvae_device_sm_count": (ctypes.c_int, []),
    "arvae_reg_loss_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "arvae_reg_loss_workspace_bytes_algo": (_sz, [_i64, _i64, _i32, _i32]),
    "arvae_reg_loss_fwdbwd_f32": (ctypes.c_int, [_vp, _i64, _i64, _vp, _i64, _i64, _c_i32p, _c_i32p, _i32,
                                                 _i64, _i64, _i64, _f, _f, _i32, _vp, _vp, _vp, _vp, _vp, _vp,
                                                 _sz, _vp]),
    "arvae_reg_loss_path_flags": (ctypes.c_int, [_i64, _i64, _i32, _i32, _vp, _c_i32p, _vp]),
    "arvae_reg_loss_scatter_bwd_f32": (ctypes.c_int, [_vp, _vp, _c_i32p, _i32, _i64, _i64, _vp, _i64, _vp]),
    "arvae_latent_head_workspace_bytes": (_sz, [_i64, _i64]),
    "arvae_latent_head_fwd_f32": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i64, _f, _f, _vp, _vp, _vp, _vp, _vp,
                                                 _vp, _sz, _vp]),
    "arvae_latent_head_bwd_f32": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _c_i32p, _i32, _f, _vp, _vp,
                                                 _i64, _i64, _vp, _vp, _vp]),
    "arvae_head_fused_workspace_bytes": (_sz, [_i64, _i32]),
    "arvae_head_fused_fwd_f32": (ctypes.c_int, [_vp, _vp, _i32, _vp, _i64, _i64, _vp, _i64, _i64, _c_i32p, _c_i32p, _i32,
                                                _f, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "arvae_head_fused_bwd_f32": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _c_i32p, _i32, _vp, _vp, _i64, _i64,
                                                _vp, _vp, _vp]),
    "arvae_reg_loss_host_f32": (ctypes.c_int, [_vp, _i64, _i64, _vp, _i64, _c_i32p, _c_i32p, _i32, _f, _f,
                                               _i32, _vp, _vp, _vp]),
    "arvae_host_release": (None, []),
    "arvae_pack_columns_f32": (ctypes.c_int, [_vp, _i64, _i64, _vp, _i64, _i64, _c_i32p, _c_i32p, _i32, _i64, _vp, _vp]),
    "arvae_attr_argsort_workspace_bytes": (_sz, [_i64]),
    "arvae_attr_argsort_f32": (ctypes.c_int, [_vp, _i64, _i64, _vp, _vp, _sz, _vp]),
    "arvae_measure_attributes_i64": (ctypes.c_int, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp]),
    "arvae_eval_metrics_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "arvae_eval_metrics_f32": (ctypes.c_int, [_vp, _i64, _i64, _vp, _i64, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp,
                                              _vp, _vp, _sz, _vp]),
    "arvae_reg_sign_matrix_i8": (ctypes.c_int, [_vp, _i64, _i64, _vp, _vp]),
    "arvae_shard_comm_bytes": (_sz, [_i64, _i32, _i32]),
    "arvae_shard_create": (ctypes.c_int, [_i32, _i32, _i64, _i32, ctypes.POINTER(_vp)]),
    "arvae_shard_ipc_handle": (ctypes.c_int, [_vp, _vp]),
    "arvae_shard_open_peers": (ctypes.c_int, [_vp, _vp]),
    "arvae_shard_set_peer": (ctypes.c_int, [_vp, _i32, _vp]),
    "arvae_shard_comm_ptr": (_vp, [_vp]),
    "arvae_shard_reg_loss_f32": (ctypes.c_int, [_vp, _vp, _i64, _i64, _vp, _i64, _i64, _c_i32p, _c_i32p, _i32,
                                                ctypes.POINTER(_i64), _f, _f, _vp, _vp, _vp, _i32, _vp]),
    "arvae_shard_reg_loss_host_f32": (ctypes.c_int, [_vp, _vp, _i64, _vp, _i64, _c_i32p, _c_i32p, _i32,
                                                     ctypes.POINTER(_i64), _f, _f, _vp, _vp, _vp]),
    "arvae_shard_status": (ctypes.c_int, [_vp, ctypes.POINTER(_i32), ctypes.POINTER(ctypes.c_uint64), _vp]),
    "arvae_shard_destroy": (ctypes.c_int, [_vp]),
    "arvae_timeline_enable": (None, [ctypes.c_int]),
    "arvae_timeline_report": (ctypes.c_int, [ctypes.c_char_p, _i32]),
    "arvae_launch_count": (_i64, [ctypes.c_int]),
    "arvae_profile_enable": (None, [ctypes.c_int]),
    "arvae_profile_pair_kernel_ms": (ctypes.c_int, [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)]),
}

_lib: Optional[ctypes.CDLL] = None


def load() -> ctypes.CDLL:
    """Load the shared library (built in-tree by ``arvae_b200.build``); raise loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"arvae_b200: {LIB_PATH} is missing. Build it with `python -m arvae_b200.build` "
                "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        if lib.arvae_version() != 200:
            raise RuntimeError("arvae_b200: libarvae_b200.so version mismatch; rebuild")
        _lib = lib
    return _lib


def last_error() -> str:
    return load().arvae_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"arvae_b200: {what} failed (rc={rc}): {last_error()}")


def i64_array(values):
    return (ctypes.c_int64 * max(len(values), 1))(*[int(v) for v in values])


def i32_array(values):
    arr = (ctypes.c_int32 * max(len(values), 1))(*[int(v) for v in values])
    return arr
