"""ctypes binding of liblist_b200.so (include/list_b200.h).

There is no fallback: if the library is missing, or a call fails, a RuntimeError is raised.
Only plain pointers/sizes cross the boundary; torch is used by the callers for device memory
and streams, never inside the library.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LIST_B200_LIB") or os.path.join(HERE, "liblist_b200.so")   # override: A/B builds only

ABI_VERSION = 3
MAX_LEVELS = 8
MAX_MAPS = 8
NUM_DISP = 7
F32, BF16 = 0, 1
OK, EINVAL, ENOMEM, ECUDA, ENOSYS = 0, -22, -12, -5, -38


class ListCtx(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("dtype", C.c_int32), ("map_size", C.c_int32), ("map_channels", C.c_int32),
        ("maps", C.c_void_p),
        ("n_levels", C.c_int32), ("reserved0", C.c_int32),
        ("vol_res", C.c_int32 * MAX_LEVELS), ("vol_ch", C.c_int32 * MAX_LEVELS),
        ("vols", C.c_void_p * MAX_LEVELS),
        ("trans_mat", C.c_void_p),
    ]


class ListLayout(C.Structure):
    _fields_ = [
        ("k_out", C.c_int32), ("k_pad", C.c_int32), ("map_off", C.c_int32), ("xyz_off", C.c_int32),
        ("vol_off", C.c_int32 * MAX_LEVELS),
    ]


class ListWeights(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("k_pad", C.c_int32), ("n0", C.c_int32), ("n1", C.c_int32), ("n2", C.c_int32),
        ("reserved0", C.c_int32),
        ("w0", C.c_void_p), ("w1", C.c_void_p), ("w2", C.c_void_p), ("w3", C.c_void_p),
        ("b0", C.c_void_p), ("b1", C.c_void_p), ("b2", C.c_void_p), ("b3", C.c_void_p),
    ]


class ListGrads(C.Structure):
    _fields_ = [
        ("d_maps", C.c_void_p), ("d_vols", C.c_void_p * MAX_LEVELS), ("d_trans_mat", C.c_void_p),
        ("d_w0", C.c_void_p), ("d_w1", C.c_void_p), ("d_w2", C.c_void_p), ("d_w3", C.c_void_p),
        ("d_b0", C.c_void_p), ("d_b1", C.c_void_p), ("d_b2", C.c_void_p), ("d_b3", C.c_void_p),
    ]


_i32, _i64, _f32, _f64, _vp, _sz = C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_void_p, C.c_size_t
_P = C.POINTER

# name -> (restype, argtypes); every symbol include/list_b200.h declares
SIGNATURES = {
    "list_b200_abi_version": (C.c_int, []),
    "list_b200_last_error": (C.c_char_p, []),
    "list_b200_device_ok": (C.c_int, []),
    "list_feature_layout": (C.c_int, [_i32, _i32, _P(_i32), _P(ListLayout), _P(_i32)]),
    "list_prep_maps": (C.c_int, [_P(_vp), _P(_i32), _P(_i32), _i32, _i32, _i32, _vp, _i32, _vp]),
    "list_prep_volume": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _i32, _vp]),
    "list_prep_maps_bwd": (C.c_int, [_vp, _P(_i32), _P(_i32), _i32, _i32, _i32, _P(_vp), _vp]),
    "list_prep_volume_bwd": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "list_grid_points": (C.c_int, [_vp, _i32, _f64, _f64, _i64, _i64, _vp]),
    "list_gather_fwd": (C.c_int, [_P(ListCtx), _vp, _i32, _vp, _i64, _i32, _i64, _vp]),
    "list_gather_grid_fwd": (C.c_int, [_P(ListCtx), _i32, _i32, _f64, _f64, _i64, _i64, _vp, _i64, _vp]),
    "list_mlp_workspace_bytes": (_sz, [_P(ListWeights), _i64]),
    "list_mlp_fwd": (C.c_int, [_P(ListWeights), _vp, _i64, _i64, _vp, _f32, _vp, _sz, _vp]),
    "list_mlp_fwd_train": (C.c_int, [_P(ListWeights), _vp, _i64, _i64, _vp, _f32, _vp, _sz, _vp]),
    "list_mlp_fwd_debug": (C.c_int, [_P(ListWeights), _vp, _i64, _i64, _vp, _f32, _vp, _vp, _vp, _vp]),
    "list_hoist_bytes": (_sz, [_P(ListCtx), _P(ListWeights)]),
    "list_hoist_layout": (C.c_int, [_P(ListCtx), _P(ListWeights), _P(_i32), _P(_i32)]),
    "list_hoist_prepare": (C.c_int, [_P(ListCtx), _P(ListWeights), _vp, _sz, _vp]),
    "list_mlp_hoisted_fwd": (C.c_int, [_P(ListWeights), _i32, _vp, _i64, _i64, _vp, _f32, _vp]),
    "list_mlp_hoisted_trace": (C.c_int, [_P(ListWeights), _i32, _vp, _i64, _i64, _vp, _f32, _vp, _vp]),
    "list_hoist_gather_grid_fwd": (C.c_int, [_P(ListCtx), _P(ListWeights), _vp, _i32, _i32, _f64, _f64, _i64, _i64, _vp,
                                             _i64, _i32, _vp]),
    "list_lines_layout": (C.c_int, [_P(ListCtx), _P(ListWeights), _P(_i32), _P(_i32), _P(_i32)]),
    "list_lines_hoist_bytes": (_sz, [_P(ListCtx), _P(ListWeights)]),
    "list_lines_prepare": (C.c_int, [_P(ListCtx), _P(ListWeights), _vp, _sz, _vp]),
    "list_lines_table_bytes": (_sz, [_P(ListCtx), _P(ListWeights), _i32, _i64, _i64]),
    "list_lines_table": (C.c_int, [_P(ListCtx), _P(ListWeights), _vp, _i32, _i32, _f64, _f64, _i64, _i64, _vp, _sz, _vp]),
    "list_lines_rest": (C.c_int, [_P(ListCtx), _P(ListWeights), _i32, _i32, _f64, _f64, _i64, _i64, _vp, _i64, _vp]),
    "list_grid_plan_bytes": (_sz, [_i32, _i64, _i64]),
    "list_grid_plan": (C.c_int, [_P(ListCtx), _P(ListWeights), _vp, _i32, _i32, _f64, _f64, _i64, _i64, _vp, _vp, _sz, _vp]),
    "list_grid_tc_fwd": (C.c_int, [_P(ListCtx), _P(ListWeights), _vp, _i32, _f64, _f64, _i64, _i64, _vp, _i64, _vp, _vp, _f32, _vp,
                                   _vp, _vp, _vp]),
    "list_sdf_workspace_bytes": (_sz, [_P(ListCtx), _P(ListWeights), _i64]),
    "list_sdf_grid_workspace_bytes": (_sz, [_P(ListCtx), _P(ListWeights), _i32, _i64]),
    "list_sdf_fwd": (C.c_int, [_P(ListCtx), _P(ListWeights), _vp, _i32, _i32, _i64, _vp, _f32, _i64, _vp, _sz, _vp]),
    "list_sdf_grid": (C.c_int, [_P(ListCtx), _P(ListWeights), _i32, _f64, _f64, _i64, _i64, _vp, _f32, _i64, _vp, _sz, _vp]),
    "list_sdf_grid_late": (C.c_int, [_P(ListCtx), _P(ListWeights), _i32, _f64, _f64, _i64, _i64, _vp, _f32, _i64, _vp, _sz, _vp,
                                     _vp, _P(_vp), _vp]),
    "list_sdf_grid_host_bytes": (_sz, [_P(_i32), _P(_i32), _i32, _i32, _i32, _P(_i32), _P(_i32), _i32, _i32, _i64, _i64]),
    "list_sdf_grid_host": (C.c_int, [_P(_vp), _P(_i32), _P(_i32), _i32, _i32, _P(_vp), _i32, _P(_i32), _P(_i32), _vp,
                                     _i32, _i32, _P(ListWeights), _i32, _f64, _f64, _i64, _i64, _f32, _i64, _vp, _vp,
                                     _sz, _vp]),
    "list_mc_workspace_bytes": (_sz, [_i32]),
    "list_mc_count": (C.c_int, [_vp, _i32, _f32, _i32, _vp, _sz, _vp, _vp]),
    "list_mc_generate": (C.c_int, [_vp, _i32, _f32, _i32, _vp, _sz, _vp, _i64, _vp, _i64, _vp]),
    "list_gemm_f32_tc": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _sz, _vp]),
    "list_bwd_workspace_bytes": (_sz, [_P(ListWeights), _i64]),
    "list_sdf_bwd": (C.c_int, [_P(ListCtx), _P(ListWeights), _vp, _i32, _i32, _i64, _vp, _i64, _vp, _vp, _P(ListGrads),
                               _vp, _sz, _vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Loads the library once; raises if it has not been built (python -m list_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m list_b200.build` "
                "(list_b200 has no CPU or PyTorch fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        got = handle.list_b200_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError(f"liblist_b200.so ABI {got} != binding ABI {ABI_VERSION}; rebuild")
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().list_b200_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != OK:
        raise RuntimeError(f"{what} failed ({rc}): {last_error()}")


def i32_array(values, n=None):
    n = n or len(values)
    arr = (C.c_int32 * n)()
    for i, v in enumerate(values):
        arr[i] = int(v)
    return arr


def ptr_array(values):
    arr = (C.c_void_p * len(values))()
    for i, v in enumerate(values):
        arr[i] = int(v)
    return arr
