"""ctypes binding of libmgb200.so (the C ABI declared in include/mgb.h).

There is no CPU fallback: if the shared library is missing, importing this
module raises, and every call that needs a GPU raises ``RuntimeError`` with the
library's error text when the device is absent.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

__all__ = ["lib", "check", "ConvDesc", "PoolDesc", "F32", "F64", "LIB_PATH", "EXPORTS"]

LIB_PATH = Path(os.environ.get("MGB200_LIB", Path(__file__).resolve().parent / "libmgb200.so"))

F32, F64 = 0, 1
MAX_SPATIAL = 3
MAX_DIMS = 8


class ConvDesc(C.Structure):
    """mirror of ``mgb_conv_desc``"""

    _fields_ = [
        ("dtype", C.c_int32),
        ("nd", C.c_int32),
        ("N", C.c_int64),
        ("C", C.c_int64),
        ("F", C.c_int64),
        ("X", C.c_int64 * MAX_SPATIAL),
        ("W", C.c_int64 * MAX_SPATIAL),
        ("stride", C.c_int64 * MAX_SPATIAL),
        ("pad", C.c_int64 * MAX_SPATIAL),
        ("dil", C.c_int64 * MAX_SPATIAL),
    ]


class PoolDesc(C.Structure):
    """mirror of ``mgb_pool_desc``"""

    _fields_ = [
        ("dtype", C.c_int32),
        ("nd", C.c_int32),
        ("mode", C.c_int32),
        ("reserved", C.c_int32),
        ("lead", C.c_int64),
        ("X", C.c_int64 * MAX_SPATIAL),
        ("P", C.c_int64 * MAX_SPATIAL),
        ("stride", C.c_int64 * MAX_SPATIAL),
    ]


_vp, _sz, _i, _i64p = C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_int64)
_cdp, _pdp = C.POINTER(ConvDesc), C.POINTER(PoolDesc)

# name -> (restype, argtypes); one entry per declaration in include/mgb.h
EXPORTS = {
    "mgb_conv_workspace_bytes": (_i, [_cdp, _i, C.POINTER(_sz)]),
    "mgb_conv_fwd": (_i, [_cdp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mgb_conv_dgrad": (_i, [_cdp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mgb_conv_wgrad": (_i, [_cdp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mgb_conv_planes_bytes": (_i, [_cdp, _i, C.POINTER(_sz)]),
    "mgb_conv_make_planes": (_i, [_cdp, _i, _vp, _vp, _vp]),
    "mgb_conv_fwd_p": (_i, [_cdp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mgb_conv_dgrad_p": (_i, [_cdp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mgb_conv_wgrad_p": (_i, [_cdp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mgb_conv_pool_supported": (_i, [_cdp, _pdp, C.POINTER(_i)]),
    "mgb_conv_pool_fwd": (_i, [_cdp, _pdp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mgb_pool_bwd_planes": (_i, [_cdp, _pdp, _vp, _vp, _vp, _vp]),
    "mgb_pool_bwd_dy_planes": (_i, [_cdp, _pdp, _vp, _vp, _vp, _vp, _vp]),
    "mgb_pool_fwd": (_i, [_pdp, _vp, _vp, _vp]),
    "mgb_pool_bwd": (_i, [_pdp, _vp, _vp, _vp, _vp]),
    "mgb_reduce_broadcast": (_i, [_i, _i, _i64p, _i, _i64p, _vp, _vp, _i, _vp]),
    "mgb_axpy": (_i, [_i, _sz, _vp, _vp, _vp]),
    "mgb_copy_cast": (_i, [_i, _i, _sz, _vp, _vp, _vp]),
    "mgb_fill": (_i, [_i, _sz, C.c_double, _vp, _vp]),
    "mgb_scale": (_i, [_i, _sz, C.c_double, _vp, _vp]),
    "mgb_relu_fwd": (_i, [_i, _sz, _vp, _vp, _vp]),
    "mgb_relu_bwd": (_i, [_i, _sz, _vp, _vp, _vp, _vp]),
    "mgb_add_broadcast": (_i, [_i, _i, _i64p, _i64p, _i64p, _vp, _vp, _vp, _vp]),
    "mgb_transpose2d": (_i, [_i, C.c_int64, C.c_int64, _vp, _vp, _vp]),
    "mgb_softmax_xent": (_i, [_i, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp, _vp]),
    "mgb_scale_by": (_i, [_i, _sz, _vp, _vp, _vp, _vp]),
    "mgb_matmul_skinny": (_i, [_i, _i, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp]),
    "mgb_einsum": (_i, [_i, _i, C.POINTER(_vp), _i, _i64p, _i64p, _i64p, _i, _i64p, _i64p, _vp, _vp]),
    "mgb_window_gather": (_i, [_i, _i, C.c_int64, _i64p, _i64p, _i64p, _i64p, _vp, _vp, _vp]),
    "mgb_batchnorm_workspace_bytes": (_i, [C.c_int64, C.c_int64, C.POINTER(_sz)]),
    "mgb_batchnorm_stats": (_i, [_i, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, _vp, _sz, _vp]),
    "mgb_batchnorm_fwd": (_i, [_i, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp, C.c_double, _vp, _vp]),
    "mgb_batchnorm_bwd_reduce": (_i, [_i, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, _vp, C.c_double, _vp, _vp,
                                      _sz, _vp]),
    "mgb_batchnorm_bwd_dx": (_i, [_i, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp, _vp, C.c_double, _vp,
                                  _vp]),
    "mgb_batchnorm_merge": (_i, [_i, C.c_int64, _vp, _vp, _vp]),
    "mgb_device_count": (_i, [C.POINTER(_i)]),
    "mgb_set_device": (_i, [_i]),
    "mgb_get_device": (_i, [C.POINTER(_i)]),
    "mgb_device_info": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_sz)]),
    "mgb_malloc_async": (_i, [C.POINTER(_vp), _sz, _vp]),
    "mgb_free_async": (_i, [_vp, _vp]),
    "mgb_host_alloc": (_i, [C.POINTER(_vp), _sz]),
    "mgb_host_free": (_i, [_vp]),
    "mgb_memcpy_h2d": (_i, [_vp, _vp, _sz, _vp]),
    "mgb_memcpy_d2h": (_i, [_vp, _vp, _sz, _vp]),
    "mgb_memcpy_d2d": (_i, [_vp, _vp, _sz, _vp]),
    "mgb_memset": (_i, [_vp, _i, _sz, _vp]),
    "mgb_stream_create": (_i, [C.POINTER(_vp)]),
    "mgb_stream_destroy": (_i, [_vp]),
    "mgb_stream_synchronize": (_i, [_vp]),
    "mgb_device_synchronize": (_i, []),
    "mgb_event_create": (_i, [C.POINTER(_vp)]),
    "mgb_event_destroy": (_i, [_vp]),
    "mgb_event_record": (_i, [_vp, _vp]),
    "mgb_event_synchronize": (_i, [_vp]),
    "mgb_event_elapsed_ms": (_i, [_vp, _vp, C.POINTER(C.c_float)]),
    "mgb_stream_wait_event": (_i, [_vp, _vp]),
    "mgb_graph_begin": (_i, [_vp]),
    "mgb_graph_end": (_i, [_vp, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]),
    "mgb_graph_launch": (_i, [_vp, _vp]),
    "mgb_graph_destroy": (_i, [_vp]),
    "mgb_launch_count": (_i, [C.POINTER(C.c_uint64)]),
    "mgb_last_error": (C.c_char_p, []),
    "mgb_version": (C.c_char_p, []),
    "mgb_comm_unique_id": (_i, [_vp]),
    "mgb_comm_init_rank": (_i, [_vp, _i, _i]),
    "mgb_comm_destroy": (_i, []),
    "mgb_allreduce_sum": (_i, [_vp, _sz, _i, _vp]),
    "mgb_peer_create": (_i, [_i, _i, _sz, _vp]),
    "mgb_peer_connect": (_i, [_vp]),
    "mgb_peer_destroy": (_i, []),
    "mgb_peer_slot_bytes": (_i, [C.POINTER(_sz)]),
    "mgb_peer_status": (_i, [C.POINTER(_i), C.POINTER(_i)]),
    "mgb_peer_allreduce_sum": (_i, [_vp, _sz, _i, _vp]),
    "mgb_conv_wgrad_dp": (_i, [_cdp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mgb_probe_mma_peak": (_i, [_i, _i, C.POINTER(C.c_double), _vp]),
}


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python mygrad_b200/csrc/build.py`). mygrad_b200 has no CPU fallback."
        )
    handle = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(handle, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return handle


lib = _load()


def check(rc: int) -> None:
    """maps a non-zero status to RuntimeError carrying mgb_last_error()"""
    if rc != 0:
        raise RuntimeError(lib.mgb_last_error().decode("utf-8", "replace"))
