"""ctypes binding of libcunumeric_b200.so (include/cunumeric_b200.h).

There is no CPU fallback: if the shared object is missing, or no CUDA device is visible, the
first compute call raises."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcunumeric_b200.so")
MAX_DIM = 4
COMM_ID_BYTES = 128


class cnb_store_t(ctypes.Structure):
    _fields_ = [
        ("ptr", ctypes.c_void_p),
        ("dtype", ctypes.c_int32),
        ("ndim", ctypes.c_int32),
        ("shape", ctypes.c_int64 * MAX_DIM),
        ("strides", ctypes.c_int64 * MAX_DIM),
    ]


class cnb_trace_record_t(ctypes.Structure):
    _fields_ = [
        ("task", ctypes.c_int32),
        ("op", ctypes.c_int32),
        ("dtype", ctypes.c_int32),
        ("kernel_kind", ctypes.c_int32),
        ("elems", ctypes.c_int64),
        ("bytes", ctypes.c_int64),
        ("ms", ctypes.c_float),
    ]


class cnb_tma_operand_t(ctypes.Structure):
    _fields_ = [
        ("base", ctypes.c_void_p),
        ("width", ctypes.c_int64),
        ("height", ctypes.c_int64),
        ("pitch_bytes", ctypes.c_int64),
        ("elem_bytes", ctypes.c_int32),
        ("box_width", ctypes.c_int32),
        ("box_height", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


class CnbError(RuntimeError):
    def __init__(self, code: int, message: str) -> None:
        super().__init__(f"[cunumeric_b200 rc={code}] {message}")
        self.code = code


_lib = None

# name -> (restype, argtypes)
_vp, _i32, _i64, _sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t
_store_p = ctypes.POINTER(cnb_store_t)
PROTOTYPES = {
    "cnb_binary_op": (_i32, [_i32, _store_p, _store_p, _store_p, _vp, _vp]),
    "cnb_unary_op": (_i32, [_i32, _store_p, _store_p, _store_p, _vp, _vp]),
    "cnb_where": (_i32, [_store_p, _store_p, _store_p, _store_p, _vp]),
    "cnb_convert": (_i32, [_i32, _store_p, _store_p, _vp]),
    "cnb_scalar_unary_red": (_i32, [_i32, _store_p, _store_p, _store_p, _vp, _vp, _vp, _vp]),
    "cnb_unary_red": (_i32, [_i32, _i32, _store_p, _store_p, _store_p, _i64, _vp]),
    "cnb_binary_red": (_i32, [_i32, _store_p, _store_p, _store_p, _vp, _vp]),
    "cnb_fill": (_i32, [_store_p, _vp, _vp]),
    "cunumeric_perform_registration": (None, []),
    "cunumeric_has_curand": (_i32, []),
    "cunumeric_register_reduction_op": (None, [_i32, _i32]),
    "cnb_registered_argval_elem": (_i32, [_i32]),
    "cnb_device_count": (_i32, []),
    "cnb_init": (_i32, [_i32]),
    "cnb_sm_count": (_i32, []),
    "cnb_malloc": (_vp, [_sz, _vp]),
    "cnb_free": (_i32, [_vp, _vp]),
    "cnb_host_alloc": (_vp, [_sz]),
    "cnb_host_free": (_i32, [_vp]),
    "cnb_memcpy_h2d": (_i32, [_vp, _vp, _sz, _vp]),
    "cnb_memcpy_d2h": (_i32, [_vp, _vp, _sz, _vp]),
    "cnb_memcpy_d2d": (_i32, [_vp, _vp, _sz, _vp]),
    "cnb_memset": (_i32, [_vp, _i32, _sz, _vp]),
    "cnb_stream_create": (_vp, []),
    "cnb_stream_destroy": (_i32, [_vp]),
    "cnb_stream_synchronize": (_i32, [_vp]),
    "cnb_device_synchronize": (_i32, []),
    "cnb_event_create": (_vp, []),
    "cnb_event_destroy": (_i32, [_vp]),
    "cnb_event_record": (_i32, [_vp, _vp]),
    "cnb_event_synchronize": (_i32, [_vp]),
    "cnb_stream_wait_event": (_i32, [_vp, _vp]),
    "cnb_event_elapsed_ms": (_i32, [_vp, _vp, ctypes.POINTER(ctypes.c_float)]),
    "cnb_mem_info": (_i32, [ctypes.POINTER(_sz), ctypes.POINTER(_sz)]),
    "cnb_launch_count": (ctypes.c_uint64, []),
    "cnb_trace_start": (_i32, [_i32]),
    "cnb_trace_stop": (_i32, []),
    "cnb_trace_get": (_i32, [_i32, ctypes.POINTER(cnb_trace_record_t)]),
    "cnb_module_load": (_i32, [_vp, ctypes.c_size_t, ctypes.POINTER(_vp)]),
    "cnb_module_get_kernel": (_i32, [_vp, ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "cnb_launch_fused": (_i32, [_vp, _vp, ctypes.c_size_t, ctypes.c_int64, ctypes.c_int64,
                                ctypes.c_int64, _i32, _i32, _i32, _vp]),
    "cnb_launch_fused_tma": (_i32, [_vp, ctypes.POINTER(cnb_tma_operand_t), _i32, _vp, _sz, _i32, _i64,
                                    _i64, _i64, _i32, _i32, _vp]),
    "cnb_copy_complement": (_i32, [_vp, _vp, _sz, _i64, _i64, _i64, _i64, _vp]),
    "cnb_last_error": (ctypes.c_char_p, []),
    "cnb_version": (ctypes.c_char_p, []),
    "cnb_argval_fold": (_i32, [_i32, _i32, _vp, _vp, _i32, _i64, _vp]),
    "cnb_comm_unique_id": (_i32, [_vp]),
    "cnb_comm_init": (_vp, [_vp, _i32, _i32]),
    "cnb_comm_destroy": (_i32, [_vp]),
    "cnb_comm_group_start": (_i32, []),
    "cnb_comm_group_end": (_i32, []),
    "cnb_comm_send": (_i32, [_vp, _vp, _sz, _i32, _vp]),
    "cnb_comm_recv": (_i32, [_vp, _vp, _sz, _i32, _vp]),
    "cnb_comm_allreduce": (_i32, [_vp, _vp, _vp, _sz, _i32, _i32, _vp]),
    "cnb_comm_allgather": (_i32, [_vp, _vp, _vp, _sz, _vp]),
}


def load():
    """dlopen the CUDA library (built in-tree by __graft_entry__.build / csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `make -C cunumeric_b200/csrc` "
                "(or __graft_entry__.build()). cunumeric_b200 has no CPU fallback.")
        _set_nccl_hint()
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def _set_nccl_hint() -> None:
    # point the dlopen in csrc/comm.cu at the NCCL that ships with torch, if any
    if os.environ.get("CNB_NCCL_LIB"):
        return
    try:
        import importlib.util

        spec = importlib.util.find_spec("nvidia.nccl")
        if spec is not None and spec.submodule_search_locations:
            cand = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["CNB_NCCL_LIB"] = cand
    except Exception:
        pass


def check(rc: int) -> None:
    if rc != 0:
        raise CnbError(rc, load().cnb_last_error().decode())


def check_ptr(p):
    if not p:
        raise CnbError(-3, load().cnb_last_error().decode())
    return p
