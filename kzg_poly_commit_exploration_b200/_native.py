"""ctypes binding of libkzg_b200.so (include/kzg_b200.h).  Fails loudly when the CUDA library is
missing — there is no Python or CPU fallback for the bulk operations."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkzg_b200.so")

KZG_OK = 0
STATUS = {
    0: "OK", 1: "DEGREE_TOO_HIGH", 2: "CONSTANT_POLY", 3: "REMAINDER_NONZERO", 4: "BAD_POINT",
    5: "BAD_LENGTH", 6: "SETUP_TOO_SHORT", 7: "BAD_ARG", 100: "CUDA", 102: "NOMEM",
}
G1_PARTIAL_BYTES = 192

_u8p = C.POINTER(C.c_uint8)
_vp = C.c_void_p
_u64 = C.c_uint64
_i32 = C.c_int32

# name -> (restype, argtypes); every symbol include/kzg_b200.h declares
SIGNATURES = {
    "kzg_ctx_create": (_i32, [C.POINTER(_vp), _i32]),
    "kzg_ctx_create_multi": (_i32, [C.POINTER(_vp), C.POINTER(_i32), _i32]),
    "kzg_ctx_device_count": (_i32, [_vp]),
    "kzg_ctx_destroy": (None, [_vp]),
    "kzg_last_error": (C.c_char_p, [_vp]),
    "kzg_status_string": (C.c_char_p, [_i32]),
    "kzg_host_alloc": (_vp, [C.c_size_t]),
    "kzg_host_free": (None, [_vp]),
    "kzg_ctx_sync": (_i32, [_vp]),
    "kzg_ctx_launch_count": (_u64, [_vp]),
    "kzg_ctx_stream": (_vp, [_vp]),
    "kzg_ctx_last_kernel_ms": (C.c_float, [_vp, _i32]),
    "kzg_ctx_timer_start": (_i32, [_vp]),
    "kzg_ctx_timer_stop": (_i32, [_vp, C.POINTER(C.c_float)]),
    "kzg_ctx_set_option": (_i32, [_vp, C.c_char_p, C.c_int64]),
    "kzg_setup_generate": (_i32, [_vp, _vp, _u64, _u64, C.POINTER(_vp)]),
    "kzg_setup_import_g1": (_i32, [_vp, _vp, _u64, C.POINTER(_vp)]),
    "kzg_setup_export_g1": (_i32, [_vp, _vp, _u64, _u64, _vp]),
    "kzg_setup_len": (_u64, [_vp]),
    "kzg_msm_plan": (_i32, [_vp, _vp, _u64, C.POINTER(_i32)]),
    "kzg_setup_generate_g2": (_i32, [_vp, _vp, _vp, _u64]),
    "kzg_setup_import_g2": (_i32, [_vp, _vp, _vp, _u64]),
    "kzg_setup_export_g2": (_i32, [_vp, _vp, _u64, _u64, _vp]),
    "kzg_setup_g2_len": (_u64, [_vp]),
    "kzg_setup_destroy": (None, [_vp]),
    "kzg_setup_prepare_msm": (_i32, [_vp, _vp, _i32]),
    "kzg_setup_prepare_fixed": (_i32, [_vp, _vp, _u64, _i32]),
    "kzg_setup_fixed_plan": (_i32, [_vp, C.POINTER(C.c_int64)]),
    "kzg_commit": (_i32, [_vp, _vp, _vp, _u64, _vp]),
    "kzg_evaluate": (_i32, [_vp, _vp, _u64, _vp, _vp]),
    "kzg_quotient": (_i32, [_vp, _vp, _u64, _vp, _vp, _vp, C.POINTER(_u64)]),
    "kzg_open": (_i32, [_vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp]),
    "kzg_msm_g1": (_i32, [_vp, _vp, _vp, _u64, _vp]),
    "kzg_commit_batch": (_i32, [_vp, _vp, _vp, _u64, _u64, _vp]),
    "kzg_open_batch": (_i32, [_vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp]),
    "kzg_commit_open_batch_dev": (_i32, [_vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp]),
    "kzg_commit_open_batch": (_i32, [_vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp]),
    "kzg_msm_g1_dev": (_i32, [_vp, _vp, _vp, _u64, _vp]),
    "kzg_msm_g1_submit": (_i32, [_vp, _vp, _vp, _u64, C.POINTER(_i32)]),
    "kzg_msm_g1_wait": (_i32, [_vp, _i32, _vp]),
    "kzg_msm_g1_submit_dev": (_i32, [_vp, _vp, _vp, _u64, C.POINTER(_i32)]),
    "kzg_msm_g1_sharded_submit_dev": (_i32, [_vp, _vp, _vp, _u64, C.POINTER(_i32)]),
    "kzg_msm_g1_sharded_submit": (_i32, [_vp, _vp, _vp, _u64, C.POINTER(_i32)]),
    "kzg_msm_g1_partial_dev": (_i32, [_vp, _vp, _vp, _u64, _vp]),
    "kzg_g1_combine_dev": (_i32, [_vp, _vp, _u64, _vp]),
    "kzg_p2p_init": (_i32, [_vp, _i32, _i32, _vp]),
    "kzg_p2p_connect": (_i32, [_vp, _vp]),
    "kzg_msm_g1_sharded_dev": (_i32, [_vp, _vp, _vp, _u64, _vp]),
    "kzg_commit_batch_dev": (_i32, [_vp, _vp, _vp, _u64, _u64, _vp]),
    "kzg_open_batch_dev": (_i32, [_vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp]),
    "kzg_evaluate_dev": (_i32, [_vp, _vp, _u64, _vp, _vp]),
    "kzg_fr_mul": (None, [_vp, _vp, _vp]),
    "kzg_fr_add": (None, [_vp, _vp, _vp]),
    "kzg_fr_sub": (None, [_vp, _vp, _vp]),
    "kzg_fr_neg": (None, [_vp, _vp]),
    "kzg_fr_pow": (None, [_vp, _u64, _vp]),
    "kzg_fr_reduce": (None, [_vp, _vp]),
    "kzg_g1_add": (_i32, [_vp, _vp, _vp]),
    "kzg_g1_sub": (_i32, [_vp, _vp, _vp]),
    "kzg_g1_mul": (_i32, [_vp, _vp, _vp]),
    "kzg_g1_from_scalar": (None, [_vp, _vp]),
    "kzg_g1_check": (_i32, [_vp]),
    "kzg_g2_from_scalar": (None, [_vp, _vp]),
    "kzg_g2_add": (_i32, [_vp, _vp, _vp]),
    "kzg_g2_sub": (_i32, [_vp, _vp, _vp]),
    "kzg_g2_neg": (_i32, [_vp, _vp]),
    "kzg_g2_check": (_i32, [_vp]),
    "kzg_pairing_equal": (_i32, [_vp, _vp, _vp, _vp, C.POINTER(_i32)]),
    "kzg_verify_proof": (_i32, [_vp, _vp, _vp, _vp, _vp, C.POINTER(_i32)]),
    "kzg_probe_imad_peak": (_i32, [_vp, C.POINTER(C.c_double)]),
    "kzg_probe_fq_mul": (_i32, [_vp, C.POINTER(C.c_double)]),
    "kzg_selftest_fq_mul": (_i32, [_vp, _vp, _vp, _u64, _vp]),
}

_lib = None


class KzgNativeError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"[{STATUS.get(status, status)}] {msg}")
        self.status = status
        self.kind = STATUS.get(status, str(status))


def lib() -> C.CDLL:
    """Load the CUDA extension; raise (never fall back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m kzg_poly_commit_exploration_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def buf(data: bytes):
    return (C.c_uint8 * len(data)).from_buffer_copy(data)


def out(n: int):
    return (C.c_uint8 * n)()
