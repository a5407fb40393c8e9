"""ctypes binding of libhashgan_b200.so (include/hashgan_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is NO CPU
fallback: if the shared object is missing or fails to load, every product entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libhashgan_b200.so")

HG_OK, HG_EINVAL, HG_ERANGE, HG_ENOMEM, HG_ECUDA, HG_ELABEL = 0, 1, 2, 3, 4, 5
FLAG_FORCE_EXACT = 1
FLAG_NO_FALLBACK = 2
FLAG_TIMING = 4
ENC_LRN = 1
ENC_CONV_TF32 = 2
ENC_CONV_TF32X3 = 8
ENC_TIMING = 16
ENC_FUSED_STAGE1 = 32
MAX_BITS = 256

_i64, _int, _u32, _vp, _sz = C.c_int64, C.c_int, C.c_uint, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); every symbol include/hashgan_b200.h declares
SIGNATURES = {
    "hg_version": (_int, []),
    "hg_crc32c": (C.c_uint32, [_vp, _sz, C.c_uint32]),
    "hg_last_error": (C.c_char_p, []),
    "hg_device_info": (_int, [C.POINTER(_int), C.POINTER(_int), C.POINTER(_int), C.POINTER(_sz)]),
    "hg_code_words": (_int, [_int]),
    "hg_label_words": (_int, [_int]),
    "hg_row_words": (_int, [_int, _int]),
    "hg_pack_rows": (_int, [_vp, _i64, _vp, _int, _i64, _int, _int, _vp, _vp, _vp]),
    "hg_pack_rows_push": (_int, [_vp, _i64, _vp, _int, _i64, _int, _int, C.POINTER(_vp), _int, _vp, _vp]),
    "hg_hamming_map_workspace_bytes": (_sz, [_i64, _i64, _int, _int, _i64]),
    "hg_hamming_map": (_int, [_vp, _i64, _vp, _i64, _int, _int, _i64, _u32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hg_hamming_map_stats": (_int, [_vp, _sz, _i64, _i64, _int, _int, _i64, C.POINTER(_i64), _vp]),
    "hg_hamming_map_phase_ms": (_int, [C.POINTER(C.c_float)]),
    "hg_select_backend": (_int, [_int, _int]),
    "hg_select_backend_for": (_int, [_i64, _i64, _int, _int, _i64]),
    "hg_select_queued_for": (_int, [_i64, _i64, _int, _int, _i64]),
    "hg_ip_map_workspace_bytes": (_sz, [_i64, _i64, _int, _int, _i64]),
    "hg_ip_map": (_int, [_vp, _vp, _i64, _vp, _vp, _i64, _int, _int, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hg_relevant_totals": (_int, [_vp, _i64, _vp, _i64, _int, _int, _vp, _vp]),
    "hg_launch_count": (_i64, [_int]),
    "hg_mean_ap": (_int, [_vp, _i64, C.POINTER(C.c_double), C.POINTER(_i64), _vp]),
    "hg_mean_ap_host": (_int, [_vp, _i64, C.POINTER(C.c_double), C.POINTER(_i64)]),
    "hg_maps_by_feature_host": (_int, [_vp, _vp, _i64, _vp, _vp, _i64, _int, _int, _int, _i64, _u32,
                                       C.POINTER(C.c_double), _vp]),
    "hg_release_cached": (_int, []),
    "hg_gemm_tf32": (_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _int, _int, _int, _int, _vp]),
    "hg_alexnet_workspace_bytes": (_sz, [_int, _u32]),
    "hg_conv_weight_pack": (_int, [_vp, _int, _int, _int, _int, _int, _vp, _vp]),
    "hg_conv1_fused_floats": (_sz, [_int]),
    "hg_conv1_fused_pack": (_int, [_vp, _int, _vp, _vp]),
    "hg_alexnet_encode": (_int, [_vp, _int, _int, _vp, _int, _u32, _vp, _vp, _sz, _vp]),
    "hg_alexnet_encode_stochastic": (_int, [_vp, _int, _int, _vp, _int, _u32, _vp, _vp, _sz, C.c_uint64, _vp]),
    "hg_transpose_f32": (_int, [_vp, _int, _int, _vp, _vp]),
    "hg_popc_peak": (_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), _int, _vp]),
    "hg_i8_peak": (_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), _int, _vp]),
    "hg_alexnet_phase_ms": (_int, [C.POINTER(C.c_float)]),
}



class AlexNetWeightsStruct(C.Structure):
    """HgAlexNetWeights of include/hashgan_b200.h."""
    _fields_ = [("conv_w", C.c_void_p * 5), ("conv_b", C.c_void_p * 5),
                ("fc6_wt", C.c_void_p), ("fc6_b", C.c_void_p), ("fc7_wt", C.c_void_p), ("fc7_b", C.c_void_p),
                ("fc8_wt", C.c_void_p), ("fc8_b", C.c_void_p), ("conv_wt", C.c_void_p * 5), ("fc_wt3", C.c_void_p * 3),
                ("conv1_fused", C.c_void_p), ("conv1_fused_wh", C.c_int)]


_lib = None


class NativeLibraryError(RuntimeError):
    pass


class HgError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libhashgan_b200 error {code}: {message}")
        self.code = code


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises NativeLibraryError when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  hashgan_b200 has no CPU fallback.")
    try:
        handle = C.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover - depends on the box
        raise NativeLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name, None)
        if fn is None:
            raise NativeLibraryError(f"{LIB_PATH} does not export {name}; rebuild the library")
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != HG_OK:
        msg = lib().hg_last_error()
        raise HgError(rc, msg.decode("utf-8", "replace") if msg else "")


def code_words(b: int) -> int:
    w = lib().hg_code_words(int(b))
    if w == 0:
        raise ValueError(f"hash length b={b} is not supported (1..{MAX_BITS})")
    return w


def label_words(L: int) -> int:
    w = lib().hg_label_words(int(L))
    if w == 0:
        raise ValueError(f"label width L={L} is not supported (1..128)")
    return w


def row_words(b: int, L: int) -> int:
    code_words(b)
    label_words(L)
    return lib().hg_row_words(int(b), int(L))
