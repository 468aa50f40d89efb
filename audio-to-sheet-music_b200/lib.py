"""ctypes binding of libathtd.so (include/athtd.h).  No torch types cross the ABI: tensors are
passed as ``data_ptr()`` integers.  The library is sm_100a-only and there is NO CPU fallback:
loading fails loudly if the shared object is missing."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libathtd.so")
_lib = None

# every symbol include/athtd.h declares: (name, restype, argtypes)
_P, _I, _L = C.c_void_p, C.c_int, C.c_long
SYMBOLS = [
    ("athtd_last_error", C.c_char_p, []),
    ("athtd_version", _I, []),
    ("athtd_param_count", _I, []),
    ("athtd_param_name", C.c_char_p, [_I]),
    ("athtd_param_numel", _L, [_I]),
    ("athtd_param_offset", _L, [_I]),
    ("athtd_params_total", _L, []),
    ("athtd_packed_bytes", _L, [_I]),
    ("athtd_pack_weights", _I, [_P, _P, _I, _P]),
    ("athtd_workspace_bytes", _L, [_I, _I, _I, _I]),
    ("athtd_plan_create", _P, [_I, _I, _I, _I, _P, _P, _P, _L, _P, _P, _P, _P]),
    ("athtd_plan_destroy", None, [_P]),
    ("athtd_plan_tokens", _I, [_P, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    ("athtd_forward", _I, [_P, _P, _P, _P, _P]),
    ("athtd_encode", _I, [_P, _P, _P]),
    ("athtd_encode_normalized", _I, [_P, _P, _P, _P]),
    ("athtd_decode", _I, [_P, _P, _P, _P]),
    ("athtd_plan_launches", _I, [_P]),
    ("athtd_plan_set_batch", _I, [_P, _I]),
    ("athtd_plan_set_profile", _I, [_P, _I]),
    ("athtd_plan_get_profile", _I, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_I)]),
    ("athtd_tap", _I, [_P, C.c_char_p, C.POINTER(_P), C.POINTER(_L), C.POINTER(_I), C.POINTER(_I * 8)]),
    ("athtd_plan_set_tc", _I, [_P, _I]),
    ("athtd_plan_set_flash", _I, [_P, _I]),
    ("athtd_plan_set_fused_dconv", _I, [_P, _I]),
    ("athtd_plan_tc_launches", _I, [_P]),
    ("athtd_plan_set_graph", _I, [_P, _I]),
    ("athtd_plan_graph_replays", _I, [_P]),
    ("athtd_attention_set_poly", _I, [_I]),
    ("athtd_set_tc_tuning", _I, [_I]),
    ("athtd_set_pdl", _I, [_I]),
    ("athtd_attention_test", _I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    ("athtd_memcpy_d2d", _I, [_P, _P, _L, _P]),
    ("athtd_sdr_sums", _I, [_P, _P, _I, _L, _P, _P]),
    ("athtd_stft_cac", _I, [_P, _I, _I, _P, _P, _P, _P, _P]),
    ("athtd_istft", _I, [_P, _I, _I, _P, _P, _P, _P, _P]),
    ("athtd_gather_chunks", _I, [_P, _L, _I, _P, _I, _I, _P, _P]),
    ("athtd_chunk_ola", _I, [_P, _L, _I, _I, _P, _P, _P, _P, _I, _L, _P, _P, _P, _P, _I, _L, _L, _P]),
    ("athtd_clap_param_count", _I, []),
    ("athtd_clap_param_name", C.c_char_p, [_I]),
    ("athtd_clap_param_numel", _L, [_I]),
    ("athtd_clap_param_offset", _L, [_I]),
    ("athtd_clap_params_total", _L, []),
    ("athtd_clap_workspace_bytes", _L, [_I, _I]),
    ("athtd_clap_text_forward", _I, [_P, _P, _P, _I, _I, _P, _P, _I, _P]),
    ("athtd_chunk_fade_add", _I, [_P, _L, _I, _I, _P, _P, _P, _P, _I, _L, _P, _P, _P, _P, _I, _L, _L, _P]),
    ("athtd_load_audio", _I, [_P, _I, _L, _P, _I, _I, _I, _I, _P, _I, _L, _P]),
    ("athtd_gemm_test", _I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
]


class AthtdError(RuntimeError):
    pass


def load():
    """dlopen the in-tree shared object and bind every exported symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AthtdError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "This package has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)     # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = "athtd") -> None:
    if status != 0:
        raise AthtdError(f"{what}: {load().athtd_last_error().decode()}")


def param_table():
    lib = load()
    return [(lib.athtd_param_name(i).decode(), lib.athtd_param_numel(i), lib.athtd_param_offset(i))
            for i in range(lib.athtd_param_count())]
