"""ctypes loader of libminlz_cuda.so (the C ABI in include/minlz_cuda.h)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MINLZ_CUDA_SO: an experiment / profiling build of the same library (build.build_variant)
SO_PATH = os.environ.get("MINLZ_CUDA_SO") or os.path.join(_HERE, "libminlz_cuda.so")

# every symbol include/minlz_cuda.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "mzcu_set_encoder_flavor": (C.c_int, [C.c_int]),
    "mzcu_get_encoder_flavor": (C.c_int, []),
    "mzcu_abi_version": (C.c_int, []),
    "mzcu_last_error": (C.c_char_p, []),
    "mzcu_device_count": (C.c_int, []),
    "mzcu_max_encoded_len": (C.c_int64, [C.c_int64]),
    "mzcu_decoded_len": (C.c_int64, [_P, C.c_size_t]),
    "mzcu_is_minlz": (C.c_int, [_P, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "mzcu_encode_blocks_dev": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "mzcu_decode_blocks_dev": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "mzcu_pack_blocks_dev": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "mzcu_encode_blocks": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "mzcu_encode_blocks_packed": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_size_t, _P]),
    "mzcu_decode_blocks": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "mzcu_crc32c_blocks_dev": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, _P]),
    "mzcu_crc32c_blocks": (C.c_int, [C.c_int, C.c_int, _P, _P, _P]),
    "mzcu_stream_encode_blocks": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_size_t, _P, _P]),
    "mzcu_stream_decode_blocks": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "mzcu_encode": (C.c_int64, [_P, C.c_size_t, _P, C.c_size_t, C.c_int]),
    "mzcu_try_encode": (C.c_int64, [_P, C.c_size_t, _P, C.c_size_t, C.c_int]),
    "mzcu_decode": (C.c_int64, [_P, C.c_size_t, _P, C.c_size_t]),
    "mzcu_encode_batch": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "mzcu_decode_batch": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "mzcu_host_alloc": (_P, [C.c_size_t]),
    "mzcu_host_free": (None, [_P]),
    "mzcu_last_kernel_ms": (C.c_float, []),
    "mzcu_stream_encode_blocks_multi": (C.c_int, [C.c_int, _P, C.c_int, C.c_int, _P, _P, _P, C.c_size_t, _P, _P, _P]),
    "mzcu_stream_decode_blocks_multi": (C.c_int, [C.c_int, _P, C.c_int, _P, _P, _P, _P, _P, _P, _P]),
    "mzcu_submit_stream_encode_blocks": (C.c_int64, [C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_size_t, _P, _P]),
    "mzcu_submit_stream_decode_blocks": (C.c_int64, [C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "mzcu_wait": (C.c_int, [C.c_int64]),
    "mzcu_set_validate": (C.c_int, [C.c_int]),
    "mzcu_get_validate": (C.c_int, []),
    "mzcu_validate_blocks_dev": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "mzcu_bind_host_to_device": (C.c_int, [C.c_int]),
}

_lib = None


def load():
    """Loads the CUDA library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                "minlz_b200: %s is missing -- build it with `python -m minlz_b200.build` "
                "(there is no CPU fallback)" % SO_PATH)
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
