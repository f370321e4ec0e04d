"""
ctypes binding of the C-ABI in include/sdft_b200.h.  Loading fails loudly when the CUDA library is
missing; nothing here computes on the CPU.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsdft_b200.so")

TYPE_SUFFIXES = ("f32f32", "f32f64", "f64f32", "f64f64")
_TD = {"f32": ctypes.c_float, "f64": ctypes.c_double}

# name -> (restype, argtypes) with TD standing for the time-domain scalar type
_P, _SZ, _I, _D, _V = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_double, None
TYPED = {
    "alloc": (_P, [_SZ]),
    "alloc_custom": (_P, [_SZ, _I, _D]),
    "alloc_batch": (_P, [_SZ, _I, _D, _SZ]),
    "free": (_V, [_P]),
    "reset": (_V, [_P]),
    "size": (_SZ, [_P]),
    "window": (_I, [_P]),
    "latency": (_D, [_P]),
    "sdft": (_V, [_P, "TD", _P]),
    "sdft_n": (_V, [_P, _SZ, _P, _P]),
    "sdft_nd": (_V, [_P, _SZ, _P, _P]),
    "isdft": ("TD", [_P, _P]),
    "isdft_n": (_V, [_P, _SZ, _P, _P]),
    "isdft_nd": (_V, [_P, _SZ, _P, _P]),
    "advance": (_V, [_P, _SZ, _P]),
    "sdft_hops": (_V, [_P, _SZ, _SZ, _P, _P, _SZ]),
    "sdft_batch": (_V, [_P, _SZ, _P, _P]),
    "isdft_batch": (_V, [_P, _SZ, _P, _P]),
    "roundtrip_n": (_V, [_P, _SZ, _P, _P]),
    "roundtrip_gain_n": (_V, [_P, _SZ, _P, _P, _P]),
    "convolve_n": (_V, [_P, _SZ, _P, _P]),
}
UNTYPED = {
    "sdft_b200_last_error": (_I, [_P]),
    "sdft_b200_last_error_string": (ctypes.c_char_p, [_P]),
    "sdft_b200_synchronize": (_I, [_P]),
    "sdft_b200_set_stream": (_I, [_P, _P]),
    "sdft_b200_set_streaming": (_I, [_P, ctypes.c_uint]),
    "sdft_b200_set_chunk": (_I, [_P, _SZ]),
    "sdft_b200_set_roi": (_I, [_P, _SZ, _SZ]),
    "sdft_b200_channels": (_SZ, [_P]),
    "sdft_b200_device": (_I, [_P]),
    "sdft_b200_table_bytes": (_SZ, [_P]),
    "sdft_b200_launch_count": (ctypes.c_ulonglong, [_P]),
    "sdft_b200_split_count": (ctypes.c_ulonglong, [_P]),
    "sdft_b200_set_profiling": (_I, [_P, _I]),
    "sdft_b200_kernel_ms": (_D, [_P, _I, ctypes.POINTER(ctypes.c_ulonglong)]),
    "sdft_b200_get_twiddles": (_I, [_P, _P, _P]),
    "sdft_b200_get_state": (_I, [_P, _SZ, ctypes.POINTER(ctypes.c_size_t), _P, _P, _P]),
    "sdft_b200_set_state": (_I, [_P, _SZ, _SZ, _P, _P]),
    "sdft_b200_measure_hbm": (_D, [_I, _P, _SZ, _I, ctypes.POINTER(ctypes.c_double)]),
    "sdft_b200_measure_dfma": (_D, [_I]),
    "sdft_b200_time_shard": (_I, [_SZ, _SZ, _SZ, _SZ] + [ctypes.POINTER(ctypes.c_size_t)] * 3),
    "sdft_b200_channel_shard": (_I, [_SZ, _SZ, _SZ] + [ctypes.POINTER(ctypes.c_size_t)] * 2),
    "sdft_b200_host_alloc": (_P, [_SZ]),
    "sdft_b200_host_free": (_V, [_P]),
    "sdft_b200_device_alloc": (_P, [_SZ]),
    "sdft_b200_device_free": (_V, [_P]),
    "sdft_b200_copy": (_I, [_P, _P, _P, _SZ]),
    "sdft_b200_version": (ctypes.c_char_p, []),
    "sdft_b200_debug_trace": (_SZ, [_P, _P, _SZ]),
}


def exported_symbols():
    """Every symbol include/sdft_b200.h declares."""
    names = list(UNTYPED)
    for sfx in TYPE_SUFFIXES:
        names += ["sdft_b200_%s_%s" % (sfx, fn) for fn in TYPED]
    return names


_lib = None


def load(path=None):
    """Loads libsdft_b200.so (building it first if the sources are newer and nvcc is present)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("SDFT_B200_LIB") or LIB_PATH
    if not os.path.exists(path):
        from . import build as _build
        _build.build()
    if not os.path.exists(path):
        raise OSError("libsdft_b200.so is missing and could not be built; sdft_b200 has no CPU fallback")
    lib = ctypes.CDLL(path)
    for name, (res, args) in UNTYPED.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    for sfx in TYPE_SUFFIXES:
        td = _TD[sfx[:3]]
        for fn, (res, args) in TYPED.items():
            f = getattr(lib, "sdft_b200_%s_%s" % (sfx, fn))
            f.restype = td if res == "TD" else res
            f.argtypes = [td if a == "TD" else a for a in args]
    _lib = lib
    return lib


def fn(lib, sfx, name):
    return getattr(lib, "sdft_b200_%s_%s" % (sfx, name))
