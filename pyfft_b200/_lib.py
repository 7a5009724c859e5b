"""ctypes binding of libb2fft.so (the C ABI declared in include/b2fft.h).

The library is built in-tree by ``pyfft_b200/build.py`` (nvcc, sm_100a).  If it is missing
this module raises -- there is deliberately no fallback implementation.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2fft.so")

OK, E_INVALID, E_CUDA, E_UNSUPPORTED = 0, -1, -2, -3
F32, F64 = 0, 1
INTERLEAVED, SPLIT = 0, 1
AXIS_X, AXIS_Y, AXIS_Z = 1, 2, 4

_lib = None
_lock = threading.Lock()

_vp, _i, _i64, _d, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/b2fft.h declares (tests check this)
SIGNATURES = {
    "b2fft_version": (_i, []),
    "b2fft_last_error": (ctypes.c_char_p, []),
    "b2fft_plan_create": (_i, [ctypes.POINTER(_vp), _i, ctypes.POINTER(_i64), _i, _i, _i, _d, _i, _i]),
    "b2fft_plan_create_ex": (_i, [ctypes.POINTER(_vp), ctypes.POINTER(_i64), _i, _i, _i, _i, _d, _i, _i, _d, _i]),
    "b2fft_plan_workspace_bytes": (_i, [_vp, _i64, ctypes.POINTER(_sz)]),
    "b2fft_plan_workspace_bytes_ex": (_i, [_vp, _i64, _i, ctypes.POINTER(_sz)]),
    "b2fft_plan_set_workspace": (_i, [_vp, _vp, _sz]),
    "b2fft_execute": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i64, _vp]),
    "b2fft_plan_destroy": (_i, [_vp]),
    "b2fft_plan_set_output_blocks": (_i, [_vp, _i, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i64, _i64]),
    "b2fft_plan_set_input_blocks": (_i, [_vp, _i, ctypes.POINTER(_vp)]),
    "b2fft_slab_plan_create": (_i, [ctypes.POINTER(_vp), ctypes.POINTER(_i64), _i, _i, _d, _i, _i, _i, _i, _i, _i, _i]),
    "b2fft_slab_plan_sizes": (_i, [_vp, ctypes.POINTER(_sz), ctypes.POINTER(_sz), ctypes.POINTER(_sz)]),
    "b2fft_slab_plan_geometry": (_i, [_vp, ctypes.POINTER(_i64)]),
    "b2fft_slab_plan_attach": (_i, [_vp, _vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    "b2fft_slab_forward": (_i, [_vp, _vp]),
    "b2fft_slab_inverse": (_i, [_vp, _vp]),
    "b2fft_slab_plan_status": (_i, [_vp, ctypes.POINTER(_i)]),
    "b2fft_slab_plan_launch_count": (_i64, [_vp]),
    "b2fft_slab_plan_describe": (_i, [_vp, ctypes.c_char_p, _sz]),
    "b2fft_slab_plan_set_trace": (_i, [_vp, _i]),
    "b2fft_slab_plan_trace": (_i, [_vp, ctypes.c_char_p, _sz]),
    "b2fft_slab_plan_destroy": (_i, [_vp]),
    "b2fft_slab_last_error": (ctypes.c_char_p, []),
    "b2fft_plan_set_outer_split": (_i, [_vp, _i64, _i64, _i64, _i64, _i64]),
    "b2fft_plan_set_exchange_ctas": (_i, [_vp, _i]),
    "b2fft_plan_set_max_ctas": (_i, [_vp, _i]),
    "b2fft_plan_set_progress": (_i, [_vp, _vp, _i64, _i, ctypes.POINTER(_i64)]),
    "b2fft_slab_plan_set_overlap": (_i, [_vp, _i]),
    "b2fft_slab_plan_set_option": (_i, [_vp, ctypes.c_char_p, _d]),
    "b2fft_slab_schedule_preview": (_i, [_i, _i, _i, _i, _i, ctypes.c_char_p, _sz]),
    "b2fft_mem_alloc": (_i, [_sz, _i, ctypes.POINTER(_vp)]),
    "b2fft_mem_free": (_i, [_vp]),
    "b2fft_ipc_export": (_i, [_vp, ctypes.c_char_p]),
    "b2fft_ipc_import": (_i, [ctypes.c_char_p, _i, ctypes.POINTER(_vp)]),
    "b2fft_ipc_release": (_i, [_vp]),
    "b2fft_stream_synchronize": (_i, [_vp]),
    "b2fft_plan_num_passes": (_i, [_vp]),
    "b2fft_plan_describe": (_i, [_vp, ctypes.c_char_p, _sz]),
    "b2fft_plan_preview": (_i, [ctypes.POINTER(_i64), _i, _i, _i, ctypes.c_char_p, _sz]),
    "b2fft_plan_launch_count": (_i64, [_vp]),
    "b2fft_num_variants": (_i, []),
    "b2fft_variant_info": (_i, [_i, ctypes.c_char_p, _sz]),
    "b2fft_run_variant": (_i, [_i, _vp, _vp, _vp, _vp, _i, _i, _i64, _i64, _i, _vp]),
    "b2fft_set_preferred_variants": (_i, [ctypes.c_char_p]),
    "b2fft_set_option": (_i, [ctypes.c_char_p, _d]),
}


def load():
    """Load libb2fft.so once; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "pyfft_b200: %s not found. Build it with `python -m pyfft_b200.build` "
                "(needs nvcc); there is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    msg = load().b2fft_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc):
    """Map C status codes onto the exception types the reference raises."""
    if rc == OK:
        return
    msg = last_error()
    if rc == E_INVALID:
        raise ValueError(msg)
    if rc == E_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError("b2fft: " + msg)
