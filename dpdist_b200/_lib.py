"""ctypes binding of libdpdist_b200.so (the C ABI in include/dpdist_b200.h).

There is no CPU fallback: if the CUDA extension cannot be loaded every op raises.
"""
import ctypes
import os
import threading

import numpy as np

from . import build as _build

_LIB = None
_LOCK = threading.Lock()

c_float_p = ctypes.POINTER(ctypes.c_float)
c_int32_p = ctypes.POINTER(ctypes.c_int32)


class HeadConfig(ctypes.Structure):
    """struct dpd_head_config (include/dpdist_b200.h)."""
    _fields_ = [("n_clouds", ctypes.c_int), ("n_query", ctypes.c_int), ("G", ctypes.c_int),
                ("C", ctypes.c_int), ("k", ctypes.c_int), ("H", ctypes.c_int), ("flags", ctypes.c_int)]


HEAD_AUTO, HEAD_SIMT, HEAD_TC, HEAD_TC_TF32 = 0, 1, 2, 3
HEAD_TRAIN = 0x10
HEAD_INPUT_GRAD = 0x20
BWD_ALL, BWD_L4, BWD_L3, BWD_L2, BWD_L1 = 0, 1, 2, 3, 4

# name -> (restype, argtypes); must list every symbol include/dpdist_b200.h declares
SIGNATURES = {
    "dpd_version": (ctypes.c_int, []),
    "dpd_last_error": (ctypes.c_char_p, []),
    "dpd_fv_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_float_p,
                                      ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "dpd_voxel_assign": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_float_p,
                                        c_float_p, c_float_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p]),
    "dpd_local_patches": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p]),
    "dpd_head_packed_bytes": (ctypes.c_size_t, [ctypes.POINTER(HeadConfig)]),
    "dpd_head_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(HeadConfig)]),
    "dpd_head_pack_weights": (ctypes.c_int, [ctypes.POINTER(HeadConfig)] + [ctypes.c_void_p] * 8 +
                              [ctypes.c_void_p, ctypes.c_void_p]),
    "dpd_head_forward": (ctypes.c_int, [ctypes.POINTER(HeadConfig), ctypes.c_void_p, ctypes.c_void_p, c_float_p,
                                        c_float_p, c_float_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
}


class ProfileEntry(ctypes.Structure):
    """struct dpd_profile_entry."""
    _fields_ = [("name", ctypes.c_char * 48), ("ms", ctypes.c_double), ("launches", ctypes.c_longlong)]


SIGNATURES.update({
    "dpd_head_backward": (ctypes.c_int, [ctypes.POINTER(HeadConfig), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_int] + [ctypes.c_void_p] * 8 + [ctypes.c_void_p, ctypes.c_size_t,
                                                                                    ctypes.c_void_p]),
    "dpd_adam_step": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                     ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                     ctypes.c_void_p]),
    "dpd_adam_lr_t": (ctypes.c_float, [ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int]),
    "dpd_adam_step_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                         ctypes.c_void_p, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_void_p]),
    "dpd_debug_tc_gemm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                         ctypes.c_int, ctypes.c_void_p]),
    "dpd_debug_tc_operand_order": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "dpd_model_forward": (ctypes.c_int, [ctypes.POINTER(HeadConfig), ctypes.c_void_p, ctypes.c_int, ctypes.c_float, c_float_p,
                                         ctypes.c_void_p, c_float_p, c_float_p, c_float_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "dpd_fv_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_float_p, ctypes.c_float,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "dpd_head_backward_inputs": (ctypes.c_int, [ctypes.POINTER(HeadConfig), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "dpd_nearest_distance": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "dpd_assemble_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p]),
    "dpd_layer_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "dpd_layer_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "dpd_layer_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_size_t, ctypes.c_void_p]),
    "dpd_gather_rows": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "dpd_relu_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "dpd_add_inplace": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "dpd_bn_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float,
                                      ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                      ctypes.c_void_p]),
    "dpd_bn_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "dpd_crc32c": (ctypes.c_uint32, [ctypes.c_char_p, ctypes.c_size_t]),
    "dpd_launch_count": (ctypes.c_longlong, []),
    "dpd_profile_enable": (ctypes.c_int, [ctypes.c_int]),
    "dpd_profile_read": (ctypes.c_int, [ctypes.POINTER(ProfileEntry), ctypes.c_int, ctypes.c_int]),
})


def profile_read(reset=True, max_entries=64):
    """-> {kernel name: (total device ms, launches)} since the last reset."""
    arr = (ProfileEntry * max_entries)()
    n = load().dpd_profile_read(arr, max_entries, int(reset))
    return {arr[i].name.decode(): (arr[i].ms, arr[i].launches) for i in range(n)}


class DPDistNativeError(RuntimeError):
    pass


def lib_path():
    return _build.LIB_PATH


ABI_VERSION = 4          # must equal DPD_ABI_VERSION in include/dpdist_b200.h


def load():
    """Load the library, (re)building it first when a compiler is present and the sources changed.  Raises if unavailable:
    there is no fallback.  Concurrent ranks are safe: the build runs under a file lock and is renamed into place."""
    global _LIB
    if _LIB is not None:
        return _LIB
    with _LOCK:
        if _LIB is not None:
            return _LIB
        path = lib_path()
        try:
            stale = _build.needs_build()
        except Exception:       # noqa: BLE001  (unreadable sources: trust a library that is there)
            stale = not os.path.exists(path)
        if stale:
            try:
                _build.build_library()
            except Exception as e:  # no nvcc / compile error
                if not os.path.exists(path):
                    raise DPDistNativeError(
                        "libdpdist_b200.so is missing and could not be built (%s). "
                        "dpdist_b200 has no CPU fallback; run `python -m dpdist_b200.build`." % e) from e
        lib = ctypes.CDLL(path)
        lib.dpd_version.restype = ctypes.c_int
        got = lib.dpd_version()
        if got != ABI_VERSION:
            raise DPDistNativeError("%s implements ABI version %d, this package binds version %d: rebuild it "
                                    "(`python -m dpdist_b200.build --force`)" % (path, got, ABI_VERSION))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
        return _LIB


def check(rc, what):
    if rc != 0:
        msg = load().dpd_last_error()
        raise DPDistNativeError("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else ""))


def fptr(a):
    """numpy float32 array -> float*"""
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_float_p)
