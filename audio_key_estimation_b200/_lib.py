"""ctypes binding of libake_b200.so (the C ABI declared in include/ake_b200.h).

There is deliberately no fallback: if the shared library is missing it is built
with nvcc, and if that is impossible the import of any compute entry point
raises.  Nothing here touches oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import threading

from . import build as _build

AKE_OK = 0
AKE_ERR_INVALID, AKE_ERR_UNSUPPORTED, AKE_ERR_CUDA, AKE_ERR_WORKSPACE = -1, -2, -3, -4
ROW_FLOATS = 35  # AKE_ROW_FLOATS: 12 key + 12 tonic + 11 genre
CQT_LOGMAG, CQT_COMPLEX = 0, 1
CQT_RECURSION_092, CQT_RECURSION_HALVE_WHILE_EVEN = 0, 1


class PcnConfig(C.Structure):
    """struct ake_pcn_config -- mirrors the `opt` fields PitchClassNet reads (models.py:260-350, 662-742)."""
    _fields_ = [(n, C.c_int32) for n in (
        "pitches", "pitch_classes", "num_layers", "kernel_size", "conv_layers", "n_filters", "head_layers",
        "time_pool_size", "genre", "max_pool", "resblock", "denseblock", "stay_sixth", "only_semitones",
        "p2pc_conv", "pc2p_mem", "local", "frames", "loc_window_size")]


_P = C.c_void_p
_SIGNATURES = {
    "ake_abi_version": (C.c_int, []),
    "ake_last_error": (C.c_char_p, []),
    "ake_launch_count": (C.c_int64, [C.c_int]),
    "ake_profile_enable": (C.c_int, [C.c_int]),
    "ake_profile_collect": (C.c_int, [_P, C.c_int, _P, _P, C.c_int]),
    "ake_pcn_create": (C.c_int, [C.POINTER(PcnConfig), C.POINTER(_P)]),
    "ake_pcn_destroy": (None, [_P]),
    "ake_pcn_num_tensors": (C.c_int, [_P]),
    "ake_pcn_tensor_name": (C.c_char_p, [_P, C.c_int]),
    "ake_pcn_tensor_shape": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int64 * 4)]),
    "ake_pcn_param_floats": (C.c_int64, [_P]),
    "ake_pcn_set_params_f32": (C.c_int, [_P, _P, C.c_int64, _P]),
    "ake_pcn_workspace_bytes": (C.c_size_t, [_P, C.c_int, C.c_int, C.c_int]),
    "ake_pcn_forward_f32": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_int, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "ake_pcn_bn_channels": (C.c_int, [_P]),
    "ake_pcn_bn_counts": (C.c_int, [_P, _P, C.c_int]),
    "ake_pcn_local_frames": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int)]),
    "ake_pcn_forward_rows_f32": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, _P, _P, C.c_size_t, _P]),
    "ake_pcn_backward_f32": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, _P, C.c_size_t, _P]),
    "ake_loss_f32": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_float, C.c_float, C.c_float, _P, _P, _P, _P, _P]),
    "ake_pcn_get_config": (C.c_int, [_P, C.POINTER(PcnConfig)]),
    "ake_pcn_get_tap": (C.c_int64, [_P, C.c_char_p, _P, C.c_int64, _P]),
    "ake_decode_f32": (C.c_int, [_P, _P, _P, C.c_int, _P, _P, _P, _P]),
    "ake_mirex_f32": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, _P, _P, _P, _P]),
    "ake_adam_step_f32": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                    C.c_float, C.c_int, _P]),
    "ake_cqt_create": (C.c_int, [C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                 C.POINTER(_P)]),
    "ake_cqt_create_ex": (C.c_int, [C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                    C.POINTER(_P)]),
    "ake_cqt_set_peak": (C.c_int, [_P, C.c_float]),
    "ake_cqt_destroy": (None, [_P]),
    "ake_cqt_n_fft": (C.c_int, [_P]),
    "ake_cqt_n_bins": (C.c_int, [_P]),
    "ake_cqt_frames": (C.c_int, [_P, C.c_int64]),
    "ake_cqt_get_bank": (C.c_int, [_P, _P, C.c_int64]),
    "ake_cqt_get_decimator": (C.c_int, [_P, _P, C.c_int]),
    "ake_cqt_workspace_bytes": (C.c_size_t, [_P, C.c_int, C.c_int64]),
    "ake_cqt_run_f32": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int, C.c_int64, C.c_int, _P, C.c_int, _P, _P,
                                  C.c_size_t, _P]),
    "ake_estimate_workspace_bytes": (C.c_size_t, [_P, _P, C.c_int, C.c_int64]),
    "ake_estimate_host_f32": (C.c_int, [_P, _P, _P, C.c_int64, _P, C.c_int, C.c_int64, _P, _P, _P, _P, _P,
                                        C.c_size_t, _P]),
    "ake_estimate_workspace_bytes_i16": (C.c_size_t, [_P, _P, C.c_int, C.c_int64]),
    "ake_estimate_host_i16": (C.c_int, [_P, _P, _P, C.c_int64, _P, C.c_int, C.c_int64, _P, _P, _P, _P, _P,
                                        C.c_size_t, _P]),
}

_lock = threading.Lock()
_lib = None


def header_symbols() -> list:
    """Every function name include/ake_b200.h declares (used by the symbol-export test)."""
    with open(os.path.join(_build.INCLUDE, "ake_b200.h")) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(ake_[a-z0-9_]+)\s*\(", text)))


def lib() -> C.CDLL:
    """Load (building if needed) libake_b200.so.  Raises if it cannot be produced."""
    global _lib
    with _lock:
        if _lib is None:
            path = os.environ.get("AKE_LIB_PATH") or _build.LIB_PATH  # AKE_LIB_PATH: a build variant under test (tools/)
            if path == _build.LIB_PATH and (not os.path.exists(path) or os.environ.get("AKE_REBUILD")):
                path = _build.build()
            handle = C.CDLL(path)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(handle, name)  # AttributeError if the library lacks a declared symbol
                fn.restype, fn.argtypes = res, args
            if handle.ake_abi_version() != 2:
                raise RuntimeError("libake_b200.so ABI version mismatch")
            _lib = handle
    return _lib


def profile_enable(on: bool) -> None:
    check(lib().ake_profile_enable(int(on)))


def profile_collect() -> dict:
    """{tag: (milliseconds, kernel launches)} accumulated since the last collect (synchronises)."""
    cap, stride = 64, 48
    tags = C.create_string_buffer(cap * stride)
    ms = (C.c_double * cap)()
    launches = (C.c_int64 * cap)()
    n = lib().ake_profile_collect(tags, stride, ms, launches, cap)
    if n < 0:
        check(n)
    return {tags.raw[i * stride:(i + 1) * stride].split(b"\0", 1)[0].decode(): (ms[i], int(launches[i])) for i in range(n)}


class AkeError(RuntimeError):
    pass


def check(rc: int) -> None:
    """Map C status codes onto the exceptions the reference surface raises."""
    if rc == AKE_OK:
        return
    msg = (lib().ake_last_error() or b"").decode()
    if rc == AKE_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == AKE_ERR_INVALID:
        raise ValueError(msg)
    raise AkeError(f"libake_b200 error {rc}: {msg}")
