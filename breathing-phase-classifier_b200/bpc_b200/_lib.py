"""ctypes binding of libbpc_b200.so (include/bpc.h).  There is no fallback: a missing library or GPU raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BPC_LIB") or os.path.join(_HERE, "libbpc_b200.so")

NUM_CHANNELS = 9
NUM_SCALARS = 36
PLANE_ROWS = 128
LIVE_ROWS = (24, 64, 12, 128, 128, 128, 120, 40, 128)     # data rows per plane (sorted-key order); the rest is one pad value
LIVE_TOTAL = 772
WAV_F32, WAV_PCM16 = 0, 1
# sorted .npz keys == channel order of the [B, 9, 128, T] tensor (reference src/dataset.py:26)
CHANNELS = ("chroma", "gammatone", "lpc", "mel", "mel_delta", "mel_delta2", "mfcc", "mod_spec", "tempogram")
SEG_NONFINITE, SEG_TUNING_EMPTY, SEG_CAND_OVERFLOW, SEG_SILENT = 1, 2, 4, 8
MIX_NONE, MIX_MIXUP, MIX_CUTMIX = 0, 1, 2


class WavInfo(C.Structure):
    """include/bpc.h::bpc_wav_info"""
    _fields_ = [("data_offset", C.c_int64), ("frames", C.c_int64), ("sr", C.c_int32), ("channels", C.c_int32),
                ("fmt", C.c_int32), ("reserved", C.c_int32)]


class Params(C.Structure):
    """bpc_params: the module constants of reference process.py:12-23."""
    _fields_ = [("sr", C.c_int32), ("n_fft", C.c_int32), ("hop", C.c_int32), ("n_mels", C.c_int32),
                ("n_mfcc", C.c_int32), ("fmax", C.c_float), ("n_gammatone", C.c_int32), ("n_lpc", C.c_int32),
                ("expected_len", C.c_int32), ("pad_scalars_to", C.c_int32)]


class BpcError(RuntimeError):
    pass


_lib = None

_SIGS = {
    "bpc_abi_version": (C.c_int, []),
    "bpc_default_params": (None, [C.POINTER(Params)]),
    "bpc_num_frames": (C.c_int, [C.POINTER(Params)]),
    "bpc_num_scalars": (C.c_int, [C.POINTER(Params)]),
    "bpc_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(Params), C.c_int, C.c_int64]),
    "bpc_destroy": (None, [C.c_void_p]),
    "bpc_last_error": (C.c_char_p, [C.c_void_p]),
    "bpc_precompute": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
    "bpc_precompute_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    "bpc_precompute_host_compact": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p]),
    "bpc_precompute_host_compact_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p,
                                                    C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]),
    "bpc_host_wait": (C.c_int, [C.c_void_p, C.c_int64]),
    "bpc_live_rows": (C.c_int, [C.c_int]),
    "bpc_expand_compact": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int]),
    "bpc_host_alloc": (C.c_void_p, [C.c_void_p, C.c_int64, C.POINTER(C.c_int)]),
    "bpc_host_free": (None, [C.c_void_p, C.c_void_p]),
    "bpc_wav_parse": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "bpc_wav_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                 C.c_void_p, C.c_void_p]),
    "bpc_resample_len": (C.c_int64, [C.c_int64, C.c_int, C.c_int]),
    "bpc_resample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "bpc_resample_filter": (C.c_int64, [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                        C.POINTER(C.c_int)]),
    "bpc_stage_logmel": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "bpc_modspec": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "bpc_channel_stats": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bpc_channel_stats_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "bpc_channel_stats_reset": (C.c_int, [C.c_void_p]),
    "bpc_debug_copy": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "bpc_set_debug": (C.c_int, [C.c_void_p, C.c_int]),
    "bpc_table_copy": (C.c_int64, [C.POINTER(Params), C.c_char_p, C.c_int, C.c_void_p, C.c_int64]),
    "bpc_chunk_size": (C.c_int, [C.c_void_p]),
    "bpc_launch_count": (C.c_int64, [C.c_void_p]),
    "bpc_set_kernel_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "bpc_kernel_times": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "bpc_kernel_name": (C.c_char_p, [C.c_int]),
    "bpc_collate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                              C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bpc_wav_load_batch": (C.c_int, [C.POINTER(C.c_char_p), C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int]),
    "bpc_npz_size": (C.c_int64, [C.c_int, C.c_int]),
    "bpc_npz_pack": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64]),
    "bpc_npz_write_batch": (C.c_int, [C.c_char_p, C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_int, C.c_int, C.c_int, C.c_void_p]),
}
EXPORTS = tuple(_SIGS)


def lib():
    """Load libbpc_b200.so once; raises BpcError when it has not been built (run __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BpcError(f"{LIB_PATH} is missing: build it with `make -C breathing-phase-classifier_b200` "
                           "(there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.bpc_abi_version() != 3:
            raise BpcError("libbpc_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def default_params(**overrides) -> Params:
    p = Params()
    lib().bpc_default_params(C.byref(p))
    for k, v in overrides.items():
        setattr(p, k, v)
    return p
