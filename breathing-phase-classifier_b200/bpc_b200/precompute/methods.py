"""Mirror of reference `src/precompute/methods.py`: same names, arguments and return shapes.

The numerical helpers run on the B200 through libbpc_b200.so (a lazily created single-segment Engine); the three
pad helpers are pure data movement and stay on the host exactly as in the reference (methods.py:24-46).
"""
from __future__ import annotations

import threading

import numpy as np

SR = 16000                      # methods.py:10-22
DURATION = 1.0
EXPECTED_LEN = int(SR * DURATION)
N_MELS = 128
N_MFCC = 40
HOP_LENGTH = 256
N_FFT = 512
FMAX = 4500
N_WORKERS = 2

DELTA_ORDER = 2
N_GAMMATONE = 64
N_LPC = 12

_engine = None
# The reference calls process_and_save_npz from N_WORKERS threads at once (core.py:33-34) and the function is
# re-entrant.  Calls on one bpc_handle must be serialised (include/bpc.h), so the per-file mirrors share one engine
# behind this lock; batched callers (core.process_dataset_threaded) own their engine and never take it.
ENGINE_LOCK = threading.RLock()


def _get_engine():
    global _engine
    with ENGINE_LOCK:
        if _engine is None:
            from ..engine import Engine
            _engine = Engine(device=0, max_batch=64, debug=True)
        return _engine


def pad_or_truncate(waveform: np.ndarray, target_len: int) -> np.ndarray:
    """methods.py:24-28."""
    n = len(waveform)
    if n >= target_len:
        return waveform[:target_len]
    return np.concatenate([waveform, np.zeros(target_len - n, dtype=np.float32)])


def pad_time(spec2d: np.ndarray, from_bins: int, T_fixed: int) -> np.ndarray:
    """methods.py:30-37."""
    t_raw = spec2d.shape[1]
    if t_raw >= T_fixed:
        return spec2d[:, :T_fixed]
    block = np.full((from_bins, T_fixed - t_raw), spec2d.min(), dtype=np.float32)
    return np.concatenate([spec2d, block], axis=1)


def pad_freq(spec2d: np.ndarray, from_bins: int, to_bins: int) -> np.ndarray:
    """methods.py:39-46."""
    if from_bins >= to_bins:
        return spec2d[:to_bins, :]
    rows = np.full((to_bins - from_bins, spec2d.shape[1]), spec2d.min(), dtype=np.float32)
    return np.concatenate([spec2d, rows], axis=0)


def _one(y: np.ndarray) -> np.ndarray:
    y = np.asarray(y, dtype=np.float32)
    if y.ndim != 1:
        raise ValueError("expected a mono waveform")
    return np.ascontiguousarray(y[None, :])


def extract_enhanced_scalar_features(y: np.ndarray, sr: int = SR) -> np.ndarray:
    """methods.py:48-114 -> float32[36] (the code emits 36 values although README / model defaults say 39)."""
    if sr != SR:
        raise ValueError("this build implements the reference sample rate (16000) only")
    with ENGINE_LOCK:
        _, scal, _ = _get_engine().precompute_host(_one(y))
    return scal[0, :36].copy()


def extract_lpc_features(y: np.ndarray, order: int = N_LPC) -> np.ndarray:
    """methods.py:116-134 -> float32 [order, n_frames] (Burg, 25 ms Hamming frames every 10 ms)."""
    if order != N_LPC:
        raise ValueError("this build implements order 12 only")
    with ENGINE_LOCK:
        eng = _get_engine()
        eng.precompute_host(_one(y))
        return eng.debug("lpc_raw", 1)[0]


def extract_gammatone_features(y: np.ndarray, sr: int = SR, n_filters: int = N_GAMMATONE) -> np.ndarray:
    """methods.py:136-140 -> float32 [64, T] = log1p(mel64 @ |STFT512|)."""
    if sr != SR or n_filters != N_GAMMATONE:
        raise ValueError("this build implements sr 16000 / 64 filters only")
    with ENGINE_LOCK:
        eng = _get_engine()
        eng.precompute_host(_one(y))
        return eng.debug("gammatone_raw", 1)[0]


def extract_spectral_modulation_features(mel_db: np.ndarray) -> np.ndarray:
    """methods.py:142-143 -> float32 [40, T]: ortho DCT-II over mel (first 40), then over time."""
    with ENGINE_LOCK:
        return _get_engine().modspec(np.asarray(mel_db, dtype=np.float32)[None])[0]
