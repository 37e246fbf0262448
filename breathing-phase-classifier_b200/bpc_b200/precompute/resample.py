"""Sample-rate conversion on load: the `librosa.load(wav_path, sr=SR)` of reference process.py:28 for files whose native
rate is not 16 kHz.  Runs on the B200 (`bpc_resample`, csrc/k_resample.cu); the filter is the 150 dB Kaiser stand-in for
libsoxr "HQ" described in include/bpc.h and oracle/resample.py."""
from __future__ import annotations

import numpy as np

from . import methods as _m


def resample(y: np.ndarray, sr_in: int, sr_out: int = _m.SR, engine=None) -> np.ndarray:
    """float32 [n] at sr_in -> float32 [ceil(n * sr_out / sr_in)] at sr_out."""
    y = np.ascontiguousarray(y, dtype=np.float32)
    if y.ndim != 1:
        raise ValueError("expected a mono waveform")
    if int(sr_in) == int(sr_out):
        return y
    if engine is not None:
        return engine.resample(y, int(sr_in), int(sr_out))
    with _m.ENGINE_LOCK:
        return _m._get_engine().resample(y, int(sr_in), int(sr_out))
