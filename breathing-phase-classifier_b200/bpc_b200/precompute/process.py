"""Mirror of reference `src/precompute/process.py`: `process_and_save_npz((file_id, wav_path, target_dir))`.

Same contract (process.py:25-108): never raises, returns (file_id, success, error), writes `<target_dir>/<file_id>.npz`
with the ten float32 arrays the reference writes (process.py:92-103).  The arithmetic is one B=1 call into the CUDA
path; `core.process_dataset_threaded` batches instead.
"""
from __future__ import annotations

import os

import numpy as np

from .methods import SR, EXPECTED_LEN, N_MELS, N_MFCC, HOP_LENGTH, N_FFT, FMAX, DELTA_ORDER, N_GAMMATONE, N_LPC  # noqa: F401
from . import methods as _m

NPZ_KEYS = ("mel", "mfcc", "chroma", "mel_delta", "mel_delta2", "gammatone", "lpc", "mod_spec", "tempogram")


def load_wav(path: str):
    """librosa.load(path, sr=16000) for PCM wav files: int16 stays int16 (the engine scales by 1/32768 on device)."""
    import scipy.io.wavfile
    sr, data = scipy.io.wavfile.read(path)
    if sr != SR:
        raise ValueError(f"{path}: sample rate {sr} != {SR}; resampling is not part of this build")
    if data.ndim > 1:                                   # librosa.load(mono=True): channel mean
        data = data.astype(np.float32).mean(axis=1) / (32768.0 if data.dtype == np.int16 else 1.0)
        return data.astype(np.float32)
    if data.dtype == np.int16:
        return data
    if data.dtype == np.int32:
        return (data.astype(np.float64) / 2147483648.0).astype(np.float32)
    if data.dtype == np.uint8:
        return ((data.astype(np.float32) - 128.0) / 128.0).astype(np.float32)
    return data.astype(np.float32)


def fit_batch(waves, length=EXPECTED_LEN):
    """Stack waveforms into one [B, length] array (pad_or_truncate on the host for ragged inputs).
    Returns int16 when every input is int16, else float32."""
    all_i16 = all(w.dtype == np.int16 for w in waves)
    out = np.zeros((len(waves), length), dtype=np.int16 if all_i16 else np.float32)
    for i, w in enumerate(waves):
        n = min(len(w), length)
        if all_i16 or w.dtype != np.int16:
            out[i, :n] = w[:n]
        else:
            out[i, :n] = w[:n].astype(np.float32) / np.float32(32768.0)
    return out


def save_npz(target_dir: str, file_id: str, feats: np.ndarray, scalars: np.ndarray) -> None:
    """feats: [9, 128, T] in sorted-key order -> the reference's .npz members."""
    from .._lib import CHANNELS
    named = {k: feats[i] for i, k in enumerate(CHANNELS)}
    np.savez(os.path.join(target_dir, file_id + ".npz"), scalars=scalars, **{k: named[k] for k in NPZ_KEYS})


def process_and_save_npz(args):
    file_id, wav_path, target_dir = args
    try:
        y = load_wav(wav_path)
        wav = fit_batch([y])
        with _m.ENGINE_LOCK:                      # the reference runs this function on two threads (core.py:33-34)
            feats, scal, status = _m._get_engine().precompute_host(wav)
        if status[0] & 1:
            raise ValueError("non-finite samples in input")
        save_npz(target_dir, file_id, feats[0], scal[0])
        return file_id, True, None
    except Exception as e:  # noqa: BLE001 -- reference convention, process.py:107-108
        return file_id, False, str(e)
