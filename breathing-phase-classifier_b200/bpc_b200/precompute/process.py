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


def _to_float32(data: np.ndarray) -> np.ndarray:
    """soundfile's integer -> float32 scaling (what librosa.load sees): int16 / 2^15, int32 (and 24-bit, which scipy
    widens to int32) / 2^31, unsigned 8-bit (x - 128) / 2^7, floats as they are."""
    if data.dtype == np.int16:
        return data.astype(np.float32) / np.float32(32768.0)
    if data.dtype == np.int32:
        return (data.astype(np.float64) / 2147483648.0).astype(np.float32)
    if data.dtype == np.uint8:
        return (data.astype(np.float32) - np.float32(128.0)) / np.float32(128.0)
    return data.astype(np.float32)


def load_wav(path: str):
    """librosa.load(path, sr=16000) for PCM wav files (process.py:28): mono int16 at 16 kHz stays int16 (the engine
    applies the 1/32768 on device); every other sample format is scaled by dtype first, multi-channel files are then
    averaged over channels (librosa.to_mono), and any other sample rate is resampled to 16 kHz (resample.py)."""
    import scipy.io.wavfile
    sr, data = scipy.io.wavfile.read(path)
    if data.ndim == 1 and data.dtype == np.int16 and sr == SR:
        return data
    y = _to_float32(data)
    if y.ndim > 1:                                      # librosa.load(mono=True): channel mean, in float32
        y = np.mean(y, axis=1, dtype=np.float32)
    if sr != SR:
        from .resample import resample
        y = resample(y, sr, SR)
    return np.ascontiguousarray(y, dtype=np.float32)


WAV_ERR_OPEN, WAV_ERR_FORMAT, WAV_ERR_UNSUPPORTED = -10, -11, -12        # include/bpc.h: bpc_wav_code


def load_wav_batch(paths, length=EXPECTED_LEN, threads=8, engine=None):
    """`librosa.load(p, sr=16000)` + `pad_or_truncate` (process.py:28-29) for many files at once.

    The library's C++ reader pool handles what the reference is fed (RIFF PCM16 mono 16 kHz) straight into one
    [n, length] int16 batch.  Files it reports as unsupported (stereo, other sample formats or rates) are decoded ON THE
    GPU when the caller passes an engine or a callable that returns one (`Engine.decode_wavs`: the file images go to the device as read, scaling /
    down-mix / resampling / padding happen there), else by `load_wav` on the host (scipy); either way the batch becomes
    float32.  Returns (batch, errors) with errors[i] = None or the failure message (the row of a failed file is zero;
    the caller reports it as `(id, False, err)`)."""
    import ctypes as C
    from .._lib import lib
    n = len(paths)
    out = np.zeros((n, length), dtype=np.int16)
    sr = np.zeros(n, dtype=np.int32); frames = np.zeros(n, dtype=np.int32); code = np.zeros(n, dtype=np.int32)
    arr = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
    rc = lib().bpc_wav_load_batch(arr, n, SR, length, out.ctypes.data, sr.ctypes.data, frames.ctypes.data,
                                  code.ctypes.data, int(threads))
    if rc != 0:
        raise RuntimeError(f"bpc_wav_load_batch failed ({rc})")
    errors = [None] * n
    slow = {}
    general = [int(i) for i in np.flatnonzero(code == WAV_ERR_UNSUPPORTED)]
    if general and engine is not None:                  # GPU-side decode of everything the PCM16 reader does not take
        if callable(engine):                            # created on first use (most datasets never need it)
            engine = engine()
        images = []
        for i in general:
            with open(paths[i], "rb") as fh:
                images.append(fh.read())
        ydev, errs = engine.decode_wavs(images, length=length, sr=SR)
        yh = ydev.cpu().numpy()
        for j, i in enumerate(general):
            if errs[j] is None:
                slow[i] = yh[j]
            else:                                       # a format the device decoder does not take either: host reader
                try:
                    slow[i] = load_wav(paths[i])
                except Exception as e:  # noqa: BLE001
                    errors[i] = str(e)
        general = []
    for i in np.flatnonzero(code):
        if code[i] == WAV_ERR_UNSUPPORTED:              # stereo, other sample formats, other rates (resampled on the device)
            if i in slow or errors[i] is not None:
                continue
            try:
                slow[i] = load_wav(paths[i])
            except Exception as e:  # noqa: BLE001
                errors[i] = str(e)
        elif code[i] == WAV_ERR_OPEN:
            errors[i] = f"[Errno 2] No such file or directory: '{paths[i]}'"
        else:
            errors[i] = f"{paths[i]}: not a RIFF/WAVE file"
    if slow:
        outf = out.astype(np.float32) / np.float32(32768.0)
        for i, w in slow.items():
            w = w.astype(np.float32) / np.float32(32768.0) if w.dtype == np.int16 else w.astype(np.float32)
            m = min(len(w), length)
            outf[i] = 0
            outf[i, :m] = w[:m]
        return outf, errors
    return out, errors


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
