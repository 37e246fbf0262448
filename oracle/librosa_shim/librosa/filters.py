"""librosa.filters subset (test infrastructure; see package docstring)."""
from __future__ import annotations

import numpy as np
import scipy.signal

from . import util

# librosa.filters.WINDOW_BANDWIDTHS["hann"]
_HANN_BANDWIDTH = 1.50018310546875


def get_window(window, Nx, *, fftbins=True):
    if callable(window):
        return window(Nx)
    if isinstance(window, (str, tuple)) or np.isscalar(window):
        return scipy.signal.get_window(window, Nx, fftbins=fftbins)
    if isinstance(window, (np.ndarray, list)):
        if len(window) == Nx:
            return np.asarray(window)
        raise ValueError("window size mismatch")
    raise ValueError(window)


def window_bandwidth(window, n=1000):
    if window in ("hann", "hanning"):
        return _HANN_BANDWIDTH
    win = get_window(window, n)
    return n * np.sum(win ** 2) / (np.sum(np.abs(win)) ** 2 + util.tiny(win))


def _hz_to_mel(frequencies, htk=False):
    frequencies = np.asanyarray(frequencies, dtype=float)
    if htk:
        return 2595.0 * np.log10(1.0 + frequencies / 700.0)
    f_min, f_sp = 0.0, 200.0 / 3
    mels = (frequencies - f_min) / f_sp
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if frequencies.ndim:
        log_t = frequencies >= min_log_hz
        mels[log_t] = min_log_mel + np.log(frequencies[log_t] / min_log_hz) / logstep
    elif frequencies >= min_log_hz:
        mels = min_log_mel + np.log(frequencies / min_log_hz) / logstep
    return mels


def _mel_to_hz(mels, htk=False):
    mels = np.asanyarray(mels, dtype=float)
    if htk:
        return 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
    f_min, f_sp = 0.0, 200.0 / 3
    freqs = f_min + f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if mels.ndim:
        log_t = mels >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (mels[log_t] - min_log_mel))
    elif mels >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (mels - min_log_mel))
    return freqs


def _mel_frequencies(n_mels=128, *, fmin=0.0, fmax=11025.0, htk=False):
    min_mel = _hz_to_mel(fmin, htk=htk)
    max_mel = _hz_to_mel(fmax, htk=htk)
    mels = np.linspace(min_mel, max_mel, n_mels)
    return _mel_to_hz(mels, htk=htk)


def mel(*, sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm="slaney", dtype=np.float32):
    """Slaney-style triangular mel bank, [n_mels, 1 + n_fft//2] (librosa.filters.mel)."""
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=dtype)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = _mel_frequencies(n_mels + 2, fmin=fmin, fmax=fmax, htk=htk)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    if norm == "slaney":
        enorm = 2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels])
        weights *= enorm[:, np.newaxis]
    elif norm is not None:
        weights = util.normalize(weights, norm=norm, axis=-1)
    return weights


def _hz_to_octs(frequencies, *, tuning=0.0, bins_per_octave=12):
    A440 = 440.0 * 2.0 ** (tuning / bins_per_octave)
    return np.log2(np.asanyarray(frequencies) / (float(A440) / 16))


def chroma(*, sr, n_fft, n_chroma=12, tuning=0.0, ctroct=5.0, octwidth=2, norm=2, base_c=True,
           dtype=np.float32):
    """librosa.filters.chroma: [n_chroma, 1 + n_fft//2]."""
    frequencies = np.linspace(0, sr, n_fft, endpoint=False)[1:]
    frqbins = n_chroma * _hz_to_octs(frequencies, tuning=tuning, bins_per_octave=n_chroma)
    frqbins = np.concatenate(([frqbins[0] - 1.5 * n_chroma], frqbins))
    binwidthbins = np.concatenate((np.maximum(frqbins[1:] - frqbins[:-1], 1.0), [1]))
    D = np.subtract.outer(frqbins, np.arange(0, n_chroma, dtype="d")).T
    n_chroma2 = np.round(float(n_chroma) / 2)
    D = np.remainder(D + n_chroma2 + 10 * n_chroma, n_chroma) - n_chroma2
    wts = np.exp(-0.5 * (2 * D / np.tile(binwidthbins, (n_chroma, 1))) ** 2)
    wts = util.normalize(wts, norm=norm, axis=0)
    if octwidth is not None:
        wts *= np.tile(np.exp(-0.5 * (((frqbins / n_chroma - ctroct) / octwidth) ** 2)), (n_chroma, 1))
    if base_c:
        wts = np.roll(wts, -3 * (n_chroma // 12), axis=0)
    return np.ascontiguousarray(wts[:, : int(1 + n_fft / 2)], dtype=dtype)


def cq_to_chroma(n_input, *, bins_per_octave=12, n_chroma=12, fmin=None, window=None, base_c=True,
                 dtype=np.float32):
    n_merge = float(bins_per_octave) / n_chroma
    if fmin is None:
        fmin = 32.70319566257483  # note_to_hz("C1")
    if np.mod(n_merge, 1) != 0:
        raise ValueError("bins_per_octave must be a multiple of n_chroma")
    cq_to_ch = np.repeat(np.eye(n_chroma), int(n_merge), axis=1)
    cq_to_ch = np.roll(cq_to_ch, -int(n_merge // 2), axis=1)
    n_octaves = np.ceil(float(n_input) / bins_per_octave)
    cq_to_ch = np.tile(cq_to_ch, int(n_octaves))[:, :n_input]
    midi_0 = np.mod(12 * (np.log2(fmin) - np.log2(440.0)) + 69, 12)
    roll = midi_0 if base_c else midi_0 - 9
    roll = int(np.round(roll * (n_chroma / 12.0)))
    cq_to_ch = np.roll(cq_to_ch, roll, axis=0).astype(dtype)
    if window is not None:
        cq_to_ch = scipy.signal.convolve(cq_to_ch, np.atleast_2d(window), mode="same")
    return cq_to_ch


def relative_bandwidth(*, freqs):
    bpo = np.empty_like(freqs)
    logf = np.log2(freqs)
    bpo[0] = 1 / (logf[1] - logf[0])
    bpo[-1] = 1 / (logf[-1] - logf[-2])
    bpo[1:-1] = 2 / (logf[2:] - logf[:-2])
    return (2.0 ** (2 / bpo) - 1) / (2.0 ** (2 / bpo) + 1)


def wavelet_lengths(*, freqs, sr, window="hann", filter_scale=1, gamma=0, alpha=None):
    freqs = np.asarray(freqs)
    if alpha is None:
        alpha = relative_bandwidth(freqs=freqs)
    else:
        alpha = np.asarray(alpha)
    gamma_ = alpha * 24.7 / 0.108 if gamma is None else gamma
    Q = float(filter_scale) / alpha
    f_cutoff = max(freqs * (1 + 0.5 * window_bandwidth(window) / Q) + 0.5 * gamma_)
    lengths = Q * sr / (freqs + gamma_ / alpha)
    return lengths, f_cutoff


def _float_window(window_spec):
    def _wrap(n, *args, **kwargs):
        n_min, n_max = int(np.floor(n)), int(np.ceil(n))
        window = get_window(window_spec, n_min)
        if len(window) < n_max:
            window = np.pad(window, [(0, n_max - n_min)], mode="constant")
        window[n_min:] = 0.0
        return window
    return _wrap


def wavelet(*, freqs, sr, window="hann", filter_scale=1, pad_fft=True, norm=1, dtype=np.complex64,
            gamma=0, alpha=None):
    lengths, _ = wavelet_lengths(freqs=freqs, sr=sr, window=window, filter_scale=filter_scale,
                                 gamma=gamma, alpha=alpha)
    filts = []
    for ilen, freq in zip(lengths, freqs):
        sig = util.phasor(np.arange(-ilen // 2, ilen // 2, dtype=float) * 2 * np.pi * freq / sr)
        sig = sig * _float_window(window)(len(sig))
        sig = util.normalize(sig, norm=norm)
        filts.append(sig)
    max_len = max(lengths)
    if pad_fft:
        max_len = int(2.0 ** (np.ceil(np.log2(max_len))))
    else:
        max_len = int(np.ceil(max_len))
    filts = np.asarray([util.pad_center(f, size=max_len) for f in filts], dtype=dtype)
    return filts, lengths
