"""librosa.feature subset (test infrastructure; see package docstring)."""
from __future__ import annotations

import numpy as np
import scipy.signal
import scipy.fftpack
import scipy.ndimage

from . import util
from . import filters
from ._core import (stft, _spectrogram, power_to_db, fft_frequencies, estimate_tuning, autocorrelate, cqt)


def melspectrogram(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None, window="hann",
                   center=True, pad_mode="constant", power=2.0, **kwargs):
    S, n_fft = _spectrogram(y=y, S=S, n_fft=n_fft, hop_length=hop_length, power=power, win_length=win_length,
                            window=window, center=center, pad_mode=pad_mode)
    mel_basis = filters.mel(sr=sr, n_fft=n_fft, **kwargs)
    return np.einsum("...ft,mf->...mt", S, mel_basis, optimize=True)


def delta(data, *, width=9, order=1, axis=-1, mode="interp", **kwargs):
    data = np.atleast_1d(data)
    if mode == "interp" and width > data.shape[axis]:
        raise ValueError("when mode='interp', width cannot exceed data.shape[axis]")
    if width < 3 or np.mod(width, 2) != 1:
        raise ValueError("width must be an odd integer >= 3")
    if order <= 0 or not isinstance(order, (int, np.integer)):
        raise ValueError("order must be a positive integer")
    kwargs.pop("deriv", None)
    kwargs.setdefault("polyorder", order)
    return scipy.signal.savgol_filter(data, width, deriv=order, axis=axis, mode=mode, **kwargs)


def mfcc(*, y=None, sr=22050, S=None, n_mfcc=20, dct_type=2, norm="ortho", lifter=0, **kwargs):
    if S is None:
        S = power_to_db(melspectrogram(y=y, sr=sr, **kwargs))
    M = scipy.fftpack.dct(S, axis=-2, type=dct_type, norm=norm)[..., :n_mfcc, :]
    if lifter != 0:
        raise NotImplementedError
    return M


def chroma_stft(*, y=None, sr=22050, S=None, norm=np.inf, n_fft=2048, hop_length=512, win_length=None,
                window="hann", center=True, pad_mode="constant", tuning=None, n_chroma=12, **kwargs):
    S, n_fft = _spectrogram(y=y, S=S, n_fft=n_fft, hop_length=hop_length, power=2, win_length=win_length,
                            window=window, center=center, pad_mode=pad_mode)
    if tuning is None:
        tuning = estimate_tuning(S=S, sr=sr, bins_per_octave=n_chroma)
    chromafb = filters.chroma(sr=sr, n_fft=n_fft, tuning=tuning, n_chroma=n_chroma, **kwargs)
    raw_chroma = np.einsum("cf,...ft->...ct", chromafb, S, optimize=True)
    return util.normalize(raw_chroma, norm=norm, axis=-2)


def chroma_cqt(*, y=None, sr=22050, C=None, hop_length=512, fmin=None, norm=np.inf, threshold=0.0, tuning=None,
               n_chroma=12, n_octaves=7, window=None, bins_per_octave=36, cqt_mode="full"):
    if bins_per_octave is None:
        bins_per_octave = n_chroma
    elif np.remainder(bins_per_octave, n_chroma) != 0:
        raise ValueError("bins_per_octave must be an integer multiple of n_chroma")
    if cqt_mode != "full":
        raise NotImplementedError(cqt_mode)
    if C is None:
        C = np.abs(cqt(y, sr=sr, hop_length=hop_length, fmin=fmin, n_bins=n_octaves * bins_per_octave,
                       bins_per_octave=bins_per_octave, tuning=tuning))
    cq_to_chr = filters.cq_to_chroma(C.shape[-2], bins_per_octave=bins_per_octave, n_chroma=n_chroma,
                                     fmin=fmin, window=window)
    chroma = np.einsum("cf,...ft->...ct", cq_to_chr, C, optimize=True)
    if threshold is not None:
        chroma[chroma < threshold] = 0.0
    return util.normalize(chroma, norm=norm, axis=-2)


def chroma_cens(*, y=None, sr=22050, C=None, hop_length=512, fmin=None, tuning=None, n_chroma=12, n_octaves=7,
                bins_per_octave=36, cqt_mode="full", window=None, norm=2, win_len_smooth=41,
                smoothing_window="hann"):
    chroma = chroma_cqt(y=y, C=C, sr=sr, hop_length=hop_length, fmin=fmin, bins_per_octave=bins_per_octave,
                        tuning=tuning, norm=None, n_chroma=n_chroma, n_octaves=n_octaves, cqt_mode=cqt_mode,
                        window=window)
    chroma = util.normalize(chroma, norm=1, axis=-2)
    QUANT_STEPS = [0.4, 0.2, 0.1, 0.05]
    QUANT_WEIGHTS = [0.25, 0.25, 0.25, 0.25]
    chroma_quant = np.zeros_like(chroma)
    for step, weight in zip(QUANT_STEPS, QUANT_WEIGHTS):
        chroma_quant += (chroma > step) * weight
    if win_len_smooth:
        win = filters.get_window(smoothing_window, win_len_smooth + 2, fftbins=False)
        win /= np.sum(win)
        win = util.expand_to(win, ndim=chroma_quant.ndim, axes=-1)
        cens = scipy.ndimage.convolve(chroma_quant, win, mode="constant")
    else:
        cens = chroma_quant
    return util.normalize(cens, norm=norm, axis=-2)


def tempogram(*, y=None, sr=22050, onset_envelope=None, hop_length=512, win_length=384, center=True,
              window="hann", norm=np.inf):
    if onset_envelope is None:
        from .onset import onset_strength
        onset_envelope = onset_strength(y=y, sr=sr, hop_length=hop_length)
    ac_window = filters.get_window(window, win_length, fftbins=True)
    n = onset_envelope.shape[-1]
    if center:
        padding = [(0, 0) for _ in onset_envelope.shape]
        padding[-1] = (int(win_length // 2),) * 2
        onset_envelope = np.pad(onset_envelope, padding, mode="linear_ramp", end_values=[0, 0])
    odf_frame = util.frame(onset_envelope, frame_length=win_length, hop_length=1)
    if center:
        odf_frame = odf_frame[..., :n]
    ac_window = util.expand_to(ac_window, ndim=odf_frame.ndim, axes=-2)
    return util.normalize(autocorrelate(odf_frame * ac_window, axis=-2), norm=norm, axis=-2)


def rms(*, y=None, S=None, frame_length=2048, hop_length=512, center=True, pad_mode="constant",
        dtype=np.float32):
    if y is None:
        raise NotImplementedError
    if center:
        y = np.pad(y, int(frame_length // 2), mode=pad_mode)
    x = util.frame(y, frame_length=frame_length, hop_length=hop_length)
    power = np.mean(util.abs2(x, dtype=dtype), axis=-2, keepdims=True)
    return np.sqrt(power)


def zero_crossing_rate(y, *, frame_length=2048, hop_length=512, center=True, threshold=1e-10):
    if center:
        y = np.pad(y, int(frame_length // 2), mode="edge")
    yf = util.frame(y, frame_length=frame_length, hop_length=hop_length)      # [frame, T]
    x = np.where(np.abs(yf) <= threshold, 0.0, yf)
    sb = np.signbit(x)
    crossings = np.zeros(yf.shape, dtype=bool)
    crossings[1:, :] = sb[1:, :] != sb[:-1, :]
    crossings[0, :] = False
    return np.mean(crossings, axis=-2, keepdims=True)


def spectral_centroid(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=512, freq=None, win_length=None,
                      window="hann", center=True, pad_mode="constant"):
    S, n_fft = _spectrogram(y=y, S=S, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window,
                            center=center, pad_mode=pad_mode)
    if freq is None:
        freq = fft_frequencies(sr=sr, n_fft=n_fft)
    if freq.ndim == 1:
        freq = util.expand_to(freq, ndim=S.ndim, axes=-2)
    return np.sum(freq * util.normalize(S, norm=1, axis=-2), axis=-2, keepdims=True)


def spectral_bandwidth(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None, window="hann",
                       center=True, pad_mode="constant", freq=None, centroid=None, norm=True, p=2):
    S, n_fft = _spectrogram(y=y, S=S, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window,
                            center=center, pad_mode=pad_mode)
    if centroid is None:
        centroid = spectral_centroid(y=y, sr=sr, S=S, n_fft=n_fft, hop_length=hop_length, freq=freq)
    if freq is None:
        freq = fft_frequencies(sr=sr, n_fft=n_fft)
    if freq.ndim == 1:
        deviation = np.abs(np.subtract.outer(centroid[..., 0, :], freq).swapaxes(-2, -1))
    else:
        deviation = np.abs(freq - centroid)
    if norm:
        S = util.normalize(S, norm=1, axis=-2)
    return np.sum(S * deviation ** p, axis=-2, keepdims=True) ** (1.0 / p)


def spectral_rolloff(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None, window="hann",
                     center=True, pad_mode="constant", freq=None, roll_percent=0.85):
    S, n_fft = _spectrogram(y=y, S=S, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window,
                            center=center, pad_mode=pad_mode)
    if freq is None:
        freq = fft_frequencies(sr=sr, n_fft=n_fft)
    if freq.ndim == 1:
        freq = util.expand_to(freq, ndim=S.ndim, axes=-2)
    total_energy = np.cumsum(S, axis=-2)
    threshold = roll_percent * total_energy[..., -1, :]
    threshold = np.expand_dims(threshold, axis=-2)
    ind = np.where(total_energy < threshold, np.nan, 1)
    return np.nanmin(ind * freq, axis=-2, keepdims=True)


def spectral_flatness(*, y=None, S=None, n_fft=2048, hop_length=512, win_length=None, window="hann",
                      center=True, pad_mode="constant", amin=1e-10, power=2.0):
    S, n_fft = _spectrogram(y=y, S=S, n_fft=n_fft, hop_length=hop_length, power=1.0, win_length=win_length,
                            window=window, center=center, pad_mode=pad_mode)
    S_thresh = np.maximum(amin, S ** power)
    gmean = np.exp(np.mean(np.log(S_thresh), axis=-2, keepdims=True))
    amean = np.mean(S_thresh, axis=-2, keepdims=True)
    return gmean / amean


def spectral_contrast(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None, window="hann",
                      center=True, pad_mode="constant", freq=None, fmin=200.0, n_bands=6, quantile=0.02,
                      linear=False):
    S, n_fft = _spectrogram(y=y, S=S, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window,
                            center=center, pad_mode=pad_mode)
    if freq is None:
        freq = fft_frequencies(sr=sr, n_fft=n_fft)
    freq = np.atleast_1d(freq)
    octa = np.zeros(n_bands + 2)
    octa[1:] = fmin * (2.0 ** np.arange(0, n_bands + 1))
    if np.any(octa[:-1] >= 0.5 * sr):
        raise ValueError("Frequency band exceeds Nyquist. Reduce either fmin or n_bands.")
    shape = list(S.shape)
    shape[-2] = n_bands + 1
    valley = np.zeros(shape)
    peak = np.zeros_like(valley)
    for k, (f_low, f_high) in enumerate(zip(octa[:-1], octa[1:])):
        current_band = np.logical_and(freq >= f_low, freq <= f_high)
        idx = np.flatnonzero(current_band)
        if k > 0:
            current_band[idx[0] - 1] = True
        if k == n_bands:
            current_band[idx[-1] + 1:] = True
        sub_band = S[..., current_band, :]
        if k < n_bands:
            sub_band = sub_band[..., :-1, :]
        idx = np.rint(quantile * np.sum(current_band))
        idx = int(np.maximum(idx, 1))
        sortedr = np.sort(sub_band, axis=-2)
        valley[..., k, :] = np.mean(sortedr[..., :idx, :], axis=-2)
        peak[..., k, :] = np.mean(sortedr[..., -idx:, :], axis=-2)
    if linear:
        return peak - valley
    return power_to_db(peak) - power_to_db(valley)
