"""librosa core subset: load, stft, power_to_db, tuning, cqt, lpc (test infrastructure; see package docstring).

Restated from the published librosa 0.10.2 algorithms.  The one piece that cannot be restated bit-for-bit is
`resample(..., res_type="soxr_hq")` (libsoxr is a C library that is absent here): it is replaced by a
zero-phase, zero-extended linear-phase half-band FIR decimator -- see `default_halfband()`.
"""
from __future__ import annotations

import warnings
import numpy as np
import scipy.signal
import scipy.io.wavfile

from . import util
from . import filters
from .filters import get_window, _hz_to_mel as hz_to_mel, _mel_to_hz as mel_to_hz, _mel_frequencies, _hz_to_octs


# ----------------------------------------------------------------------------------------------- loading
def load(path, *, sr=22050, mono=True, offset=0.0, duration=None, dtype=np.float32, res_type="soxr_hq"):
    """librosa.load for PCM wav files (soundfile semantics: int16 / 32768 -> float32)."""
    native_sr, data = scipy.io.wavfile.read(path)
    if data.dtype == np.int16:
        y = data.astype(np.float32) / np.float32(32768.0)
    elif data.dtype == np.int32:
        y = (data.astype(np.float64) / 2147483648.0).astype(np.float32)
    elif data.dtype == np.uint8:
        y = ((data.astype(np.float32) - 128.0) / 128.0).astype(np.float32)
    else:
        y = data.astype(np.float32)
    if y.ndim > 1:
        y = y.T
        if mono:
            y = np.mean(y, axis=0)
    if sr is not None and native_sr != sr:
        # libsoxr "HQ" is not available: the shared Kaiser stand-in of oracle/resample.py (parity unpinned for this stage)
        from oracle.resample import resample as _resample_on_load
        y = _resample_on_load(np.asarray(y, dtype=np.float32), int(native_sr), int(sr))
    return np.asarray(y, dtype=dtype), (sr if sr is not None else native_sr)


# ------------------------------------------------------------------------------------------------- STFT
def fft_frequencies(*, sr=22050, n_fft=2048):
    return np.fft.rfftfreq(n=n_fft, d=1.0 / sr)


def mel_frequencies(n_mels=128, *, fmin=0.0, fmax=11025.0, htk=False):
    return _mel_frequencies(n_mels, fmin=fmin, fmax=fmax, htk=htk)


def hz_to_octs(frequencies, *, tuning=0.0, bins_per_octave=12):
    return _hz_to_octs(frequencies, tuning=tuning, bins_per_octave=bins_per_octave)


def hz_to_midi(frequencies):
    return 12 * (np.log2(np.asanyarray(frequencies)) - np.log2(440.0)) + 69


def note_to_hz(note):
    if note != "C1":
        raise NotImplementedError(note)
    return 440.0 * 2.0 ** ((24 - 69) / 12.0)


def stft(y, *, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True, dtype=None,
         pad_mode="constant", out=None):
    """float32 signal -> window (float64) * frames -> float64 rfft -> complex64, [1 + n_fft//2, n_frames]."""
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    fft_window = get_window(window, win_length, fftbins=True)
    fft_window = util.pad_center(fft_window, size=n_fft)
    y = np.asarray(y)
    if center:
        if pad_mode != "constant":
            raise NotImplementedError(pad_mode)
        y = np.pad(y, int(n_fft // 2), mode="constant")
    if dtype is None:
        dtype = util.dtype_r2c(y.dtype)
    y_frames = util.frame(y, frame_length=n_fft, hop_length=hop_length)          # [n_fft, T]
    spec = np.fft.rfft(fft_window[:, None] * y_frames, axis=0)                    # float64 arithmetic
    return spec.astype(dtype)


def _spectrogram(*, y=None, S=None, n_fft=2048, hop_length=512, power=1, win_length=None, window="hann",
                 center=True, pad_mode="constant"):
    if S is not None:
        if n_fft is None or n_fft // 2 + 1 != S.shape[-2]:
            n_fft = 2 * (S.shape[-2] - 1)
    else:
        S = np.abs(stft(y, n_fft=n_fft, hop_length=hop_length, win_length=win_length, center=center,
                        window=window, pad_mode=pad_mode)) ** power
    return S, n_fft


def power_to_db(S, *, ref=1.0, amin=1e-10, top_db=80.0):
    S = np.asarray(S)
    magnitude = np.abs(S) if np.issubdtype(S.dtype, np.complexfloating) else S
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


# ----------------------------------------------------------------------------------------------- tuning
def _parabolic_interpolation(x, *, axis=-2):
    """numba stencil in librosa: `2 * x[0]` and `/ 2` promote float32 operands to float64."""
    xi = np.swapaxes(x, -1, axis)
    shifts = np.zeros(x.shape, dtype=x.dtype)
    shiftsi = np.swapaxes(shifts, -1, axis)
    c = xi[..., 1:-1]
    up = xi[..., 2:]
    dn = xi[..., :-2]
    a = (up + dn).astype(np.float64) - 2.0 * c.astype(np.float64)
    b = (up - dn).astype(np.float64) / 2.0
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.where(np.abs(b) >= np.abs(a), 0.0, -b / a)
    shiftsi[..., 1:-1] = s
    return shifts


def piptrack(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=None, fmin=150.0, fmax=4000.0,
             threshold=0.1, win_length=None, window="hann", center=True, pad_mode="constant", ref=None):
    S, n_fft = _spectrogram(y=y, S=S, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                            window=window, center=center, pad_mode=pad_mode)
    S = np.abs(S)
    fmin = np.maximum(fmin, 0)
    fmax = np.minimum(fmax, float(sr) / 2)
    fft_freqs = fft_frequencies(sr=sr, n_fft=n_fft)
    avg = np.gradient(S, axis=-2)
    shift = _parabolic_interpolation(S, axis=-2)
    dskew = 0.5 * avg * shift
    pitches = np.zeros_like(S)
    mags = np.zeros_like(S)
    freq_mask = (fmin <= fft_freqs) & (fft_freqs < fmax)
    freq_mask = util.expand_to(freq_mask, ndim=S.ndim, axes=-2)
    if ref is None:
        ref = np.max
    if callable(ref):
        ref_value = threshold * ref(S, axis=-2)
        ref_value = np.expand_dims(ref_value, -2)
    else:
        ref_value = np.abs(ref)
    idx = np.nonzero(freq_mask & util.localmax(S * (S > ref_value), axis=-2))
    pitches[idx] = (idx[-2] + shift[idx]) * float(sr) / n_fft
    mags[idx] = S[idx] + dskew[idx]
    return pitches, mags


def pitch_tuning(frequencies, *, resolution=0.01, bins_per_octave=12):
    frequencies = np.atleast_1d(frequencies)
    frequencies = frequencies[frequencies > 0]
    if not np.any(frequencies):
        warnings.warn("Trying to estimate tuning from empty frequency set.", stacklevel=2)
        return 0.0
    residual = np.mod(bins_per_octave * hz_to_octs(frequencies), 1.0)
    residual[residual >= 0.5] -= 1.0
    bins = np.linspace(-0.5, 0.5, int(np.ceil(1.0 / resolution)) + 1)
    counts, tuning = np.histogram(residual, bins)
    return tuning[np.argmax(counts)]


def estimate_tuning(*, y=None, sr=22050, S=None, n_fft=2048, resolution=0.01, bins_per_octave=12, **kwargs):
    pitch, mag = piptrack(y=y, sr=sr, S=S, n_fft=n_fft, **kwargs)
    pitch_mask = pitch > 0
    threshold = np.median(mag[pitch_mask]) if pitch_mask.any() else 0.0
    return pitch_tuning(pitch[(mag >= threshold) & pitch_mask], resolution=resolution,
                        bins_per_octave=bins_per_octave)


# ------------------------------------------------------------------------------------- autocorrelation
def autocorrelate(y, *, max_size=None, axis=-1):
    if max_size is None:
        max_size = y.shape[axis]
    max_size = int(min(max_size, y.shape[axis]))
    n_pad = 2 * y.shape[axis] - 1
    powspec = util.abs2(np.fft.rfft(y, n=n_pad, axis=axis))
    autocorr = np.fft.irfft(powspec, n=n_pad, axis=axis)
    subslice = [slice(None)] * autocorr.ndim
    subslice[axis] = slice(max_size)
    return autocorr[tuple(subslice)]


# ------------------------------------------------------------------------ soxr_hq stand-in (decimate by 2)
_HALFBAND = None


def default_halfband(numtaps=127, passband=0.60, atten_db=150.0):
    """Linear-phase half-band low-pass for 2:1 decimation.

    soxr 'HQ' is a linear-phase FIR with a pass-band flat to ~1e-6 up to 0.913 of the new Nyquist and ~120 dB
    rejection; the CQT only consumes content below ~0.55 of each stage's new Nyquist (every octave's filters sit
    at 0.26-0.52 of it and the sparsified bases have compact support), so any linear-phase decimator that is flat
    there and rejects the band that aliases onto it is interchangeable with soxr to ~1e-6.  This Kaiser design
    (cutoff at the new Nyquist) is flat to <1e-7 below `passband` and >`atten_db` dB down above 2-`passband`
    (both relative to the new Nyquist).  Coefficients are float64; DC gain is normalised to exactly 1.
    """
    width = 2.0 * (1.0 - passband) / 2.0      # transition width as a fraction of the *old* Nyquist
    beta = scipy.signal.kaiser_beta(atten_db)
    taps = scipy.signal.firwin(numtaps, 0.5, window=("kaiser", beta), pass_zero=True, fs=2.0)
    del width
    return taps / np.sum(taps)


def set_halfband(taps):
    global _HALFBAND
    _HALFBAND = None if taps is None else np.asarray(taps, dtype=np.float64)


def get_halfband():
    global _HALFBAND
    if _HALFBAND is None:
        _HALFBAND = default_halfband()
    return _HALFBAND


def resample(y, *, orig_sr, target_sr, res_type="soxr_hq", fix=True, scale=False, axis=-1):
    """Only the 2:1 decimation the CQT octave recursion uses.  out[n] = sum_k h[k] * y[2n + c - k], zero extension,
    c = (numtaps-1)/2 (zero delay), n_out = ceil(n_in / 2); `scale=True` divides by sqrt(ratio)."""
    if not (orig_sr == 2 and target_sr == 1 and y.ndim == 1):
        raise NotImplementedError("oracle shim: only the CQT's 2:1 decimation is restated")
    h = get_halfband()
    c = (len(h) - 1) // 2
    n_out = int(np.ceil(y.shape[-1] * 0.5))
    full = np.convolve(y.astype(np.float64), h)          # full[m] = sum_k h[k] y[m-k]
    idx = 2 * np.arange(n_out) + c
    y_hat = full[idx]
    if scale:
        y_hat = y_hat / np.sqrt(0.5)
    return np.asarray(y_hat, dtype=y.dtype)


# ------------------------------------------------------------------------------------------------- CQT
def _vqt_filter_fft(sr, freqs, filter_scale, norm, sparsity, *, window="hann", gamma=0.0,
                    dtype=np.complex64, alpha=None):
    basis, lengths = filters.wavelet(freqs=freqs, sr=sr, filter_scale=filter_scale, norm=norm, pad_fft=True,
                                     window=window, gamma=gamma, alpha=alpha)
    n_fft = basis.shape[1]
    basis *= lengths[:, np.newaxis] / float(n_fft)
    fft_basis = np.fft.fft(basis, n=n_fft, axis=1)[:, : (n_fft // 2) + 1]
    fft_basis = util.sparsify_rows(fft_basis, quantile=sparsity, dtype=dtype)
    return fft_basis, n_fft, lengths


def _cqt_response(y, n_fft, hop_length, fft_basis, mode, *, dtype=None):
    D = stft(y, n_fft=n_fft, hop_length=hop_length, window="ones", pad_mode=mode, dtype=dtype)
    return np.asarray(fft_basis.dot(D), dtype=D.dtype)


def vqt(y, *, sr=22050, hop_length=512, fmin=None, n_bins=84, intervals="equal", gamma=None,
        bins_per_octave=12, tuning=0.0, filter_scale=1, norm=1, sparsity=0.01, window="hann", scale=True,
        pad_mode="constant", res_type="soxr_hq", dtype=None):
    n_octaves = int(np.ceil(float(n_bins) / bins_per_octave))
    n_filters = min(bins_per_octave, n_bins)
    if fmin is None:
        fmin = note_to_hz("C1")
    if tuning is None:
        tuning = estimate_tuning(y=y, sr=sr, bins_per_octave=bins_per_octave)
    if dtype is None:
        dtype = util.dtype_r2c(y.dtype)
    fmin = fmin * 2.0 ** (tuning / bins_per_octave)
    # interval_frequencies(intervals="equal", sort=True)
    ratios = 2.0 ** (np.arange(0, bins_per_octave, dtype=float) / bins_per_octave)
    all_ratios = np.multiply.outer(2.0 ** np.arange(n_octaves, dtype=float), ratios).flatten()[:n_bins]
    freqs = np.sort(all_ratios) * fmin
    alpha = filters.relative_bandwidth(freqs=freqs)
    lengths, filter_cutoff = filters.wavelet_lengths(freqs=freqs, sr=sr, window=window,
                                                     filter_scale=filter_scale, gamma=gamma, alpha=alpha)
    nyquist = sr / 2.0
    if filter_cutoff > nyquist:
        raise ValueError("Wavelet basis with max frequency would exceed the Nyquist frequency")
    # __early_downsample_count
    count1 = max(0, int(np.ceil(np.log2(nyquist / filter_cutoff)) - 1) - 1)
    num_twos = 0
    h = hop_length
    while h % 2 == 0 and h > 0:
        num_twos += 1
        h //= 2
    count2 = max(0, num_twos - n_octaves + 1)
    if min(count1, count2) > 0:
        raise NotImplementedError("oracle shim: early down-sampling never triggers at sr=16000")
    if num_twos < n_octaves - 1:
        raise ValueError("hop_length must be a positive integer multiple of 2^(n_octaves-1)")

    vqt_resp = []
    my_y, my_sr, my_hop = y, sr, hop_length
    for i in range(n_octaves):
        sl = slice(-n_filters, None) if i == 0 else slice(-n_filters * (i + 1), -n_filters * i)
        fft_basis, n_fft, _ = _vqt_filter_fft(my_sr, freqs[sl], filter_scale, norm, sparsity, window=window,
                                              gamma=gamma, dtype=dtype, alpha=alpha[sl])
        fft_basis = fft_basis * np.sqrt(sr / my_sr)
        vqt_resp.append(_cqt_response(my_y, n_fft, my_hop, fft_basis.astype(dtype), pad_mode, dtype=dtype))
        if my_hop % 2 == 0:
            my_hop //= 2
            my_sr /= 2.0
            my_y = resample(my_y, orig_sr=2, target_sr=1, res_type=res_type, scale=True)
    # __trim_stack
    max_col = min(c.shape[-1] for c in vqt_resp)
    V = np.empty((n_bins, max_col), dtype=dtype, order="F")
    end = n_bins
    for c in vqt_resp:
        n_oct = c.shape[-2]
        if end < n_oct:
            V[:end, :] = c[-end:, :max_col]
        else:
            V[end - n_oct: end, :] = c[:, :max_col]
        end -= n_oct
    if scale:
        lengths, _ = filters.wavelet_lengths(freqs=freqs, sr=sr, window=window, filter_scale=filter_scale,
                                             gamma=gamma, alpha=alpha)
        V /= np.sqrt(lengths)[:, None]
    return V


def cqt(y, *, sr=22050, hop_length=512, fmin=None, n_bins=84, bins_per_octave=12, tuning=0.0, filter_scale=1,
        norm=1, sparsity=0.01, window="hann", scale=True, pad_mode="constant", res_type="soxr_hq", dtype=None):
    return vqt(y=y, sr=sr, hop_length=hop_length, fmin=fmin, n_bins=n_bins, intervals="equal", gamma=0,
               bins_per_octave=bins_per_octave, tuning=tuning, filter_scale=filter_scale, norm=norm,
               sparsity=sparsity, window=window, scale=scale, pad_mode=pad_mode, res_type=res_type, dtype=dtype)


# ------------------------------------------------------------------------------------------------- LPC
def lpc(y, *, order, axis=-1):
    """Burg's method as in librosa.core.audio.__lpc (Marple 1980, section III), dtype of y."""
    y = np.asarray(y)
    if y.ndim != 1:
        raise NotImplementedError
    dtype = y.dtype
    ar_coeffs = np.zeros(order + 1, dtype=dtype)
    ar_coeffs[0] = 1
    ar_coeffs_prev = ar_coeffs.copy()
    epsilon = util.tiny(np.zeros(1, dtype=dtype))
    fwd = y[1:]
    bwd = y[:-1]
    den = np.sum(fwd ** 2 + bwd ** 2, axis=0)
    for i in range(order):
        k = np.sum(bwd * fwd, axis=0)
        k = k * -2
        k = k / (den + epsilon)
        ar_coeffs_prev, ar_coeffs = ar_coeffs, ar_coeffs_prev
        for j in range(1, i + 2):
            ar_coeffs[j] = ar_coeffs_prev[j] + k * ar_coeffs_prev[i - j + 1]
        fwd_tmp = fwd
        fwd = fwd + k * bwd
        bwd = bwd + k * fwd_tmp
        q = 1.0 - k ** 2
        den = q * den - bwd[-1] ** 2 - fwd[0] ** 2
        fwd = fwd[1:]
        bwd = bwd[:-1]
    return ar_coeffs
