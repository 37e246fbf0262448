"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the `--precompute` hot path of dohyeoplim/breathing-phase-classifier.

PARITY UNPINNED: the reference ships no tests / golden vectors and its numerics live in librosa 0.10.2.post1,
which cannot be installed here.  This file restates `src/precompute/process.py` and `src/precompute/methods.py`
call-for-call on top of `oracle/librosa_shim` (a numpy/scipy restatement of the librosa entry points).  In the build
container `oracle/check_against_reference.py` runs the reference's *own, unmodified* `process_and_save_npz` over the
same shim and requires this restatement to be bit-identical to it; the committed fixtures under `tests/golden/` come
from that run.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import this
module.  The product path never does.

Every function cites the reference lines it follows (paths relative to /root/reference/src/precompute/).
"""
from __future__ import annotations

import os
import sys
import warnings
from dataclasses import dataclass

import numpy as np
import scipy.signal
import scipy.stats
from scipy.signal import find_peaks
from scipy.fftpack import dct

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "librosa_shim")
if _SHIM not in sys.path:
    sys.path.insert(0, _SHIM)
import librosa  # noqa: E402  (the shim)

warnings.filterwarnings("ignore")

CHANNEL_KEYS = ("mel", "mfcc", "chroma", "mel_delta", "mel_delta2", "gammatone", "lpc", "mod_spec", "tempogram")
SORTED_KEYS = tuple(sorted(CHANNEL_KEYS))     # the order dataset.py:26 stacks them in


@dataclass(frozen=True)
class Params:
    """Module constants of the reference: process.py:12-23, methods.py:10-22."""
    sr: int = 16000
    duration: float = 1.0
    n_mels: int = 128
    n_mfcc: int = 40
    hop: int = 256
    n_fft: int = 512
    fmax: float = 4500
    n_gammatone: int = 64
    n_lpc: int = 12

    @property
    def expected_len(self) -> int:
        return int(self.sr * self.duration)

    @property
    def t_fixed(self) -> int:          # process.py:30
        return self.expected_len // self.hop + 1


DEFAULT = Params()


# ------------------------------------------------------------------------------------------ methods.py:24-46
def fit_length(wave, n):
    """methods.py:24-28 pad_or_truncate."""
    if len(wave) >= n:
        return wave[:n]
    return np.concatenate([wave, np.zeros(n - len(wave), dtype=np.float32)])


def fit_time(a, rows, t_fixed):
    """methods.py:30-37 pad_time: truncate, else pad columns with the array minimum."""
    t_raw = a.shape[1]
    if t_raw >= t_fixed:
        return a[:, :t_fixed]
    fill = np.full((rows, t_fixed - t_raw), a.min(), dtype=np.float32)
    return np.concatenate([a, fill], axis=1)


def fit_rows(a, rows_from, rows_to):
    """methods.py:39-46 pad_freq: truncate, else pad rows with the array minimum."""
    if rows_from >= rows_to:
        return a[:rows_to, :]
    fill = np.full((rows_to - rows_from, a.shape[1]), a.min(), dtype=np.float32)
    return np.concatenate([a, fill], axis=0)


def z_all(a):
    """process.py:36-38,60,65,70,76: whole-array z-score."""
    return (a - a.mean()) / (a.std() + 1e-8)


def z_rows(a):
    """process.py:47,55: row-wise z-score."""
    return (a - a.mean(axis=1, keepdims=True)) / (a.std(axis=1, keepdims=True) + 1e-8)


# ------------------------------------------------------------------------------------------ methods.py:48-114
def scalar_features(y, p: Params = DEFAULT, debug=None):
    """methods.py:48-114 extract_enhanced_scalar_features -> float32[36]."""
    sr, hop = p.sr, p.hop
    out = []
    rms = librosa.feature.rms(y=y, hop_length=hop)[0]                                  # :52
    zcr = librosa.feature.zero_crossing_rate(y=y, hop_length=hop)[0]                   # :53
    out += [np.mean(rms), np.std(rms), np.max(rms), np.min(rms),
            np.mean(zcr), np.std(zcr), np.max(zcr), np.min(zcr)]                       # :54-57

    cent = librosa.feature.spectral_centroid(y=y, sr=sr, hop_length=hop)[0]            # :59
    bw = librosa.feature.spectral_bandwidth(y=y, sr=sr, hop_length=hop)[0]             # :60
    roll = librosa.feature.spectral_rolloff(y=y, sr=sr, roll_percent=0.85)[0]          # :61 (hop 512!)
    flat = librosa.feature.spectral_flatness(y=y, hop_length=hop)[0]                   # :62
    contrast = librosa.feature.spectral_contrast(y=y, sr=sr, hop_length=hop)           # :63
    nyq = sr / 2
    out += [np.mean(cent) / nyq, np.std(cent) / nyq, scipy.stats.skew(cent),
            np.mean(bw) / nyq, np.std(bw) / nyq,
            np.mean(roll) / nyq, np.std(roll) / nyq,
            np.mean(flat), np.std(flat),
            np.mean(contrast), np.std(contrast)]                                       # :64-70

    env = np.abs(scipy.signal.hilbert(y))                                              # :72
    env_mean = np.mean(env)
    env_std = np.std(env)
    env_snr = env_mean / (env_std + 1e-8)
    peaks, props = find_peaks(env, height=env_mean, distance=sr // 10)                 # :76
    n_peaks = len(peaks)
    heights = props["peak_heights"] if n_peaks > 0 else [0]
    out += [env_mean, env_std, env_snr, n_peaks, np.mean(heights),
            np.std(heights) if n_peaks > 1 else 0]                                     # :79-82

    mag = np.abs(librosa.stft(y, n_fft=p.n_fft, hop_length=hop))                       # :84
    low_bins = int(1000 * p.n_fft / sr)
    low = np.sum(mag[:low_bins, :] ** 2)
    tot = np.sum(mag ** 2)
    low_ratio = low / (tot + 1e-8)                                                     # :85-88

    mel = librosa.feature.melspectrogram(y=y, sr=sr, n_mels=p.n_mels, hop_length=hop)  # :90 (n_fft 2048)
    mel_db = librosa.power_to_db(mel, ref=np.max)
    flux = np.sqrt(np.sum(np.diff(mel_db, axis=1) ** 2, axis=0))                       # :92
    out += [low_ratio, np.mean(flux), np.std(flux), np.max(flux)]

    out += [scipy.stats.skew(y), scipy.stats.kurtosis(y),
            np.percentile(np.abs(y), 90), np.percentile(np.abs(y), 10)]                # :98-103

    ac = np.correlate(y, y, mode="full")[len(y) - 1:]                                  # :105
    ac = ac / ac[0]
    first_min = np.argmin(ac[: sr // 20]) if len(ac) > sr // 20 else len(ac) // 2
    out += [ac[sr // 100] if len(ac) > sr // 100 else 0,
            ac[sr // 50] if len(ac) > sr // 50 else 0,
            first_min / sr]                                                            # :108-112
    if debug is not None:
        debug.update(rms=rms, zcr=zcr, centroid=cent, bandwidth=bw, rolloff=roll, flatness=flat,
                     contrast=contrast, envelope=env, peaks=peaks, flux=flux, mel2048_db=mel_db,
                     first_min_idx=int(first_min), n_peaks=int(n_peaks), autocorr=ac[: sr // 20].copy())
    return np.array(out, dtype=np.float32)                                             # :114


# ----------------------------------------------------------------------------------------- methods.py:116-143
def lpc_frames(y, p: Params = DEFAULT):
    """methods.py:116-134 extract_lpc_features -> float32 [order, n_frames]."""
    order = p.n_lpc
    emph = np.append(y[0], y[1:] - 0.97 * y[:-1])
    flen = int(0.025 * p.sr)
    fshift = int(0.010 * p.sr)
    rows = []
    for start in range(0, len(emph) - flen, fshift):
        fr = emph[start:start + flen] * np.hamming(flen)
        try:
            rows.append(librosa.lpc(fr, order=order)[1:])
        except Exception:
            rows.append(np.zeros(order))
    if not rows:
        return np.zeros((order, 1), dtype=np.float32)
    return np.array(rows, dtype=np.float32).T


def gammatone_frames(y, p: Params = DEFAULT):
    """methods.py:136-140: the 'gammatone' channel is log1p(mel64 @ |STFT512|)."""
    bank = librosa.filters.mel(sr=p.sr, n_fft=p.n_fft, n_mels=p.n_gammatone)
    mag = np.abs(librosa.stft(y, n_fft=p.n_fft, hop_length=p.hop))
    return np.log1p(np.dot(bank, mag))


def modulation_frames(mel_db):
    """methods.py:142-143: 2-D DCT (mel axis, keep 40; then time axis)."""
    return dct(dct(mel_db, axis=0, norm="ortho")[:40, :], axis=1, norm="ortho")


# ------------------------------------------------------------------------------------------ process.py:25-103
def segment_features(y, p: Params = DEFAULT, debug=None):
    """process.py:29-103 for one waveform already loaded as float32.

    Returns (channels: dict key -> float32 [128, t_fixed], scalars float32 [36]).  If `debug` is a dict it receives
    the un-normalised intermediates the GPU parity tests compare (dB spectra, tuning, onset envelope ...).
    """
    y = fit_length(np.asarray(y, dtype=np.float32), p.expected_len)                     # :29
    T = p.t_fixed
    H = p.n_mels
    sr, hop, n_fft = p.sr, p.hop, p.n_fft

    mel_pow = librosa.feature.melspectrogram(y=y, sr=sr, n_fft=n_fft, hop_length=hop, n_mels=H, fmax=p.fmax)
    mel_db = librosa.power_to_db(mel_pow, ref=np.max)                                   # :33
    d1 = librosa.feature.delta(mel_db, order=1)
    d2 = librosa.feature.delta(mel_db, order=2)
    ch = {}
    ch["mel"] = fit_time(z_all(mel_db), H, T)                                           # :36,39
    ch["mel_delta"] = fit_time(z_all(d1), H, T)
    ch["mel_delta2"] = fit_time(z_all(d2), H, T)

    mf = librosa.feature.mfcc(y=y, sr=sr, n_mfcc=p.n_mfcc, hop_length=hop, n_fft=n_fft)  # :43
    mf_all = np.vstack([mf, librosa.feature.delta(mf, order=1), librosa.feature.delta(mf, order=2)])
    ch["mfcc"] = fit_rows(fit_time(z_rows(mf_all), mf_all.shape[0], T), mf_all.shape[0], H)

    mag = np.abs(librosa.stft(y, n_fft=n_fft, hop_length=hop))                          # :51
    c_stft = librosa.feature.chroma_stft(S=mag, sr=sr, hop_length=hop)                  # :52
    c_cens = librosa.feature.chroma_cens(y=y, sr=sr, hop_length=hop)                    # :53
    c_all = np.vstack([c_stft, c_cens])
    ch["chroma"] = fit_rows(fit_time(z_rows(c_all), 24, T), 24, H)

    gam = gammatone_frames(y, p)                                                        # :59
    ch["gammatone"] = fit_rows(fit_time(z_all(gam), p.n_gammatone, T), p.n_gammatone, H)

    lp = lpc_frames(y, p)                                                               # :64
    ch["lpc"] = fit_rows(fit_time(z_all(lp), p.n_lpc, T), p.n_lpc, H)

    mod = modulation_frames(mel_db)                                                     # :69
    ch["mod_spec"] = fit_rows(fit_time(z_all(mod), 40, T), 40, H)

    onset = librosa.onset.onset_strength(y=y, sr=sr, hop_length=hop)                    # :74
    tg = librosa.feature.tempogram(onset_envelope=onset, sr=sr, hop_length=hop)         # :75
    ch["tempogram"] = fit_rows(fit_time(z_all(tg), tg.shape[0], T), tg.shape[0], H)

    scal = scalar_features(y, p, debug=debug)                                           # :80
    ch = {k: v.astype(np.float32) for k, v in ch.items()}                               # :82-90

    if debug is not None:
        with np.errstate(divide="ignore"):
            debug.update(
                y=y, stft512_mag=mag, mel_db=mel_db, mel_delta_raw=d1, mel_delta2_raw=d2, mfcc_raw=mf_all,
                chroma_stft_raw=c_stft, chroma_cens_raw=c_cens, gammatone_raw=gam, lpc_raw=lp, mod_spec_raw=mod,
                onset_env=onset, tempogram_raw=tg,
                stft512_db=librosa.power_to_db(mag ** 2, ref=np.max),
                tuning12=float(librosa.estimate_tuning(S=mag, sr=sr, bins_per_octave=12)),
                tuning36=float(librosa.estimate_tuning(y=y, sr=sr, bins_per_octave=36)),
            )
    return ch, scal


def logmel_stage(y, p: Params = DEFAULT):
    """BASELINE config 2: log-power STFT [257,T] (power_to_db(|X|^2, ref=max)) + mel / mel_delta / mel_delta2."""
    dbg = {}
    ch, _ = segment_features(y, p, debug=dbg)
    return dbg["stft512_db"], np.stack([ch["mel"], ch["mel_delta"], ch["mel_delta2"]])


def stack_sorted(ch):
    """dataset.py:25-26,48: channels stacked in sorted-key order -> float32 [9, 128, T]."""
    return np.stack([ch[k] for k in SORTED_KEYS], axis=0).astype(np.float32)


def save_segment_npz(path, ch, scal):
    """process.py:92-103: the on-disk contract."""
    np.savez(path, mel=ch["mel"], mfcc=ch["mfcc"], chroma=ch["chroma"], mel_delta=ch["mel_delta"],
             mel_delta2=ch["mel_delta2"], gammatone=ch["gammatone"], lpc=ch["lpc"], mod_spec=ch["mod_spec"],
             tempogram=ch["tempogram"], scalars=scal)


def process_wav(file_id, wav_path, target_dir, p: Params = DEFAULT):
    """process.py:25-108 including the never-raise error convention."""
    try:
        y, _ = librosa.load(wav_path, sr=p.sr)
        ch, scal = segment_features(y, p)
        save_segment_npz(os.path.join(target_dir, file_id + ".npz"), ch, scal)
        return file_id, True, None
    except Exception as e:  # noqa: BLE001  (reference catches everything, process.py:107)
        return file_id, False, str(e)


# --------------------------------------------------------------------------------- SURVEY 8(d) synthetic input
def synth_segment(i, length=16000, sr=16000, seed=20250101):
    """Synthetic 'breathing-like' segment i: low-passed noise burst under a breath envelope, quantised to PCM16.

    Deterministic in (seed, i).  Returns float32 [length] that is exactly int16 / 32768 (the wav path)."""
    rng = np.random.default_rng([seed, int(i)])
    dur = length / sr
    t = np.arange(length) / sr
    cutoff = float(np.exp(rng.uniform(np.log(120.0), np.log(600.0))))
    b, a = scipy.signal.butter(2, cutoff / (sr / 2))
    x = scipy.signal.lfilter(b, a, rng.standard_normal(length + 2000))[2000:]
    x = x / (np.std(x) + 1e-12)
    x = x + 10 ** (-50 / 20) * rng.standard_normal(length)
    w = rng.uniform(0.4, 1.0) * dur
    t0 = rng.uniform(0.0, dur - w)
    env = np.full(length, 0.15)
    inside = (t >= t0) & (t <= t0 + w)
    env[inside] = 0.15 + 0.85 * np.sin(np.pi * (t[inside] - t0) / w) ** 2
    x = x * env
    target_rms = float(np.exp(rng.normal(np.log(0.02), 0.8)))
    x = x * (target_rms / (np.sqrt(np.mean(x ** 2)) + 1e-12))
    peak = np.max(np.abs(x))
    if peak >= 0.95:
        x = x * (0.95 / peak)
    q = np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)
    return (q.astype(np.float32) / np.float32(32768.0)).astype(np.float32)


def synth_batch(start, count, length=16000, sr=16000, seed=20250101):
    return np.stack([synth_segment(start + k, length, sr, seed) for k in range(count)])
