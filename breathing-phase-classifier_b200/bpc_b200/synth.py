"""Synthetic 'breathing-like' segments: low-passed noise burst under a breath envelope, quantised to
PCM16 so the data follows the wav path exactly.  Host-side data generation for bench.py / examples; deterministic in
(seed, index).  `tests/test_host.py` pins it against the oracle's own copy of the generator."""
from __future__ import annotations

import numpy as np
import scipy.signal


def synth_pcm16(i: int, length: int = 16000, sr: int = 16000, seed: int = 20250101) -> np.ndarray:
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
    return np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)


def synth_batch_pcm16(start: int, count: int, length: int = 16000) -> np.ndarray:
    return np.stack([synth_pcm16(start + k, length) for k in range(count)])
