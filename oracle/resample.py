"""CPU oracle (test infrastructure, never on the product path) of the sample-rate conversion on load.

Reference: src/precompute/process.py:28 `y, _ = librosa.load(wav_path, sr=SR)`, i.e. librosa.resample(y, orig_sr,
target_sr, res_type="soxr_hq") followed by util.fix_length(ceil(n * target_sr / orig_sr)) for every file whose native
rate is not 16 kHz.  **libsoxr 0.1.3 (env.yaml:256-257) is an un-vendored C dependency that cannot be restated bit for
bit** ("parity unpinned" for this one stage, exactly like the CQT's soxr decimator): what is
restated is soxr HQ's published contract -- a linear-phase FIR, pass band flat up to 0.9125 of the lower Nyquist rate,
stop band from that Nyquist on, rejection far beyond 16-bit PCM -- as ONE explicit filter that oracle and device share:

    p / q = sr_out / sr_in in lowest terms, r = min(1, p / q)
    fc    = 0.5 * 0.95625 * r                (cut-off = centre of the transition band, cycles per INPUT sample)
    delta = 0.5 * 0.0875 * r                 (transition width), A = 150 dB, beta = 0.1102 (A - 8.7)
    N     = ceil((A - 7.95) / (14.36 delta)) (Kaiser's length estimate), half = (N + 1) // 2
    h(u)  = 2 fc sinc(2 fc u) I0(beta sqrt(1 - (u / half)^2)) / I0(beta),  |u| <= half
    out[m] = sum_k x[k] h(m q / p - k),  m < ceil(n_in p / q),  x = 0 outside the signal,

evaluated as a polyphase filter (phase f = (m q) mod p, every phase's 2 half coefficients scaled to unit DC gain) with
float64 accumulation and one rounding to float32.  The device kernel is csrc/k_resample.cu, its table
csrc/tables.cpp::resample_filter (pinned against `polyphase_table` by tests/test_host.py).
"""
from __future__ import annotations

import math

import numpy as np


def polyphase_table(sr_in: int, sr_out: int):
    """-> (p, q, half, tab[p, 2 half] float64) with tab[f][j] = h(f / p + half - 1 - j) / sum_j(...)."""
    g = math.gcd(int(sr_in), int(sr_out))
    p, q = int(sr_out) // g, int(sr_in) // g
    r = min(1.0, p / q)
    fc = 0.5 * 0.95625 * r
    delta = 0.5 * 0.0875 * r
    atten = 150.0
    beta = 0.1102 * (atten - 8.7)
    ntaps = int(math.ceil((atten - 7.95) / (14.36 * delta)))
    half = (ntaps + 1) // 2
    f = np.arange(p, dtype=np.float64)[:, None] / p
    j = np.arange(2 * half, dtype=np.float64)[None, :]
    u = f + (half - 1) - j
    xr = u / half
    w = np.where(np.abs(xr) <= 1.0, np.i0(beta * np.sqrt(np.maximum(0.0, 1.0 - xr * xr))) / np.i0(beta), 0.0)
    tab = 2.0 * fc * np.sinc(2.0 * fc * u) * w
    tab /= tab.sum(axis=1, keepdims=True)
    return p, q, half, tab


def resample(y: np.ndarray, sr_in: int, sr_out: int) -> np.ndarray:
    """float32 [n_in] -> float32 [ceil(n_in * sr_out / sr_in)]."""
    y = np.asarray(y, dtype=np.float32)
    if sr_in == sr_out:
        return y.copy()
    p, q, half, tab = polyphase_table(sr_in, sr_out)
    n_in = len(y)
    n_out = -((-n_in * sr_out) // sr_in)
    m = np.arange(n_out, dtype=np.int64)
    t = m * q
    k0 = t // p - half + 1
    ph = t - (t // p) * p
    lo = int(min(0, k0.min())) if n_out else 0
    hi = int(max(n_in, (k0.max() + 2 * half) if n_out else 0))
    xp = np.zeros(hi - lo, dtype=np.float64)
    xp[-lo:-lo + n_in] = y
    out = np.empty(n_out, dtype=np.float64)
    win = np.lib.stride_tricks.sliding_window_view(xp, 2 * half)
    for f in range(p):
        sel = np.flatnonzero(ph == f)
        if len(sel):
            out[sel] = win[k0[sel] - lo] @ tab[f]
    return out.astype(np.float32)
