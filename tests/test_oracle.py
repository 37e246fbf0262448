"""CPU tests of the oracle (test infrastructure): shim pieces against independent implementations, the pipeline
restatement against the committed golden vectors, and (build container only) against the reference's own code."""
import os
import sys

import numpy as np
import pytest

from oracle import pipeline as P
import librosa                                    # the shim (oracle/librosa_shim), put on sys.path by oracle.pipeline
from librosa import filters as F, _core as C


def test_shim_is_the_shim():
    assert librosa.__version__.endswith("shim")


@pytest.mark.parametrize("n_fft,n_mels,fmax", [(512, 128, 4500.0), (512, 128, 8000.0), (512, 64, 8000.0),
                                                (2048, 128, 8000.0)])
def test_mel_bank_vs_torchaudio(n_fft, n_mels, fmax):
    import torchaudio
    ours = F.mel(sr=16000, n_fft=n_fft, n_mels=n_mels, fmax=fmax)
    ta = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, 0.0, fmax, n_mels, 16000, norm="slaney",
                                               mel_scale="slaney").T.numpy()
    assert ours.dtype == np.float32 and ours.shape == (n_mels, n_fft // 2 + 1)
    assert np.abs(ours - ta).max() < 5e-7


def test_mel_frequencies_known_values():
    # librosa docstring: mel_frequencies(n_mels=40) -> 0., 85.317, 170.635, ...
    f = librosa.mel_frequencies(40, fmin=0.0, fmax=11025.0)
    assert np.allclose(f[:3], [0.0, 85.317, 170.635], atol=1e-3)


def test_dct_vs_torchaudio():
    import scipy.fftpack
    import torchaudio
    x = np.random.default_rng(0).standard_normal((128, 7)).astype(np.float32)
    a = scipy.fftpack.dct(x, axis=0, type=2, norm="ortho")[:40]
    d = torchaudio.functional.create_dct(40, 128, "ortho").numpy()          # [128, 40]
    assert np.abs(a - d.T @ x).max() < 1e-4


def test_delta_edge_rule():
    # savgol(mode='interp') with polyorder == deriv: the 4 edge values equal the first / last interior value
    x = np.random.default_rng(1).standard_normal((5, 63)).astype(np.float32)
    for order in (1, 2):
        d = librosa.feature.delta(x, order=order)
        assert d.dtype == np.float32
        assert np.abs(d[:, :4] - d[:, 4:5]).max() < 1e-5 and np.abs(d[:, -4:] - d[:, -5:-4]).max() < 1e-5
    c1 = np.arange(-4, 5) / 60.0
    ref = np.array([np.dot(c1, x[0, t - 4:t + 5]) for t in range(4, 59)])
    assert np.abs(librosa.feature.delta(x, order=1)[0, 4:59] - ref).max() < 1e-6


def test_power_to_db_semantics():
    S = np.array([[1e-12, 1.0, 100.0]], dtype=np.float32)
    out = librosa.power_to_db(S, ref=np.max)
    assert np.allclose(out, [[-80.0, -20.0, 0.0]], atol=1e-5)
    out = librosa.power_to_db(S, ref=1.0)
    assert np.allclose(out, [[-60.0, 0.0, 20.0]], atol=1e-5)


def test_stft_is_float64_fft_rounded_to_complex64():
    y = P.synth_segment(3)
    S = librosa.stft(y, n_fft=512, hop_length=256)
    assert S.dtype == np.complex64 and S.shape == (257, 63)
    yp = np.pad(y.astype(np.float64), 256)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(512) / 512)
    t = 17
    ref = np.fft.rfft(yp[t * 256:t * 256 + 512] * w)
    assert np.abs(S[:, t] - ref.astype(np.complex64)).max() <= 1e-6 * np.abs(ref).max()


def test_halfband_is_flat_and_rejects():
    import scipy.signal
    h = C.default_halfband()
    assert len(h) == 127 and abs(h.sum() - 1.0) < 1e-12
    w, H = scipy.signal.freqz(h, worN=8192, fs=2.0)           # w in units of the old Nyquist
    mag = np.abs(H)
    assert np.abs(mag[w <= 0.30] - 1.0).max() < 1e-6           # pass-band: 0.6 of the new Nyquist
    assert mag[w >= 0.70].max() < 10 ** (-120 / 20)            # band that aliases onto the CQT octave
    assert np.abs(h[63 + 2::2]).max() < 1e-15                  # half-band: even offsets vanish


def test_cqt_basis_is_octave_invariant():
    freqs = 440.0 * 2.0 ** ((24 - 69) / 12.0) * 2.0 ** (np.arange(252) / 36)
    alpha = F.relative_bandwidth(freqs=freqs)
    top, n_fft, _ = C._vqt_filter_fft(16000, freqs[-36:], 1, 1, 0.01, alpha=alpha[-36:])
    low, _, _ = C._vqt_filter_fft(16000 / 16, freqs[-36 * 5:-36 * 4], 1, 1, 0.01, alpha=alpha[-36 * 5:-36 * 4])
    assert n_fft == 512
    assert np.abs(top.toarray() - low.toarray()).max() < 1e-6
    nz = np.nonzero(np.abs(top.toarray()).sum(0))[0]
    assert 60 <= nz.min() and nz.max() <= 144 and np.diff(top.indptr).max() <= 20


def test_pipeline_against_golden(golden):
    names = list(golden["names"])
    pcm = golden["pcm16"]
    for gi in range(len(names)):
        y = pcm[gi].astype(np.float32) / np.float32(32768.0)
        dbg = {}
        ch, sc = P.segment_features(y, debug=dbg)
        for k in P.CHANNEL_KEYS:
            ref = golden[f"{gi}/{k}"]
            assert ch[k].shape == (128, 63) and ch[k].dtype == np.float32
            assert np.abs(ch[k] - ref).max() < 2e-5, (names[gi], k)
        ref = golden[f"{gi}/scalars"]
        assert sc.shape == (36,) and sc.dtype == np.float32
        assert np.allclose(sc, ref, rtol=1e-5, atol=1e-7)
        ints = golden[f"{gi}/dbg/ints"]
        assert dbg["n_peaks"] == ints[0] and dbg["first_min_idx"] == ints[1]
        assert sc[22] == ints[0] and sc[35] == np.float32(ints[1] / 16000)


def test_degenerate_inputs():
    ch, sc = P.segment_features(np.zeros(16000, dtype=np.float32))
    for k in ("mel", "mel_delta", "mel_delta2", "gammatone", "mod_spec"):
        assert np.all(np.isfinite(ch[k]))
    assert np.isnan(sc[29]) and np.isnan(sc[33]) and sc[35] == 0.0 and sc[22] == 0.0
    short = P.synth_segment(5)[:9000]
    ch2, _ = P.segment_features(short)
    ch3, _ = P.segment_features(np.concatenate([short, np.zeros(7000, np.float32)]))
    assert all(np.array_equal(ch2[k], ch3[k]) for k in ch2)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/precompute"), reason="needs /root/reference")
def test_pipeline_bit_identical_to_reference_code(tmp_path):
    """The reference's own, unmodified process.py (over the shim) vs oracle/pipeline.py: bit-for-bit."""
    import glob
    sys.path.insert(0, "/root/reference")
    try:
        from src.precompute.process import process_and_save_npz
        wav = sorted(glob.glob("/root/reference/input/train/*.wav"))[123]
        (tmp_path / "ref").mkdir(); (tmp_path / "ora").mkdir()
        assert process_and_save_npz(("x", wav, str(tmp_path / "ref")))[1]
        assert P.process_wav("x", wav, str(tmp_path / "ora"))[1]
        a = np.load(tmp_path / "ref" / "x.npz"); b = np.load(tmp_path / "ora" / "x.npz")
        assert sorted(a.files) == sorted(b.files)
        for k in a.files:
            assert np.array_equal(a[k], b[k], equal_nan=True), k
    finally:
        sys.path.remove("/root/reference")
