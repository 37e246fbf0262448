"""CPU tests of the host side: C-ABI exports, constant tables against the oracle shim, mirrors of the reference helpers,
the .npz / Dataset contract, the gloo (world_size 2) statistics exchange.  No compute call is made here."""
import ctypes
import os
import re

import numpy as np
import pytest

import bpc_b200
from bpc_b200 import _lib as L
from bpc_b200.precompute import methods as M, process as PR, core as CO
from oracle import pipeline as P
from librosa import filters as F, _core as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "bpc.h")).read()
    declared = set(re.findall(r"\b(bpc_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    lib = ctypes.CDLL(bpc_b200.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/bpc.h but not exported"
    assert declared == set(bpc_b200.EXPORTS), declared ^ set(bpc_b200.EXPORTS)
    assert bpc_b200.lib().bpc_abi_version() == 3


def test_default_params_are_the_reference_constants():
    p = bpc_b200.default_params()
    assert (p.sr, p.n_fft, p.hop, p.n_mels, p.n_mfcc, p.n_gammatone, p.n_lpc, p.expected_len) == \
           (16000, 512, 256, 128, 40, 64, 12, 16000)
    assert p.fmax == 4500.0 and p.pad_scalars_to == 0
    assert bpc_b200.lib().bpc_num_frames(ctypes.byref(p)) == 63
    assert bpc_b200.lib().bpc_num_scalars(ctypes.byref(p)) == 36
    p.pad_scalars_to = 39
    assert bpc_b200.lib().bpc_num_scalars(ctypes.byref(p)) == 39
    assert (M.SR, M.N_FFT, M.HOP_LENGTH, M.N_MELS, M.N_MFCC, M.FMAX, M.N_GAMMATONE, M.N_LPC) == \
           (16000, 512, 256, 128, 40, 4500, 64, 12)


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(bpc_b200.BpcError, match="no CUDA device|CPU fallback"):
        bpc_b200.Engine(device=0)


def test_unsupported_params_are_rejected():
    # expected_len = 16000 d with d = 2^a 3^b 5^c <= 32 is the long mode (BASELINE config 4); anything else is refused
    for bad in (16000 * 7, 16000 * 33, 24000, 0):
        p = bpc_b200.default_params(expected_len=bad)
        h = ctypes.c_void_p()
        rc = bpc_b200.lib().bpc_create(ctypes.byref(h), ctypes.byref(p), 0, 16)
        assert rc == -4 and not h.value, bad
        assert b"expected_len" in bpc_b200.lib().bpc_last_error(None)
    p = bpc_b200.default_params(expected_len=32000)
    assert bpc_b200.lib().bpc_num_frames(ctypes.byref(p)) == 126


@pytest.mark.parametrize("name,args", [("mel_a", (512, 128, 4500.0)), ("mel_b", (512, 128, 8000.0)),
                                        ("mel_c", (512, 64, 8000.0)), ("mel_d", (2048, 128, 8000.0))])
def test_mel_tables_bit_exact(name, args):
    n_fft, n_mels, fmax = args
    assert np.array_equal(bpc_b200.table(name), F.mel(sr=16000, n_fft=n_fft, n_mels=n_mels, fmax=fmax))


def test_mel_d_band_is_the_dense_bank_and_bank_conflict_free():
    """The band form of the n_fft 2048 mel bank that k_frame2048 walks (starts moved down by r < 32 with r leading zero
    weights): every row reproduces the dense bank exactly, the 32 rows a warp reads together start in 32 different
    shared-memory banks, the padded bands stay inside the |X| row and its zero tail (1092 words)."""
    import bpc_b200
    dense = bpc_b200.table("mel_d")
    band = bpc_b200.table("mel_d_band")
    start, count, w = band[:, 0].astype(int), band[:, 1].astype(int), band[:, 2:]
    assert start.min() >= 0 and count.max() <= 80
    for r in range(128):
        row = np.zeros(1025 + 128, np.float32)
        row[start[r]:start[r] + 80] = w[r]
        assert np.array_equal(row[:1025], dense[r]) and not row[1025:].any()
        assert not w[r, count[r]:].any()
    for g in range(4):
        sl = slice(32 * g, 32 * g + 32)
        assert len(set((start[sl] % 32).tolist())) == 32
        rounds = (count[sl].max() + 3) // 4 * 4                       # the kernel walks a group's bands in rounds of four
        assert (start[sl] + rounds).max() <= 1092
    nat = [int(np.flatnonzero(dense[r])[0]) for r in range(128)]
    cnt = [int(np.flatnonzero(dense[r])[-1]) - nat[r] + 1 for r in range(128)]
    # the padding costs at most two more rounds of four taps per frame than the natural layout (23)
    assert sum((count[32 * g:32 * g + 32].max() + 3) // 4 for g in range(4)) <= 2 + sum((max(cnt[32 * g:32 * g + 32]) + 3) // 4 for g in range(4))


def test_window_and_misc_tables():
    import scipy.fftpack
    import scipy.signal
    for n in (512, 2048, 384):
        assert np.array_equal(bpc_b200.table(f"hann{n}"), scipy.signal.get_window("hann", n))
    assert np.array_equal(bpc_b200.table("hamming400"), np.hamming(400))
    assert np.array_equal(bpc_b200.table("hist_edges"), np.linspace(-0.5, 0.5, 101))
    assert np.abs(bpc_b200.table("halfband") - C.default_halfband()).max() < 1e-14
    d = scipy.fftpack.dct(np.eye(128), type=2, norm="ortho", axis=0)[:40]
    assert np.abs(bpc_b200.table("dct_mel") - d).max() < 1e-7
    d = scipy.fftpack.dct(np.eye(63), type=2, norm="ortho", axis=0)
    assert np.abs(bpc_b200.table("dct_time") - d).max() < 1e-7


@pytest.mark.parametrize("ti", [0, 13, 50, 99])
def test_chroma_and_cqt_tables(ti):
    edges = np.linspace(-0.5, 0.5, 101)
    assert np.array_equal(bpc_b200.table("chroma", ti), F.chroma(sr=16000, n_fft=512, tuning=edges[ti]))
    fmin = 440.0 * 2.0 ** ((24 - 69) / 12.0) * 2.0 ** (edges[ti] / 36)
    ratios = 2.0 ** (np.arange(0, 36, dtype=float) / 36)
    freqs = np.sort(np.multiply.outer(2.0 ** np.arange(7, dtype=float), ratios).flatten()) * fmin
    alpha = F.relative_bandwidth(freqs=freqs)
    lengths, _ = F.wavelet_lengths(freqs=freqs, sr=16000, alpha=alpha)
    fb, _, _ = C._vqt_filter_fft(16000, freqs[-36:], 1, 1, 0.01, alpha=alpha[-36:])
    mine = bpc_b200.table("cqt_basis", ti)
    mine = mine[..., 0] + 1j * mine[..., 1]
    ref = fb.toarray()
    assert np.array_equal(mine != 0, ref != 0)
    assert np.abs(mine - ref).max() < 1e-7
    assert np.abs(bpc_b200.table("cqt_sqrt_len", ti) - np.sqrt(lengths)).max() < 1e-9


def test_synth_generators_agree():
    from bpc_b200.synth import synth_pcm16
    for i in (0, 7, 1234):
        q = synth_pcm16(i)
        assert q.dtype == np.int16 and q.shape == (16000,)
        assert np.array_equal(q.astype(np.float32) / np.float32(32768.0), P.synth_segment(i))


def test_pad_helpers_match_reference_semantics():
    a = np.arange(12, dtype=np.float32).reshape(3, 4) - 5
    assert np.array_equal(M.pad_time(a, 3, 2), a[:, :2])
    out = M.pad_time(a, 3, 6)
    assert out.shape == (3, 6) and np.all(out[:, 4:] == a.min())
    assert np.array_equal(M.pad_freq(a, 3, 2), a[:2])
    out = M.pad_freq(a, 3, 5)
    assert out.shape == (5, 4) and np.all(out[3:] == a.min()) and out.dtype == np.float32
    y = np.ones(10, dtype=np.float32)
    assert np.array_equal(M.pad_or_truncate(y, 4), y[:4])
    z = M.pad_or_truncate(y, 13)
    assert z.shape == (13,) and np.all(z[10:] == 0) and z.dtype == np.float32
    assert np.array_equal(M.pad_time(a, 3, 6), P.fit_time(a, 3, 6)) and np.array_equal(M.pad_freq(a, 3, 5), P.fit_rows(a, 3, 5))


def test_wav_name_mapping():
    assert CO.wav_name_for("steth_20180814_09_37_11_I_004", "train") == "steth_20180814_09_37_11_004.wav"
    assert CO.wav_name_for("steth_20180814_09_37_11_E_004", "train") == "steth_20180814_09_37_11_004.wav"
    assert CO.wav_name_for("steth_20190713_09_58_25_007.wav", "test") == "steth_20190713_09_58_25_007.wav"
    assert CO.wav_name_for("abc", "test") == "abc.wav"


def test_npz_contract_and_dataset_order(tmp_path):
    feats = np.arange(9 * 128 * 63, dtype=np.float32).reshape(9, 128, 63)
    scal = np.arange(36, dtype=np.float32)
    PR.save_npz(str(tmp_path), "id_1.wav", feats, scal)
    d = np.load(tmp_path / "id_1.wav.npz")
    assert set(d.files) == set(P.CHANNEL_KEYS) | {"scalars"}
    excluded = {"scalars", "sr", "hop_length", "n_fft"}                 # dataset.py:8
    names = sorted(k for k in d.files if k not in excluded)             # dataset.py:25-26
    assert tuple(names) == bpc_b200.CHANNELS == P.SORTED_KEYS
    stacked = np.stack([d[k] for k in names]).astype(np.float32)        # dataset.py:48
    assert np.array_equal(stacked, feats) and d["scalars"].shape == (36,)
    assert all(d[k].dtype == np.float32 and d[k].shape == (128, 63) for k in names)


def test_fit_batch_and_load_wav(tmp_path):
    import scipy.io.wavfile
    q = (np.random.default_rng(0).standard_normal(12000) * 3000).astype(np.int16)
    scipy.io.wavfile.write(tmp_path / "a.wav", 16000, q)
    w = PR.load_wav(str(tmp_path / "a.wav"))
    assert w.dtype == np.int16 and np.array_equal(w, q)
    b = PR.fit_batch([w, np.concatenate([q, q])])
    assert b.dtype == np.int16 and b.shape == (2, 16000)
    assert np.array_equal(b[0, :12000], q) and np.all(b[0, 12000:] == 0) and np.array_equal(b[1], np.concatenate([q, q])[:16000])
    mixed = PR.fit_batch([w, w.astype(np.float32) / 32768])
    assert mixed.dtype == np.float32 and np.array_equal(mixed[0], mixed[1])
    # librosa.load semantics for the other PCM layouts: scale by dtype first, then average the channels
    st16 = np.stack([q, -q // 2], axis=1).astype(np.int16)
    scipy.io.wavfile.write(tmp_path / "st16.wav", 16000, st16)
    w = PR.load_wav(str(tmp_path / "st16.wav"))
    assert w.dtype == np.float32 and np.array_equal(w, np.mean(st16.astype(np.float32) / np.float32(32768), axis=1, dtype=np.float32))
    st32 = (st16.astype(np.int32) << 16)
    scipy.io.wavfile.write(tmp_path / "st32.wav", 16000, st32)
    assert np.array_equal(PR.load_wav(str(tmp_path / "st32.wav")), w)         # same samples at 32 bits: same waveform
    u8 = np.stack([(q >> 8) + 128, 128 - (q >> 9)], axis=1).astype(np.uint8)
    scipy.io.wavfile.write(tmp_path / "st8.wav", 16000, u8)
    w8 = PR.load_wav(str(tmp_path / "st8.wav"))
    assert np.abs(w8).max() <= 1.0 and np.array_equal(w8, np.mean((u8.astype(np.float32) - 128) / 128, axis=1, dtype=np.float32))
    m32 = (q.astype(np.int32) << 16)
    scipy.io.wavfile.write(tmp_path / "m32.wav", 16000, m32)
    assert np.array_equal(PR.load_wav(str(tmp_path / "m32.wav")), q.astype(np.float32) / np.float32(32768))
    fid, ok, err = PR.process_and_save_npz(("nope", str(tmp_path / "missing.wav"), str(tmp_path)))
    assert fid == "nope" and ok is False and isinstance(err, str)      # never raises (process.py:107-108)


def test_resample_filter_table_matches_oracle():
    """process.py:28 resample-on-load: the C++ polyphase table of bpc_resample against oracle/resample.py (host only)."""
    from oracle import resample as R
    from bpc_b200.engine import resample_filter
    for sr_in, sr_out in ((8000, 16000), (48000, 16000), (44100, 16000), (22050, 16000), (11025, 16000)):
        p, q, half, tab = resample_filter(sr_in, sr_out)
        rp, rq, rhalf, rtab = R.polyphase_table(sr_in, sr_out)
        assert (p, q, half) == (rp, rq, rhalf) and tab.shape == rtab.shape
        assert np.abs(tab - rtab).max() < 2e-14 and np.allclose(tab.sum(axis=1), 1.0, atol=1e-14)
        assert bpc_b200.lib().bpc_resample_len(12345, sr_in, sr_out) == -((-12345 * sr_out) // sr_in)
    # the oracle's own contract: pass band flat to float32 rounding, stop band rejected far below 16-bit PCM
    t = np.arange(48000) / 48000.0
    for f0, want in ((1000.0, 1.0), (7000.0, 1.0), (9000.0, 0.0)):
        o = R.resample(np.sin(2 * np.pi * f0 * t).astype(np.float32), 48000, 16000)
        assert len(o) == 16000
        ref = want * np.sin(2 * np.pi * f0 * np.arange(16000) / 16000.0)
        assert np.abs(o[2000:14000] - ref[2000:14000]).max() < 2e-7, f0
    assert bpc_b200.lib().bpc_resample_filter(0, 16000, None, 0, None, None, None) == -1


def test_expand_compact_and_live_rows():
    """Compact host layout (include/bpc.h): 772 data rows + 9 pad values per segment <-> the full [9,128,T] planes."""
    from bpc_b200 import shards
    lib = bpc_b200.lib()
    assert [lib.bpc_live_rows(c) for c in range(9)] == list(L.LIVE_ROWS) and sum(L.LIVE_ROWS) == L.LIVE_TOTAL == 772
    assert lib.bpc_live_rows(9) == -1 and lib.bpc_live_rows(-1) == -1
    for T, n, threads in ((63, 7, 3), (126, 2, 1), (63, 0, 2)):
        F, _ = _fake_rows(n, T=T, seed=11)
        rows, pad = shards.compact_from_full(F)
        assert rows.shape == (n, 772, T) and pad.shape == (n, 9)
        assert np.array_equal(bpc_b200.expand_compact(rows, pad, threads=threads), F)
    with pytest.raises(ValueError):
        bpc_b200.expand_compact(np.zeros((2, 700, 63), np.float32), np.zeros((2, 9), np.float32))
    assert lib.bpc_expand_compact(None, None, 1, 63, None, 1) == -1


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from bpc_b200.stats import allreduce_stats, finalize_stats, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    data = rng.standard_normal((10, 45, 7))                              # 10 "segments", 45 rows
    lo, hi = shard_range(10, rank, world)
    mine = data[lo:hi]
    st = torch.zeros(45, 5, dtype=torch.float64)
    st[:, 0] = mine.shape[0] * 7
    st[:, 1] = torch.from_numpy(mine.sum(axis=(0, 2)))
    st[:, 2] = torch.from_numpy((mine ** 2).sum(axis=(0, 2)))
    st[:, 3] = torch.from_numpy(mine.min(axis=(0, 2)))
    st[:, 4] = torch.from_numpy(mine.max(axis=(0, 2)))
    allreduce_stats(st, dist)
    f = finalize_stats(st)
    ok = (np.allclose(f["mean"].numpy(), data.mean(axis=(0, 2))) and np.allclose(f["std"].numpy(), data.std(axis=(0, 2)))
          and np.allclose(f["min"].numpy(), data.min(axis=(0, 2))) and np.allclose(f["max"].numpy(), data.max(axis=(0, 2)))
          and float(f["count"][0]) == 70.0)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_stats_allreduce_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_shard_range_partitions():
    from bpc_b200.stats import shard_range
    for n, w in [(1000000, 8), (10, 3), (7, 8)]:
        parts = [shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))


def test_register_fft_index_logic_on_host(tmp_path):
    """csrc/fft_reg.cuh is __host__ __device__ and lane-explicit: the four-step index logic, the compile-time twiddles
    and the real-input split are run on the CPU for every lane of a team and compared with a long-double DFT."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = tmp_path / "fft_host_test"
    src = os.path.join(ROOT, "tests", "host", "fft_host_test.cpp")
    inc = os.path.join(ROOT, "breathing-phase-classifier_b200", "csrc")
    subprocess.run([nvcc, "-x", "cu", "-std=c++20", "-Wno-deprecated-gpu-targets", "-I", inc, src, "-o", str(exe)],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    errs = [float(x.split()[-1]) for x in out.strip().splitlines()]
    assert len(errs) == 3 and errs[0] < 1e-10 and errs[1] < 1e-9 and errs[2] < 1e-9, out


def test_radix20_hilbert_fft_on_host(tmp_path):
    """csrc/fft20.cuh (the three radix-20 passes each way of k_hilbert, padded storage, digit-reversed spectrum, pair
    split) is __host__ __device__: run every butterfly on the CPU and compare with scipy.signal.hilbert (methods.py:72)."""
    import shutil
    import subprocess
    import scipy.signal
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = tmp_path / "fft20_host_test"
    src = os.path.join(ROOT, "tests", "host", "fft20_host_test.cpp")
    inc = os.path.join(ROOT, "breathing-phase-classifier_b200", "csrc")
    subprocess.run([nvcc, "-x", "cu", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", inc, src, "-o", str(exe)],
                   check=True, capture_output=True)
    rng = np.random.default_rng(5)
    for trial in range(2):
        y = (rng.standard_normal(16000) * (np.hanning(16000) if trial else 1.0)).astype(np.float32)
        y.tofile(tmp_path / "in.f32")
        subprocess.run([str(exe), str(tmp_path / "in.f32"), str(tmp_path / "out.f32")], check=True)
        h = np.fromfile(tmp_path / "out.f32", dtype=np.float32)
        ref = np.imag(scipy.signal.hilbert(y.astype(np.float64)))
        ref32 = np.imag(scipy.signal.hilbert(y))                      # what the reference runs: float32 pocketfft
        scale = np.abs(ref).max()
        assert np.abs(h - ref).max() / scale < 6e-7
        assert np.abs(h - ref).max() < 2.0 * np.abs(ref32 - ref).max() + 1e-7 * scale


def test_c64_abs_f32_algorithm():
    """csrc/fft.cuh::c64_abs_f32 (np.abs of a complex64 in float32 pairs, used by the STFT / CQT / Hilbert kernels) emulated
    operation by operation in numpy (float32 FMAs are exact in float64): equal to (float)sqrt((double)re^2 + (double)im^2)
    with a correctly rounded rsqrt, and within one ulp in <= 5 per million when the rsqrt is off by 2 ulp (the hardware
    rsqrt.approx bound)."""
    f32, f64 = np.float32, np.float64

    def emul(re, im, ulps):
        a, b = np.abs(re), np.abs(im)
        x, y = np.maximum(a, b).astype(f64), np.minimum(a, b).astype(f64)
        p = (x * x).astype(f32); pe = (x * x - p.astype(f64)).astype(f32)
        q = (y * y).astype(f32); qe = (y * y - q.astype(f64)).astype(f32)
        hi = (p + q).astype(f32)
        lo = ((((p - hi).astype(f32) + q).astype(f32)) + (pe + qe).astype(f32)).astype(f32)
        rs = (1.0 / np.sqrt(hi.astype(f64))).astype(f32)
        for _ in range(abs(ulps)):
            rs = np.nextafter(rs, f32(np.inf) if ulps > 0 else f32(0))
        r0 = (hi * rs).astype(f32).astype(f64)
        res = ((hi.astype(f64) - r0 * r0).astype(f32) + lo).astype(f32)
        return (res.astype(f64) * (f32(0.5) * rs).astype(f32).astype(f64) + r0).astype(f32)

    rng = np.random.default_rng(1)
    n = 1_000_000
    re = (rng.standard_normal(n) * 10.0 ** rng.uniform(-8, 4, n)).astype(f32)
    for im in ((rng.standard_normal(n) * 10.0 ** rng.uniform(-8, 4, n)).astype(f32), (re * rng.uniform(0.5, 2.0, n)).astype(f32)):
        ref = np.sqrt(re.astype(f64) ** 2 + im.astype(f64) ** 2).astype(f32)
        assert np.array_equal(emul(re, im, 0), ref)
        for u in (2, -2):
            out = emul(re, im, u)
            bad = out != ref
            assert bad.sum() <= 5
            assert np.all(np.abs(out[bad].astype(f64) - ref[bad]) <= np.spacing(ref[bad]).astype(f64))


def test_div_fast_algorithm():
    """csrc/k_lpc.cu::div_fast (the Burg reflection coefficient: reciprocal seed with 2^-23 relative error, two Newton
    steps, one residual correction) emulated with exact rational FMAs: within one ulp of the IEEE quotient."""
    from fractions import Fraction as Fr

    def fma(a, b, c):
        return float(Fr(a) * Fr(b) + Fr(c))

    rng = np.random.default_rng(2)
    worst = 0.0
    for _ in range(3000):
        a = float(rng.standard_normal() * 10.0 ** rng.uniform(-30, 30))
        b = float(abs(rng.standard_normal()) * 10.0 ** rng.uniform(-30, 30)) + 1e-300
        r = (1.0 / b) * (1.0 + float(rng.choice([-1.0, 1.0])) * 2.0 ** -23)
        e = fma(-b, r, 1.0); r = fma(r, e, r)
        e = fma(-b, r, 1.0); r = fma(r, e, r)
        q = a * r
        q = fma(r, fma(-b, q, a), q)
        exact = a / b
        worst = max(worst, abs(q - exact) / np.spacing(abs(exact)))
    assert worst <= 1.0


# ------------------------------------------------------------------------------------- output writer / packed shard
def _fake_rows(n, T=63, S=36, seed=3):
    """Random planes with the structure the path produces: rows live[c]..127 of a plane hold its minimum (pad_freq)."""
    rng = np.random.default_rng(seed)
    F = rng.standard_normal((n, 9, 128, T)).astype(np.float32)
    for c, lv in enumerate(L.LIVE_ROWS):
        if lv < 128:
            F[:, c, lv:] = F[:, c, :lv].min(axis=(1, 2))[:, None, None]
    return F, rng.standard_normal((n, S)).astype(np.float32)


def test_cpp_npz_writer_is_readable_by_numpy_and_zipfile(tmp_path):
    """bpc_npz_pack / bpc_npz_write_batch against np.savez of the same arrays (process.py:92-103)."""
    import io
    import zipfile
    from bpc_b200 import shards
    F, S = _fake_rows(5)
    raw = shards.npz_bytes(F[0], S[0])
    assert len(raw) == bpc_b200.lib().bpc_npz_size(63, 36)
    z = zipfile.ZipFile(io.BytesIO(raw))
    assert z.testzip() is None                                            # CRC-32 of every member checks out
    assert [i.filename for i in z.infolist()] == [k + ".npy" for k in PR.NPZ_KEYS] + ["scalars.npy"]
    ref = io.BytesIO()
    np.savez(ref, **{k: F[0, bpc_b200.CHANNELS.index(k)] for k in PR.NPZ_KEYS}, scalars=S[0])
    a, b = np.load(io.BytesIO(raw)), np.load(io.BytesIO(ref.getvalue()))
    assert a.files == b.files
    for k in a.files:
        assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape and np.array_equal(a[k], b[k])
    for m in z.infolist():                                                # same .npy bytes as numpy writes
        assert z.read(m.filename) == zipfile.ZipFile(io.BytesIO(ref.getvalue())).read(m.filename)
    status = np.array([0, 0, 1, 0, 2], dtype=np.int32)                    # bit 0 = non-finite input -> failure tuple
    res = shards.write_npz_batch(str(tmp_path), [f"id_{i}" for i in range(5)], F, S, status, threads=3)
    assert [ok for _, ok, _ in res] == [True, True, False, True, True] and "non-finite" in res[2][2]
    assert not (tmp_path / "id_2.npz").exists()
    d = np.load(tmp_path / "id_4.npz")
    assert np.array_equal(d["tempogram"], F[4, 8]) and np.array_equal(d["scalars"], S[4])
    res = shards.write_npz_batch(str(tmp_path / "missing_dir"), ["x"], F[:1], S[:1])
    assert res[0][1] is False and "open" in res[0][2]
    odd = shards.npz_bytes(F[1, :, :, :7].copy(), S[1, :5].copy())       # other shapes keep the 64-byte npy alignment
    d = np.load(io.BytesIO(odd))
    assert d["mel"].shape == (128, 7) and d["scalars"].shape == (5,) and np.array_equal(d["lpc"], F[1, 2, :, :7])


def _reference_ds_items(df, feature_dir, is_training):
    """What dataset.py:16-57 does for every row (restated: np.load per item, sorted non-excluded keys, stack)."""
    import torch
    out = []
    for _, row in df.iterrows():
        d = np.load(os.path.join(feature_dir, row["ID"] + ".npz"))
        names = sorted(k for k in d.keys() if k not in {"scalars", "sr", "hop_length", "n_fft"})
        f = torch.from_numpy(np.stack([d[k] for k in names], axis=0).astype(np.float32))
        s = torch.from_numpy(d["scalars"].astype(np.float32))
        out.append((f, s, torch.tensor(1.0 if row["Target"] == "E" else 0.0) if is_training else row["ID"]))
    return out


@pytest.mark.parametrize("compact", [True, False])
@pytest.mark.parametrize("is_training", [True, False])
def test_packed_shard_ds_matches_per_file_ds(tmp_path, is_training, compact):
    import pandas as pd
    import torch
    from bpc_b200 import shards
    F, S = _fake_rows(12, seed=5)
    ids = [f"steth_{i:03d}_{'EI'[i % 2]}_1" for i in range(12)]
    per_file = tmp_path / "npz"; per_file.mkdir()
    shards.write_npz_batch(str(per_file), ids, F, S)
    with shards.ShardWriter(str(tmp_path / "packed"), 12, 63, 36, compact=compact) as w:
        w.append(ids[:7], F[:7], S[:7])
        w.append(ids[7:], F[7:], S[7:], np.zeros(5, np.int32))
    assert shards.is_packed(str(tmp_path / "packed")) and not shards.is_packed(str(per_file))
    assert os.path.exists(tmp_path / "packed" / ("rows.npy" if compact else "feats.npy"))
    if compact:                                                           # 67 % of the bytes of the full layout
        assert os.path.getsize(tmp_path / "packed" / "rows.npy") < 0.68 * 12 * 9 * 128 * 63 * 4
    order = [5, 0, 11, 3, 3, 8]                                           # a shuffled subset, as a split would give
    df = pd.DataFrame({"ID": [ids[i] for i in order], "Target": ["EI"[i % 2] for i in order]})
    ds = shards.PackedDS(df, str(tmp_path / "packed"), is_training)
    assert ds.feature_names == list(bpc_b200.CHANNELS) and ds.n_features == 9 and ds.scalar_dim == 36
    ref = _reference_ds_items(df, str(per_file), is_training)
    assert len(ds) == len(ref)
    for i, (f, s, y) in enumerate(ref):
        g = ds[i]
        assert torch.equal(g[0], f) and torch.equal(g[1], s)
        assert (torch.equal(g[2], y) if is_training else g[2] == y)
    fb, sb, yb = shards.collate_fn([ds[i] for i in range(4)])
    assert fb.shape == (4, 9, 128, 63) and sb.shape == (4, 36)
    assert (yb.shape == (4,)) if is_training else (yb == [ids[i] for i in order[:4]])
    loader = torch.utils.data.DataLoader(ds, batch_size=4, shuffle=False, collate_fn=shards.collate_fn)
    fb2 = next(iter(loader))[0]
    assert torch.equal(fb2, fb)
    with pytest.raises(ValueError):
        with shards.ShardWriter(str(tmp_path / "short"), 3, 63, 36) as w:
            w.append(ids[:2], F[:2], S[:2])


def test_rand_bbox_matches_reference_expression():
    """augmentation.py:11-25 restated with the same numpy calls and the same RNG order."""
    from bpc_b200.resident import rand_bbox
    for seed in range(20):
        lam = float(np.random.RandomState(seed).beta(1.0, 1.0))
        rs = np.random.RandomState(seed + 100)
        got = rand_bbox(63, 128, lam, rs)
        rs = np.random.RandomState(seed + 100)
        cut_rat = np.sqrt(1. - lam); cut_w = np.int32(63 * cut_rat); cut_h = np.int32(128 * cut_rat)
        cx = rs.randint(63); cy = rs.randint(128)
        want = (np.clip(cx - cut_w // 2, 0, 63), np.clip(cy - cut_h // 2, 0, 128),
                np.clip(cx + cut_w // 2, 0, 63), np.clip(cy + cut_h // 2, 0, 128))
        assert got == tuple(int(v) for v in want)


def test_cpp_wav_reader_matches_scipy_and_reports_per_file(tmp_path):
    """bpc_wav_load_batch + pad_or_truncate against scipy.io.wavfile (what the oracle's librosa.load shim reads)."""
    import scipy.io.wavfile
    rng = np.random.default_rng(1)
    sig = {n: (rng.standard_normal(n) * 4000).astype(np.int16) for n in (16000, 12345, 20000, 1)}
    paths = []
    for n, q in sig.items():
        scipy.io.wavfile.write(tmp_path / f"m{n}.wav", 16000, q)
        paths.append(str(tmp_path / f"m{n}.wav"))
    scipy.io.wavfile.write(tmp_path / "stereo.wav", 16000, np.stack([sig[16000], sig[16000] // 2], axis=1))
    scipy.io.wavfile.write(tmp_path / "sr8k.wav", 8000, sig[12345])
    scipy.io.wavfile.write(tmp_path / "f32.wav", 16000, (sig[12345] / 32768.0).astype(np.float32))
    (tmp_path / "junk.wav").write_bytes(b"not a wave file at all")
    # a file with an extra chunk before "data" and an odd-sized chunk (must be skipped with its pad byte)
    raw = open(paths[0], "rb").read()
    extra = b"LIST" + (5).to_bytes(4, "little") + b"abcde\x00"
    patched = raw[:12] + raw[12:36] + extra + raw[36:]
    patched = patched[:4] + (len(patched) - 8).to_bytes(4, "little") + patched[8:]
    (tmp_path / "chunks.wav").write_bytes(patched)
    paths += [str(tmp_path / k) for k in ("stereo.wav", "sr8k.wav", "f32.wav", "junk.wav", "missing.wav", "chunks.wav")]
    batch, errs = PR.load_wav_batch(paths, 16000, threads=3)
    assert batch.dtype == np.float32                                       # stereo / float rows forced the float path
    for i, n in enumerate(sig):
        want = M.pad_or_truncate(sig[n].astype(np.float32) / np.float32(32768.0), 16000)
        assert errs[i] is None and np.array_equal(batch[i], want), n
    st = PR.load_wav(paths[4])                                              # librosa.load(mono=True): channel mean
    assert errs[4] is None and np.array_equal(batch[4], st[:16000])
    # an 8 kHz file is resampled on the device (bpc_resample); without a GPU that is a per-file error, never a raise
    assert errs[5] is None or "CUDA" in errs[5] or "bpc_create" in errs[5], errs[5]
    assert errs[6] is None and np.array_equal(batch[6], M.pad_or_truncate((sig[12345] / 32768.0).astype(np.float32), 16000))
    assert "RIFF" in errs[7] and "No such file" in errs[8]
    assert errs[9] is None and np.array_equal(batch[9], batch[0])
    only16, errs = PR.load_wav_batch(paths[:4], 16000)
    assert only16.dtype == np.int16 and errs == [None] * 4 and np.array_equal(only16[0], sig[16000])
    assert np.array_equal(only16[1, :12345], sig[12345]) and not only16[1, 12345:].any()


def test_host_only_entry_points_reject_bad_arguments(tmp_path):
    """No exceptions cross the ABI: bad arguments come back as BPC_ERR_ARG (-1) / per-file codes."""
    import ctypes as C
    lib = bpc_b200.lib()
    f = np.zeros((9, 128, 63), np.float32); s = np.zeros(36, np.float32); out = np.zeros(16, np.uint8)
    assert lib.bpc_npz_size(0, 36) == -1 and lib.bpc_npz_size(63, 0) == -1
    assert lib.bpc_npz_pack(f.ctypes.data, s.ctypes.data, 63, 36, out.ctypes.data, out.nbytes) == -1     # buffer too small
    assert lib.bpc_npz_pack(None, s.ctypes.data, 63, 36, out.ctypes.data, out.nbytes) == -1
    ok = np.zeros(1, np.int32)
    ids = (C.c_char_p * 1)(b"x")
    assert lib.bpc_npz_write_batch(None, ids, f.ctypes.data, s.ctypes.data, None, 1, 63, 36, 2, ok.ctypes.data) == -1
    sr = np.zeros(1, np.int32); fr = np.zeros(1, np.int32); code = np.zeros(1, np.int32); pcm = np.zeros(16000, np.int16)
    assert lib.bpc_wav_load_batch(None, 1, 16000, 16000, pcm.ctypes.data, sr.ctypes.data, fr.ctypes.data, code.ctypes.data, 2) == -1
    paths = (C.c_char_p * 1)(str(tmp_path / "nope.wav").encode())
    assert lib.bpc_wav_load_batch(paths, 1, 16000, 0, pcm.ctypes.data, sr.ctypes.data, fr.ctypes.data, code.ctypes.data, 2) == -1
    assert lib.bpc_wav_load_batch(paths, 1, 16000, 16000, pcm.ctypes.data, sr.ctypes.data, fr.ctypes.data, code.ctypes.data, 2) == 0
    assert code[0] == -10                                                  # BPC_WAV_ERR_OPEN, reported per file
    assert lib.bpc_table_copy(C.byref(bpc_b200.default_params()), b"no_such_table", 0, out.ctypes.data, 16) < 0
    assert lib.bpc_num_frames(None) == -1 and lib.bpc_num_scalars(None) == -1
    assert lib.bpc_create(None, C.byref(bpc_b200.default_params()), 0, 16) == -1
    assert b"NULL" in lib.bpc_last_error(None)


def test_wav_parse_walks_chunks_without_touching_samples():
    """bpc_wav_parse (host half of the GPU-side decode): formats, WAVE_FORMAT_EXTENSIBLE, a junk chunk of odd size,
    a truncated data chunk, rejections."""
    import ctypes as C
    from bpc_b200 import _lib as L
    import wavutil as W
    lib = L.lib()
    for kind, fmt in (("u8", 1), ("pcm16", 2), ("pcm24", 3), ("pcm32", 4), ("f32", 5), ("f64", 6)):
        for ch in (1, 2, 3):
            x = W.samples(kind, 1000 + ch, ch, 7)
            for ext, junk in ((False, False), (True, True)):
                img = W.image(kind, x, 22050, extensible=ext, junk=junk)
                info = L.WavInfo()
                assert lib.bpc_wav_parse(img, len(img), C.byref(info)) == 0
                assert (info.sr, info.channels, info.fmt, info.frames) == (22050, ch, fmt, 1000 + ch)
                assert img[info.data_offset:info.data_offset + 16] == W.payload(kind, x)[:16]
    x = W.samples("pcm16", 1000, 2, 1)
    img = W.image("pcm16", x, 16000, cut=402)                       # 100 whole frames + half a frame missing
    info = L.WavInfo()
    assert lib.bpc_wav_parse(img, len(img), C.byref(info)) == 0 and info.frames == 899
    assert lib.bpc_wav_parse(b"RIFX" + img[4:], len(img), C.byref(info)) == -11
    assert lib.bpc_wav_parse(img[:30], 30, C.byref(info)) == -11
    bad = bytearray(W.image("pcm16", x, 16000)); bad[20] = 2                     # ADPCM
    assert lib.bpc_wav_parse(bytes(bad), len(bad), C.byref(info)) == -12
    assert lib.bpc_wav_parse(None, 0, C.byref(info)) == -1


def _burg_lag_products(x, M=12, floor=1e-4):
    """numpy restatement of csrc/k_lpc.cu::k_lpc_fast: Burg's recursion with the numerator as a quadratic form in the
    lag-product matrix Phi, librosa's own denominator recursion from the two edge errors.  Returns (a, min den_i / den_0)."""
    N = len(x)
    Phi = np.zeros((M + 1, M + 1))
    for v in range(M + 1):
        Phi[0, v] = np.dot(x[M:N], x[M - v:N - v])
    for u in range(M):
        for v in range(u, M):
            Phi[u + 1, v + 1] = Phi[u, v] + x[M - 1 - u] * x[M - 1 - v] - x[N - 1 - u] * x[N - 1 - v]
    Phi = np.triu(Phi) + np.triu(Phi, 1).T
    a = np.zeros(M + 1); a[0] = 1.0
    den = den0 = 2 * np.dot(x, x) - x[0] ** 2 - x[N - 1] ** 2
    gain = 1.0
    if den0 == 0.0:
        return a, 1.0
    for i in range(M):
        G = Phi[:i + 1, 1:i + 2].T @ a[:i + 1]                       # G[v-1] = sum_j a_j Phi[j][v], v = 1 .. i+1
        num = float(np.dot(a[i::-1][:i + 1], G))                      # sum_v a_{i+1-v} G[v]
        for n in range(i + 1, M):                                     # the pairs n = i+1 .. M-1, explicitly
            j = np.arange(i + 1)
            num += np.dot(a[:i + 1], x[n - j]) * np.dot(a[:i + 1], x[n - 1 - i + j])
        k = -2.0 * num / (den + np.finfo(np.float64).tiny)
        an = a.copy()
        an[1:i + 2] = a[1:i + 2] + k * a[i::-1][:i + 1]
        a = an
        j = np.arange(i + 2)
        fe, be = np.dot(a[:i + 2], x[i + 1 - j]), np.dot(a[:i + 2], x[N - 1 - (i + 1) + j])
        den = (1.0 - k * k) * den - be ** 2 - fe ** 2
        gain = min(gain, den / den0)
    return a, gain


def test_burg_from_lag_products_matches_direct_recursion(golden):
    """wherever the inverse prediction gain stays above the kernel's floor of 1e-4 the
    lag-product form of Burg's method agrees with librosa's direct recursion to 1e-8 (measured 3e-9); below it the
    frame is one the kernel hands to the direct method (a pure tone is such a frame)."""
    import librosa
    from oracle import pipeline as P
    ys = [q.astype(np.float32) / np.float32(32768.0) for q in golden["pcm16"][:4]] + [P.synth_segment(1000 + i) for i in range(3)]
    worst, kept, dropped = 0.0, 0, 0
    for y in ys:
        emph = np.append(y[0], y[1:] - 0.97 * y[:-1])
        for start in range(0, len(emph) - 400, 160 * 5):
            fr = np.asarray(emph[start:start + 400] * np.hamming(400), dtype=np.float64)
            a, gain = _burg_lag_products(fr)
            if gain > 1e-4:
                kept += 1
                worst = max(worst, float(np.abs(a - librosa.lpc(fr, order=12)).max()))
            else:
                dropped += 1
    assert kept > 100 and worst < 1e-8, (kept, dropped, worst)
    t = np.arange(400) / 16000.0
    tone = np.round(0.5 * np.sin(2 * np.pi * 440.0 * t) * 32768.0) / 32768.0 * np.hamming(400)
    assert _burg_lag_products(tone)[1] < 1e-4                         # queued for the direct recursion
    assert np.array_equal(_burg_lag_products(np.zeros(400))[0], np.eye(13)[0])


def test_sliding_dft_of_the_low_cqt_octaves_matches_fft():
    """DESIGN section 4 (r02-h), csrc/k_cens.cu::k_cens_lo: the rectangular-window STFT-512 of a 1000 / 500 / 250
    sample signal at hop 16 / 8 / 4, bins 60 .. 148, carried from frame to frame by
    X_{j+1}[k] = W^{-hk} (X_j[k] + sum_n (x[s_j + 512 + n] - x[s_j + n]) W^{nk}) after a start-up of 16 hops of 16
    samples from the all-zero window; in the 250-sample octave every frame is a pure rotation of the first."""
    rng = np.random.default_rng(3)
    ks = np.arange(60, 149)
    W = np.exp(-2j * np.pi * np.arange(512) / 512)
    for n, h in ((1000, 16), (500, 8), (250, 4)):
        x = rng.standard_normal(n).astype(np.float32).astype(np.float64)
        xp = np.concatenate([np.zeros(256), x, np.zeros(768)])
        ref = np.stack([np.fft.rfft(xp[t * h:t * h + 512])[ks] for t in range(63)])
        xx = np.concatenate([x, np.zeros(2048)])
        d = np.array([xx[i] - (xx[i - 512] if i >= 512 else 0.0) for i in range(256 + 63 * h)])
        X = np.zeros(len(ks), complex)
        for j in range(16):                                           # start-up: 16 hops of 16 samples
            X = np.conj(W[(16 * ks) % 512]) * (X + sum(d[16 * j + m] * W[(m * ks) % 512] for m in range(16)))
        out = []
        for t in range(63):
            out.append(X)
            X = np.conj(W[(h * ks) % 512]) * (X + sum(d[256 + t * h + m] * W[(m * ks) % 512] for m in range(h)))
        err = np.abs(np.array(out) - ref).max() / np.abs(ref).max()
        assert err < 1e-12, (n, h, err)
        if n == 250:
            assert np.all(d[256:] == 0.0)                             # nothing enters or leaves the window any more
