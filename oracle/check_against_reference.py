"""TEST INFRASTRUCTURE ONLY.  Build-container script (needs /root/reference; never runs on the GPU box).

1. Runs the reference's own, unmodified `src/precompute/process.py::process_and_save_npz` (imported from
   /root/reference) on top of `oracle/librosa_shim` for a set of real fixture wavs and synthetic wavs.
2. Requires `oracle/pipeline.py` (the restatement that travels to the GPU box) to reproduce those `.npz` files
   bit-for-bit.
3. Writes the golden fixtures `tests/golden/golden_segments.npz` (inputs as PCM16 + every output + the debug
   intermediates the GPU parity tests use).

Usage:  python oracle/check_against_reference.py [--n-real 24] [--n-synth 8] [--write-golden]
"""
from __future__ import annotations

import argparse
import glob
import os
import sys
import tempfile

import numpy as np
import scipy.io.wavfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "librosa_shim"))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-real", type=int, default=24)
    ap.add_argument("--n-synth", type=int, default=8)
    ap.add_argument("--n-golden-real", type=int, default=5)
    ap.add_argument("--n-golden-synth", type=int, default=3)
    ap.add_argument("--write-golden", action="store_true")
    args = ap.parse_args()

    if not os.path.isdir(REF):
        raise SystemExit("needs /root/reference (build container only)")
    sys.path.insert(0, REF)
    from src.precompute.process import process_and_save_npz          # the reference's own code, unmodified
    from oracle import pipeline as P

    wavs = sorted(glob.glob(os.path.join(REF, "input/train/*.wav")))
    step = max(1, len(wavs) // max(1, args.n_real))
    picks = wavs[::step][: args.n_real]

    tmp = tempfile.mkdtemp(prefix="bpc_oracle_")
    items = []                                                        # (name, wav_path)
    for w in picks:
        items.append((os.path.basename(w)[:-4], w))
    for i in range(args.n_synth):
        y = P.synth_segment(i)
        q = np.round(y * 32768.0).astype(np.int16)
        path = os.path.join(tmp, f"synth_{i:04d}.wav")
        scipy.io.wavfile.write(path, 16000, q)
        items.append((f"synth_{i:04d}", path))

    ref_dir = os.path.join(tmp, "ref")
    ora_dir = os.path.join(tmp, "ora")
    os.makedirs(ref_dir)
    os.makedirs(ora_dir)
    worst = 0.0
    for name, path in items:
        fid, ok, err = process_and_save_npz((name, path, ref_dir))
        assert ok, (fid, err)
        fid, ok, err = P.process_wav(name, path, ora_dir)
        assert ok, (fid, err)
        a = np.load(os.path.join(ref_dir, name + ".npz"))
        b = np.load(os.path.join(ora_dir, name + ".npz"))
        assert sorted(a.files) == sorted(b.files)
        for k in a.files:
            assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, (name, k)
            same = np.array_equal(a[k], b[k], equal_nan=True)
            if not same:
                d = float(np.nanmax(np.abs(a[k].astype(np.float64) - b[k].astype(np.float64))))
                worst = max(worst, d)
                print(f"MISMATCH {name} {k}: max abs diff {d:g}")
    print(f"compared {len(items)} segments x 10 arrays; worst abs diff {worst:g}")
    if worst != 0.0:
        raise SystemExit("oracle/pipeline.py is not bit-identical to the reference run over the shim")
    print("OK: restatement is bit-identical to the reference's own process.py over the shim")

    if args.write_golden:
        gold = items[: args.n_golden_real] + [it for it in items if it[0].startswith("synth_")][: args.n_golden_synth]
        out = {"names": np.array([g[0] for g in gold])}
        pcm = []
        for gi, (name, path) in enumerate(gold):
            sr, q = scipy.io.wavfile.read(path)
            assert sr == 16000 and q.dtype == np.int16 and q.shape == (16000,)
            pcm.append(q)
            a = np.load(os.path.join(ref_dir, name + ".npz"))           # the reference's own output
            for k in a.files:
                out[f"{gi}/{k}"] = a[k]
            dbg = {}
            P.segment_features(q.astype(np.float32) / np.float32(32768.0), debug=dbg)
            for k in ("mel_db", "stft512_db", "mfcc_raw", "gammatone_raw", "mod_spec_raw", "lpc_raw", "onset_env",
                      "chroma_stft_raw", "chroma_cens_raw", "flux", "centroid", "rolloff", "contrast"):
                out[f"{gi}/dbg/{k}"] = np.asarray(dbg[k], dtype=np.float32)
            out[f"{gi}/dbg/ints"] = np.array([dbg["n_peaks"], dbg["first_min_idx"]], dtype=np.int64)
            out[f"{gi}/dbg/tuning"] = np.array([dbg["tuning12"], dbg["tuning36"]], dtype=np.float64)
        out["pcm16"] = np.stack(pcm)
        dst = os.path.join(ROOT, "tests", "golden", "golden_segments.npz")
        np.savez_compressed(dst, **out)
        print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
