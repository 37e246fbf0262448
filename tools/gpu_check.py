"""Per-stage parity report: CUDA path (through the C ABI) vs the CPU oracle on golden + synthetic segments.

Run on the GPU box:  python tools/gpu_check.py [--n-synth 8]
Prints one line per stage with the worst absolute / relative deviation; exits non-zero if a gate fails.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "breathing-phase-classifier_b200"))


def compare(n_synth=8, verbose=True, real_inputs=False):
    import torch
    import bpc_b200
    from oracle import pipeline as P

    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden_segments.npz"))
    pcm = gold["pcm16"]
    if real_inputs:                                     # 56 more real fixture segments, inputs only (tests/golden/README.md)
        pcm = np.load(os.path.join(ROOT, "tests", "golden", "real_inputs_pcm16.npz"))["pcm16"]
    ys = [q.astype(np.float32) / np.float32(32768.0) for q in pcm]
    ys += [P.synth_segment(1000 + i) for i in range(n_synth)]
    Y = np.stack(ys)
    B = len(Y)

    eng = bpc_b200.Engine(device=0, max_batch=max(B, 16), debug=True)
    wav = torch.from_numpy(Y).cuda()
    feats, scal, status = eng.precompute(wav)
    torch.cuda.synchronize()
    feats = feats.cpu().numpy(); scal = scal.cpu().numpy(); status = status.cpu().numpy()
    dbg = {k: eng.debug(k, B) for k in ("mel_db", "mfcc_raw", "gammatone_raw", "mod_spec_raw", "chroma_stft_raw",
                                         "chroma_cens_raw", "lpc_raw", "onset_env", "tuning", "ints", "mag512")}
    rows = []
    worst = {}

    def upd(name, a, b):
        a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
        d = np.abs(a - b)
        bad = ~np.isfinite(d) & ~(np.isnan(a) & np.isnan(b))
        d = np.where(np.isnan(a) & np.isnan(b), 0.0, d)
        err = float(np.max(np.where(bad, np.inf, d))) if d.size else 0.0
        worst[name] = max(worst.get(name, 0.0), err)
        return err

    edges = np.linspace(-0.5, 0.5, 101)
    tun_ok = [0, 0]
    scal_rel = np.zeros(36)
    scal_excess = np.full(36, -np.inf)          # max over segments of |d| - (1e-4 |ref| + 2e-6): <= 0 means inside the gate
    ints_ok = 0
    for i in range(B):
        d = {}
        ch, sc = P.segment_features(Y[i], debug=d)
        ref = P.stack_sorted(ch)
        t12 = int(np.argmin(np.abs(edges[:100] - d["tuning12"]))); t36 = int(np.argmin(np.abs(edges[:100] - d["tuning36"])))
        ok12 = t12 == dbg["tuning"][i, 0]; ok36 = t36 == dbg["tuning"][i, 1]
        tun_ok[0] += ok12; tun_ok[1] += ok36
        upd("stft512_mag", dbg["mag512"][i][:, :257].T, d["stft512_mag"])
        upd("mel_db", dbg["mel_db"][i], d["mel_db"])
        upd("mfcc_raw", dbg["mfcc_raw"][i], d["mfcc_raw"])
        upd("gammatone_raw", dbg["gammatone_raw"][i], d["gammatone_raw"])
        upd("mod_spec_raw", dbg["mod_spec_raw"][i], d["mod_spec_raw"])
        upd("lpc_raw", dbg["lpc_raw"][i], d["lpc_raw"])
        upd("onset_env", dbg["onset_env"][i], d["onset_env"])
        if ok12:
            upd("chroma_stft_raw", dbg["chroma_stft_raw"][i], d["chroma_stft_raw"])
        if ok36:
            upd("chroma_cens_raw", dbg["chroma_cens_raw"][i], d["chroma_cens_raw"])
        for c, key in enumerate(P.SORTED_KEYS):
            if key == "chroma" and not (ok12 and ok36):
                continue
            upd("ch:" + key, feats[i, c], ref[c])
        ints_ok += int(dbg["ints"][i, 0] == d["n_peaks"] and dbg["ints"][i, 1] == d["first_min_idx"])
        with np.errstate(divide="ignore", invalid="ignore"):
            rel = np.abs(scal[i, :36].astype(np.float64) - sc.astype(np.float64)) / np.maximum(np.abs(sc.astype(np.float64)), 1e-30)
        rel = np.where(np.isnan(sc) & np.isnan(scal[i, :36]), 0.0, rel)
        rel = np.where(sc == scal[i, :36], 0.0, rel)
        scal_rel = np.maximum(scal_rel, rel)
        ad = np.abs(scal[i, :36].astype(np.float64) - sc.astype(np.float64))
        ad = np.where(np.isnan(sc) & np.isnan(scal[i, :36]), 0.0, ad)
        ad = np.where(np.isfinite(ad), ad, np.inf)
        scal_excess = np.maximum(scal_excess, ad - (1e-4 * np.abs(np.nan_to_num(sc.astype(np.float64))) + 2e-6))
        if verbose and i < 2:
            print(f"seg {i}: status={status[i]} tuning gpu={dbg['tuning'][i].tolist()} ref=[{t12},{t36}] "
                  f"ints gpu={dbg['ints'][i].tolist()} ref=[{d['n_peaks']},{d['first_min_idx']}]")
            print("  scal gpu", np.array2string(scal[i, :36], precision=5, max_line_width=200))
            print("  scal ref", np.array2string(sc, precision=5, max_line_width=200))
    if verbose:
        for k, v in worst.items():
            print(f"{k:20s} max abs err {v:.3e}")
        print("scalar max rel err per index:")
        print(np.array2string(scal_rel, precision=2, max_line_width=200))
        print("scalar max(|d| - (1e-4 |ref| + 2e-6)) per index (<= 0 passes):")
        print(np.array2string(scal_excess, precision=2, max_line_width=200))
        print(f"tuning agreement: 12-bpo {tun_ok[0]}/{B}, 36-bpo {tun_ok[1]}/{B};  integer outputs exact: {ints_ok}/{B}")
    return dict(worst=worst, scal_rel=scal_rel, scal_excess=scal_excess, tun_ok=tun_ok, ints_ok=ints_ok, B=B, status=status)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-synth", type=int, default=8)
    a = ap.parse_args()
    compare(a.n_synth)
