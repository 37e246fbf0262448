"""A/B of the CQT low-octave path: k_cens_lo (sliding DFT, BPC_CENS_LO=1) against all seven octaves by FFT (=0).

Run on the GPU box:  python tools/cens_ab.py
Each variant runs in its own process (the switch is read once per process); prints how many chroma_cens values differ
between the two and by how much.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "breathing-phase-classifier_b200"))


def child(out):
    import torch
    import bpc_b200
    from oracle import pipeline as P
    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden_segments.npz"))["pcm16"]
    real = np.load(os.path.join(ROOT, "tests", "golden", "real_inputs_pcm16.npz"))["pcm16"]
    ys = [q.astype(np.float32) / np.float32(32768.0) for q in list(gold) + list(real)]
    ys += [P.synth_segment(1000 + i) for i in range(64)]
    ys += [np.zeros(16000, np.float32)]
    click = np.zeros(16000, np.float32); click[8000] = 1.0
    ys += [click]
    Y = np.stack(ys)
    eng = bpc_b200.Engine(device=0, max_batch=len(Y), debug=True)
    feats, scal, status = eng.precompute(torch.from_numpy(Y).cuda())
    torch.cuda.synchronize()
    np.savez(out, chroma=feats[:, 0].cpu().numpy(), raw=eng.debug("chroma_cens_raw", len(Y)))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
        sys.exit(0)
    res = {}
    for v in ("1", "0"):
        out = f"/tmp/cens_ab_{v}.npz"
        subprocess.run([sys.executable, __file__, out], check=True, env=dict(os.environ, BPC_CENS_LO=v))
        res[v] = np.load(out)
    for k in ("raw", "chroma"):
        a, b = res["1"][k], res["0"][k]
        same = (a == b) | (np.isnan(a) & np.isnan(b))
        d = np.abs(a.astype(np.float64) - b.astype(np.float64))
        d = np.where(np.isnan(d), 0.0 if True else 0.0, d)
        per_seg = (~same).reshape(len(a), -1).sum(1)
        print(f"{k}: {int((~same).sum())} of {a.size} values differ (segments touched: {int((per_seg > 0).sum())} of {len(a)}), "
              f"max |delta| {d.max():.3e}, nan mismatch {int((np.isnan(a) != np.isnan(b)).sum())}")
