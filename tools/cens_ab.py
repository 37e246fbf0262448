"""A/B of CQT / chroma_cens / LPC variants on the GPU box: every variant runs in its own process and dumps the chroma plane
the raw chroma_cens rows, the LPC plane and the raw LPC coefficients of 130 one-second segments (+ 6 five-second segments with --long); the first variant is the
reference the others are compared with, value by value.

usage: python tools/cens_ab.py [--long] VARIANT [VARIANT ...]
  VARIANT = name[:KEY=VAL[,KEY=VAL ...]]   environment of the child; LIB=<path> loads that library instead of the
                                           in-tree libbpc_b200.so (e.g. gpurun_variants/lib_prev.so)
  e.g.  python tools/cens_ab.py fft7:BPC_CENS_LO=0,LIB=gpurun_variants/lib_prev.so new new_nolo:BPC_CENS_LO=0
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "breathing-phase-classifier_b200"))


def child(out, long_mode):
    import torch
    import bpc_b200
    from bpc_b200 import _lib
    if os.environ.get("BPC_AB_LIB"):
        _lib.LIB_PATH = os.path.join(ROOT, os.environ["BPC_AB_LIB"])
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
    res = dict(chroma=feats[:, 0].cpu().numpy(), raw=eng.debug("chroma_cens_raw", len(Y)),
               lpc=feats[:, 2].cpu().numpy(), lpc_raw=eng.debug("lpc_raw", len(Y)))
    if long_mode:
        d = 5
        YL = np.stack([np.concatenate([ys[(7 * i + j) % 128] for j in range(d)]) for i in range(6)])
        engl = bpc_b200.Engine(device=0, max_batch=len(YL), params=_lib.default_params(expected_len=16000 * d))
        fl, sl, stl = engl.precompute(torch.from_numpy(YL).cuda())
        torch.cuda.synchronize()
        res["chroma_long"] = fl[:, 0].cpu().numpy()
    np.savez(out, **res)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2], sys.argv[3] == "1")
        sys.exit(0)
    args = sys.argv[1:]
    long_mode = "--long" in args
    variants = [a for a in args if a != "--long"]
    res = {}
    for v in variants:
        name, _, envs = v.partition(":")
        env = dict(os.environ)
        for kv in filter(None, envs.split(",")):
            k, _, val = kv.partition("=")
            env["BPC_AB_LIB" if k == "LIB" else k] = val
        out = f"/tmp/cens_ab_{name}.npz"
        subprocess.run([sys.executable, __file__, "--child", out, "1" if long_mode else "0"], check=True, env=env)
        res[name] = np.load(out)
    ref = variants[0].partition(":")[0]
    for v in variants[1:]:
        name = v.partition(":")[0]
        for k in res[ref].files:
            a, b = res[name][k], res[ref][k]
            same = (a == b) | (np.isnan(a) & np.isnan(b))
            d = np.abs(a.astype(np.float64) - b.astype(np.float64))
            d = np.where(np.isnan(d), 0.0, d)
            per_seg = (~same).reshape(len(a), -1).sum(1)
            print(f"{name} vs {ref} {k}: {int((~same).sum())} of {a.size} values differ (segments touched: "
                  f"{int((per_seg > 0).sum())} of {len(a)}), max |delta| {d.max():.3e}, "
                  f"nan mismatch {int((np.isnan(a) != np.isnan(b)).sum())}")
