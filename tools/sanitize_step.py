"""Small workload that touches every kernel of the path, for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_step.py
1 s engine (float32 aligned input, ragged PCM16 through k_ingest, a silent and a constant segment so that k_lpc_redo
runs), BASELINE configs[1] (k_logmel_fused), the 2 s long mode, the wav decoder and the resampler."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "breathing-phase-classifier_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, bpc_b200
from bpc_b200.synth import synth_batch_pcm16
import wavutil as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
pcm = synth_batch_pcm16(0, B)
pcm[1] = 0; pcm[2] = 8000
eng = bpc_b200.Engine(device=0, max_batch=B)
wav = (torch.from_numpy(pcm).cuda().float() / 32768.0).contiguous()
f, s, st = eng.precompute(wav)
f2, s2, st2 = eng.precompute(torch.from_numpy(pcm[:, :15000].copy()).cuda())
eng.stage_logmel(wav)
eng.stage_logmel(torch.from_numpy(pcm).cuda())
fh, sh, sth = eng.precompute_host(pcm)
imgs = [W.image(k, W.samples(k, 9000, c, 3), 16000) for k in W.FMT for c in (1, 2)] + [W.image("pcm16", W.samples("pcm16", 9000, 1, 4), 22050)]
y, errs = eng.decode_wavs(imgs)
engl = bpc_b200.Engine(device=0, max_batch=4, params=bpc_b200.default_params(expected_len=32000))
fl, sl, stl = engl.precompute(torch.from_numpy(np.concatenate([pcm[:4], pcm[4:8]], axis=1)).cuda())
torch.cuda.synchronize()
print("ok", bool(torch.isfinite(f[0]).all()), bool(np.isfinite(fh[0]).all()), bool(torch.isfinite(fl).all()), errs.count(None), eng.launch_count())
