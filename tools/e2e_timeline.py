"""Per-piece device timeline of one bpc_precompute_host_compact call (BPC_HOST_TRACE=2).  usage: python tools/e2e_timeline.py [B]"""
import os, sys
os.environ["BPC_HOST_TRACE"] = "2"
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "breathing-phase-classifier_b200"))
import torch, bpc_b200
from bpc_b200.synth import synth_batch_pcm16
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
pcm = np.tile(synth_batch_pcm16(0, 256), (B // 256, 1))
eng = bpc_b200.Engine(device=0, max_batch=B)
h_in = eng.host_empty(pcm.shape, np.int16); h_in[:] = pcm
rows = eng.host_empty((B, 772, 63)); pad = eng.host_empty((B, 9)); sc = eng.host_empty((B, 36)); st = eng.host_empty((B,), np.int32)
for i in range(3):
    print(f"--- call {i}", file=sys.stderr, flush=True)
    eng.precompute_host_compact(h_in, rows, pad, sc, st)
