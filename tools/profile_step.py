"""Small fixed workload for ncu: `steps` passes of the full path over one chunk-sized batch (default 592 segments)."""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "breathing-phase-classifier_b200"))
import torch, bpc_b200
from bpc_b200.synth import synth_batch_pcm16
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=592); ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
eng = bpc_b200.Engine(device=0, max_batch=a.batch)
base = synth_batch_pcm16(0, 64)
wav = (torch.from_numpy(np.tile(base, ((a.batch + 63) // 64, 1))[:a.batch]).cuda().float() / 32768.0).contiguous()
for _ in range(a.steps):
    f, s, st = eng.precompute(wav)
torch.cuda.synchronize()
print("ok", eng.launch_count())          # (no torch reduction here: it would show up in the capture with 1.2 GB of reads)
