"""Small fixed workload for ncu: BASELINE configs[1] (bpc_stage_logmel: log-power STFT + mel / delta / delta2) on 4096 segments."""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "breathing-phase-classifier_b200")]
import bpc_b200
from bpc_b200.synth import synth_batch_pcm16
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=4096); ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--pcm", type=int, default=0)
a = ap.parse_args()
eng = bpc_b200.Engine(device=0, max_batch=a.batch)
base = synth_batch_pcm16(0, 64)
pcm = torch.from_numpy(np.tile(base, ((a.batch + 63) // 64, 1))[:a.batch]).cuda()
wav = pcm if a.pcm else (pcm.float() / 32768.0).contiguous()
for _ in range(a.steps):
    eng.stage_logmel(wav, want_stft=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    eng.stage_logmel(wav, want_stft=True)
e1.record(); torch.cuda.synchronize()
print("config2 ms per step", e0.elapsed_time(e1) / a.steps, "pcm" if a.pcm else "f32")
