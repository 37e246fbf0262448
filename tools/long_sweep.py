"""Per-kernel times of the long mode: python tools/long_sweep.py d [segments]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "breathing-phase-classifier_b200"))
import torch, bpc_b200
from bpc_b200.synth import synth_batch_pcm16
d = int(sys.argv[1]); B = int(sys.argv[2]) if len(sys.argv) > 2 else max(1, 4096 // d)
eng = bpc_b200.Engine(device=0, max_batch=B, params=bpc_b200.default_params(expected_len=16000 * d))
base = synth_batch_pcm16(0, 64 * d if 64 * d <= 1920 else 1920)
need = B * d
pcm = np.tile(base, ((need + len(base) - 1) // len(base), 1))[:need].reshape(B, d * 16000)
wav = (torch.from_numpy(pcm).cuda().float() / 32768.0).contiguous()
for _ in range(2): eng.precompute(wav)
torch.cuda.synchronize()
eng.kernel_times(); eng.set_kernel_timing(True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); eng.precompute(wav); b.record(); torch.cuda.synchronize()
print(f"d={d} B={B} chunk={eng.chunk} T={eng.T}: {a.elapsed_time(b):.2f} ms, {B / a.elapsed_time(b) * 1e3:.0f} seg/s, {B * d / a.elapsed_time(b) * 1e3:.0f} audio-s/s")
for k, v in eng.kernel_times().items(): print(f"   {k:28s} {v[0]:9.3f} ms  ({v[1]} launches)")
