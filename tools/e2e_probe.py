"""Break-down of the host path (bpc_precompute_host): run with BPC_HOST_TRACE=1 and different BPC_* settings."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "breathing-phase-classifier_b200"))
import torch, bpc_b200
from bpc_b200.synth import synth_batch_pcm16
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = bpc_b200.Engine(device=0, max_batch=B)
pcm = np.tile(synth_batch_pcm16(0, 64), (B // 64, 1))
h_in = torch.from_numpy(pcm).pin_memory()
h_f = torch.empty((B, 9, 128, 63), dtype=torch.float32).pin_memory()
h_s = torch.empty((B, eng.nscal), dtype=torch.float32).pin_memory()
h_st = torch.empty((B,), dtype=torch.int32).pin_memory()
# plain D2H bandwidth of this box for reference
d = torch.empty((592, 9, 128, 63), dtype=torch.float32, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): h_f[:592].copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"plain pinned D2H of one chunk: {d.numel()*4/dt/1e9:.1f} GB/s ({dt*1e3:.2f} ms)")
for _ in range(2): eng.precompute_host(h_in.numpy(), h_f.numpy(), h_s.numpy(), h_st.numpy())
t0 = time.perf_counter()
for _ in range(3): eng.precompute_host(h_in.numpy(), h_f.numpy(), h_s.numpy(), h_st.numpy())
dt = (time.perf_counter() - t0) / 3
print(f"e2e {B/dt:.0f} seg/s ({dt*1e3:.2f} ms per call)")
