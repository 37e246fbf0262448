"""Host-path sweep on one GPU: e2e segments/s of bpc_precompute_host_compact for piece sizes / schedules (the BPC_*
environment switches are read when an Engine first uses the host path).  usage: python tools/e2e_sweep.py [B]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "breathing-phase-classifier_b200"))
import torch, bpc_b200
from bpc_b200.synth import synth_batch_pcm16
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
pcm = np.tile(synth_batch_pcm16(0, 256), (B // 256, 1))
configs = [dict(BPC_HOST_CHUNK=str(c), BPC_D2H_MODE=m, BPC_RAMP=r) for c in (148, 296, 444, 592) for m in ("contig", "2d") for r in ("1", "0")]
for cfg in configs:
    os.environ.update(cfg)
    eng = bpc_b200.Engine(device=0, max_batch=B)
    h_in = eng.host_empty(pcm.shape, np.int16); h_in[:] = pcm
    rows = eng.host_empty((B, 772, 63)); pad = eng.host_empty((B, 9)); sc = eng.host_empty((B, 36)); st = eng.host_empty((B,), np.int32)
    for _ in range(2): eng.precompute_host_compact(h_in, rows, pad, sc, st)
    t0 = time.perf_counter()
    for _ in range(5): eng.precompute_host_compact(h_in, rows, pad, sc, st)
    dt = (time.perf_counter() - t0) / 5
    print(f"{cfg}  {dt*1e3:.2f} ms  {B/dt:.0f} seg/s", flush=True)
    eng.close()
