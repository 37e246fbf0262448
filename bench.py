#!/usr/bin/env python
"""bench.py -- segments/s of the full 9-channel + 36-scalar precompute (BASELINE.json metric) on N B200s.

A "step" is one pass of the hot path over one batch of B synthetic segments per GPU (weak scaling: every rank owns its
own batch, the only collective is one all-reduce of the dataset-level channel statistics at the end of the timed region).

    python bench.py --gpus 1 --steps 5 --warmup 3            # our arm (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --steps 3 --warmup 1    # the reference CPU path (oracle port) on the host cores

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "breathing-phase-classifier_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

L = 16000
T = 63
ALG_BYTES_PER_SEG = L * 4 + 9 * 128 * T * 4 + 36 * 4          # 354,448 B
METRIC = "segments/sec full 9-ch+36-scalar precompute"
WORKLOAD = ("full 9-channel [9,128,63] + 36-scalar precompute of 1 s / 16 kHz synthetic breathing-like segments "
            "(BASELINE configs[2], per-GPU shard)")


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------- CPU (oracle port) leg
def _cpu_worker(idx_range):
    from oracle import pipeline as P          # the oracle: only executed as the CPU baseline / reference arm
    try:                                      # one process per core: keep BLAS / OpenMP from oversubscribing
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    lo, hi = idx_range
    n = 0
    for i in range(lo, hi):
        P.segment_features(P.synth_segment(i))
        n += 1
    return n


def cpu_port_throughput(n_segments: int, cores: int, pool=None):
    """Wall-clock segments/s of the oracle port (restatement of process.py:25-103) over `cores` processes."""
    import multiprocessing as mp
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(cores)
    per = max(1, n_segments // cores)
    chunks = [(k * per, (k + 1) * per) for k in range(cores)]
    t0 = time.perf_counter()
    done = sum(pool.map(_cpu_worker, chunks))
    dt = time.perf_counter() - t0
    if own:
        pool.close()
    return done / dt, done, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = host_cores()
    per_step = 4 * cores
    pool = mp.get_context("fork").Pool(cores)
    pool.map(_cpu_worker, [(0, 1)] * cores)                       # import + first-call warm-up
    for _ in range(args.warmup):
        cpu_port_throughput(per_step, cores, pool)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        _, n, _ = cpu_port_throughput(per_step, cores, pool)
        done += n
    dt = time.perf_counter() - t0
    pool.close()
    v = done / dt
    # the reference's own degree of parallelism, N_WORKERS = 2 (core.py:12,33) -- two worker
    # processes here (the reference uses two GIL-bound threads, which can only be slower)
    pool2 = mp.get_context("fork").Pool(2)
    pool2.map(_cpu_worker, [(0, 1)] * 2)
    v2, n2, dt2 = cpu_port_throughput(16, 2, pool2)
    pool2.close()
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "segments/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "segments_per_step": per_step, "seq_len": L},
        "cpu_baseline": {"value": v, "unit": "segments/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} synthetic segments per step through oracle/pipeline.py "
                                   "(numpy/scipy restatement of the reference's librosa path; the reference itself "
                                   "is Python and cannot travel to the GPU box)",
                         "n_workers_2": {"value": v2, "segments": n2,
                                         "note": "same port with the reference's N_WORKERS = 2 (core.py:12)"}},
        "e2e": {"value": v, "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms",
                                          "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import bpc_b200
    from bpc_b200.synth import synth_batch_pcm16

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    B = args.batch
    eng = bpc_b200.Engine(device=local, max_batch=B)

    # synthetic input: `distinct` seeded segments (host generator == the parity-test generator), tiled to B, PCM16
    distinct = B if args.distinct <= 0 else min(B, args.distinct)
    base = synth_batch_pcm16(rank * 100000, distinct)
    pcm = np.tile(base, ((B + distinct - 1) // distinct, 1))[:B]
    wav_f32 = (torch.from_numpy(pcm).to(dev).float() / 32768.0).contiguous()      # resident in HBM before timing
    feats = torch.empty((B, 9, 128, T), dtype=torch.float32, device=dev)
    scal = torch.empty((B, eng.nscal), dtype=torch.float32, device=dev)
    status = torch.empty((B,), dtype=torch.int32, device=dev)

    def step():
        eng.precompute(wav_f32, feats, scal, status)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    if dist is not None:                                 # warm-up of the one collective too (NCCL sets a new kind of
        from bpc_b200.stats import allreduce_stats       # collective up lazily: ~30 ms on its first call)
        allreduce_stats(eng.channel_stats_device().clone(), dist)
    barrier()
    eng.reset_stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    if dist is not None:                                 # config 3: the one collective, the exchange of the channel statistics
        allreduce_stats(eng.channel_stats_device(), dist)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0
    # Per-kernel durations for the roofline leg: the same K steps again with CUDA events around every launch.  The
    # production step above forks independent kernels onto side streams, where in-stream events of overlapping
    # kernels would not be attributable, so this instrumented pass runs the launch sequence on one stream.
    eng.kernel_times()                                   # drop anything recorded so far
    eng.set_kernel_timing(True)
    i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    i0.record()
    for _ in range(args.steps):
        step()
    i1.record()
    barrier()
    serial_ms = i0.elapsed_time(i1) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    eng.set_kernel_timing(False)
    ktimes = eng.kernel_times()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * args.steps * B / (ms_max * 1e-3)

    # ---- multi-GPU parity probe (outside every timed region): all ranks precompute the SAME 64 segments; rank 0 checks
    # that every rank produced bit-identical outputs and that 8 of them match the CPU oracle within the test gates
    probe = synth_batch_pcm16(777000, 64)
    pf, ps, pst = eng.precompute(torch.from_numpy(probe).to(dev))
    torch.cuda.synchronize()
    chk = torch.cat([pf.view(torch.int32).to(torch.int64).sum(dim=(1, 2, 3)),
                     ps.contiguous().view(torch.int32).to(torch.int64).sum(dim=1), pst.to(torch.int64)])
    chks = [chk]
    if dist is not None:
        chks = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(chks, chk)
    parity_probe = None
    if rank == 0:
        same = all(bool(torch.equal(c, chks[0])) for c in chks)
        from oracle import pipeline as P_oracle           # the checker, never the thing measured
        pf_h, ps_h = pf[:8].cpu().numpy(), ps[:8].cpu().numpy()
        worst_plane, worst_excess, ints_ok = 0.0, -1.0, 0
        for i in range(8):
            ch, sc = P_oracle.segment_features(probe[i].astype(np.float32) / np.float32(32768.0))
            ref = P_oracle.stack_sorted(ch)
            worst_plane = max(worst_plane, float(np.abs(pf_h[i] - ref).max()))
            d_abs = np.abs(ps_h[i, :36].astype(np.float64) - sc.astype(np.float64))
            worst_excess = max(worst_excess, float(np.max(d_abs - (1e-4 * np.abs(sc.astype(np.float64)) + 2e-6))))
            ints_ok += int(ps_h[i, 22] == sc[22] and ps_h[i, 35] == sc[35])
        ok = same and worst_plane < 2e-4 and worst_excess <= 0.0 and ints_ok == 8 and int(pst.abs().sum()) == 0
        parity_probe = {"result": "ok" if ok else "FAILED", "ranks_bit_identical": same, "ranks": world,
                        "probe_segments": 64, "oracle_checked": 8, "worst_plane_abs": worst_plane,
                        "scalars_inside_1e-4_rel_plus_2e-6": worst_excess <= 0.0, "integer_outputs_exact": ints_ok}
    del pf, ps, pst

    # ---- end-to-end through the reference-facing host calls: pinned host PCM16 in, pinned host float32 out.
    # Headline: bpc_precompute_host_compact -- the compact host layout ([B,772,63] data rows + [B,9] pad values, what the
    # packed shard / PackedDS consume; buffers from bpc_host_alloc, NUMA-local to the GPU).  Next to it the same call
    # into the full [B,9,128,63] tensor (bpc_precompute_host: same bytes over PCIe + a host-side fill of the pad rows).
    h_in = eng.host_empty((B, L), np.int16)
    h_in[:] = pcm
    h_rows = eng.host_empty((B, 772, T), np.float32)
    h_pad = eng.host_empty((B, 9), np.float32)
    h_scal = eng.host_empty((B, eng.nscal), np.float32)
    h_stat = eng.host_empty((B,), np.int32)
    e2e_steps = max(1, min(args.steps, 5))                 # the two comparison legs (synchronous call, full layout)
    stream_n = max(1, min(args.steps, 20))                 # the headline leg runs the K steps of the run (at most 20)

    def timed_host(fn):
        fn()                                                              # warm-up (allocates the slots)
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    dt_sync = timed_host(lambda: eng.precompute_host_compact(h_in, h_rows, h_pad, h_scal, h_stat))
    e2e_sync = world * e2e_steps * B / dt_sync
    # Headline: the streaming form of the same call (bpc_precompute_host_compact_begin / bpc_host_wait), the way a caller
    # works through a dataset batch after batch -- step k + 1 is enqueued before step k is waited for, on two sets of
    # pinned output buffers, so the head and the tail of a call overlap with its neighbours.  Every step's H2D and D2H still
    # happen inside the timed region, and the last wait is inside it too.
    out_sets = [(h_rows, h_pad, h_scal, h_stat),
                (eng.host_empty((B, 772, T), np.float32), eng.host_empty((B, 9), np.float32),
                 eng.host_empty((B, eng.nscal), np.float32), eng.host_empty((B,), np.int32))]

    def stream_steps(n_steps):
        prev = None
        for k in range(n_steps):
            tk = eng.precompute_host_compact_begin(h_in, *out_sets[k % 2])
            if prev is not None:
                eng.host_wait(prev)
            prev = tk
        eng.host_wait(prev)

    stream_steps(2)                                                        # warm-up (second buffer set gets touched)
    barrier()
    t0 = time.perf_counter()
    stream_steps(stream_n)
    torch.cuda.synchronize()
    tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt_c = float(tt.item())
    e2e_value = world * stream_n * B / dt_c
    stream_check = bool(np.array_equal(out_sets[0][0][:64], out_sets[1][0][:64]) and np.array_equal(out_sets[0][2], out_sets[1][2]))
    d2h_step = int(B * (772 * T * 4 + 9 * 4 + eng.nscal * 4 + 4))
    # the compact result must be the full-layout result: expand 16 segments on the host and compare with the device path
    e2e_check = bool(np.array_equal(bpc_b200.expand_compact(h_rows[:16], h_pad[:16]), feats[:16].cpu().numpy())
                     and np.array_equal(h_scal[:16], scal[:16].cpu().numpy()))
    h_feats = eng.host_empty((B, 9, 128, T), np.float32)
    dt_f = timed_host(lambda: eng.precompute_host(h_in, h_feats, h_scal, h_stat))
    e2e_full = world * e2e_steps * B / dt_f
    # concurrent D2H ceiling of this box, measured live: every rank copies the same bytes a step moves, no kernels
    d_dummy = torch.empty(d2h_step // 4, dtype=torch.float32, device=dev)
    h_dummy = torch.from_numpy(h_feats.reshape(-1)[: d2h_step // 4])
    for _ in range(2):
        h_dummy.copy_(d_dummy, non_blocking=True)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(5):
        h_dummy.copy_(d_dummy, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    tt = torch.tensor([c0.elapsed_time(c1) / 5], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    d2h_ceiling_gbs = world * d2h_step / (float(tt.item()) * 1e-3) / 1e9
    d2h_gbs = world * stream_n * d2h_step / dt_c / 1e9
    del d_dummy

    # ---- BASELINE configs[3]: long-segment sweep (1 s - 30 s), equal total samples per point (B / d segments of d seconds
    # per GPU).  Every rank runs it on its own shard (no data-path collective: segments are independent at every length);
    # a point's time is the MAX over ranks of the CUDA-event time, its throughput the whole job's.  The collective that
    # combines the ranks sits outside the try block, so a failing rank cannot leave the others waiting.
    sweep_all = None
    if not args.no_extras:
        lens = (1, 2, 5, 10, 30)
        sw = torch.zeros(len(lens) + 1, dtype=torch.float64, device=dev)      # ms per point + a failure flag
        frames = {}
        err = None
        try:
            for i, d in enumerate(lens):
                Bd = max(1, B // d)
                e_d = eng if d == 1 else bpc_b200.Engine(device=local, max_batch=Bd,
                                                         params=bpc_b200.default_params(expected_len=L * d))
                wd = wav_f32[:Bd * d].reshape(Bd, d * L).contiguous()
                fd = torch.empty((Bd, 9, 128, e_d.T), dtype=torch.float32, device=dev)
                sd = torch.empty((Bd, e_d.nscal), dtype=torch.float32, device=dev)
                td = torch.empty((Bd,), dtype=torch.int32, device=dev)
                e_d.precompute(wd, fd, sd, td)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(2):
                    e_d.precompute(wd, fd, sd, td)
                b.record()
                torch.cuda.synchronize()
                sw[i] = a.elapsed_time(b) / 2
                frames[d] = e_d.T
                if d != 1:
                    e_d.close()
                del fd, sd, td, wd
        except Exception as e:
            err = repr(e)
            sw[len(lens)] = 1.0
        if dist is not None:
            dist.all_reduce(sw, op=dist.ReduceOp.MAX)
        sw_h = sw.cpu().tolist()
        if sw_h[len(lens)] > 0:
            sweep_all = {"error": err or "another rank failed"}
        else:
            sweep_all = {"ranks": world}
            for i, d in enumerate(lens):
                Bd = max(1, B // d)
                sweep_all[f"{d}s"] = {"segments": world * Bd, "frames": frames[d], "ms": sw_h[i],
                                      "segments_per_s": world * Bd / (sw_h[i] * 1e-3),
                                      "audio_seconds_per_s": world * Bd * d / (sw_h[i] * 1e-3),
                                      "alg_bytes_per_segment": 4 * L * d + 9 * 128 * frames[d] * 4 + 144}

    # ---- side measurements (not the headline): BASELINE configs[1] stage and the HBM-bound batch-assembly kernel
    extras = {}
    if rank == 0 and not args.no_extras:
        def timed(fn, reps):
            fn(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        ms2 = timed(lambda: eng.stage_logmel(wav_f32, want_stft=True), 3)
        bytes2 = L * 4 + 257 * T * 4 + 3 * 128 * T * 4                       # 225,532 B
        extras["config2_logmel"] = {"segments_per_s": B / (ms2 * 1e-3), "ms_per_step": ms2, "batch": B,
                                    "alg_bytes_per_segment": bytes2,
                                    "achieved_gbs": B * bytes2 / (ms2 * 1e-3) / 1e9,
                                    "outputs": "log-power STFT [B,257,63] + mel/mel_delta/mel_delta2 [B,3,128,63]"}
        from bpc_b200.resident import ResidentDS
        ds = ResidentDS(eng, feats, scal, torch.zeros(B, device=dev))
        gen = torch.Generator().manual_seed(0)
        nb = min(B, 2048)
        idx = torch.randperm(B, generator=gen)[:nb].to(dev)
        perm = torch.randperm(nb, generator=gen)
        seg_bytes = 9 * 128 * T * 4
        for name, kw, streams in (("gather", {}, 2), ("mixup", {"mix": "mixup", "perm": perm}, 3),
                                  ("cutmix", {"mix": "cutmix", "perm": perm}, 2)):
            msb = timed(lambda: ds.batch(idx, rng=np.random.RandomState(1), **kw), 10)
            extras["collate_" + name] = {"batch": nb, "ms": msb, "segments_per_s": nb / (msb * 1e-3),
                                         "alg_bytes": nb * seg_bytes * streams,
                                         "achieved_gbs": nb * seg_bytes * streams / (msb * 1e-3) / 1e9}
        del ds
        # BASELINE configs[4]: precompute of 1024 segments feeding CNN8 / VGG forward (eval mode, random-init weights of
        # the reference architectures rebuilt in tools/consumer_models.py; fp32 and bf16-autocast forwards)
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import consumer_models as CM
            nb5 = min(B, 1024)
            w5 = wav_f32[:nb5].contiguous()
            f5, s5, _ = eng.precompute(w5)
            ms_pre = timed(lambda: eng.precompute(w5, feats[:nb5], scal[:nb5], status[:nb5]), 5)
            c5 = {"batch": nb5, "precompute_ms": ms_pre, "precompute_segments_per_s": nb5 / (ms_pre * 1e-3)}
            for name in ("CNN8", "VGG"):
                net = getattr(CM, name)(9, eng.nscal).to(dev).eval()
                with torch.no_grad():
                    ms_f = timed(lambda: net(f5, s5), 5)
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        ms_a = timed(lambda: net(f5, s5), 5)
                    ok = bool(torch.isfinite(net(f5, s5)).all())
                c5[name] = {"params": CM.n_params(net), "forward_ms_fp32": ms_f, "forward_ms_bf16_autocast": ms_a,
                            "finite": ok}
                del net
            extras["config5_precompute_plus_forward"] = c5
        except Exception as e:                           # the side measurement must not take the bench line down
            extras["config5_precompute_plus_forward"] = {"error": repr(e)}

        # BASELINE configs[3]: the long-segment sweep measured above on every rank
        if sweep_all is not None:
            extras["config4_long_segment_sweep"] = sweep_all

        # BASELINE configs[0] through OUR entry point: wav files + CSV rows -> process_dataset_threaded -> .npz files
        # (reader pool -> bpc_precompute_host -> writer pool), next to the same rows into one packed shard
        try:
            import contextlib, shutil, tempfile
            import pandas as pd
            import scipy.io.wavfile
            from bpc_b200.precompute import core as CO
            nf = 1024
            root = tempfile.mkdtemp(prefix="bpc_cfg1_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            try:
                wdir = os.path.join(root, "train"); os.makedirs(wdir)
                ids = [f"steth_{i:05d}_{'EI'[i % 2]}_001" for i in range(nf)]
                for i, fid in enumerate(ids):
                    scipy.io.wavfile.write(os.path.join(wdir, CO.wav_name_for(fid, "train")), 16000, pcm[i % len(pcm)])
                df = pd.DataFrame({"ID": ids, "Target": ["EI"[i % 2] for i in range(nf)]})
                c1 = {"files": nf, "storage": "tmpfs" if root.startswith("/dev/shm") else "tmp"}
                for mode, packed in (("npz", False), ("packed_shard", True)):
                    out = os.path.join(root, "out_" + mode); os.makedirs(out)
                    with contextlib.redirect_stdout(sys.stderr):
                        CO.process_dataset_threaded(df.iloc[:64], wdir, out, "train", engine=eng, packed=packed)   # warm
                        t0 = time.perf_counter()
                        res = CO.process_dataset_threaded(df, wdir, out, "train", engine=eng, packed=packed)
                        dtf = time.perf_counter() - t0
                    c1[mode] = {"files_per_s": nf / dtf, "ok": int(sum(1 for r in res if r[1]))}
                extras["config1_files_through_process_dataset_threaded"] = c1
            finally:
                shutil.rmtree(root, ignore_errors=True)
        except Exception as e:
            extras["config1_files_through_process_dataset_threaded"] = {"error": repr(e)}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
        top = max(ktimes.items(), key=lambda kv: kv[1][0]) if ktimes else ("none", (0.0, 1))
        # Counters of the committed ncu capture (tools/ncu_counters.py -> profiles/kernel_counters.json): DRAM bytes and
        # FP64 thread instructions per segment and kernel.  They are properties of the code; the rates below combine
        # them with the CUDA-event times of THIS run.
        counters = {}
        try:
            counters = json.load(open(os.path.join(ROOT, "profiles", "kernel_counters.json")))
        except Exception:
            pass
        ck = counters.get("kernels", {})
        timing_ids = {"k_stft512": ["k_stft512"], "k_spec512_consumers": ["k_spec512_consumers", "k_spec512_light"],
                      "k_frame2048": ["k_frame2048"], "k_even2048": ["k_even2048"], "k_cens": ["k_cens"], "k_cens_dec": ["k_cens_dec"], "k_cens_lo": ["k_cens_lo"],
                      "k_time_basic+k_autocorr": ["k_time_basic", "k_autocorr"], "k_hilbert": ["k_hilbert"],
                      "k_lpc": ["k_lpc", "k_lpc_fast", "k_lpc_redo"], "k_stats": ["k_stats_scalars", "k_stats"], "k_seg2048": ["k_seg2048"],
                      "k_ingest": ["k_ingest"]}
        segs_per_launch = min(B, eng.chunk)
        traffic = None
        parts = [ck[k]["dram_bytes_per_segment"] for k in timing_ids.get(top[0], []) if k in ck]
        if parts:
            traffic = float(sum(parts)) * segs_per_launch
        sm_clock_hz = 1e6 * float((clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0)
        fp64_peak_inst = 148 * 64 * sm_clock_hz                       # FP64 lanes: 64 thread instructions / clk / SM
        fp64 = None
        if ck:
            seg_per_s_gpu = value / world
            per_kernel = {}
            for tname, (tms, tcnt) in ktimes.items():
                names = [k for k in timing_ids.get(tname, []) if k in ck]
                if not names or tms <= 0:
                    continue
                flop = sum(ck[k]["fp64_flop_per_segment"] for k in names)
                inst = sum(sum(ck[k]["fp64_thread_inst_per_segment"].values()) for k in names)
                sec = tms / args.steps * 1e-3
                per_kernel[tname] = {"fp64_flop_per_segment": flop, "ms_per_step": tms / args.steps,
                                     "achieved_tflops": flop * B / sec / 1e12,
                                     "frac_of_fma_peak": flop * B / sec / (2 * fp64_peak_inst),
                                     "pipe_busy_frac": inst * B / sec / fp64_peak_inst}
            tot_flop = counters.get("total_fp64_flop_per_segment", 0.0)
            tot_inst = counters.get("total_fp64_inst_per_segment", 0.0)
            fp64 = {"bound": "fp64 pipe (the roofline that binds this step: every FFT, the Burg recursion and the "
                             "decimator accumulate in FP64 because librosa / scipy do)",
                    "flop_per_segment": tot_flop, "thread_inst_per_segment": tot_inst,
                    "achieved_tflops": tot_flop * seg_per_s_gpu / 1e12,
                    "peak_tflops": 2 * fp64_peak_inst / 1e12,
                    "frac": tot_flop * seg_per_s_gpu / (2 * fp64_peak_inst),
                    "pipe_busy_frac": tot_inst * seg_per_s_gpu / fp64_peak_inst,
                    "peak_source": f"148 SMs x 64 FP64 lanes x {sm_clock_hz / 1e6:.0f} MHz x 2 (FMA); pipe_busy counts every "
                                   "FP64 thread instruction (DFMA, DMUL or DADD) as one lane-slot",
                    "counters_source": counters.get("source"), "per_kernel": per_kernel}
        total_k = sum(v[0] for v in ktimes.values()) or 1.0
        avg_ms = top[1][0] / max(1, top[1][1])
        achieved = ALG_BYTES_PER_SEG * segs_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        step_gbs = value / world * ALG_BYTES_PER_SEG / 1e9
        cores = host_cores()
        cpu = None
        if world == 1 and not args.no_cpu:
            import multiprocessing as mp
            n_cpu = max(32 * cores, 256)                     # ~20-40 core-seconds of the oracle port
            pool = mp.get_context("fork").Pool(cores)
            pool.map(_cpu_worker, [(0, 1)] * cores)          # imports + first-call warm-up outside the timed sample
            v, n, dtc = cpu_port_throughput(n_cpu, cores, pool)
            pool.close()
            cpu = {"value": v, "unit": "segments/s", "cores": cores, "kind": "port",
                   "sample": f"{n} synthetic segments of the same generator through oracle/pipeline.py "
                             f"(numpy/scipy restatement of the reference librosa path), {dtc:.1f} s wall"}
        line = {
            "metric": METRIC, "value": value, "unit": "segments/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "segments_per_step_per_gpu": B, "seq_len": L, "frames": T,
                       "distinct_segments": distinct, "input_dtype_resident": "f32",
                       "segments_total_timed": int(world * args.steps * B),
                       "l2_policy": f"inputs ({B * L * 4 / 1e6:.0f} MB) and outputs ({B * 9 * 128 * T * 4 / 1e6:.0f} MB) "
                                    "per step exceed the 126 MB L2",
                       "parallelism": f"dp{world} (batch-sharded, one NCCL all-reduce of channel statistics)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "kernel": top[0],
                         "kernel_avg_ms": avg_ms, "kernel_share_of_step": top[1][0] / total_k,
                         "segments_per_launch": segs_per_launch, "alg_bytes_per_segment": ALG_BYTES_PER_SEG,
                         "peak_source": peak_src,
                         "whole_step": {"achieved": step_gbs, "frac": step_gbs / peak},
                         "traffic_total": (counters.get("total_dram_bytes_per_segment", 0.0) * B) if ck else None,
                         "traffic_total_over_algorithmic": (counters.get("total_dram_bytes_per_segment", 0.0) / ALG_BYTES_PER_SEG) if ck else None,
                         "fp64": fp64,
                         "kernel_timing": "instrumented single-stream pass of the same steps, CUDA events around "
                                          "every launch; the timed step overlaps independent kernels on side streams",
                         "single_stream_ms_per_step": serial_ms,
                         "kernel_ms_per_step": {k: v[0] / args.steps for k, v in ktimes.items()}},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "segments/s", "h2d_bytes_per_step": int(B * L * 2),
                    "d2h_bytes_per_step": d2h_step, "steps": stream_n,
                    "input": "pinned host PCM16 [B,16000] -> bpc_precompute_host_compact -> pinned host float32 rows "
                             "[B,772,63] + pad [B,9] + scalars [B,36] + status [B] (the compact host layout of "
                             "include/bpc.h: every data row of the nine planes + one pad_freq constant per plane; "
                             "bpc_expand_compact / PackedDS rebuild [9,128,63] bit-identically)",
                    "host_buffers": f"bpc_host_alloc (NUMA node {getattr(eng, 'host_numa_node', -1)}; -1 = single-node box)",
                    "d2h_gbs": d2h_gbs, "d2h_ceiling_gbs": d2h_ceiling_gbs,
                    "d2h_frac_of_ceiling": d2h_gbs / d2h_ceiling_gbs if d2h_ceiling_gbs else None,
                    "d2h_ceiling_how": "all ranks copy d2h_bytes_per_step device->pinned host concurrently, 5 times, no "
                                       "kernels running; aggregate bytes / slowest rank (tools/d2h_ceiling.py is the long form)",
                    "matches_device_path": e2e_check and stream_check,
                    "call": "bpc_precompute_host_compact_begin(step k + 1) then bpc_host_wait(step k): two sets of pinned "
                            "output buffers, every step's copies and the final wait inside the timed region",
                    "synchronous_call": {"value": e2e_sync, "unit": "segments/s", "steps": e2e_steps,
                                         "call": "bpc_precompute_host_compact, one blocking call per step"},
                    "full_layout": {"value": e2e_full, "unit": "segments/s", "steps": e2e_steps,
                                    "call": "bpc_precompute_host -> pinned host float32 [B,9,128,63]: the same bytes over "
                                            "PCIe, the 380 constant pad rows per segment written by the library's host threads"}},
            "parity_probe": parity_probe,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "extras": extras,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="segments per step per GPU")
    ap.add_argument("--distinct", type=int, default=0,
                    help="distinct synthetic segments tiled to the batch (0 = every segment of the batch is distinct)")
    ap.add_argument("--no-extras", action="store_true", help="skip the config-2 / collate side measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (used under ncu only)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
