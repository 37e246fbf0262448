"""Concurrent device->host copy ceiling of a box: what bpc_precompute_host's e2e throughput is bounded by.

    python tools/d2h_ceiling.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
        tools/d2h_ceiling.py                                      # N ranks copying at the same time

No kernels run.  Every rank copies the bytes one 4096-segment step of the host path moves (772 data rows x 63 frames x
4 B + 9 pad values per segment = 798 MB) from its GPU into pinned host memory, K times, all ranks released from one
barrier; CUDA events per rank, aggregate = total bytes / slowest rank.  Variants:

    default     pinned buffer from cudaHostAlloc as the caller's thread happens to be placed
    numa        the thread is first bound (sched_setaffinity + set_mempolicy) to the NUMA node the GPU's PCIe root
                hangs off, then the buffer is allocated and touched
    pieces      the same bytes as 14 pieces of 296 segments on one stream (the granularity of the pipeline)
    strided     the six cudaMemcpy2DAsync row runs per piece of round 1 (full [B,9,128,63] host layout)
    h2d         host->device of the step's PCM16 input (131 MB) alone
    bidir       D2H of the output with the H2D of the next step's input in flight on a second stream

It also prints, per rank, the GPU's PCI address, its NUMA node / local CPUs (/sys/bus/pci/devices/*/numa_node) and the
node the pinned pages landed on (move_pages), which is what decides between "the box" and "the code".
"""
from __future__ import annotations

import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

SEG_ROWS_LIVE = 772
T = 63
B = int(os.environ.get("D2H_B", "4096"))
K = int(os.environ.get("D2H_K", "6"))
PIECE = 296

libc = ctypes.CDLL("libc.so.6", use_errno=True)
SYS_mbind, SYS_set_mempolicy, SYS_move_pages = 237, 238, 279         # x86_64
MPOL_DEFAULT, MPOL_PREFERRED, MPOL_BIND = 0, 1, 2


def pci_of(local: int) -> str:
    p = torch.cuda.get_device_properties(local)
    try:
        return f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except AttributeError:
        import subprocess
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                             capture_output=True, text=True).stdout.strip().lower()
        return out[-12:] if len(out) >= 12 else out


def sysfs(path: str, default: str = "?") -> str:
    try:
        return open(path).read().strip()
    except OSError:
        return default


def parse_cpulist(s: str):
    cpus = []
    for part in s.split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def page_nodes(t: torch.Tensor, samples: int = 64):
    """NUMA node of `samples` pages spread over the tensor's memory (move_pages with nodes = NULL only queries)."""
    nbytes = t.numel() * t.element_size()
    page = 4096
    addrs = [(t.data_ptr() + (i * (nbytes - page) // max(1, samples - 1))) & ~(page - 1) for i in range(samples)]
    arr = (ctypes.c_void_p * samples)(*addrs)
    status = (ctypes.c_int * samples)()
    rc = libc.syscall(SYS_move_pages, 0, ctypes.c_ulong(samples), arr, None, status, 0)
    if rc != 0:
        return {"error": ctypes.get_errno()}
    hist = {}
    for s in status:
        hist[int(s)] = hist.get(int(s), 0) + 1
    return hist


def bind_to_node(node: int, cpus) -> str:
    notes = []
    try:
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            notes.append(f"affinity->{len(allowed)} cpus")
        else:
            notes.append("no allowed cpu on that node")
    except OSError as e:
        notes.append(f"affinity failed {e}")
    if node >= 0:
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(64))
        notes.append("mempolicy preferred ok" if rc == 0 else f"set_mempolicy errno {ctypes.get_errno()}")
    return ", ".join(notes)


def main() -> int:
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    cudart = ctypes.CDLL("libcudart.so.12")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def gather(x: float):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if dist is None:
            return [x]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    pci = pci_of(local)
    node = int(sysfs(f"/sys/bus/pci/devices/{pci}/numa_node", "-1") or -1)
    cpulist = sysfs(f"/sys/bus/pci/devices/{pci}/local_cpulist", "")
    nodes_online = sysfs("/sys/devices/system/node/online")
    info = {"rank": rank, "pci": pci, "gpu_numa_node": node, "gpu_local_cpulist": cpulist,
            "nodes_online": nodes_online, "affinity_cpus": len(os.sched_getaffinity(0)),
            "cpu_now": os.sched_getcpu() if hasattr(os, "sched_getcpu") else -1}

    compact_bytes = B * (SEG_ROWS_LIVE * T * 4 + 36)
    full_elems = B * 9 * 128 * T
    d_compact = torch.empty(compact_bytes // 4, dtype=torch.float32, device=dev).normal_()
    d_full = torch.empty(full_elems, dtype=torch.float32, device=dev).normal_()
    d_in = torch.empty(B * 16000, dtype=torch.int16, device=dev)
    s2 = torch.cuda.Stream()

    def timed(fn, reps=K):
        fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        if dist is not None:
            dist.barrier()
        return ms

    results = {}

    def report(name, nbytes, ms):
        per = gather(nbytes / (ms * 1e-3) / 1e9)
        slow = gather(ms)
        results[name] = {"per_rank_gbs": [round(x, 2) for x in per], "aggregate_gbs": round(world * nbytes / (max(slow) * 1e-3) / 1e9, 2),
                         "ms_max": round(max(slow), 3), "bytes_per_rank": nbytes}

    def run_suite(tag: str):
        h_compact = torch.empty(compact_bytes // 4, dtype=torch.float32).pin_memory()
        h_compact.fill_(0)                                           # touch every page under the current policy
        h_in = torch.zeros(B * 16000, dtype=torch.int16).pin_memory()
        info[f"pages_{tag}"] = page_nodes(h_compact)
        report(f"{tag}:contig", compact_bytes, timed(lambda: h_compact.copy_(d_compact, non_blocking=True)))

        def pieces():
            per = compact_bytes // 4 // B * PIECE
            for off in range(0, compact_bytes // 4, per):
                e = min(off + per, compact_bytes // 4)
                h_compact[off:e].copy_(d_compact[off:e], non_blocking=True)
        report(f"{tag}:pieces296", compact_bytes, timed(pieces))
        report(f"{tag}:h2d_pcm16", B * 32000, timed(lambda: d_in.copy_(h_in, non_blocking=True)))

        def bidir():
            with torch.cuda.stream(s2):
                d_in.copy_(h_in, non_blocking=True)
            h_compact.copy_(d_compact, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s2)
        report(f"{tag}:bidir(d2h bytes)", compact_bytes, timed(bidir))
        return h_compact

    run_suite("default")

    # the six 2D row runs of the round-1 host path into the full [B,9,128,63] layout
    h_full = torch.empty(full_elems, dtype=torch.float32).pin_memory()
    h_full.fill_(0)
    live = [24, 64, 12, 128, 128, 128, 120, 40, 128]
    runs = []
    for c, lv in enumerate(live):
        s, e = c * 128, c * 128 + lv
        if runs and runs[-1][1] == s:
            runs[-1][1] = e
        else:
            runs.append([s, e])
    st = torch.cuda.current_stream().cuda_stream
    pitch = 9 * 128 * T * 4

    def strided():
        for off in range(0, B, PIECE):
            n = min(PIECE, B - off)
            for s, e in runs:
                o = (off * 9 * 128 + s) * T * 4
                rc = cudart.cudaMemcpy2DAsync(ctypes.c_void_p(h_full.data_ptr() + o), ctypes.c_size_t(pitch),
                                              ctypes.c_void_p(d_full.data_ptr() + o), ctypes.c_size_t(pitch),
                                              ctypes.c_size_t((e - s) * T * 4), ctypes.c_size_t(n), 2, ctypes.c_void_p(st))
                assert rc == 0, rc
    report("default:strided2d", B * SEG_ROWS_LIVE * T * 4, timed(strided))
    report("default:full_planes", full_elems * 4, timed(lambda: h_full.copy_(d_full, non_blocking=True)))
    del h_full

    info["bind"] = bind_to_node(node, parse_cpulist(cpulist))
    run_suite("numa")

    infos = [info]
    if dist is not None:
        infos = [None] * world
        dist.all_gather_object(infos, info)
    if rank == 0:
        print(json.dumps({"world": world, "B": B, "reps": K, "ranks": infos, "results": results}))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
