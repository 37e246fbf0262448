"""Mirror of reference `src/precompute/core.py`: `precompute()` and `process_dataset_threaded(df, audio_dir,
target_dir, dataset_name)` with the same ID -> wav-path mapping, success / failure tally and messages.

Instead of one Python call per file on two threads (core.py:33-34), rows are decoded by a reader pool, stacked
into [B, 16000] PCM16 batches (`bpc_wav_load_batch`, the C++ reader pool of the library) and sent through ONE C-ABI call per batch (`bpc_precompute_host`, pinned double
buffering); the `.npz` files are serialised by the C++ writer pool of the library (`bpc_npz_write_batch`).  With
`packed=True` the same rows go into one packed shard instead (bpc_b200/shards.py), which `PackedDS` reads in place of
the reference `DS`.
"""
from __future__ import annotations

import os
import re
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import methods as _m
from .process import load_wav, fit_batch, load_wav_batch
from ..shards import ShardWriter, write_npz_batch

SR = 16000                      # core.py:9-17
DURATION = 1.0
EXPECTED_LEN = int(SR * DURATION)
N_WORKERS = 2
TRAIN_CSV_PATH = "input/train.csv"
TEST_CSV_PATH = "input/test.csv"
TRAIN_AUDIO_DIR = "input/train"
TEST_AUDIO_DIR = "input/test"
PRECOMP_DIR = "input/precomputed/"
BATCH = 512
IO_THREADS = 8


def _print_error(msg):
    try:
        from rich.console import Console
        Console().print(f"[bold red]❌ {msg}[/bold red]")
    except Exception:
        print(f"❌ {msg}")


def _print_success(msg):
    try:
        from rich.console import Console
        Console().print(f"[bold green]✅ {msg}[/bold green]")
    except Exception:
        print(f"✅ {msg}")


def wav_name_for(file_id: str, dataset_name: str) -> str:
    """core.py:23-26."""
    if dataset_name == "train":
        return re.sub(r"_[EI]_", "_", file_id) + ".wav"
    return file_id if file_id.endswith(".wav") else (file_id + ".wav")


def _try_load(path):
    try:
        return load_wav(path), None
    except Exception as e:  # noqa: BLE001
        return None, str(e)


def process_dataset_threaded(df, audio_dir, target_dir, dataset_name, engine=None, batch=BATCH, packed=False):
    """core.py:19-45.  Returns the (file_id, success, error) tuples in row order (the reference prints / tallies them).

    Three stages run concurrently, one batch apart: the C++ reader pool decodes batch k+1, the GPU path computes batch k
    and the C++ writer pool serialises batch k-1 (ctypes releases the GIL during all three)."""
    items = [(fid, os.path.join(audio_dir, wav_name_for(fid, dataset_name))) for fid in df["ID"].tolist()]
    # A batched caller owns its engine (no debug planes, chunks of the batch size, no ENGINE_LOCK); the shared
    # single-segment debug engine of methods.py serves the per-file mirrors only.
    own_engine = engine is None
    if own_engine:
        from ..engine import Engine
        engine = Engine(device=0, max_batch=max(1, min(batch, max(1, len(items)))), debug=False)
    eng = engine
    results = [None] * len(items)
    starts = list(range(0, len(items), batch))
    shard = ShardWriter(os.path.join(target_dir, dataset_name), len(items), eng.T, eng.nscal) if packed and items else None

    # The reader thread decodes the rare general files (stereo, 8 / 24 / 32-bit, float, other rates) on the GPU through a
    # handle of its own (calls on one handle are serialised by its owner; handles are independent, include/bpc.h).
    dec = []

    def _decoder():
        if not dec:
            from ..engine import Engine
            dec.append(Engine(device=eng.device, max_batch=1, params=eng.params))
        return dec[0]

    def _load(lo):
        return load_wav_batch([p for _, p in items[lo:lo + batch]], eng.L, threads=IO_THREADS, engine=_decoder)

    def _write(pos, ids, feats, scal, status):
        for p, r in zip(pos, write_npz_batch(target_dir, ids, feats, scal, status, threads=IO_THREADS)):
            results[p] = r

    with ThreadPoolExecutor(max_workers=2) as pool:
        nxt = pool.submit(_load, starts[0]) if starts else None
        pending_write = None
        for k, lo in enumerate(starts):
            wavs, errs = nxt.result()
            nxt = pool.submit(_load, starts[k + 1]) if k + 1 < len(starts) else None
            good_rows = [i for i, e in enumerate(errs) if e is None]
            for i, e in enumerate(errs):
                if e is not None:
                    results[lo + i] = (items[lo + i][0], False, e)
            if not good_rows:
                continue
            batch_wav = wavs if len(good_rows) == len(errs) else np.ascontiguousarray(wavs[good_rows])
            pos = [lo + i for i in good_rows]
            ids = [items[p][0] for p in pos]
            if packed:
                # the library writes this batch straight into the shard's memory-mapped arrays, in the compact host
                # layout (data rows + one pad value per plane: nothing is expanded on the way to disk)
                n = len(ids)
                r_out, p_out, s_out, st_out = shard.reserve(n)
                eng.precompute_host_compact(batch_wav, r_out, p_out, s_out, st_out)
                bad = np.flatnonzero(np.asarray(st_out) & 1)
                if len(bad):                          # rare: drop the rows whose input was not finite
                    keep = [i for i in range(n) if not (st_out[i] & 1)]
                    kept = [np.asarray(a)[keep] for a in (r_out, p_out, s_out, st_out)]
                    shard.unreserve(n)
                    for dst, src in zip(shard.reserve(len(keep)), kept):
                        dst[:] = src
                    shard.commit([ids[i] for i in keep])
                else:
                    shard.commit(ids)
                badset = set(int(i) for i in bad)
                for i in range(n):
                    results[pos[i]] = (ids[i], False, "non-finite samples in input") if i in badset else (ids[i], True, None)
            else:
                feats, scal, status = eng.precompute_host(batch_wav)
                if pending_write is not None:
                    pending_write.result()
                pending_write = pool.submit(_write, pos, ids, feats, scal, status)
        if pending_write is not None:
            pending_write.result()
    if shard is not None:
        shard.close()
    for d in dec:
        d.close()
    if own_engine:
        eng.close()
    successful = failed = 0
    for fid, ok, err in results:
        if ok:
            successful += 1
        else:
            failed += 1
            _print_error(f"{fid}: {err}")
    _print_success(f"{successful} 성공, {failed} 실패")
    return results


def precompute(packed=False):
    """core.py:47-56."""
    import pandas as pd
    os.makedirs(PRECOMP_DIR, exist_ok=True)
    train_df = pd.read_csv(TRAIN_CSV_PATH)
    test_df = pd.read_csv(TEST_CSV_PATH)
    process_dataset_threaded(train_df, TRAIN_AUDIO_DIR, PRECOMP_DIR, "train", packed=packed)
    process_dataset_threaded(test_df, TEST_AUDIO_DIR, PRECOMP_DIR, "test", packed=packed)
    _print_success("완료")
