"""Mirror of reference `src/precompute/core.py`: `precompute()` and `process_dataset_threaded(df, audio_dir,
target_dir, dataset_name)` with the same ID -> wav-path mapping, success / failure tally and messages.

Instead of one Python call per file on two threads (core.py:33-34), rows are decoded by a small thread pool, stacked
into [B, 16000] PCM16 batches and sent through ONE C-ABI call per batch (`bpc_precompute_host`, pinned double
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
from .process import load_wav, fit_batch
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
BATCH = 1024
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
    items = [(row["ID"], os.path.join(audio_dir, wav_name_for(row["ID"], dataset_name))) for _, row in df.iterrows()]
    eng = engine if engine is not None else _m._get_engine()
    successful = failed = 0
    results = []
    shard_rows = []                                   # packed mode: rows are collected and written at the end
    with ThreadPoolExecutor(max_workers=IO_THREADS) as pool:
        for lo in range(0, len(items), batch):
            part = items[lo:lo + batch]
            loaded = list(pool.map(_try_load, [p for _, p in part]))
            good = [(fid, w) for (fid, _), (w, err) in zip(part, loaded) if w is not None]
            for (fid, _), (w, err) in zip(part, loaded):
                if w is None:
                    results.append((fid, False, err))
            if not good:
                continue
            feats, scal, status = eng.precompute_host(fit_batch([w for _, w in good], eng.L))
            ids = [fid for fid, _ in good]
            if packed:
                keep = [i for i in range(len(ids)) if not (status[i] & 1)]
                results.extend((ids[i], False, "non-finite samples in input") for i in range(len(ids)) if status[i] & 1)
                shard_rows.append(([ids[i] for i in keep], feats[keep], scal[keep], status[keep]))
                results.extend((ids[i], True, None) for i in keep)
            else:
                results.extend(write_npz_batch(target_dir, ids, feats, scal, status, threads=IO_THREADS))
    if packed and shard_rows:
        shard_dir = os.path.join(target_dir, dataset_name)
        with ShardWriter(shard_dir, sum(len(r[0]) for r in shard_rows), eng.T, eng.nscal) as w:
            for ids, f, s, st in shard_rows:
                w.append(ids, f, s, st)
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
