"""Mirror of reference `src/precompute/core.py`: `precompute()` and `process_dataset_threaded(df, audio_dir,
target_dir, dataset_name)` with the same ID -> wav-path mapping, success / failure tally and messages.

Instead of one Python call per file on two threads (core.py:33-34), rows are decoded by a small thread pool, stacked
into [B, 16000] PCM16 batches and sent through ONE C-ABI call per batch (`bpc_precompute_host`, pinned double
buffering); `.npz` files are written by the same pool.
"""
from __future__ import annotations

import os
import re
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import methods as _m
from .process import load_wav, fit_batch, save_npz

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


def process_dataset_threaded(df, audio_dir, target_dir, dataset_name, engine=None, batch=BATCH):
    items = [(row["ID"], os.path.join(audio_dir, wav_name_for(row["ID"], dataset_name))) for _, row in df.iterrows()]
    eng = engine if engine is not None else _m._get_engine()
    successful = failed = 0
    results = []
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

            def _write(i):
                fid = good[i][0]
                try:
                    if status[i] & 1:
                        raise ValueError("non-finite samples in input")
                    save_npz(target_dir, fid, feats[i], scal[i])
                    return fid, True, None
                except Exception as e:  # noqa: BLE001
                    return fid, False, str(e)
            results.extend(pool.map(_write, range(len(good))))
    for fid, ok, err in results:
        if ok:
            successful += 1
        else:
            failed += 1
            _print_error(f"{fid}: {err}")
    _print_success(f"{successful} 성공, {failed} 실패")
    return results


def precompute():
    """core.py:47-56."""
    import pandas as pd
    os.makedirs(PRECOMP_DIR, exist_ok=True)
    train_df = pd.read_csv(TRAIN_CSV_PATH)
    test_df = pd.read_csv(TEST_CSV_PATH)
    process_dataset_threaded(train_df, TRAIN_AUDIO_DIR, PRECOMP_DIR, "train")
    process_dataset_threaded(test_df, TEST_AUDIO_DIR, PRECOMP_DIR, "test")
    _print_success("완료")
