"""Packed output container of the precompute path and readers that stand in for the reference's `DS`.

The reference writes one `.npz` (a zip of ten `.npy` members) per segment (process.py:92-103) and `DS.__getitem__`
re-opens one file per item (dataset.py:37-57).  At the 1M-segment scale of BASELINE config 3 that is a million small
files and a million zip parses per epoch, so next to the drop-in per-segment writer (`write_npz_batch`, the C++ pool in
csrc/npz.cpp) the path can write a *packed shard*:

    <dir>/feats.npy      float32 [N, 9, 128, T]   channel axis in sorted-key order (== dataset.py:25-26)      (v1), or
    <dir>/rows.npy       float32 [N, 772, T]      the data rows of the nine planes back to back  }  compact (v2, default):
    <dir>/pad.npy        float32 [N, 9]           the constant of every plane's pad rows         }  what bpc_precompute_host_compact
                                                  (pad_freq, methods.py:39-46)                       writes, 67 % of the bytes
    <dir>/scalars.npy    float32 [N, S]
    <dir>/status.npy     int32   [N]              per-segment status bits (include/bpc.h)
    <dir>/index.json     {"ids": [...], "channels": [...], "T": T, "S": S, "params": {...}}

All arrays are plain `.npy` files, so `np.load(mmap_mode="r")` maps them without a copy; a compact shard rebuilds the
[9, 128, T] planes of an item when it is read (`PackedShard.feats[i]`, bit-identical to the full layout).  `PackedDS` has the
constructor, attributes and item layout of the reference `DS` (dataset.py:7-57); `collate_fn` is dataset.py:59-73.
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from . import _lib as L

EXCLUDED_KEYS = {"scalars", "sr", "hop_length", "n_fft"}          # dataset.py:8
INDEX_NAME = "index.json"


# ------------------------------------------------------------------------------------------- per-segment .npz
def npz_bytes(feats: np.ndarray, scalars: np.ndarray) -> bytes:
    """One segment ([9,128,T] sorted-key slab + [S] scalars) -> the bytes of the reference's `.npz`."""
    feats = np.ascontiguousarray(feats, dtype=np.float32)
    scalars = np.ascontiguousarray(scalars, dtype=np.float32)
    if feats.ndim != 3 or feats.shape[:2] != (L.NUM_CHANNELS, L.PLANE_ROWS) or scalars.ndim != 1:
        raise ValueError("feats must be [9, 128, T] and scalars [S]")
    T, S = feats.shape[2], scalars.shape[0]
    n = L.lib().bpc_npz_size(T, S)
    buf = np.empty(n, dtype=np.uint8)
    got = L.lib().bpc_npz_pack(feats.ctypes.data, scalars.ctypes.data, T, S, buf.ctypes.data, n)
    if got != n:
        raise L.BpcError(f"bpc_npz_pack returned {got}, expected {n}")
    return buf.tobytes()


def write_npz_batch(target_dir: str, file_ids, feats: np.ndarray, scalars: np.ndarray, status=None, threads: int = 8):
    """Write `<target_dir>/<id>.npz` for every row with the C++ writer pool -> list of (file_id, ok, error) in the
    reference's return convention (process.py:105-108)."""
    feats = np.ascontiguousarray(feats, dtype=np.float32)
    scalars = np.ascontiguousarray(scalars, dtype=np.float32)
    n = len(file_ids)
    if feats.shape[0] != n or scalars.shape[0] != n:
        raise ValueError("one feats / scalars row per file id")
    T, S = feats.shape[3], scalars.shape[1]
    ids = (C.c_char_p * n)(*[os.fsencode(str(f)) for f in file_ids])
    ok = np.zeros(n, dtype=np.int32)
    st = None if status is None else np.ascontiguousarray(status, dtype=np.int32)
    rc = L.lib().bpc_npz_write_batch(os.fsencode(target_dir), ids, feats.ctypes.data, scalars.ctypes.data,
                                     None if st is None else st.ctypes.data, n, T, S, int(threads), ok.ctypes.data)
    if rc != 0:
        raise L.BpcError(f"bpc_npz_write_batch failed ({rc})")
    why = {-1: "non-finite samples in input", -2: "cannot open output file", -3: "short write"}
    return [(fid, bool(o == 1), None if o == 1 else why.get(int(o), f"writer error {o}")) for fid, o in zip(file_ids, ok)]


# ------------------------------------------------------------------------------------------------ packed shard
def compact_from_full(feats: np.ndarray):
    """[n, 9, 128, T] -> (rows [n, 772, T], pad [n, 9]): keep the data rows, and the first pad row's first value."""
    feats = np.asarray(feats, dtype=np.float32)
    n, _, _, T = feats.shape
    rows = np.empty((n, L.LIVE_TOTAL, T), dtype=np.float32)
    pad = np.zeros((n, L.NUM_CHANNELS), dtype=np.float32)
    r0 = 0
    for c, lv in enumerate(L.LIVE_ROWS):
        rows[:, r0:r0 + lv] = feats[:, c, :lv]
        if lv < L.PLANE_ROWS:
            pad[:, c] = feats[:, c, lv, 0]
        r0 += lv
    return rows, pad


class ShardWriter:
    """Append batches of precompute output to a packed shard of `capacity` segments; `close()` writes the index.
    compact=True (default) stores the compact host layout (rows + pad), compact=False the full [N, 9, 128, T] tensor."""

    def __init__(self, target_dir: str, capacity: int, T: int, S: int = L.NUM_SCALARS, params: dict | None = None,
                 compact: bool = True):
        os.makedirs(target_dir, exist_ok=True)
        self.dir, self.capacity, self.T, self.S = target_dir, int(capacity), int(T), int(S)
        self.compact = bool(compact)
        fmt = np.lib.format
        if self.compact:
            self.rows = fmt.open_memmap(os.path.join(target_dir, "rows.npy"), mode="w+", dtype=np.float32,
                                        shape=(self.capacity, L.LIVE_TOTAL, self.T))
            self.pad = fmt.open_memmap(os.path.join(target_dir, "pad.npy"), mode="w+", dtype=np.float32,
                                       shape=(self.capacity, L.NUM_CHANNELS))
        else:
            self.feats = fmt.open_memmap(os.path.join(target_dir, "feats.npy"), mode="w+", dtype=np.float32,
                                         shape=(self.capacity, L.NUM_CHANNELS, L.PLANE_ROWS, self.T))
        self.scalars = fmt.open_memmap(os.path.join(target_dir, "scalars.npy"), mode="w+", dtype=np.float32,
                                       shape=(self.capacity, self.S))
        self.status = fmt.open_memmap(os.path.join(target_dir, "status.npy"), mode="w+", dtype=np.int32,
                                      shape=(self.capacity,))
        self.ids: list[str] = []
        self.params = params or {}

    def append(self, file_ids, feats, scalars, status=None):
        """feats: full [n, 9, 128, T] planes (compacted here when the shard is compact)."""
        n = len(file_ids)
        views = self.reserve(n)
        if self.compact:
            r, p = compact_from_full(feats)
            views[0][:] = r
            views[1][:] = p
        else:
            views[0][:] = feats
        views[-2][:] = scalars
        views[-1][:] = 0 if status is None else status
        self.commit(file_ids)

    def reserve(self, n: int):
        """Views of the next n rows for a producer that writes in place -- (rows, pad, scalars, status) for a compact
        shard, (feats, scalars, status) otherwise; follow with commit()."""
        lo = len(self.ids)
        if lo + n > self.capacity:
            raise ValueError("shard capacity exceeded")
        self._reserved = n
        if self.compact:
            return self.rows[lo:lo + n], self.pad[lo:lo + n], self.scalars[lo:lo + n], self.status[lo:lo + n]
        return self.feats[lo:lo + n], self.scalars[lo:lo + n], self.status[lo:lo + n]

    def unreserve(self, n: int):
        self._reserved = 0

    def commit(self, file_ids):
        if len(file_ids) != getattr(self, "_reserved", -1):
            raise ValueError("commit() must name exactly the reserved rows")
        self.ids.extend(str(f) for f in file_ids)
        self._reserved = 0

    def close(self, allow_short: bool = True):
        """Write the index.  A shard may hold fewer rows than its capacity (failed files): readers use the first
        len(ids) rows of the arrays."""
        if len(self.ids) > self.capacity or (not allow_short and len(self.ids) != self.capacity):
            raise ValueError(f"shard holds {len(self.ids)} of {self.capacity} segments")
        arrays = [self.rows, self.pad] if self.compact else [self.feats]
        for a in arrays + [self.scalars, self.status]:
            a.flush()
        with open(os.path.join(self.dir, INDEX_NAME), "w") as f:
            json.dump({"ids": self.ids, "rows": len(self.ids), "channels": list(L.CHANNELS), "T": self.T, "S": self.S,
                       "params": self.params, "layout": "compact" if self.compact else "full",
                       "live_rows": list(L.LIVE_ROWS),
                       "format": "bpc_b200 packed shard v2" if self.compact else "bpc_b200 packed shard v1"}, f)
        if self.compact:
            del self.rows, self.pad
        else:
            del self.feats
        del self.scalars, self.status

    def __enter__(self):
        return self

    def __exit__(self, et, ev, tb):
        if et is None:
            self.close(allow_short=False)


class CompactFeats:
    """Read-only [N, 9, 128, T] view over a compact shard: indexing with an int, a slice or an index array rebuilds
    the planes of the selected segments (bpc_expand_compact on the host)."""

    def __init__(self, rows, pad):
        self.rows, self.pad = rows, pad
        self.shape = (rows.shape[0], L.NUM_CHANNELS, L.PLANE_ROWS, rows.shape[2])
        self.dtype = np.dtype(np.float32)

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, idx):
        from .engine import expand_compact
        one = isinstance(idx, (int, np.integer))
        sel = [int(idx)] if one else idx
        out = expand_compact(np.ascontiguousarray(self.rows[sel]), np.ascontiguousarray(self.pad[sel]),
                             threads=1 if one else 4)
        return out[0] if one else out


class PackedShard:
    """Read side: memory-mapped arrays + id -> row map."""

    def __init__(self, shard_dir: str):
        with open(os.path.join(shard_dir, INDEX_NAME)) as f:
            self.index = json.load(f)
        self.ids = self.index["ids"]
        self.row = {fid: i for i, fid in enumerate(self.ids)}
        self.channels = list(self.index["channels"])
        n = len(self.ids)
        if self.index.get("layout", "full") == "compact":
            self.rows = np.load(os.path.join(shard_dir, "rows.npy"), mmap_mode="r")[:n]
            self.pad = np.load(os.path.join(shard_dir, "pad.npy"), mmap_mode="r")[:n]
            self.feats = CompactFeats(self.rows, self.pad)
        else:
            self.feats = np.load(os.path.join(shard_dir, "feats.npy"), mmap_mode="r")[:n]
        self.scalars = np.load(os.path.join(shard_dir, "scalars.npy"), mmap_mode="r")[:n]
        self.status = np.load(os.path.join(shard_dir, "status.npy"), mmap_mode="r")[:n]

    def __len__(self):
        return len(self.ids)


def is_packed(feature_dir: str) -> bool:
    return os.path.exists(os.path.join(feature_dir, INDEX_NAME))


class PackedDS:
    """Stand-in for the reference `DS(data_frame, feature_dir, is_training)` (dataset.py:7-57) over a packed shard:
    same attributes (`feature_names` sorted, `n_features`, `scalar_dim`) and the same items
    (features [9,128,T] float32, scalars [S] float32, label tensor | file id).  Works with torch's DataLoader."""
    EXCLUDED_KEYS = EXCLUDED_KEYS

    def __init__(self, data_frame, feature_dir: str, is_training: bool):
        self.df = data_frame.reset_index(drop=True)
        self.feature_dir = feature_dir
        self.is_training = is_training
        if len(self.df) == 0:
            raise ValueError
        self.shard = PackedShard(feature_dir)
        self.feature_names = sorted(k for k in self.shard.channels if k not in self.EXCLUDED_KEYS)
        self._chan = [self.shard.channels.index(k) for k in self.feature_names]
        self.n_features = len(self.feature_names)
        self.scalar_dim = int(self.shard.scalars.shape[1])
        self._ids = self.df["ID"].tolist()
        self._targets = self.df["Target"].tolist() if is_training else None
        print(f"#Features: {self.n_features} - {', '.join(self.feature_names)}")
        print(f"#Scalars: {self.scalar_dim}")

    def __len__(self):
        return len(self.df)

    def __getitem__(self, idx):
        import torch
        file_id = self._ids[idx]
        r = self.shard.row[file_id]                      # KeyError ~ the reference's FileNotFoundError
        f = np.array(self.shard.feats[r], dtype=np.float32)
        if self._chan != list(range(f.shape[0])):
            f = f[self._chan]
        features = torch.from_numpy(f)
        scalars = torch.from_numpy(np.array(self.shard.scalars[r], dtype=np.float32))
        if self.is_training:
            label = 1.0 if self._targets[idx] == "E" else 0.0
            return features, scalars, torch.tensor(label, dtype=torch.float32)
        return features, scalars, file_id


def collate_fn(batch):
    """dataset.py:59-73."""
    import torch
    feats, scals, labs_or_ids = zip(*batch)
    features = torch.stack(list(feats), dim=0)
    scalars = torch.stack(list(scals), dim=0)
    if isinstance(labs_or_ids[0], torch.Tensor):
        return features, scalars, torch.stack(list(labs_or_ids), dim=0)
    return features, scalars, list(labs_or_ids)
