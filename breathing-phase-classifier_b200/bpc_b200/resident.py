"""GPU-resident dataset + batch assembly.

The reference keeps features on disk, re-reads one `.npz` per item in 8 DataLoader workers (dataloaders.py:21-54,
dataset.py:37-57), stacks them in `collate_fn` (dataset.py:59-73), copies the batch to the GPU and then applies
CutMix / MixUp (augmentation.py:5-44, train.py:76-89).  Here the `[N, 9, 128, T]` output of the precompute path stays
in HBM (1M segments of 1 s are 290 GB: 36 GB per GPU on 8 GPUs) and a batch is ONE gather kernel (`bpc_collate`) that
also applies the mix.  The random draws are made on the host with the reference's own calls, in the reference's order.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def rand_bbox(W: int, H: int, lam: float, rng=np.random):
    """augmentation.py:11-25 -> (bbx1, bby1, bbx2, bby2)."""
    cut_rat = np.sqrt(1.0 - lam)
    cut_w = np.int32(W * cut_rat)
    cut_h = np.int32(H * cut_rat)
    cx = rng.randint(W)
    cy = rng.randint(H)
    bbx1 = int(np.clip(cx - cut_w // 2, 0, W))
    bby1 = int(np.clip(cy - cut_h // 2, 0, H))
    bbx2 = int(np.clip(cx + cut_w // 2, 0, W))
    bby2 = int(np.clip(cy + cut_h // 2, 0, H))
    return bbx1, bby1, bbx2, bby2


class ResidentDS:
    """Device-resident feature store with the batch layout of `DS` + `collate_fn`.

    feats [N, 9, 128, T] / scalars [N, S] are CUDA tensors (e.g. straight from `Engine.precompute`), labels a float
    tensor [N] (1.0 = "E", dataset.py:52) or None for a test set, ids the file ids."""

    def __init__(self, engine, feats, scalars, labels=None, ids=None):
        import torch
        if not feats.is_cuda or feats.dim() != 4 or feats.shape[1:3] != (L.NUM_CHANNELS, L.PLANE_ROWS):
            raise ValueError("feats must be a CUDA tensor [N, 9, 128, T]")
        if feats.shape[3] != engine.T or scalars.shape[1] != engine.nscal:
            raise ValueError("store shape does not match the engine geometry")
        self.engine = engine
        self.feats = feats.contiguous()
        self.scalars = scalars.contiguous()
        self.labels = None if labels is None else labels.to(feats.device, torch.float32)
        self.ids = list(ids) if ids is not None else None
        self.feature_names = list(L.CHANNELS)
        self.n_features = L.NUM_CHANNELS
        self.scalar_dim = int(scalars.shape[1])

    def __len__(self):
        return int(self.feats.shape[0])

    @classmethod
    def from_shard(cls, engine, data_frame, feature_dir: str, is_training: bool, chunk: int = 2048):
        """Load the rows of `data_frame` (the reference's `DS(data_frame, feature_dir, is_training)` arguments,
        dataset.py:10) from a packed shard into HBM, in data-frame order; labels as in dataset.py:52."""
        import torch
        from .shards import PackedShard
        shard = PackedShard(feature_dir)
        ids = data_frame["ID"].tolist()
        rows = np.asarray([shard.row[i] for i in ids], dtype=np.int64)
        dev = torch.device(f"cuda:{engine.device}")
        feats = torch.empty((len(rows),) + tuple(shard.feats.shape[1:]), dtype=torch.float32, device=dev)
        scal = torch.empty((len(rows), shard.scalars.shape[1]), dtype=torch.float32, device=dev)
        for lo in range(0, len(rows), chunk):                      # staged through pinned memory chunk by chunk
            sel = rows[lo:lo + chunk]
            feats[lo:lo + len(sel)].copy_(torch.from_numpy(np.ascontiguousarray(shard.feats[sel])).pin_memory(), non_blocking=True)
            scal[lo:lo + len(sel)].copy_(torch.from_numpy(np.ascontiguousarray(shard.scalars[sel])).pin_memory(), non_blocking=True)
        torch.cuda.synchronize(dev)
        labels = None
        if is_training:
            labels = torch.tensor([1.0 if t == "E" else 0.0 for t in data_frame["Target"].tolist()], dtype=torch.float32, device=dev)
        return cls(engine, feats, scal, labels, ids)

    def _collate(self, ia, ib, mode, lam, box):
        import torch
        n = int(ia.numel())
        dev = self.feats.device
        out_f = torch.empty((n,) + tuple(self.feats.shape[1:]), dtype=torch.float32, device=dev)
        out_s = torch.empty((n, self.scalar_dim), dtype=torch.float32, device=dev)
        x1, y1, x2, y2 = box
        st = torch.cuda.current_stream(dev).cuda_stream
        eng = self.engine
        rc = eng._lib.bpc_collate(eng._h, self.feats.data_ptr(), self.scalars.data_ptr(), len(self), ia.data_ptr(),
                                  None if ib is None else ib.data_ptr(), n, mode, float(lam), y1, y2, x1, x2,
                                  out_f.data_ptr(), out_s.data_ptr(), C.c_void_p(st))
        if rc != 0:
            raise L.BpcError(f"bpc_collate failed ({rc}): {eng._lib.bpc_last_error(eng._h).decode()}")
        return out_f, out_s

    def batch(self, indices, mix: str | None = None, alpha: float = 1.0, rng=np.random, perm=None):
        """(features [n,9,128,T], scalars [n,S], labels [n] | ids) for store rows `indices`.

        mix = "cutmix": augmentation.py:5-33 (features only; labels mixed with the box-area lam);
        mix = "mixup":  train.py:80-86 (features, scalars and labels).  `perm` overrides torch.randperm(n)."""
        import torch
        dev = self.feats.device
        ia = torch.as_tensor(indices, dtype=torch.int64, device=dev).contiguous()
        n = int(ia.numel())
        if mix is None:
            f, s = self._collate(ia, None, L.MIX_NONE, 1.0, (0, 0, 0, 0))
            if self.labels is not None:
                return f, s, self.labels[ia]
            return f, s, [self.ids[i] for i in ia.tolist()] if self.ids is not None else ia
        if self.labels is None:
            raise ValueError("mixing needs labels")
        perm = torch.randperm(n).to(dev) if perm is None else torch.as_tensor(perm, dtype=torch.int64, device=dev)
        ib = ia[perm].contiguous()
        lam = float(rng.beta(alpha, alpha))
        lab = self.labels[ia]
        if mix == "cutmix":
            T = int(self.feats.shape[3])
            box = rand_bbox(T, L.PLANE_ROWS, lam, rng)
            f, s = self._collate(ia, ib, L.MIX_CUTMIX, lam, box)
            lam = 1 - ((box[2] - box[0]) * (box[3] - box[1]) / (T * L.PLANE_ROWS))
        elif mix == "mixup":
            f, s = self._collate(ia, ib, L.MIX_MIXUP, lam, (0, 0, 0, 0))
        else:
            raise ValueError("mix must be None, 'cutmix' or 'mixup'")
        return f, s, lam * lab + (1 - lam) * lab[perm]

    def batches(self, batch_size: int, shuffle: bool = False, drop_last: bool = False, generator=None):
        """Epoch iterator in the order a `DataLoader(DS, batch_size, shuffle)` would draw."""
        import torch
        n = len(self)
        order = torch.randperm(n, generator=generator) if shuffle else torch.arange(n)
        for lo in range(0, n, batch_size):
            idx = order[lo:lo + batch_size]
            if drop_last and idx.numel() < batch_size:
                break
            yield self.batch(idx)


def cutmix_data(features, labels, alpha=1.0, device="cuda", engine=None, indices=None, rng=np.random):
    """Mirror of augmentation.py:5-33 on an already collated CUDA batch: same arguments, same 4-tuple."""
    import torch
    ds = ResidentDS(engine or _engine_for(features), features, _no_scalars(features, engine), labels)
    n = features.size(0)
    indices = torch.randperm(n).to(features.device) if indices is None else indices.to(features.device)
    lam = float(rng.beta(alpha, alpha))
    W, H = features.size(3), features.size(2)
    box = rand_bbox(W, H, lam, rng)
    ia = torch.arange(n, device=features.device)
    mixed, _ = ds._collate(ia, indices.to(torch.int64).contiguous(), L.MIX_CUTMIX, lam, box)
    lam = 1 - ((box[2] - box[0]) * (box[3] - box[1]) / (W * H))
    return mixed, lam * labels + (1 - lam) * labels[indices], indices, lam


def mixup_data(features, labels, alpha=1.0, device="cuda", engine=None, indices=None, rng=np.random):
    """Mirror of augmentation.py:36-44."""
    import torch
    ds = ResidentDS(engine or _engine_for(features), features, _no_scalars(features, engine), labels)
    n = features.size(0)
    indices = torch.randperm(n).to(features.device) if indices is None else indices.to(features.device)
    lam = float(rng.beta(alpha, alpha))
    ia = torch.arange(n, device=features.device)
    mixed, _ = ds._collate(ia, indices.to(torch.int64).contiguous(), L.MIX_MIXUP, lam, (0, 0, 0, 0))
    return mixed, lam * labels + (1 - lam) * labels[indices], indices, lam


def _engine_for(features):
    from .precompute.methods import _get_engine
    return _get_engine()


def _no_scalars(features, engine):
    import torch
    eng = engine or _engine_for(features)
    return torch.zeros((features.size(0), eng.nscal), dtype=torch.float32, device=features.device)
