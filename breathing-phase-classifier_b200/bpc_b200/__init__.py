"""bpc_b200 -- B200-native `--precompute` path of dohyeoplim/breathing-phase-classifier.

Layout mirrors the reference (`src/precompute/{core,process,methods}.py`); the arithmetic lives in
libbpc_b200.so (hand-written sm_100a CUDA behind the C ABI of include/bpc.h).  No CPU fallback exists.
"""
from ._lib import (BpcError, Params, default_params, lib, CHANNELS, NUM_CHANNELS, NUM_SCALARS, PLANE_ROWS,
                   WAV_F32, WAV_PCM16, EXPORTS, LIB_PATH)
from .engine import Engine, table, expand_compact
from .shards import ShardWriter, PackedShard, PackedDS, collate_fn, write_npz_batch, npz_bytes
from .resident import ResidentDS, cutmix_data, mixup_data

__all__ = ["BpcError", "Params", "default_params", "lib", "Engine", "table", "expand_compact", "ShardWriter", "PackedShard", "PackedDS", "collate_fn", "write_npz_batch", "npz_bytes",
           "ResidentDS", "cutmix_data", "mixup_data", "CHANNELS", "NUM_CHANNELS",
           "NUM_SCALARS", "PLANE_ROWS", "WAV_F32", "WAV_PCM16", "EXPORTS", "LIB_PATH"]
