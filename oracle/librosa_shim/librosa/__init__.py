"""TEST INFRASTRUCTURE ONLY -- numpy/scipy restatement of the librosa 0.10.2.post1 entry points
that dohyeoplim/breathing-phase-classifier's `src/precompute` calls.

PARITY UNPINNED: librosa / soundfile / soxr are not installable in the build image and the reference
ships no golden vectors, so this shim restates the published librosa 0.10.2 algorithms (the version
pinned by reference `env.yaml:156`).  Cross-checks that *are* available here (torchaudio mel banks,
transformers.audio_utils chroma/mel banks, scipy) are exercised by `tests/test_oracle_shim.py`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import
this package.  The product path (bpc_b200 + libbpc_b200.so) never does.

Call sites in the reference that this shim serves (reference file:line):
  process.py:28 load            process.py:32 feature.melspectrogram   process.py:33 power_to_db
  process.py:34 feature.delta   process.py:43 feature.mfcc             process.py:51 stft
  process.py:52 feature.chroma_stft   process.py:53 feature.chroma_cens
  process.py:74 onset.onset_strength  process.py:75 feature.tempogram
  methods.py:52 feature.rms     methods.py:53 feature.zero_crossing_rate
  methods.py:59-63 feature.spectral_{centroid,bandwidth,rolloff,flatness,contrast}
  methods.py:126 lpc            methods.py:137 filters.mel
"""
from __future__ import annotations

import numpy as np
import scipy.signal

from . import util
from . import filters
from ._core import (
    load, stft, power_to_db, lpc, fft_frequencies, mel_frequencies, hz_to_mel, mel_to_hz,
    hz_to_octs, hz_to_midi, note_to_hz, estimate_tuning, piptrack, pitch_tuning, autocorrelate,
    resample, cqt, vqt, get_window, _spectrogram, set_halfband,
)
from . import feature
from . import onset

__version__ = "0.10.2.post1-shim"
