"""librosa.onset subset (test infrastructure; see package docstring)."""
from __future__ import annotations

import numpy as np

from ._core import power_to_db
from .feature import melspectrogram


def onset_strength(*, y=None, sr=22050, S=None, lag=1, max_size=1, ref=None, detrend=False, center=True,
                   feature=None, aggregate=None, n_fft=2048, hop_length=512, **kwargs):
    """onset_strength -> onset_strength_multi(channels=None)[..., 0, :] with the default mel feature."""
    if feature is not None or aggregate is not None or ref is not None or max_size != 1 or detrend:
        raise NotImplementedError
    kwargs.setdefault("fmax", 0.5 * sr)
    if S is None:
        S = np.abs(melspectrogram(y=y, sr=sr, n_fft=n_fft, hop_length=hop_length, **kwargs))
        S = power_to_db(S)
    S = np.atleast_2d(S)
    ref = S
    onset_env = S[..., lag:] - ref[..., :-lag]
    onset_env = np.maximum(0.0, onset_env)
    # util.sync(onset_env, [slice(None)], aggregate=np.mean, pad=True, axis=-2): one mean over the mel axis,
    # stored in the dtype of onset_env
    agg = np.empty((1, onset_env.shape[-1]), dtype=onset_env.dtype)
    agg[0] = np.mean(onset_env, axis=-2)
    onset_env = agg
    pad_width = lag
    if center:
        pad_width += n_fft // (2 * hop_length)
    onset_env = np.pad(onset_env, [(0, 0), (int(pad_width), 0)], mode="constant")
    if center:
        onset_env = onset_env[..., : S.shape[-1]]
    return onset_env[0]
