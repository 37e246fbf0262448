"""librosa.util subset (test infrastructure; see package docstring)."""
from __future__ import annotations

import numpy as np
import scipy.sparse


def tiny(x):
    x = np.asarray(x)
    if np.issubdtype(x.dtype, np.floating) or np.issubdtype(x.dtype, np.complexfloating):
        dtype = x.dtype
    else:
        dtype = np.dtype(np.float32)
    return np.finfo(dtype).tiny


def dtype_r2c(d, default=np.complex64):
    mapping = {np.dtype(np.float32): np.complex64, np.dtype(np.float64): np.complex128}
    dt = np.dtype(d)
    if dt.kind == "c":
        return dt
    return np.dtype(mapping.get(dt, default))


def pad_center(data, *, size, axis=-1, **kwargs):
    kwargs.setdefault("mode", "constant")
    n = data.shape[axis]
    lpad = int((size - n) // 2)
    lengths = [(0, 0)] * data.ndim
    lengths[axis] = (lpad, int(size - n - lpad))
    if lpad < 0:
        raise ValueError("target size smaller than input")
    return np.pad(data, lengths, **kwargs)


def fix_length(data, *, size, axis=-1, **kwargs):
    kwargs.setdefault("mode", "constant")
    n = data.shape[axis]
    if n > size:
        sl = [slice(None)] * data.ndim
        sl[axis] = slice(0, size)
        return data[tuple(sl)]
    if n < size:
        lengths = [(0, 0)] * data.ndim
        lengths[axis] = (0, size - n)
        return np.pad(data, lengths, **kwargs)
    return data


def expand_to(x, *, ndim, axes):
    try:
        axes = tuple(axes)
    except TypeError:
        axes = (axes,)
    shape = [1] * ndim
    for i, ax in enumerate(axes):
        shape[ax] = x.shape[i]
    return x.reshape(shape)


def frame(x, *, frame_length, hop_length, axis=-1):
    """Strided framing: for axis=-1 the result is [..., frame_length, n_frames]."""
    x = np.asarray(x)
    if axis != -1 and axis != x.ndim - 1:
        raise NotImplementedError
    if x.shape[-1] < frame_length:
        raise ValueError("input too short for framing")
    xw = np.lib.stride_tricks.sliding_window_view(x, frame_length, axis=-1)  # [..., n, frame_length]
    xw = xw[..., ::hop_length, :]
    return np.swapaxes(xw, -1, -2)


def normalize(S, *, norm=np.inf, axis=0, threshold=None, fill=None):
    """librosa.util.normalize: the norm is taken in float64, the quotient is stored in S.dtype."""
    if threshold is None:
        threshold = tiny(S)
    if norm is None:
        return S
    mag = np.abs(S).astype(float)
    if norm == np.inf:
        length = np.max(mag, axis=axis, keepdims=True)
    elif norm == -np.inf:
        length = np.min(mag, axis=axis, keepdims=True)
    elif norm == 0:
        length = np.sum(mag > 0, axis=axis, keepdims=True, dtype=mag.dtype)
    elif np.issubdtype(type(norm), np.number) and norm > 0:
        length = np.sum(mag ** norm, axis=axis, keepdims=True) ** (1.0 / norm)
    else:
        raise ValueError(norm)
    small_idx = length < threshold
    Snorm = np.empty_like(S)
    if fill is None:
        length[small_idx] = 1.0
        Snorm[:] = S / length
    else:
        raise NotImplementedError
    return Snorm


def localmax(x, *, axis=0):
    xi = np.swapaxes(x, -1, axis)
    lmax = np.zeros(x.shape, dtype=bool)
    lmaxi = np.swapaxes(lmax, -1, axis)
    lmaxi[..., 1:-1] = (xi[..., 1:-1] > xi[..., :-2]) & (xi[..., 1:-1] >= xi[..., 2:])
    lmaxi[..., -1] = xi[..., -1] > xi[..., -2]
    return lmax


def abs2(x, dtype=None):
    if np.iscomplexobj(x):
        y = x.real ** 2 + x.imag ** 2
        return y if dtype is None else y.astype(dtype)
    return np.square(x, dtype=dtype)


def phasor(angles):
    return np.cos(angles) + 1j * np.sin(angles)


def sparsify_rows(x, *, quantile=0.01, dtype=None):
    if x.ndim == 1:
        x = x.reshape((1, -1))
    if dtype is None:
        dtype = x.dtype
    x_sparse = scipy.sparse.lil_matrix(x.shape, dtype=dtype)
    mags = np.abs(x)
    norms = np.sum(mags, axis=1, keepdims=True)
    mag_sort = np.sort(mags, axis=1)
    cumulative_mag = np.cumsum(mag_sort / norms, axis=1)
    threshold_idx = np.argmin(cumulative_mag < quantile, axis=1)
    for i, j in enumerate(threshold_idx):
        idx = np.where(mags[i] >= mag_sort[i, j])
        x_sparse[i, idx] = x[i, idx]
    return x_sparse.tocsr()


def sync_mean_all(data, axis=-2):
    """util.sync(data, [slice(None)], aggregate=np.mean, axis=-2): one aggregate over the whole axis."""
    return np.mean(data, axis=axis, keepdims=True).astype(data.dtype)
