"""Host-side engine: owns one bpc_handle (one per device / stream) and moves torch / numpy buffers across the C ABI.

PyTorch is used for device memory and streams only; every computation happens inside libbpc_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def _check(handle, rc, what):
    if rc != 0:
        msg = L.lib().bpc_last_error(handle)
        raise L.BpcError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


class Engine:
    """One B200, one handle.  `precompute` (device tensors), `precompute_host` (host arrays), `stage_logmel`."""

    def __init__(self, device: int = 0, max_batch: int = 4096, params: L.Params | None = None, debug: bool = False):
        self._lib = L.lib()
        self.params = params if params is not None else L.default_params()
        self._h = C.c_void_p()
        rc = self._lib.bpc_create(C.byref(self._h), C.byref(self.params), int(device), int(max_batch))
        if rc != 0:
            msg = self._lib.bpc_last_error(None)
            raise L.BpcError(f"bpc_create failed ({rc}): {msg.decode() if msg else ''}")
        self.device = int(device)
        self.T = self._lib.bpc_num_frames(C.byref(self.params))
        self.nscal = self._lib.bpc_num_scalars(C.byref(self.params))
        self.L = int(self.params.expected_len)
        self.chunk = int(self._lib.bpc_chunk_size(self._h))
        if debug:
            self.set_debug(True)

    # ------------------------------------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.bpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_debug(self, on: bool):
        _check(self._h, self._lib.bpc_set_debug(self._h, int(bool(on))), "bpc_set_debug")

    # ------------------------------------------------------------------------------------------- device tensors
    @staticmethod
    def _wav_dtype(t):
        import torch
        if t.dtype == torch.float32:
            return L.WAV_F32
        if t.dtype == torch.int16:
            return L.WAV_PCM16
        raise TypeError("wav must be float32 or int16")

    def precompute(self, wav, feats=None, scalars=None, status=None, stream=None):
        """wav: cuda tensor [B, L_in] float32 / int16 -> (feats [B,9,128,T], scalars [B,S], status [B]) on device."""
        import torch
        if not wav.is_cuda or wav.dim() != 2:
            raise ValueError("wav must be a 2-D CUDA tensor; host arrays go through precompute_host")
        wav = wav.contiguous()
        B, L_in = wav.shape
        dev = wav.device
        if feats is None:
            feats = torch.empty((B, L.NUM_CHANNELS, L.PLANE_ROWS, self.T), dtype=torch.float32, device=dev)
        if scalars is None:
            scalars = torch.empty((B, self.nscal), dtype=torch.float32, device=dev)
        if status is None:
            status = torch.empty((B,), dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream if stream is None else stream
        rc = self._lib.bpc_precompute(self._h, wav.data_ptr(), self._wav_dtype(wav), B, L_in, feats.data_ptr(),
                                      scalars.data_ptr(), status.data_ptr(), C.c_void_p(st))
        _check(self._h, rc, "bpc_precompute")
        return feats, scalars, status

    def stage_logmel(self, wav, want_stft=True, stream=None):
        """BASELINE config 2: (stft_db [B,257,T] or None, mel3 [B,3,128,T])."""
        import torch
        wav = wav.contiguous()
        B, L_in = wav.shape
        dev = wav.device
        stft = torch.empty((B, 257, self.T), dtype=torch.float32, device=dev) if want_stft else None
        mel3 = torch.empty((B, 3, L.PLANE_ROWS, self.T), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream if stream is None else stream
        rc = self._lib.bpc_stage_logmel(self._h, wav.data_ptr(), self._wav_dtype(wav), B, L_in,
                                        stft.data_ptr() if want_stft else None, mel3.data_ptr(), C.c_void_p(st))
        _check(self._h, rc, "bpc_stage_logmel")
        return stft, mel3

    def modspec(self, mel_db):
        """methods.py:142-143 on [n, 128, T] mel_db (numpy or cuda tensor) -> same kind, [n, 40, T]."""
        import torch
        is_np = isinstance(mel_db, np.ndarray)
        x = torch.from_numpy(np.ascontiguousarray(mel_db, dtype=np.float32)).cuda(self.device) if is_np else mel_db.contiguous()
        if x.dim() != 3 or x.shape[1] != L.PLANE_ROWS or x.shape[2] != self.T:
            raise ValueError(f"mel_db must be [n, 128, {self.T}]")
        out = torch.empty((x.shape[0], 40, self.T), dtype=torch.float32, device=x.device)
        st = torch.cuda.current_stream(x.device).cuda_stream
        _check(self._h, self._lib.bpc_modspec(self._h, x.data_ptr(), x.shape[0], out.data_ptr(), C.c_void_p(st)),
               "bpc_modspec")
        return out.cpu().numpy() if is_np else out

    def resample(self, y, sr_in: int, sr_out: int = 16000):
        """librosa.load's rate conversion (process.py:28) on the device: numpy or cuda float32 [n] -> same kind,
        [ceil(n * sr_out / sr_in)]."""
        import torch
        is_np = isinstance(y, np.ndarray)
        x = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)).cuda(self.device) if is_np else y.contiguous()
        if x.dim() != 1 or x.dtype != torch.float32:
            raise ValueError("y must be a 1-D float32 waveform")
        n_out = int(self._lib.bpc_resample_len(int(x.numel()), int(sr_in), int(sr_out)))
        if n_out < 0:
            raise ValueError("bad sample rates")
        out = torch.empty((n_out,), dtype=torch.float32, device=x.device)
        st = torch.cuda.current_stream(x.device).cuda_stream
        rc = self._lib.bpc_resample(self._h, x.data_ptr(), int(x.numel()), int(sr_in), int(sr_out), out.data_ptr(), n_out,
                                    C.c_void_p(st))
        _check(self._h, rc, "bpc_resample")
        return out.cpu().numpy() if is_np else out

    def decode_wavs(self, images, length: int | None = None, sr: int = 16000):
        """GPU-side `librosa.load(path, sr=sr)` + `pad_or_truncate` (process.py:28-29) for a list of RIFF/WAVE file
        images (bytes as read from disk): the host walks the chunk headers (`bpc_wav_parse`), the images go to the
        device as they are, `bpc_wav_decode` scales / down-mixes / pads there, files of another rate are resampled on
        the device (`bpc_resample`).  Returns (cuda float32 [n, length], errors) with errors[i] = None or a message
        (the row of a failed file is zero)."""
        import torch
        length = int(self.L if length is None else length)
        n = len(images)
        infos = (L.WavInfo * max(n, 1))()
        offs = np.zeros(max(n, 1), dtype=np.int64)
        errors = [None] * n
        pos = 0
        for i, img in enumerate(images):
            rc = self._lib.bpc_wav_parse(img, len(img), C.byref(infos[i]))
            if rc != 0:
                errors[i] = "not a RIFF/WAVE file" if rc == -11 else f"unsupported wav format ({rc})"
            offs[i] = pos
            pos += (len(img) + 15) & ~15
        dev = torch.device("cuda", self.device)
        y = torch.zeros((n, length), dtype=torch.float32, device=dev)
        good = [i for i in range(n) if errors[i] is None]
        if not good:
            return y, errors
        blob_h = torch.empty(max(pos, 16), dtype=torch.uint8, pin_memory=True)
        bh = blob_h.numpy()
        for i in good:
            bh[offs[i]:offs[i] + len(images[i])] = np.frombuffer(images[i], dtype=np.uint8)
        blob = blob_h.to(dev, non_blocking=True)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        same = [i for i in good if infos[i].sr == sr]
        other = [i for i in good if infos[i].sr != sr]

        def _decode(idx, out, out_len):
            inf = (L.WavInfo * len(idx))(*[infos[i] for i in idx])
            of = np.ascontiguousarray(offs[idx])
            _check(self._h, self._lib.bpc_wav_decode(self._h, blob.data_ptr(), int(blob.numel()), of.ctypes.data, C.byref(inf), len(idx),
                                                     int(out_len), out.data_ptr(), st), "bpc_wav_decode")

        if same:
            if len(same) == n:
                _decode(same, y, length)
            else:
                tmp = torch.empty((len(same), length), dtype=torch.float32, device=dev)
                _decode(same, tmp, length)
                y[torch.as_tensor(same, device=dev)] = tmp
        for i in other:                                   # decode at the file's own rate, then convert
            frames = int(infos[i].frames)
            if frames == 0:
                continue
            raw = torch.empty((1, frames), dtype=torch.float32, device=dev)
            _decode([i], raw, frames)
            r = self.resample(raw[0], int(infos[i].sr), sr)
            m = min(int(r.numel()), length)
            y[i, :m] = r[:m]
        torch.cuda.current_stream(dev).synchronize()      # the pinned blob may be released
        return y, errors

    # ---------------------------------------------------------------------------------------------- host arrays
    def precompute_host(self, wav: np.ndarray, feats=None, scalars=None, status=None):
        """wav: numpy [B, L_in] float32 / int16 (pageable or pinned) -> numpy (feats, scalars, status)."""
        wav = np.ascontiguousarray(wav)
        if wav.ndim != 2:
            raise ValueError("wav must be [B, L_in]")
        if wav.dtype == np.float32:
            dt = L.WAV_F32
        elif wav.dtype == np.int16:
            dt = L.WAV_PCM16
        else:
            raise TypeError("wav must be float32 or int16")
        B, L_in = wav.shape
        if feats is None:
            feats = np.empty((B, L.NUM_CHANNELS, L.PLANE_ROWS, self.T), dtype=np.float32)
        if scalars is None:
            scalars = np.empty((B, self.nscal), dtype=np.float32)
        if status is None:
            status = np.empty((B,), dtype=np.int32)
        for name, a, shape, dtp in (("feats", feats, (B, L.NUM_CHANNELS, L.PLANE_ROWS, self.T), np.float32),
                                    ("scalars", scalars, (B, self.nscal), np.float32), ("status", status, (B,), np.int32)):
            if tuple(a.shape) != shape or a.dtype != dtp or not a.flags.c_contiguous or not a.flags.writeable:
                raise ValueError(f"{name} must be a writable C-contiguous {np.dtype(dtp).name} array of shape {shape}")
        rc = self._lib.bpc_precompute_host(self._h, wav.ctypes.data, dt, B, L_in, feats.ctypes.data,
                                           scalars.ctypes.data, status.ctypes.data)
        _check(self._h, rc, "bpc_precompute_host")
        return feats, scalars, status

    def precompute_host_compact(self, wav: np.ndarray, rows=None, pad=None, scalars=None, status=None):
        """wav: numpy [B, L_in] float32 / int16 -> (rows [B, 772, T], pad [B, 9], scalars, status): the compact host
        layout of include/bpc.h (only the data rows of every plane + one pad value per plane; `expand_compact` or
        `PackedDS` rebuild the [9, 128, T] planes).  One contiguous device->host copy per piece, no host fill."""
        wav = np.ascontiguousarray(wav)
        if wav.ndim != 2:
            raise ValueError("wav must be [B, L_in]")
        if wav.dtype == np.float32:
            dt = L.WAV_F32
        elif wav.dtype == np.int16:
            dt = L.WAV_PCM16
        else:
            raise TypeError("wav must be float32 or int16")
        B, L_in = wav.shape
        if rows is None:
            rows = np.empty((B, L.LIVE_TOTAL, self.T), dtype=np.float32)
        if pad is None:
            pad = np.empty((B, L.NUM_CHANNELS), dtype=np.float32)
        if scalars is None:
            scalars = np.empty((B, self.nscal), dtype=np.float32)
        if status is None:
            status = np.empty((B,), dtype=np.int32)
        for name, a, shape, dtp in (("rows", rows, (B, L.LIVE_TOTAL, self.T), np.float32),
                                    ("pad", pad, (B, L.NUM_CHANNELS), np.float32),
                                    ("scalars", scalars, (B, self.nscal), np.float32), ("status", status, (B,), np.int32)):
            if tuple(a.shape) != shape or a.dtype != dtp or not a.flags.c_contiguous or not a.flags.writeable:
                raise ValueError(f"{name} must be a writable C-contiguous {np.dtype(dtp).name} array of shape {shape}")
        rc = self._lib.bpc_precompute_host_compact(self._h, wav.ctypes.data, dt, B, L_in, rows.ctypes.data,
                                                   pad.ctypes.data, scalars.ctypes.data, status.ctypes.data)
        _check(self._h, rc, "bpc_precompute_host_compact")
        return rows, pad, scalars, status

    def precompute_host_compact_begin(self, wav: np.ndarray, rows, pad, scalars, status) -> int:
        """Streaming form (include/bpc.h): enqueue the call behind whatever is still in flight and return a ticket;
        `host_wait(ticket)` returns once the outputs are in `rows` / `pad` / `scalars` / `status`.  All five arrays
        belong to the library until then (the engine keeps them alive).  Alternate two sets of output buffers --
        begin(k + 1), then wait(k) -- to keep the GPU and the copy engines busy across calls."""
        if wav.ndim != 2 or not wav.flags.c_contiguous or wav.dtype not in (np.float32, np.int16):
            raise ValueError("wav must be a C-contiguous [B, L_in] float32 / int16 array")
        dt = L.WAV_F32 if wav.dtype == np.float32 else L.WAV_PCM16
        B, L_in = wav.shape
        for name, a, shape, dtp in (("rows", rows, (B, L.LIVE_TOTAL, self.T), np.float32),
                                    ("pad", pad, (B, L.NUM_CHANNELS), np.float32),
                                    ("scalars", scalars, (B, self.nscal), np.float32), ("status", status, (B,), np.int32)):
            if tuple(a.shape) != shape or a.dtype != dtp or not a.flags.c_contiguous or not a.flags.writeable:
                raise ValueError(f"{name} must be a writable C-contiguous {np.dtype(dtp).name} array of shape {shape}")
        ticket = C.c_int64(0)
        rc = self._lib.bpc_precompute_host_compact_begin(self._h, wav.ctypes.data, dt, B, L_in, rows.ctypes.data,
                                                         pad.ctypes.data, scalars.ctypes.data, status.ctypes.data,
                                                         C.byref(ticket))
        _check(self._h, rc, "bpc_precompute_host_compact_begin")
        self._pending = getattr(self, "_pending", {})
        self._pending[ticket.value] = (wav, rows, pad, scalars, status)
        return ticket.value

    def host_wait(self, ticket: int = -1) -> None:
        """Block until every output of the call with this ticket (default: of every call) is in its host buffers."""
        _check(self._h, self._lib.bpc_host_wait(self._h, int(ticket)), "bpc_host_wait")
        pend = getattr(self, "_pending", {})
        for t in [t for t in pend if ticket < 0 or t <= ticket]:
            del pend[t]

    def host_empty(self, shape, dtype=np.float32) -> np.ndarray:
        """Pinned host array on the NUMA node of this engine's GPU (bpc_host_alloc); lives as long as the engine or until
        `host_free(arr)`.  `arr.bpc_numa_node` is not available on ndarrays, so the node is kept in `self.host_numa_node`."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        node = C.c_int(-1)
        ptr = self._lib.bpc_host_alloc(self._h, max(n, 1), C.byref(node))
        if not ptr:
            raise L.BpcError("bpc_host_alloc failed")
        self.host_numa_node = int(node.value)
        buf = (C.c_char * max(n, 1)).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self._host_ptrs = getattr(self, "_host_ptrs", {})
        self._host_ptrs[arr.ctypes.data] = ptr
        return arr

    def host_free(self, arr: np.ndarray):
        ptr = getattr(self, "_host_ptrs", {}).pop(arr.ctypes.data, None)
        if ptr is not None and self._h.value:
            self._lib.bpc_host_free(self._h, C.c_void_p(ptr))

    # ------------------------------------------------------------------------------------------------ statistics
    def channel_stats(self) -> np.ndarray:
        out = np.empty((9 + self.nscal, 5), dtype=np.float64)
        _check(self._h, self._lib.bpc_channel_stats(self._h, out.ctypes.data), "bpc_channel_stats")
        return out

    def channel_stats_device(self):
        """torch float64 view [(9+S), 5] of the device accumulator (for an NCCL all-reduce)."""
        import torch
        ptr = C.c_void_p()
        rows = C.c_int64()
        _check(self._h, self._lib.bpc_channel_stats_device(self._h, C.byref(ptr), C.byref(rows)), "stats_device")
        n = rows.value * 5

        class _Iface:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr.value, False), "version": 2}
        return torch.as_tensor(_Iface(), device=f"cuda:{self.device}").view(rows.value, 5)

    def reset_stats(self):
        _check(self._h, self._lib.bpc_channel_stats_reset(self._h), "bpc_channel_stats_reset")

    def set_kernel_timing(self, on: bool):
        _check(self._h, self._lib.bpc_set_kernel_timing(self._h, int(bool(on))), "bpc_set_kernel_timing")

    def kernel_times(self):
        """{kernel name: (total ms, launches)} since the last call (synchronises the device)."""
        n = 13                                           # BPC_NUM_KERNEL_IDS
        ms = np.zeros(n, dtype=np.float64)
        cnt = np.zeros(n, dtype=np.int64)
        _check(self._h, self._lib.bpc_kernel_times(self._h, ms.ctypes.data, cnt.ctypes.data, n), "bpc_kernel_times")
        return {self._lib.bpc_kernel_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i]}

    def launch_count(self) -> int:
        return int(self._lib.bpc_launch_count(self._h))

    # ----------------------------------------------------------------------------------------------------- debug
    _DEBUG_SHAPES = {
        "mag512": lambda s: ((s.T, 260), np.float32), "mel_db": lambda s: ((128, s.T), np.float32),
        "mfcc_raw": lambda s: ((120, s.T), np.float32), "gammatone_raw": lambda s: ((64, s.T), np.float32),
        "mod_spec_raw": lambda s: ((40, s.T), np.float32), "chroma_stft_raw": lambda s: ((12, s.T), np.float32),
        "chroma_cens_raw": lambda s: ((12, s.T), np.float32),
        "lpc_raw": lambda s: ((12, (s.L - 400 + 159) // 160), np.float32),
        "onset_env": lambda s: ((s.T,), np.float32), "tuning": lambda s: ((2,), np.int32),
        "ints": lambda s: ((2,), np.int32),
    }

    def debug(self, what: str, n: int) -> np.ndarray:
        """Raw intermediate `what` of the first n segments of the last chunk (needs set_debug(True) before the call)."""
        shape, dt = self._DEBUG_SHAPES[what](self)
        out = np.empty((n,) + shape, dtype=dt)
        got = C.c_int64()
        rc = self._lib.bpc_debug_copy(self._h, what.encode(), out.ctypes.data, out.nbytes, C.byref(got))
        _check(self._h, rc, "bpc_debug_copy")
        if got.value != out.nbytes:
            raise L.BpcError(f"debug {what}: wanted {out.nbytes} bytes, got {got.value}")
        return out


def expand_compact(rows: np.ndarray, pad: np.ndarray, out: np.ndarray | None = None, threads: int = 4) -> np.ndarray:
    """Compact host layout (rows [n, 772, T], pad [n, 9]) -> the full [n, 9, 128, T] tensor (bpc_expand_compact; host
    only, no GPU involved).  Bit-identical to what `precompute_host` returns."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    pad = np.ascontiguousarray(pad, dtype=np.float32)
    if rows.ndim != 3 or rows.shape[1] != L.LIVE_TOTAL or pad.shape != (rows.shape[0], L.NUM_CHANNELS):
        raise ValueError("rows must be [n, 772, T] and pad [n, 9]")
    n, _, T = rows.shape
    if out is None:
        out = np.empty((n, L.NUM_CHANNELS, L.PLANE_ROWS, T), dtype=np.float32)
    if out.shape != (n, L.NUM_CHANNELS, L.PLANE_ROWS, T) or out.dtype != np.float32 or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous float32 [n, 9, 128, T] array")
    rc = L.lib().bpc_expand_compact(rows.ctypes.data, pad.ctypes.data, n, T, out.ctypes.data, int(threads))
    if rc != 0:
        raise L.BpcError(f"bpc_expand_compact failed ({rc})")
    return out


def resample_filter(sr_in: int, sr_out: int):
    """Host-side polyphase table of bpc_resample (no GPU needed) -> (p, q, half, tab [p, 2 half] float64)."""
    lib = L.lib()
    p, q, half = C.c_int(), C.c_int(), C.c_int()
    cap = 1 << 22
    buf = np.empty(cap, dtype=np.float64)
    n = lib.bpc_resample_filter(int(sr_in), int(sr_out), buf.ctypes.data, cap, C.byref(p), C.byref(q), C.byref(half))
    if n < 0:
        raise L.BpcError(f"bpc_resample_filter failed ({n})")
    return p.value, q.value, half.value, buf[:n].reshape(p.value, 2 * half.value).copy()


def table(name: str, tuning_idx: int = 50, params: L.Params | None = None) -> np.ndarray:
    """Host-side constant table by name (no GPU needed); see include/bpc.h::bpc_table_copy."""
    p = params if params is not None else L.default_params()
    T = p.expected_len // p.hop + 1
    shapes = {"mel_a": ((128, 257), np.float32), "mel_b": ((128, 257), np.float32), "mel_c": ((64, 257), np.float32),
              "mel_d": ((128, 1025), np.float32), "mel_d_band": ((128, 82), np.float32), "dct_mel": ((40, 128), np.float32), "dct_time": ((T, T), np.float32),
              "hann512": ((512,), np.float64), "hann2048": ((2048,), np.float64), "hann384": ((384,), np.float64),
              "hamming400": ((400,), np.float64), "chroma": ((12, 257), np.float32),
              "hist_edges": ((101,), np.float64), "halfband": ((127,), np.float64),
              "cqt_basis": ((36, 257, 2), np.float32), "cqt_sqrt_len": ((252,), np.float64)}
    shape, dt = shapes[name]
    out = np.empty(shape, dtype=dt)
    n = L.lib().bpc_table_copy(C.byref(p), name.encode(), int(tuning_idx), out.ctypes.data, out.size)
    if n != out.size:
        raise L.BpcError(f"bpc_table_copy({name}) returned {n}, expected {out.size}")
    return out
