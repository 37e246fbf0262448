"""RIFF/WAVE images for the decode tests: any sample format / channel count, optional junk chunk and truncation."""
import struct

import numpy as np

FMT = {"u8": (1, 8), "pcm16": (1, 16), "pcm24": (1, 24), "pcm32": (1, 32), "f32": (3, 32), "f64": (3, 64)}


def samples(kind: str, frames: int, channels: int, seed: int) -> np.ndarray:
    """[frames, channels] array in the storage type of `kind` (pcm24: int32 values in [-2^23, 2^23))."""
    rng = np.random.default_rng(seed)
    if kind == "u8":
        return rng.integers(0, 256, (frames, channels)).astype(np.uint8)
    if kind == "pcm16":
        return rng.integers(-32768, 32768, (frames, channels)).astype(np.int16)
    if kind == "pcm24":
        return rng.integers(-(1 << 23), 1 << 23, (frames, channels)).astype(np.int32)
    if kind == "pcm32":
        return rng.integers(-(1 << 31), 1 << 31, (frames, channels)).astype(np.int32)
    return (rng.standard_normal((frames, channels)) * 0.3).astype(np.float32 if kind == "f32" else np.float64)


def payload(kind: str, x: np.ndarray) -> bytes:
    if kind == "pcm24":
        b = x.astype("<i4").tobytes()
        a = np.frombuffer(b, dtype=np.uint8).reshape(-1, 4)[:, :3]
        return a.tobytes()
    return x.astype(x.dtype.newbyteorder("<")).tobytes()


def image(kind: str, x: np.ndarray, sr: int, extensible=False, junk=False, cut=0) -> bytes:
    tag, bits = FMT[kind]
    channels = x.shape[1]
    data = payload(kind, x)
    align = channels * bits // 8
    if extensible:
        sub = struct.pack("<H", tag) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
        fmt = struct.pack("<HHIIHHHHI", 0xFFFE, channels, sr, sr * align, align, bits, 22, bits, 0) + sub
    else:
        fmt = struct.pack("<HHIIHH", tag, channels, sr, sr * align, align, bits)
    chunks = b"fmt " + struct.pack("<I", len(fmt)) + fmt
    if junk:
        chunks += b"LIST" + struct.pack("<I", 5) + b"hello" + b"\x00"          # odd size: one pad byte
    chunks += b"data" + struct.pack("<I", len(data)) + data
    img = b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks
    return img[:len(img) - cut] if cut else img
