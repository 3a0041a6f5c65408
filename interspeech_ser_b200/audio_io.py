"""Minimal host audio ingest replacing `librosa.load(path, sr=16000)` (preprocess_speech.py:47).

RIFF/WAVE PCM (8/16/24/32-bit int, 32/64-bit float) -> mono float32 in [-1, 1]; other sample rates are
resampled to 16 kHz with a polyphase filter (scipy) like librosa's default would. librosa / soundfile are
not installed in this image."""
from __future__ import annotations

import struct
from typing import Tuple

import numpy as np


def read_bytes(path: str) -> bytes:
    """The I/O half of read_wav (releases the GIL: this is what the CLI's worker threads run)."""
    with open(path, "rb") as fh:
        return fh.read()


def read_wav(path: str) -> Tuple[np.ndarray, int]:
    return decode_wav(read_bytes(path), path)


def decode_wav(data: bytes, path: str = "<bytes>", keep_int16: bool = False) -> Tuple[np.ndarray, int]:
    """RIFF/WAVE bytes -> (mono float32, sample rate). One pass over the samples, no copy of the payload.
    keep_int16: mono 16-bit PCM is returned as the int16 samples themselves (a zero-copy view of `data`); the GPU
    kernels apply the 1 / 32768 scaling on load (include/serenc.h SERENC_WAV_I16), so host decode work and the
    upload are halved with bit-identical results."""
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    view = memoryview(data)
    pos = 12
    fmt = None
    payload = None
    while pos + 8 <= len(data):
        cid = data[pos:pos + 4]
        size = struct.unpack_from("<I", data, pos + 4)[0]
        body = view[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, sr, _, _, bits = struct.unpack("<HHIIHH", body[:16])
            if tag == 0xFFFE and len(body) >= 26:  # WAVE_FORMAT_EXTENSIBLE: sub-format GUID starts with the real tag
                tag = struct.unpack("<H", body[24:26])[0]
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            payload = body
        pos += 8 + size + (size & 1)
    if fmt is None or payload is None:
        raise ValueError(f"{path}: missing fmt/data chunk")
    tag, ch, sr, bits = fmt
    if tag == 1:
        if bits == 16:
            if keep_int16 and ch == 1:
                return np.frombuffer(payload[: len(payload) // 2 * 2], dtype="<i2"), int(sr)
            x = np.frombuffer(payload[: len(payload) // 2 * 2], dtype="<i2").astype(np.float32)
            x *= np.float32(1.0 / 32768.0)   # exact (power of two), in place
        elif bits == 8:
            x = (np.frombuffer(payload, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        elif bits == 32:
            x = np.frombuffer(payload, dtype="<i4").astype(np.float32) / 2147483648.0
        elif bits == 24:
            b = np.frombuffer(payload[: len(payload) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            v = np.where(v & 0x800000, v - 0x1000000, v)
            x = v.astype(np.float32) / 8388608.0
        else:
            raise ValueError(f"{path}: unsupported PCM width {bits}")
    elif tag == 3:
        x = np.frombuffer(payload, dtype="<f4" if bits == 32 else "<f8").astype(np.float32)
    else:
        raise ValueError(f"{path}: unsupported WAVE format tag {tag}")
    if ch > 1:
        x = x[: len(x) // ch * ch].reshape(-1, ch).mean(axis=1)  # librosa.load(mono=True)
    return np.ascontiguousarray(x, dtype=np.float32), int(sr)


def load_audio(path: str, sr: int = 16000) -> Tuple[np.ndarray, int]:
    """librosa.load(path, sr=16000) stand-in: mono float32 at `sr`."""
    return load_audio_bytes(read_bytes(path), path, sr)


def load_audio_bytes(data: bytes, path: str = "<bytes>", sr: int = 16000, keep_int16: bool = False) -> Tuple[np.ndarray, int]:
    x, file_sr = decode_wav(data, path, keep_int16)
    if file_sr != sr:
        if x.dtype == np.int16:
            x = x.astype(np.float32) * np.float32(1.0 / 32768.0)
        from math import gcd

        from scipy.signal import resample_poly

        g = gcd(sr, file_sr)
        x = resample_poly(x.astype(np.float64), sr // g, file_sr // g).astype(np.float32)
    return x, sr


def write_wav(path: str, x: np.ndarray, sr: int = 16000) -> None:
    """16-bit PCM mono writer (tests / synthetic corpora)."""
    pcm = np.clip(np.round(np.asarray(x, dtype=np.float64) * 32768.0), -32768, 32767).astype("<i2").tobytes()
    with open(path, "wb") as fh:
        fh.write(b"RIFF" + struct.pack("<I", 36 + len(pcm)) + b"WAVE")
        fh.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, sr, sr * 2, 2, 16))
        fh.write(b"data" + struct.pack("<I", len(pcm)) + pcm)
