"""Audio front-end of the EMAGE path: the step right before the hot path.

The reference calls `librosa.load(path, sr=16000)` (test_emage_audio.py:17): decode, mix down to mono, resample
to 16 kHz float32 in [-1, 1].  librosa / soundfile / ffmpeg are not available offline, so this is a small
stand-alone reader for PCM / IEEE-float WAV files with a polyphase resampler (scipy).  It is not sample-identical
to librosa's default `soxr_hq` resampler; files already at 16 kHz are returned exactly as librosa would.
"""
from __future__ import annotations

import struct
from fractions import Fraction

import numpy as np


def _read_wav(path):
    with open(path, "rb") as f:
        head = f.read(12)
        if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
            raise ValueError(f"{path}: not a RIFF/WAVE file (compressed formats such as MP3 need an external decoder)")
        fmt = data = None
        while True:
            chunk = f.read(8)
            if len(chunk) < 8:
                break
            cid, size = chunk[:4], struct.unpack("<I", chunk[4:])[0]
            body = f.read(size + (size & 1))
            if cid == b"fmt ":
                fmt = body[:size]
            elif cid == b"data":
                data = body[:size]
        if fmt is None or data is None:
            raise ValueError(f"{path}: missing fmt or data chunk")
    tag, channels, rate, _, _, bits = struct.unpack("<HHIIHH", fmt[:16])
    if tag == 0xFFFE and len(fmt) >= 26:                     # WAVE_FORMAT_EXTENSIBLE: real tag in the sub-format GUID
        tag = struct.unpack("<H", fmt[24:26])[0]
    if tag == 1:                                             # integer PCM
        if bits == 8:
            x = (np.frombuffer(data, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(data, dtype="<i2").astype(np.float32) / 32768.0
        elif bits == 24:
            b = np.frombuffer(data, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            x = (np.where(v >= 1 << 23, v - (1 << 24), v)).astype(np.float32) / float(1 << 23)
        elif bits == 32:
            x = np.frombuffer(data, dtype="<i4").astype(np.float32) / float(1 << 31)
        else:
            raise ValueError(f"{path}: unsupported PCM width {bits}")
    elif tag == 3:                                           # IEEE float
        x = np.frombuffer(data, dtype="<f4" if bits == 32 else "<f8").astype(np.float32)
    else:
        raise ValueError(f"{path}: unsupported WAV format tag {tag}")
    return x.reshape(-1, channels), rate


def load_audio(path, sr: int = 16000) -> np.ndarray:
    """Mono float32 waveform at `sr` Hz."""
    x, rate = _read_wav(path)
    mono = x.mean(axis=1).astype(np.float32)
    if rate != sr:
        from scipy.signal import resample_poly
        ratio = Fraction(sr, rate)
        mono = resample_poly(mono, ratio.numerator, ratio.denominator).astype(np.float32)
    return mono
