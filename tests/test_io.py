"""Callers either side of the hot path: npz writer / linear upsampling and the WAV reader."""
import os
import struct
import wave

import numpy as np
import pytest

from pantomatrix_b200 import audio_io, motion_io


def test_time_upsample_matches_reference_golden(golden_dir):
    """Golden produced by executing the reference's own time_upsample_numpy (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(golden_dir, "upsample.npz"))
    for k in (1, 2, 3):
        got = motion_io.time_upsample_numpy(g["x"], k)
        np.testing.assert_allclose(got, g[f"k{k}"], rtol=0, atol=1e-6)
    assert motion_io.time_upsample_numpy(g["x"], 2).shape == (2, 14, 5)


def test_time_upsample_is_piecewise_linear():
    x = np.random.default_rng(0).standard_normal((9, 4))
    got = motion_io.time_upsample_numpy(x, 4)
    pos = np.linspace(0, 8, 36)
    want = np.stack([np.interp(pos, np.arange(9), x[:, c]) for c in range(4)], axis=1)
    np.testing.assert_allclose(got, want, atol=1e-12)
    assert np.array_equal(got[0], x[0]) and np.allclose(got[-1], x[-1])


def test_beat_format_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    poses, expr, trans = rng.standard_normal((20, 165)).astype(np.float32), rng.standard_normal((20, 100)).astype(np.float32), rng.standard_normal((20, 3)).astype(np.float32)
    path = str(tmp_path / "clip_output.npz")
    motion_io.beat_format_save(path, poses, expressions=expr, trans=trans, upsample=1)
    raw = np.load(path, allow_pickle=True)
    assert raw["betas"].shape == (300,) and str(raw["model"]) == "smplx2020" and str(raw["gender"]) == "neutral"
    assert int(raw["mocap_frame_rate"]) == 30
    back = motion_io.beat_format_load(path)
    assert np.array_equal(back["poses"], poses) and np.array_equal(back["expressions"], expr) and np.array_equal(back["trans"], trans)
    mask = [j % 2 == 0 for j in range(55)]
    sel = motion_io.select_with_mask(poses, mask)
    assert sel.shape == (20, 28 * 3) and np.array_equal(motion_io.select_with_mask(motion_io.recover_from_mask(sel, mask), mask), sel)
    motion_io.beat_format_save(path, sel, mask=mask, trans=trans, upsample=2)
    assert np.load(path)["poses"].shape == (40, 165)
    with pytest.raises(NotImplementedError):
        motion_io.beat_format_save(path, poses)                 # trans=None needs the licensed SMPL-X model


def test_wav_reader_and_resampling(tmp_path):
    sr, n = 48000, 48000
    t = np.arange(n) / sr
    stereo = np.stack([np.sin(2 * np.pi * 440 * t), 0.5 * np.sin(2 * np.pi * 220 * t)], axis=1)
    path = str(tmp_path / "a.wav")
    with wave.open(path, "wb") as w:
        w.setnchannels(2), w.setsampwidth(2), w.setframerate(sr)
        w.writeframes((stereo * 32767).astype("<i2").tobytes())
    x = audio_io.load_audio(path, sr=16000)
    assert x.dtype == np.float32 and abs(len(x) - 16000) <= 1 and np.abs(x).max() <= 1.0
    want = (stereo.mean(1))[::3][:len(x)]
    assert np.abs(x[100:-100] - want[100:-100]).max() < 2e-2        # band-limited resampling of two low tones
    same = audio_io.load_audio(path, sr=48000)
    np.testing.assert_allclose(same, ((stereo * 32767).astype("<i2") / 32768).mean(1).astype(np.float32), atol=1e-6)
    p32 = str(tmp_path / "f.wav")
    data = stereo[:, 0].astype("<f4").tobytes()
    with open(p32, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 3, 1, 16000, 64000, 4, 32)
                + b"data" + struct.pack("<I", len(data)) + data)
    np.testing.assert_array_equal(audio_io.load_audio(p32), stereo[:, 0].astype(np.float32))
    bad = str(tmp_path / "mp3.wav")
    open(bad, "wb").write(b"ID3\x04" + b"\0" * 64)
    with pytest.raises(ValueError):
        audio_io.load_audio(bad)
