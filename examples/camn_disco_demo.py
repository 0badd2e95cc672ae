#!/usr/bin/env python
"""Audio folder -> BEAT-format npz with CaMN or DisCo on the B200 path (reference test_camn_audio.py / test_disco_audio.py).

    python examples/camn_disco_demo.py --model camn --checkpoint /path/to/camn_audio --audio_folder ./wavs
    python examples/camn_disco_demo.py --model disco --synthetic --audio_folder ./wavs

These models emit the upper body + hands only and no translation; like the reference demos this needs the SMPL-X body
model to place the pelvis when writing the npz, which is not available offline - so here the pelvis translation is
written as zeros (pass --trans-zero explicitly to acknowledge)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from models.camn_audio import CamnAudioModel  # noqa: E402
from models.disco_audio import DiscoAudioModel  # noqa: E402
from pantomatrix_b200.audio_io import load_audio  # noqa: E402
from pantomatrix_b200.motion_io import beat_format_save  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", choices=["camn", "disco"], default="camn")
    ap.add_argument("--audio_folder", default="./examples/audio")
    ap.add_argument("--save_folder", default="./examples/motion")
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--trans-zero", action="store_true")
    args = ap.parse_args()
    if not args.trans_zero:
        ap.error("the npz needs a pelvis translation; without the SMPL-X model files pass --trans-zero to write zeros")
    device = torch.device("cuda")
    if args.synthetic:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from synthetic_models import build_lstm_product
        model = build_lstm_product(args.model, device=device)
    else:
        cls = CamnAudioModel if args.model == "camn" else DiscoAudioModel
        model = cls.from_pretrained(args.checkpoint).to(device).eval()
    os.makedirs(args.save_folder, exist_ok=True)
    sr, fps, seed_frames = model.cfg.audio_sr, model.cfg.pose_fps, model.cfg.seed_frames
    frames, t0 = 0, time.time()
    for name in sorted(f for f in os.listdir(args.audio_folder) if f.endswith(".wav")):
        audio = torch.from_numpy(load_audio(os.path.join(args.audio_folder, name), sr=sr)).unsqueeze(0).to(device)
        aa = model(audio, torch.zeros(1, 1, dtype=torch.long, device=device), seed_frames=seed_frames)["motion_axis_angle"]
        t = aa.shape[1]
        beat_format_save(os.path.join(args.save_folder, os.path.splitext(name)[0] + "_output.npz"),
                         aa.cpu().numpy().reshape(t, -1), upsample=30 // fps, trans=np.zeros((t, 3), dtype=np.float32))
        frames += t
    print(f"generate total {frames / fps:.2f} seconds motion in {time.time() - t0:.2f} seconds, saved in {args.save_folder}")


if __name__ == "__main__":
    main()
