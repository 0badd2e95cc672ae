"""Golden for EmageAudioModel.forward() itself (one window, user masked_motion / mask), with use_audio=True and with
the training-time ablation use_audio=False (M.py:310-311), from the UNMODIFIED reference in /root/reference:

    python tests/golden/make_golden_forward.py        -> tests/golden/case_forward.npz

Inputs are regenerated from seeds by the tests; outputs are stored every 3rd frame (all 256 channels)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from make_golden import build_reference, import_reference  # noqa: E402

BS, N, T = 2, 34112, 64


def inputs():
    sys.path.insert(0, ROOT)
    from oracle.weights import synth_audio
    g = torch.Generator().manual_seed(17)
    audio = torch.from_numpy(synth_audio(BS, N, 55))
    motion = torch.randn(BS, T, 337, generator=g) * 0.3
    mask = (torch.rand(BS, T, 337, generator=g) > 0.3).float()
    return audio, motion, mask


def main():
    ref = import_reference()
    model, _, _, _ = build_reference(ref, seed=0)
    audio, motion, mask = inputs()
    spk = torch.zeros(BS, 1, dtype=torch.long)
    out = {}
    with torch.no_grad():
        for tag, flag in (("audio", True), ("noaudio", False)):
            res = model.forward(audio, spk, motion, mask, use_audio=flag)
            for k, v in res.items():
                out[f"{tag}_{k}"] = v.numpy().astype(np.float32)[:, ::3]
                if k.startswith("cls_"):
                    out[f"{tag}_idx_{k}"] = v.argmax(-1).numpy().astype(np.int16)
    np.savez_compressed(os.path.join(HERE, "case_forward.npz"), **out)
    print({k: v.shape for k, v in out.items()})
    print("body logits changed by the ablation:", float(np.abs(out["audio_cls_upper"] - out["noaudio_cls_upper"]).max()),
          "face unchanged:", float(np.abs(out["audio_cls_face"] - out["noaudio_cls_face"]).max()))


if __name__ == "__main__":
    main()
