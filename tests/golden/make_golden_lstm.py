"""Golden fixtures for the CaMN / DisCo paths from the UNMODIFIED reference modules (build container only).

    python tests/golden/make_golden_lstm.py

Adds the `camn` / `disco` state-dict manifests to state_dict_manifest.json and writes case_camn.npz, case_disco.npz:
outputs of reference CamnAudioModel / DiscoAudioModel forward() (test_camn_audio.py:16-21 call form) on the synthetic
checkpoint (oracle/weights.py) and synthetic audio, with and without a user seed motion."""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PM_REFERENCE", "/root/reference")


def main():
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]
    sys.path.insert(0, REF)
    stub = types.ModuleType("omegaconf")
    stub.OmegaConf = type("OmegaConf", (), {})
    sys.modules.setdefault("omegaconf", stub)
    from models.camn_audio import CamnAudioConfig, CamnAudioModel
    from models.disco_audio import DiscoAudioConfig, DiscoAudioModel
    sys.path.remove(REF)
    sys.path.insert(0, ROOT)
    from oracle.weights import LSTM_CFG, load_synthetic, synth_audio

    path = os.path.join(HERE, "state_dict_manifest.json")
    manifest = json.load(open(path))
    for kind, cls, ccls in (("camn", CamnAudioModel, CamnAudioConfig), ("disco", DiscoAudioModel, DiscoAudioConfig)):
        model = cls(ccls(**LSTM_CFG)).eval()
        manifest[kind] = [(k, list(v.shape)) for k, v in model.state_dict().items()]
        load_synthetic(model, 0, kind)
        bs, n = 3, 48000                                      # 3 s -> 44 frames @ 15 fps
        audio = torch.from_numpy(synth_audio(bs, n, 4242))
        spk = torch.zeros(bs, 1, dtype=torch.long)
        g = np.random.Generator(np.random.PCG64(5))
        seed_motion = torch.from_numpy(g.standard_normal((bs, 30, 258)).astype(np.float32) * 0.3)   # shorter than t
        out = {"bs": np.int64(bs), "n_samples": np.int64(n), "audio_seed": np.int64(4242), "seed_motion": seed_motion.numpy()}
        with torch.no_grad():
            a = model(audio, spk, seed_frames=4, seed_motion=None)
            b = model(audio, spk, seed_frames=4, seed_motion=seed_motion)
        out["motion"] = a["motion"].reshape(bs, a["motion"].shape[1], -1).numpy()
        out["motion_axis_angle"] = a["motion_axis_angle"].numpy()
        out["seeded_motion"] = b["motion"].reshape(bs, b["motion"].shape[1], -1).numpy()
        if kind == "disco":
            out["audio_fea_c"], out["audio_fea_r"] = a["audio_fea_c"].numpy(), a["audio_fea_r"].numpy()
        np.savez_compressed(os.path.join(HERE, f"case_{kind}.npz"), **out)
        print(kind, len(manifest[kind]), "tensors,", sum(int(np.prod(s)) for _, s in manifest[kind]) / 1e6, "M params, motion",
              out["motion"].shape, "std", float(out["motion"].std()), "time-std", float(out["motion"].std(1).mean()))
    json.dump(manifest, open(path, "w"), indent=0)


if __name__ == "__main__":
    main()
