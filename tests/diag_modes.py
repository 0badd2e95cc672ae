"""Diagnostic (not a test): free-running BASELINE batch (32 clips x 300 frames) in every precision mode vs the CPU
oracle: how many emitted codes agree, pose error on the frames whose codes agree.  JSON on stdout."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import geodesic_deg  # noqa: E402
from synthetic_models import build_product  # noqa: E402
from oracle import emage_oracle as O  # noqa: E402
from oracle.weights import make_checkpoint, synth_audio  # noqa: E402
from pantomatrix_b200.emage_audio import engine  # noqa: E402
from pantomatrix_b200.pipeline import generate  # noqa: E402

PARTS = ("face", "upper", "hands", "lower")


def main():
    bs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    torch.set_num_threads(min(16, os.cpu_count()))
    audio = torch.from_numpy(synth_audio(bs, 160000, 1234))
    sd, cfg, vq = make_checkpoint(0)
    with torch.no_grad():
        want_lat, want = O.emage_generate(sd, cfg, vq, audio, torch.zeros(bs, 1, dtype=torch.long))
    want_face = O.l2_argmin(want_lat["rec_face"], vq["face"][0]["quantizer.embedding.weight"])
    model, vqm = build_product(0)
    out = {"clips": bs, "frames": int(want["motion_axis_angle"].shape[1]), "modes": {}}
    modes = ("fp32", "fp16x3", "bf16x6", "bf16x3", "bf16")
    for mode in modes:
        engine.set_precision(mode)
        lat, pred = generate(model, vqm, audio.cuda())
        rec = {}
        ok = torch.ones(bs, lat["rec_face"].shape[1], dtype=torch.bool)
        for p in PARTS[1:]:
            same = lat["cls_" + p].argmax(-1).cpu() == want_lat["cls_" + p].argmax(-1)
            rec["codes_equal_" + p] = [int(same.sum()), same.numel()]
            ok &= same
        face = vqm.vq_model_face._index_of(lat["rec_face"]).cpu() == want_face
        rec["codes_equal_face"] = [int(face.sum()), face.numel()]
        ok &= face
        # first frame (per clip) at which any code differs: everything after it follows a different seed
        first_bad = torch.where(ok.all(1), torch.full((bs,), ok.shape[1]), (~ok).float().argmax(1))
        rec["clips_fully_identical"] = int(ok.all(1).sum())
        rec["median_first_divergent_frame"] = float(first_bad.float().median())
        geo = geodesic_deg(pred["motion_axis_angle"].cpu().reshape(bs, -1, 55, 3), want["motion_axis_angle"].reshape(bs, -1, 55, 3))
        near = torch.nn.functional.max_pool1d((~ok).float().unsqueeze(1), 19, 1, 9)[:, 0] > 0
        good = ~near
        rec["geodesic_deg_median_all"] = float(geo.median())
        rec["geodesic_deg_max_where_codes_agree"] = float(geo[good].max()) if good.any() else None
        rec["max_abs_latent_err"] = max(float((lat["rec_" + p].cpu() - want_lat["rec_" + p]).abs().max()) for p in PARTS)
        out["modes"][mode] = rec
    engine.set_precision(engine.DEFAULT_PRECISION)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
