"""Diagnostic (not a test): how far the CUDA path and the fp32 CPU oracle each are from a float64 run of
the same model, and how many emitted indices each gets "wrong" against float64.  Run on the GPU box:
    python tests/diag_precision.py [bs]   -> JSON on stdout (committed under profiles/)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from synthetic_models import build_product  # noqa: E402
from oracle import emage_oracle as O  # noqa: E402
from oracle.weights import make_checkpoint, synth_audio  # noqa: E402
from pantomatrix_b200.pipeline import generate  # noqa: E402

PARTS = ("face", "upper", "hands", "lower")


def main():
    bs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    torch.set_num_threads(os.cpu_count())
    audio = torch.from_numpy(synth_audio(bs, 160000, 1234))
    spk = torch.zeros(bs, 1, dtype=torch.long)
    sd32, cfg, vq32 = make_checkpoint(0)
    sd64, _, vq64 = make_checkpoint(0, dtype=torch.float64)
    model, vqm = build_product(0)
    tr32, tr64 = [], []
    with torch.no_grad():
        lat32, _ = O.emage_generate(sd32, cfg, vq32, audio, spk, trace=tr32)
        lat64, _ = O.emage_generate(sd64, cfg, vq64, audio.double(), spk, trace=tr64)
    out = {"bs": bs, "windows": []}
    # teacher-forced on the fp64 run's window inputs, so the three implementations see identical inputs
    for w in tr64:
        a, m, k = w["audio"].float(), w["motion"].float(), w["mask"].float()
        with torch.no_grad():
            o32 = O.emage_forward(sd32, a, spk, m, k)
        og = model.forward(a.cuda(), spk.cuda(), m.cuda(), k.cuda())
        rec = {}
        for p in PARTS:
            for kind in ("rec_", "cls_"):
                t64 = w["out"][kind + p]
                rec[kind + p] = {"gpu_max": (og[kind + p].cpu().double() - t64).abs().max().item(),
                                 "cpu32_max": (o32[kind + p].double() - t64).abs().max().item(),
                                 "gpu_rms": (og[kind + p].cpu().double() - t64).pow(2).mean().sqrt().item(),
                                 "cpu32_rms": (o32[kind + p].double() - t64).pow(2).mean().sqrt().item()}
            i64 = w["out"]["cls_" + p].argmax(-1)
            rec["idx_" + p] = {"n": i64.numel(), "gpu_flips": int((og["cls_" + p].argmax(-1).cpu() != i64).sum()),
                               "cpu32_flips": int((o32["cls_" + p].argmax(-1) != i64).sum())}
        cb64 = vq64["face"][0]["quantizer.embedding.weight"]
        f64 = O.l2_argmin(w["out"]["rec_face"], cb64)
        fg = vqm.vq_model_face._index_of(og["rec_face"]).cpu()
        f32 = O.l2_argmin(o32["rec_face"], cb64.float())
        rec["idx_face_l2"] = {"n": f64.numel(), "gpu_flips": int((fg != f64).sum()), "cpu32_flips": int((f32 != f64).sum())}
        out["windows"].append(rec)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
