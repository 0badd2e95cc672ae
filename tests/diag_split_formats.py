"""Diagnostic (not a test, CPU only): which operand format gives fp32-quality GEMMs with the fewest tensor-core
products?  Runs the product's host schedule with the kernel wrappers replaced by tests/fake_ops.py (the same
stand-ins the host-logic tests use) and swaps the emulated operand split:

  bf16x6   3 bf16 planes, 6 products (the shipped default)        bf16x3   2 bf16 planes, 3 products
  fp16x3   2 fp16 planes, 3 products; weights scaled per tensor by a power of two into the top of the fp16 range,
           activations scaled by (none | a fixed 2^4 | a per-tensor power of two); "nothing-scaled" = raw fp16 casts

Every window of a 2-clip case is teacher-forced on the float64 oracle's inputs; reported: max / rms error of the
latents and logits against float64 and the number of emitted codes that differ.  Products are accumulated in fp32
by torch (no tensor-core truncation), so this isolates the REPRESENTATION error of each format.

    python tests/diag_split_formats.py  ->  JSON lines (committed as profiles/split_formats_r1.json)"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fake_ops  # noqa: E402
from synthetic_models import build_product  # noqa: E402
from oracle import emage_oracle as O  # noqa: E402
from oracle.weights import make_checkpoint, synth_audio  # noqa: E402

PARTS = ("face", "upper", "hands", "lower")
FP16_TOP = 32768.0
SEEN = {"act_max": 0.0, "w_max": 0.0, "w_min_tensor_max": float("inf")}


def _pow2_scale(v):
    m = float(v.abs().max())
    return 1.0 if m == 0.0 or not math.isfinite(m) else 2.0 ** math.floor(math.log2(FP16_TOP / m))


FLUSH = {"on": False}      # emulate a tensor core that flushes fp16 subnormal operands to zero (worst case)


def make_splitter(fmt, act_policy):
    """-> split(v, nsplit, is_weight): list of fp32 tensors whose sum approximates v the way the format would."""
    def split(v, nsplit, is_weight):
        m = float(v.abs().max()) if v.numel() else 0.0
        if is_weight:
            SEEN["w_max"] = max(SEEN["w_max"], m)
            SEEN["w_min_tensor_max"] = min(SEEN["w_min_tensor_max"], m)
        else:
            SEEN["act_max"] = max(SEEN["act_max"], m)
        if fmt == "bf16":
            planes, rem = [], v
            for _ in range(nsplit):
                p = rem.to(torch.bfloat16).float()
                planes.append(p)
                rem = rem - p
            return planes
        if act_policy == "raw":                      # nothing scaled at all, weights included
            scale = 1.0
        else:
            scale = _pow2_scale(v) if (is_weight or act_policy == "dyn") else (float(2 ** int(act_policy[6:])) if act_policy.startswith("static") else 1.0)
        planes, rem = [], v * scale
        for _ in range(nsplit):
            p = rem.to(torch.float16).float()           # saturates to inf on overflow, like the hardware convert
            if FLUSH["on"]:
                p = torch.where(p.abs() < 2.0 ** -14, torch.zeros_like(p), p)
            planes.append(p / scale)
            rem = rem - p
        return planes
    return split


def install(split):
    import pantomatrix_b200.ops as real
    from pantomatrix_b200.emage_audio import engine, modeling
    for name in dir(fake_ops):
        if not name.startswith("_") and callable(getattr(fake_ops, name)) and hasattr(real, name):
            setattr(real, name, getattr(fake_ops, name))
    modeling._require_cuda = lambda module, what: torch.device("cpu")
    fake_ops.PLANE_DTYPE = torch.float32
    real._PLANE_DTYPE = torch.float32             # the emulated planes are stored de-scaled in fp32
    fake_ops._split = lambda v, nsplit: split(v, nsplit, False)

    class EmuPackedW(real.PackedW):
        def __init__(self, w, nsplit):
            taps, cout, cin = w.shape
            bn = 64 if cout <= 64 else 128
            self.taps, self.cout, self.cin = taps, cout, cin
            self.w_rows, self.ldw = real._round_up(cout, bn), real._round_up(cin, 8)
            full = torch.zeros(taps, self.w_rows, self.ldw)
            full[:, :cout, :cin] = w
            self.t = torch.stack(split(full, nsplit, True)).contiguous()
            self.acc_scale = 1.0                     # the emulated planes are stored de-scaled
    real.PackedW = EmuPackedW
    fake_ops.PackedW = EmuPackedW
    return engine


def main():
    bs, n = 2, 70000
    audio = torch.from_numpy(synth_audio(bs, n, 1234))
    spk = torch.zeros(bs, 1, dtype=torch.long)
    sd64, cfg, vq64 = make_checkpoint(0, dtype=torch.float64)
    tr64 = []
    with torch.no_grad():
        O.emage_generate(sd64, cfg, vq64, audio.double(), spk, trace=tr64)
    modes = [("fp32", 0, "bf16", None), ("bf16x3", 2, "bf16", None), ("bf16x6", 3, "bf16", None),
             ("fp16x3/nothing-scaled", 2, "fp16", "raw"), ("fp16x3/act-unscaled", 2, "fp16", "none"), ("fp16x3/act-x16", 2, "fp16", "static4"),
             ("fp16x3/act-per-tensor", 2, "fp16", "dyn"),
             ("fp16x3/act-unscaled, subnormals flushed", 2, "fp16", "none"), ("fp16x3/act-x16, subnormals flushed", 2, "fp16", "static4"),
             ("fp16x3/act-per-tensor, subnormals flushed", 2, "fp16", "dyn"),
             ("fp16x3/act-x64, subnormals flushed", 2, "fp16", "static6"), ("fp16x3/act-x256, subnormals flushed", 2, "fp16", "static8"),
             ("fp16x3/act-x256", 2, "fp16", "static8")]
    for name, nsplit, fmt, pol in modes:
        FLUSH["on"] = "flushed" in name
        engine = install(make_splitter(fmt, pol))
        engine._STATE["nsplit"] = nsplit
        model, vqm = build_product(seed=0, device="cpu")
        worst = {"rec": 0.0, "cls": 0.0}
        sq, cnt, flips, total, face_flips = 0.0, 0, 0, 0, 0
        for w in tr64:
            a, m, k = w["audio"].float(), w["motion"].float(), w["mask"].float()
            with torch.no_grad():
                og = model.forward(a, spk, m, k)
            for p in PARTS:
                for kind in ("rec", "cls"):
                    d = og[f"{kind}_{p}"].double() - w["out"][f"{kind}_{p}"]
                    worst[kind] = max(worst[kind], d.abs().max().item())
                    if kind == "rec":
                        sq += d.pow(2).sum().item()
                        cnt += d.numel()
                i64 = w["out"]["cls_" + p].argmax(-1)
                flips += int((og["cls_" + p].argmax(-1) != i64).sum())
                total += i64.numel()
            cb64 = vq64["face"][0]["quantizer.embedding.weight"]
            face_flips += int((O.l2_argmin(og["rec_face"].double(), cb64) != O.l2_argmin(w["out"]["rec_face"], cb64)).sum())
        print(json.dumps({"mode": name, "tensor_core_products": {0: 0, 2: 3, 3: 6}[nsplit], "windows": len(tr64),
                          "rec_max_err": worst["rec"], "rec_rms_err": math.sqrt(sq / cnt), "cls_max_err": worst["cls"],
                          "argmax_flips": flips, "face_l2_flips": face_flips, "codes": total,
                          "largest_gemm_input": SEEN["act_max"], "largest_weight": SEEN["w_max"]}), flush=True)


if __name__ == "__main__":
    main()
