"""Generate the golden fixtures that pin oracle/emage_oracle.py to the REAL reference.

Run in the build container only (it imports the unmodified reference from /root/reference,
which does not exist on the GPU box):

    python tests/golden/make_golden.py

Writes, next to this file:
  state_dict_manifest.json   key names + shapes of the reference checkpoints (the HF layout
                             the drop-in modules must load with strict=True)
  case_*.npz                 outputs of reference EmageAudioModel.inference() + the final
                             EmageVQModel.decode(get_global_motion=True), driven exactly like
                             /root/reference/test_emage_audio.py:16-47, on synthetic weights
                             (oracle/weights.py) and synthetic audio.

The only accommodation made to import the reference is a stub `omegaconf` module (imported at
configuration_emage_audio.py:2 but unused when config_obj is None); `from_pretrained` is not
used (it needs the HF hub).  Nothing of the reference is copied.
"""
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PM_REFERENCE", "/root/reference")


def import_reference():
    # The repo root also has a `models/` shim package; the reference tree must win here.
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]
    sys.path.insert(0, REF)
    stub = types.ModuleType("omegaconf")
    stub.OmegaConf = type("OmegaConf", (), {})
    sys.modules.setdefault("omegaconf", stub)
    import models.emage_audio as ref
    assert os.path.realpath(ref.__file__).startswith(os.path.realpath(REF)), ref.__file__
    sys.path.remove(REF)
    return ref


def build_reference(ref, seed):
    sys.path.insert(0, ROOT)
    from oracle.weights import EMAGE_CFG, VQ_CFGS, load_synthetic
    model = load_synthetic(ref.EmageAudioModel(ref.EmageAudioConfig(**EMAGE_CFG)).eval(), seed, "emage")
    vq = {p: load_synthetic(ref.EmageVQVAEConv(ref.EmageVQVAEConvConfig(**VQ_CFGS[p])).eval(), seed, "vq_" + p)
          for p in ("face", "upper", "hands", "lower")}
    glob = load_synthetic(ref.EmageVAEConv(ref.EmageVAEConvConfig(**VQ_CFGS["global"])).eval(), seed, "vq_global")
    vqm = ref.EmageVQModel(face_model=vq["face"], upper_model=vq["upper"], lower_model=vq["lower"],
                           hands_model=vq["hands"], global_model=glob).eval()
    return model, vqm, vq, glob


def drive_like_demo(model, vqm, audio, masked_motion=None, mask=None):
    """Same call sequence as reference test_emage_audio.py:16-47 (minus librosa / npz)."""
    speaker_id = torch.zeros(audio.shape[0], 1).long()
    trans = torch.zeros(1, 1, 3)
    cfg = model.cfg
    with torch.no_grad():
        lat = model.inference(audio, speaker_id, vqm, masked_motion=masked_motion, mask=mask)
        pick = lambda p, l, c: lat["rec_" + p] if l > 0 and c == 0 else None
        index = lambda p, c: torch.max(F.log_softmax(lat["cls_" + p], dim=2), dim=2)[1] if c > 0 else None
        pred = vqm.decode(
            face_latent=pick("face", cfg.lf, cfg.cf), upper_latent=pick("upper", cfg.lu, cfg.cu),
            lower_latent=pick("lower", cfg.ll, cfg.cl), hands_latent=pick("hands", cfg.lh, cfg.ch),
            face_index=index("face", cfg.cf), upper_index=index("upper", cfg.cu),
            lower_index=index("lower", cfg.cl), hands_index=index("hands", cfg.ch),
            get_global_motion=True, ref_trans=trans[:, 0])
    return lat, pred


# name -> (bs, n_samples, seed-motion frames or 0).  Frame count L = n*30//16000.
CASES = {
    "tail11": (2, 70000, 0),      # L=131: 2 full windows + tail window of 11 frames
    "clip10s": (1, 160000, 0),    # BASELINE configs[0]: L=300, 4 windows + tail of 60
    "drop_tail": (2, 66134, 0),   # L=124: remain==0 -> output is 120 frames (tail dropped)
    "short40": (2, 21600, 0),     # L=40 < 64: single tail window
    "seeded": (2, 70000, 8),      # user-supplied masked_motion/mask for the first 8 frames
}


def main():
    sys.path.insert(0, ROOT)
    from oracle.weights import synth_audio
    ref = import_reference()
    model, vqm, vq, glob = build_reference(ref, seed=0)

    manifest = {"emage": [(k, list(v.shape)) for k, v in model.state_dict().items()]}
    for p, m in vq.items():
        manifest["vq_" + p] = [(k, list(v.shape)) for k, v in m.state_dict().items()]
    manifest["vq_global"] = [(k, list(v.shape)) for k, v in glob.state_dict().items()]
    with open(os.path.join(HERE, "state_dict_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=0)

    for name, (bs, n, n_seed) in CASES.items():
        audio = torch.from_numpy(synth_audio(bs, n, seed=1234))
        mm = mk = None
        if n_seed:
            g = np.random.Generator(np.random.PCG64(99))
            aa = torch.from_numpy(g.standard_normal((bs, n_seed, 55, 3)).astype(np.float32) * 0.3)
            sys.path.insert(0, ROOT)
            from oracle.emage_oracle import axis_angle_to_rot6d
            mm = torch.cat([axis_angle_to_rot6d(aa).reshape(bs, n_seed, 330),
                            torch.from_numpy(g.standard_normal((bs, n_seed, 7)).astype(np.float32) * 0.1)], -1)
            mk = torch.zeros(bs, n_seed, 337)
            mk[:, :, 300:] = 1.0      # partially masked seed: exercises the per-element `where`
        lat, pred = drive_like_demo(model, vqm, audio, mm, mk)
        out = {"audio_seed": np.int64(1234), "bs": np.int64(bs), "n_samples": np.int64(n)}
        if mm is not None:
            out["masked_motion"], out["mask"] = mm.numpy(), mk.numpy()
        for p in ("face", "upper", "hands", "lower"):
            out["idx_cls_" + p] = lat["cls_" + p].argmax(-1).numpy().astype(np.int16)
            out["rec_" + p] = lat["rec_" + p].numpy().astype(np.float32)[:, ::7]      # every 7th frame
            out["cls_" + p] = lat["cls_" + p].numpy().astype(np.float32)[:, ::7]
        cb = vq["face"].quantizer.embedding.weight
        z = lat["rec_face"].reshape(-1, 256)
        d = (z ** 2).sum(1, keepdim=True) + (cb ** 2).sum(1) - 2 * z @ cb.t()
        out["idx_l2_face"] = d.argmin(1).reshape(lat["rec_face"].shape[:2]).numpy().astype(np.int16)
        for k in ("expression", "motion_axis_angle", "trans", "all_motion4inference"):
            out[k] = pred[k].numpy().astype(np.float32)
        np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim},
              "distinct idx:", {p: len(np.unique(out["idx_cls_" + p])) for p in ("upper", "hands", "lower")},
              "face-l2:", len(np.unique(out["idx_l2_face"])))


def upsample_golden():
    """Golden for the output writer's linear upsampler: executes the REFERENCE's own time_upsample_numpy.  The
    module emage_utils/motion_io.py cannot be imported here (it imports smplx at line 3), so only that function
    is compiled out of the reference source with ast - nothing is copied into the repo."""
    import ast
    src = open(os.path.join(REF, "emage_utils", "motion_io.py")).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "time_upsample_numpy"][0]
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "ref_motion_io", "exec"), ns)
    x = np.random.default_rng(7).standard_normal((2, 7, 5))
    np.savez(os.path.join(HERE, "upsample.npz"), x=x, **{f"k{k}": ns["time_upsample_numpy"](x, k) for k in (1, 2, 3)})


if __name__ == "__main__":
    main()
    upsample_golden()
