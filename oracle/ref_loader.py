"""Import and drive the UNMODIFIED reference (byte-compiled into oracle/_ref by oracle/make_ref.py, or a live
/root/reference tree) exactly like its demo does (test_emage_audio.py:16-47).  TEST / MEASUREMENT INFRASTRUCTURE:
used by bench.py's reference arm and by tests; never by the product.

Accommodations: a stub `omegaconf` module (imported at configuration_emage_audio.py:2, unused
when config_obj is None) and no `from_pretrained` (needs the HF hub): modules are built from Config(**dims) and filled
through load_state_dict(strict=True) with the synthetic checkpoint of oracle/weights.py.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")


def staged_available() -> bool:
    return os.path.exists(os.path.join(STAGED, "MANIFEST.json"))


def _verify(root):
    with open(os.path.join(root, "MANIFEST.json")) as f:
        files = json.load(f)["files"]
    for rel, rec in files.items():
        with open(os.path.join(root, rel), "rb") as f:
            if hashlib.sha256(f.read()).hexdigest() != rec["sha256"]:
                raise RuntimeError(f"oracle/_ref/{rel} does not match its manifest hash: re-run oracle/make_ref.py")


def import_reference(root: str | None = None):
    """Returns the reference's `models` namespace with emage_audio / camn_audio / disco_audio imported from `root`
    (default: the staged copy).  The repo root has a `models/` shim package of its own: the reference tree must win
    here, so cached `models*` modules are dropped first and restored by nobody (callers are short-lived processes)."""
    root = root or STAGED
    if root == STAGED:
        _verify(root)
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]
    stub = types.ModuleType("omegaconf")
    stub.OmegaConf = type("OmegaConf", (), {})
    sys.modules.setdefault("omegaconf", stub)
    saved = list(sys.path)
    repo = os.path.dirname(HERE)       # this repo's own `models/` shim must not be visible while the reference imports
    sys.path[:] = [root] + [p for p in saved if os.path.realpath(p or ".") != os.path.realpath(repo)]
    try:
        import models.emage_audio as emage
        import models.camn_audio as camn
        import models.disco_audio as disco
    finally:
        sys.path[:] = saved
    for m in (emage, camn, disco):
        assert os.path.realpath(m.__file__).startswith(os.path.realpath(root)), m.__file__
    return types.SimpleNamespace(emage=emage, camn=camn, disco=disco)


def build_emage(ref, seed=0):
    from oracle.weights import EMAGE_CFG, VQ_CFGS, load_synthetic
    e = ref.emage
    model = load_synthetic(e.EmageAudioModel(e.EmageAudioConfig(**EMAGE_CFG)).eval(), seed, "emage")
    vq = {p: load_synthetic(e.EmageVQVAEConv(e.EmageVQVAEConvConfig(**VQ_CFGS[p])).eval(), seed, "vq_" + p)
          for p in ("face", "upper", "hands", "lower")}
    glob = load_synthetic(e.EmageVAEConv(e.EmageVAEConvConfig(**VQ_CFGS["global"])).eval(), seed, "vq_global")
    vqm = e.EmageVQModel(face_model=vq["face"], upper_model=vq["upper"], lower_model=vq["lower"],
                         hands_model=vq["hands"], global_model=glob).eval()
    return model, vqm


def build_lstm(ref, kind, seed=0):
    from oracle.weights import LSTM_CFG, load_synthetic
    mod = ref.camn if kind == "camn" else ref.disco
    cls, ccls = ((mod.CamnAudioModel, mod.CamnAudioConfig) if kind == "camn" else (mod.DiscoAudioModel, mod.DiscoAudioConfig))
    return load_synthetic(cls(ccls(**LSTM_CFG)).eval(), seed, kind)


def drive_like_demo(model, vqm, audio):
    """The timed span of the reference demo, test_emage_audio.py:32-47 (inference + index selection + final decode)."""
    import torch
    import torch.nn.functional as F
    speaker_id = torch.zeros(audio.shape[0], 1).long()
    trans = torch.zeros(1, 1, 3)
    cfg = model.cfg
    with torch.no_grad():
        lat = model.inference(audio, speaker_id, vqm, masked_motion=None, mask=None)
        pick = lambda p, l, c: lat["rec_" + p] if l > 0 and c == 0 else None
        index = lambda p, c: torch.max(F.log_softmax(lat["cls_" + p], dim=2), dim=2)[1] if c > 0 else None
        pred = vqm.decode(
            face_latent=pick("face", cfg.lf, cfg.cf), upper_latent=pick("upper", cfg.lu, cfg.cu),
            lower_latent=pick("lower", cfg.ll, cfg.cl), hands_latent=pick("hands", cfg.lh, cfg.ch),
            face_index=index("face", cfg.cf), upper_index=index("upper", cfg.cu),
            lower_index=index("lower", cfg.cl), hands_index=index("hands", cfg.ch),
            get_global_motion=True, ref_trans=trans[:, 0])
    return lat, pred
