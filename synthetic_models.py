"""Benchmark / test support: the drop-in modules filled with the deterministic synthetic checkpoint.

There is no network, so the Hugging Face checkpoints the reference downloads are not available; every tensor of a
reference-layout state_dict is drawn from the name-keyed generator in oracle/weights.py (pure data generation - no
oracle arithmetic) and loaded through load_state_dict(strict=True), i.e. through the checkpoint boundary.
Used by bench.py, __graft_entry__.smoke(), tools/ and tests/; the product package never imports it."""
from __future__ import annotations

from oracle.weights import EMAGE_CFG, LSTM_CFG, VQ_CFGS, load_synthetic


def build_product(seed=0, device="cuda"):
    """(EmageAudioModel, EmageVQModel) of pantomatrix_b200.emage_audio with synthetic weights."""
    from pantomatrix_b200.emage_audio import (EmageAudioConfig, EmageAudioModel, EmageVAEConv, EmageVAEConvConfig,
                                              EmageVQModel, EmageVQVAEConv, EmageVQVAEConvConfig)
    model = load_synthetic(EmageAudioModel(EmageAudioConfig(**EMAGE_CFG)), seed, "emage").to(device).eval()
    vq = {p: load_synthetic(EmageVQVAEConv(EmageVQVAEConvConfig(**VQ_CFGS[p])), seed, "vq_" + p).to(device).eval()
          for p in ("face", "upper", "hands", "lower")}
    glob = load_synthetic(EmageVAEConv(EmageVAEConvConfig(**VQ_CFGS["global"])), seed, "vq_global").to(device).eval()
    vqm = EmageVQModel(face_model=vq["face"], upper_model=vq["upper"], lower_model=vq["lower"],
                       hands_model=vq["hands"], global_model=glob).to(device).eval()
    return model, vqm


def build_lstm_product(kind, seed=0, device="cuda"):
    """CamnAudioModel ("camn") or DiscoAudioModel ("disco") of pantomatrix_b200.lstm_audio with synthetic weights."""
    from pantomatrix_b200.lstm_audio import CamnAudioConfig, CamnAudioModel, DiscoAudioConfig, DiscoAudioModel
    cls, ccls = (CamnAudioModel, CamnAudioConfig) if kind == "camn" else (DiscoAudioModel, DiscoAudioConfig)
    return load_synthetic(cls(ccls(**LSTM_CFG)), seed, kind).to(device).eval()
