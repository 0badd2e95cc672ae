"""Deterministic synthetic checkpoints for the EMAGE hot path (TEST INFRASTRUCTURE).

There is no network, so neither the Hugging Face checkpoint the reference downloads
(/root/reference/test_emage_audio.py:82-93) nor any dataset is available.  Every
tensor of a reference-layout ``state_dict`` is therefore drawn here from a
name-keyed, platform-independent generator (numpy PCG64 seeded by CRC32(name) ^ seed),
so this container, the GPU box, the oracle and the CUDA path all see identical weights
without shipping 0.6 GB of fixtures.

Why every tensor is re-drawn (pitfall 1): under the reference's
default init ``nn.TransformerDecoder`` deep-copies one layer (all 15 decoder layers
identical), BatchNorm running stats are (0, 1) and the codebooks are U(+-1/256), so
weight-routing and BN-folding bugs would be invisible and the argmax collapses to a
handful of codes.

Nothing in here is imported by the product path (pantomatrix_b200/); only tests,
bench.py and __graft_entry__.smoke() use it.
"""
from __future__ import annotations

import zlib

import numpy as np

# Dimensions of the shipped model (reference configs/emage_audio.yaml:24-52) and of the
# VQ-VAE checkpoints (not in the repo; SURVEY.md section 0: vae_length == codebook == 256,
# vae_test_dim forced by modeling_emage_audio.py:100-107,136-168,197).
EMAGE_CFG = dict(
    pose_fps=30, motion_f=256, pose_dims=330, pose_rep="smplx", audio_rep="wave16k",
    audio_sr=16000, audio_fps=16000, audio_norm=False, audio_f=256, speaker_f=768,
    speaker_dims=1, hidden_size=768, seed_frames=4, pose_length=64, stride=20,
    test_length=64, joint_mask=None, vae_codebook_size=256,
    ll=3, lf=3, lu=3, lh=3, cl=1, cf=0, cu=1, ch=1,
)

VQ_CFGS = {
    "face": dict(vae_layer=2, vae_length=256, vae_test_dim=106, vae_codebook_size=256, vae_quantizer_lambda=1.0),
    "upper": dict(vae_layer=2, vae_length=256, vae_test_dim=78, vae_codebook_size=256, vae_quantizer_lambda=1.0),
    "hands": dict(vae_layer=2, vae_length=256, vae_test_dim=180, vae_codebook_size=256, vae_quantizer_lambda=1.0),
    "lower": dict(vae_layer=4, vae_length=256, vae_test_dim=61, vae_codebook_size=256, vae_quantizer_lambda=1.0),
    "global": dict(vae_layer=4, vae_length=256, vae_test_dim=61, vae_codebook_size=256, vae_quantizer_lambda=1.0),
}


LSTM_CFG = dict(
    pose_fps=15, motion_f=256, pose_dims=258, pose_rep="smplx", body_dims=78, hands_dims=180, audio_rep="wave16k",
    audio_sr=16000, audio_fps=16000, audio_norm=False, audio_f=128, speaker_f=16, speaker_dims=1, hidden_size=512,
    n_layer=4, dropout_prob=0.1, seed_frames=4, joint_mask="local_upper",
)      # configs/camn_audio.yaml:27-47 == configs/disco_audio.yaml:26-46 (model block)


# Gains that keep the time-varying (audio-driven) part of the activations comparable to the
# constant part, so emitted code indices are diverse (first matching substring wins).
_GAINS = (
    ("feat_extractor.0.conv1.weight", 10.0),        # raw audio is only +-0.1
    ("feat_extractor.0.downsample.0.weight", 10.0),
    ("audio_encoder_", 1.2),
    ("motion_encoder.", 1.4),
    ("_cls", 1.5),
    ("body_out.fc", 3.0),                           # CaMN / DisCo rot6d heads
    ("hands_out.fc", 3.0),
    ("audio_face_motion_proj", 4.0),                # memory scale => audio-driven face variation
    ("face_motion_decoder.", 2.0),
)


def _rng(name: str, seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([zlib.crc32(name.encode()), seed & 0xFFFFFFFF]))


def synth_tensor(name: str, shape, seed: int = 0, tag: str = "") -> np.ndarray | None:
    """One tensor of a synthetic checkpoint, chosen by the role its key name implies.

    Returns None for entries that must keep their analytic value (positional-encoding
    buffer, BatchNorm batch counter)."""
    shape = tuple(int(s) for s in shape)
    g = _rng(tag + "/" + name, seed)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked" or name.endswith("position_embeddings.pe"):
        return None
    if leaf == "running_var":
        return g.uniform(0.5, 1.5, shape).astype(np.float32)
    if leaf == "running_mean":
        return (0.1 * g.standard_normal(shape)).astype(np.float32)
    if "quantizer.embedding" in name:                      # VQ codebook rows
        return g.standard_normal(shape).astype(np.float32)
    if "_motion_decoder.weight_" in name:                  # LSTM matrices: keep the gates away from saturation
        fan_in = int(shape[1])
        return (g.standard_normal(shape) * (1.6 / np.sqrt(fan_in))).astype(np.float32)
    if name in ("body_out.fc2.bias", "hands_out.fc2.bias"):      # rot6d heads: O(1) vectors -> well-conditioned rotations
        return (0.6 * g.standard_normal(shape)).astype(np.float32)
    if "speaker_embedding" in name:
        return (0.2 * g.standard_normal(shape)).astype(np.float32)
    if name == "mask_embedding":
        return (768 ** -0.5 * g.standard_normal(shape)).astype(np.float32)
    if len(shape) == 1:
        if leaf == "weight":                               # LayerNorm / BatchNorm gain
            return (1.0 + 0.1 * g.standard_normal(shape)).astype(np.float32)
        return (0.02 * g.standard_normal(shape)).astype(np.float32)   # every bias / beta
    fan_in = int(np.prod(shape[1:]))
    gain = 1.0
    for pat, gn in _GAINS:
        if pat in name:
            gain = gn
            break
    return (g.standard_normal(shape) * (gain / np.sqrt(fan_in))).astype(np.float32)


def synth_state_dict(manifest, seed: int = 0, tag: str = ""):
    """manifest: iterable of (name, shape) -> {name: np.ndarray or None}."""
    return {name: synth_tensor(name, shape, seed, tag) for name, shape in manifest}


def load_synthetic(module, seed: int = 0, tag: str = ""):
    """Fill any torch module that exposes the reference key layout (the reference classes,
    or this repo's drop-in classes) with the synthetic checkpoint.  Uses only
    state_dict()/load_state_dict(strict=True), i.e. the checkpoint boundary."""
    import torch

    sd = module.state_dict()
    new = {}
    for name, t in sd.items():
        arr = synth_tensor(name, t.shape, seed, tag)
        new[name] = t.clone() if arr is None else torch.from_numpy(arr).to(t.dtype)
    module.load_state_dict(new, strict=True)
    return module


def synth_audio(bs: int, n_samples: int, seed: int = 1234) -> np.ndarray:
    """Synthetic 16 kHz audio, U(-0.1, 0.1)."""
    g = np.random.Generator(np.random.PCG64([0xA0D10, seed & 0xFFFFFFFF]))
    return ((g.random((bs, n_samples), dtype=np.float32) * 2.0 - 1.0) * 0.1).astype(np.float32)


def load_manifest():
    """Key names + shapes of the reference checkpoints, recorded from the live reference
    modules by tests/golden/make_golden.py."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "tests", "golden", "state_dict_manifest.json")
    with open(path) as f:
        return json.load(f)


def make_checkpoint(seed: int = 0, dtype=None):
    """Flat reference-layout checkpoints for the oracle: (emage_sd, emage_cfg, vq) with
    vq = {part: (state_dict, cfg)}.  Entries the generator leaves analytic are filled from
    their definition (positional table P.py:328-340, zero batch counters)."""
    import torch
    from .emage_oracle import pos_table

    dtype = dtype or torch.float32
    man = load_manifest()

    def build(tag):
        sd = {}
        for name, shape in man[tag]:
            arr = synth_tensor(name, shape, seed, tag)
            if arr is not None:
                sd[name] = torch.from_numpy(arr).to(dtype)
            elif name.endswith("position_embeddings.pe"):
                period = EMAGE_CFG["pose_length"]
                sd[name] = pos_table(shape[2], period).repeat(shape[1] // period, 1).unsqueeze(0).to(dtype)
            else:
                sd[name] = torch.zeros(shape, dtype=torch.long)
        return sd

    vq = {p: (build("vq_" + p), dict(VQ_CFGS[p])) for p in VQ_CFGS}
    return build("emage"), dict(EMAGE_CFG), vq


def make_lstm_checkpoint(kind: str, seed: int = 0, dtype=None):
    """Flat reference-layout checkpoint of CamnAudioModel (kind="camn") or DiscoAudioModel ("disco")."""
    import torch
    dtype = dtype or torch.float32
    sd = {}
    for name, shape in load_manifest()[kind]:
        arr = synth_tensor(name, shape, seed, kind)
        sd[name] = torch.from_numpy(arr).to(dtype) if arr is not None else torch.zeros(shape, dtype=torch.long)
    return sd, dict(LSTM_CFG)
