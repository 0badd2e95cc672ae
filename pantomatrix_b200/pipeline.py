"""Caller plumbing of the hot path: the timed span of the reference demo
(/root/reference/test_emage_audio.py:16-47, twin train_emage_audio.py:33-102) without audio file
decoding and npz writing: inference() -> indices from the concatenated logits -> full-length
decode(get_global_motion=True)."""
from __future__ import annotations

import torch

from . import ops
from .emage_audio.engine import PARTS, select_inputs


@torch.no_grad()
def generate(model, motion_vq, audio, speaker_id=None, masked_motion=None, mask=None, ref_trans=None):
    """audio (bs, n) float32 16 kHz.  Returns (latent_dict, pred_dict) like T.py:32 and T.py:44-47."""
    dev = next(model.parameters()).device
    bs = audio.shape[0]
    if speaker_id is None:
        speaker_id = torch.zeros(bs, 1, dtype=torch.long, device=dev)                  # T.py:19
    lat = model.inference(audio, speaker_id, motion_vq, masked_motion=masked_motion, mask=mask)
    cfg = model.cfg.to_dict()
    idx = {p: ops.row_argmax(lat["cls_" + p]) for p in PARTS}                          # T.py:39-42
    index, latent = select_inputs(cfg, lat, idx)
    if ref_trans is None:
        ref_trans = torch.zeros(1, 3, device=dev)                                       # trans[:,0], T.py:30,47
    pred = motion_vq.decode(
        face_latent=latent["face"], upper_latent=latent["upper"], lower_latent=latent["lower"],
        hands_latent=latent["hands"], face_index=index["face"], upper_index=index["upper"],
        lower_index=index["lower"], hands_index=index["hands"], get_global_motion=True, ref_trans=ref_trans)
    return lat, pred
