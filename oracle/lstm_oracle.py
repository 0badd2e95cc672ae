"""CPU oracle for the CaMN / DisCo audio->motion paths (BASELINE configs[2], [3]).  TEST INFRASTRUCTURE ONLY.

Functional restatement of
  C.py = /root/reference/models/camn_audio/modeling_camn_audio.py   (CamnAudioModel.forward 237-281)
  D.py = /root/reference/models/disco_audio/modeling_disco_audio.py (DiscoAudioModel.forward 220-267)
on flat checkpoint dicts.  torch.nn.LSTM (not vendored in the reference tree) is restated from its documented
equations: i,f,g,o gate order, h_t = o * tanh(c_t), c_t = f * c_{t-1} + i * g, layer l > 0 reads the
concatenated [forward | backward] outputs of layer l-1, dropout inactive in eval mode.
Pinned to the unmodified reference modules by tests/golden/make_golden_lstm.py -> tests/golden/case_camn.npz,
case_disco.npz (checked in tests/test_oracle_golden.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .emage_oracle import _bn, _conv, mlp, rot6d_to_axis_angle

# WavEncoder of CaMN/DisCo (C.py:135-146): (cin, cout, stride, first padding); a block has a downsample branch
# iff stride != 1 or cin != cout (C.py:113-118).
WAV_BLOCKS_LSTM = ((1, 32, 5, 1600), (32, 32, 6, 0), (32, 32, 1, 7), (32, 64, 6, 0), (64, 64, 1, 7), (64, 128, 6, 0))

# joint masks (C.py:19-29)
LOCAL_UPPER = tuple(j for j in range(55) if j in (3, 6, 9) or 12 <= j <= 21 or j >= 25)
LOCAL_FULL = tuple(range(1, 55))
MASKS = {"local_upper": LOCAL_UPPER, "local_full": LOCAL_FULL}


def wav_encoder(sd, p, wav):
    x = wav.unsqueeze(1)
    for i, (cin, cout, stride, pad) in enumerate(WAV_BLOCKS_LSTM):
        q = f"{p}.feat_extractor.{i}"
        y = F.leaky_relu(_bn(sd, q + ".bn1", _conv(sd, q + ".conv1", x, stride, pad)), 0.01)
        y = _bn(sd, q + ".bn2", _conv(sd, q + ".conv2", y, 1, 7))
        has_ds = stride != 1 or cin != cout
        sc = _bn(sd, q + ".downsample.1", _conv(sd, q + ".downsample.0", x, stride, pad)) if has_ds else x
        x = F.leaky_relu(y + sc, 0.01)
    return x.transpose(1, 2)


def bilstm(sd, p, x, n_layer, hidden):
    """nn.LSTM(batch_first=True, bidirectional=True), zero initial state.  x (bs,T,in) -> (bs,T,2*hidden)."""
    bs, T, _ = x.shape
    for layer in range(n_layer):
        outs = []
        for suffix in ("", "_reverse"):
            w_ih, w_hh = sd[f"{p}.weight_ih_l{layer}{suffix}"], sd[f"{p}.weight_hh_l{layer}{suffix}"]
            b = sd[f"{p}.bias_ih_l{layer}{suffix}"] + sd[f"{p}.bias_hh_l{layer}{suffix}"]
            xp = x @ w_ih.t() + b
            h = torch.zeros(bs, hidden, dtype=x.dtype)
            c = torch.zeros(bs, hidden, dtype=x.dtype)
            seq = [None] * T
            order = range(T) if suffix == "" else range(T - 1, -1, -1)
            for t in order:
                g = xp[:, t] + h @ w_hh.t()
                i, f, gg, o = g.split(hidden, dim=1)
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
                h = torch.sigmoid(o) * torch.tanh(c)
                seq[t] = h
            outs.append(torch.stack(seq, dim=1))
        x = torch.cat(outs, dim=2)
    return x


def _seed_features(cfg, bs, t, seed_frames, seed_motion, dtype):
    """C.py:249-263 / D.py:232-246."""
    dims = cfg["pose_dims"] + 1
    if seed_motion is None:
        seed = torch.zeros(bs, t, dims, dtype=dtype)
        seed[:, :seed_frames, -1] = 1
        return seed
    t_m = seed_motion.shape[1]
    pad = torch.zeros(bs, t_m, dims, dtype=dtype)
    pad[:, :seed_frames, :-1] = seed_motion[:, :seed_frames]
    pad[:, :seed_frames, -1] = 1
    if t_m > t:
        pad = pad[:, :t]
    elif t_m < t:
        pad = torch.cat((pad, pad[:, -(t - t_m):]), 1)
    return pad


def _to_axis_angle(cfg, rot6d, bs, t):
    """rot6d (bs,t,J,6) -> (bs,t,165) with zeros at unselected joints (C.py:274-277)."""
    joints = MASKS[cfg["joint_mask"]]
    aa = rot6d_to_axis_angle(rot6d.reshape(-1, len(joints), 6)).reshape(bs, t, len(joints), 3)
    full = torch.zeros(bs, t, 55, 3, dtype=aa.dtype)
    full[:, :, list(joints)] = aa
    return full.reshape(bs, t, 165)


def camn_forward(sd, cfg, audio, speaker_id, seed_frames=4, seed_motion=None):
    a = wav_encoder(sd, "audio_encoder", audio)
    bs, t, _ = a.shape
    spk = sd["speaker_embedding.weight"][speaker_id].repeat(1, t, 1)
    seed = _seed_features(cfg, bs, t, seed_frames, seed_motion, a.dtype)
    in_fea = torch.cat((a, spk, seed), dim=2)
    H = cfg["hidden_size"]
    y = bilstm(sd, "body_motion_decoder", in_fea, cfg["n_layer"], H)
    body = mlp(sd, "body_out", y[:, :, :H] + y[:, :, H:])
    y = bilstm(sd, "hands_motion_decoder", torch.cat((in_fea, body), dim=2), cfg["n_layer"], H)
    hands = mlp(sd, "hands_out", y[:, :, :H] + y[:, :, H:])
    motion = torch.cat([body.reshape(bs, t, -1, 6), hands.reshape(bs, t, -1, 6)], dim=2)      # recombine, C.py:227-234
    return {"motion": motion, "motion_axis_angle": _to_axis_angle(cfg, motion, bs, t)}


def disco_forward(sd, cfg, audio, speaker_id, seed_frames=4, seed_motion=None):
    a = wav_encoder(sd, "audio_encoder", audio)
    bs, t, _ = a.shape
    spk = sd["speaker_embedding.weight"][speaker_id].repeat(1, t, 1)
    seed = _seed_features(cfg, bs, t, seed_frames, seed_motion, a.dtype)
    c1, c2, r = mlp(sd, "audio_encoder_c1", a), mlp(sd, "audio_encoder_c2", a), mlp(sd, "audio_encoder_r", a)
    w = torch.softmax(mlp(sd, "selector", a), dim=2)
    fea_c = w[:, :, 0:1] * c1 + w[:, :, 1:2] * c2
    in_fea = torch.cat((fea_c, r, spk, seed), dim=2)
    H = cfg["hidden_size"]
    y = bilstm(sd, "body_motion_decoder", in_fea, cfg["n_layer"], H)
    motion = mlp(sd, "body_out", y[:, :, :H] + y[:, :, H:])
    return {"motion": motion, "motion_axis_angle": _to_axis_angle(cfg, motion.reshape(bs, t, -1, 6), bs, t),
            "audio_fea_c": fea_c, "audio_fea_r": r}
