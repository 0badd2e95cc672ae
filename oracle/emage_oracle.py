"""CPU oracle for the EMAGE audio->motion inference hot path.  TEST INFRASTRUCTURE ONLY.

A functional restatement (plain torch CPU ops on a flat ``{key: tensor}`` checkpoint
dict, no nn.Module graph) of what the reference computes on this path:

  M.py = /root/reference/models/emage_audio/modeling_emage_audio.py
  P.py = /root/reference/models/emage_audio/processing_emage_audio.py
  T.py = /root/reference/test_emage_audio.py

Only tests/, bench.py's cpu_baseline / ``--impl reference`` leg and
__graft_entry__.smoke() may import this file.  The product path (pantomatrix_b200/)
never does; it fails loudly when its CUDA library is missing.

Pinning: the reference holds no golden vectors.  This oracle is
pinned instead against outputs of the *unmodified reference modules* imported from
/root/reference in the build container; tests/golden/make_golden.py is the generating
script and tests/test_oracle_golden.py the check (runs without the reference tree).

All heavy arithmetic of the reference is torch.nn library code that is not vendored in
/root/reference (Conv1d, BatchNorm1d, Linear, LayerNorm, MultiheadAttention,
Transformer{En,De}coderLayer, Embedding; reference pins torch==2.0.0,
pre-requirements.txt:2).  Their published semantics are restated below with
torch.nn.functional primitives; every function cites the reference call site it follows.

dtype: everything follows the dtype of the checkpoint dict (float32 = the reference's
arithmetic, float64 = tie-break / error-analysis runs).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------
# Rotation conversions (P.py:6-104).  Same formula order as the reference so that float32
# results agree to rounding; masked assignments are written as torch.where.
# ----------------------------------------------------------------------------------------


def _sqrt_pos(x):
    """P.py:10-14: sqrt(x) where x > 0 else 0."""
    return torch.where(x > 0, torch.sqrt(torch.clamp(x, min=0)), torch.zeros_like(x))


def _sign_like(a, b):
    """P.py:6-8: flip a where sign bits of a and b differ (a<0) != (b<0)."""
    return torch.where((a < 0) != (b < 0), -a, a)


def rot6d_to_matrix(d6):
    """P.py:49-55 (Gram-Schmidt, rows b1,b2,b3)."""
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = a1 / torch.clamp(torch.linalg.vector_norm(a1, dim=-1, keepdim=True), min=1e-12)
    b2 = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
    b2 = b2 / torch.clamp(torch.linalg.vector_norm(b2, dim=-1, keepdim=True), min=1e-12)
    b3 = torch.linalg.cross(b1, b2, dim=-1)
    return torch.stack((b1, b2, b3), dim=-2)


def matrix_to_quat(m):
    """P.py:16-29."""
    m00, m11, m22 = m[..., 0, 0], m[..., 1, 1], m[..., 2, 2]
    w = 0.5 * _sqrt_pos(1 + m00 + m11 + m22)
    x = 0.5 * _sqrt_pos(1 + m00 - m11 - m22)
    y = 0.5 * _sqrt_pos(1 - m00 + m11 - m22)
    z = 0.5 * _sqrt_pos(1 - m00 - m11 + m22)
    x = _sign_like(x, m[..., 2, 1] - m[..., 1, 2])
    y = _sign_like(y, m[..., 0, 2] - m[..., 2, 0])
    z = _sign_like(z, m[..., 1, 0] - m[..., 0, 1])
    return torch.stack((w, x, y, z), -1)


def _sin_half_over_angle(half, ang):
    """P.py:35-43 / 66-74: sin(a/2)/a with the |a|<1e-6 Taylor branch 0.5 - a^2/48."""
    small = ang.abs() < 1e-6
    safe = torch.where(small, torch.ones_like(ang), ang)
    return torch.where(small, 0.5 - (ang * ang) / 48, torch.sin(half) / safe)


def quat_to_axis_angle(q):
    """P.py:31-44."""
    n = torch.linalg.vector_norm(q[..., 1:], dim=-1, keepdim=True)
    half = torch.atan2(n, q[..., :1])
    ang = 2 * half
    return q[..., 1:] / _sin_half_over_angle(half, ang)


def rot6d_to_axis_angle(d6):
    """P.py:57-58."""
    return quat_to_axis_angle(matrix_to_quat(rot6d_to_matrix(d6)))


def axis_angle_to_quat(aa):
    """P.py:63-78."""
    ang = torch.linalg.vector_norm(aa, dim=-1, keepdim=True)
    half = 0.5 * ang
    return torch.cat([torch.cos(half), aa * _sin_half_over_angle(half, ang)], dim=-1)


def quat_to_matrix(q):
    """P.py:80-98."""
    r, i, j, k = torch.unbind(q, -1)
    s2 = 2.0 / (q * q).sum(-1)
    o = torch.stack((
        1 - s2 * (j * j + k * k), s2 * (i * j - k * r), s2 * (i * k + j * r),
        s2 * (i * j + k * r), 1 - s2 * (i * i + k * k), s2 * (j * k - i * r),
        s2 * (i * k - j * r), s2 * (j * k + i * r), 1 - s2 * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def axis_angle_to_rot6d(aa):
    """P.py:100-104,60-61: first two matrix rows."""
    m = quat_to_matrix(axis_angle_to_quat(aa))
    return m[..., :2, :].reshape(*m.shape[:-2], 6)


def rot6d_condition(raw, bs, t):
    """Conditioning of the Gram-Schmidt step (P.py:49-55) per joint: kappa = 1 / min(|a1|, |b2 before its
    normalisation|).  A decoder-output perturbation eps moves the emitted rotation by ~ eps * kappa rad, so
    two fp32 evaluations of the same decoder cannot agree better than that.  raw: {part: decoder output}
    as returned in vq_decode()["_raw"].  Returns (bs, t, 55); joints without a decoder (eyes) get 1."""
    d6 = torch.zeros(bs, t, 55, 6, dtype=torch.float64)
    d6[..., 0] = 1.0
    d6[..., 4] = 1.0
    if "upper" in raw:
        d6[:, :, list(UPPER_JOINTS)] = raw["upper"].double().reshape(bs, t, 13, 6)
    if "hands" in raw:
        d6[:, :, list(HANDS_JOINTS)] = raw["hands"].double().reshape(bs, t, 30, 6)
    if "lower" in raw:
        d6[:, :, list(LOWER_JOINTS)] = raw["lower"].double()[:, :, :54].reshape(bs, t, 9, 6)
    if "face" in raw:
        d6[:, :, JAW_JOINT] = raw["face"].double()[:, :, :6]
    a1, a2 = d6[..., :3], d6[..., 3:]
    n1 = a1.norm(dim=-1)
    b1 = a1 / n1.clamp(min=1e-30).unsqueeze(-1)
    b2 = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
    return 1.0 / torch.minimum(n1, b2.norm(dim=-1)).clamp(min=1e-30)


# SMPL-X joint bookkeeping (M.py:75-90,181): which of the 55 joints each body part owns.
UPPER_JOINTS = (3, 6, 9, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21)
LOWER_JOINTS = (0, 1, 2, 4, 5, 7, 8, 10, 11)
HANDS_JOINTS = tuple(range(25, 55))
JAW_JOINT = 22


# ----------------------------------------------------------------------------------------
# Layer primitives
# ----------------------------------------------------------------------------------------


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def _conv(sd, p, x, stride=1, pad=0):
    return F.conv1d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=pad)


def _bn(sd, p, x, eps=1e-5):
    """nn.BatchNorm1d in eval mode (running statistics), P.py:270,274,279."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], training=False, eps=eps)


def _ln(sd, p, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def mlp(sd, p, x):
    """P.py:316-326: Linear, LeakyReLU(0.1), Linear."""
    return _lin(sd, p + ".fc2", F.leaky_relu(_lin(sd, p + ".fc1", x), 0.1))


# WavEncoder geometry (P.py:300-307): (stride, first-conv padding, has downsample)
WAV_BLOCKS = ((5, 1600, True), (6, 0, True), (1, 7, False), (6, 0, True), (1, 7, False), (3, 0, True))


def wav_encoder(sd, p, wav):
    """P.py:263-314.  wav (bs, n) -> (bs, T', out_dim)."""
    x = wav.unsqueeze(1)
    for i, (stride, pad, has_ds) in enumerate(WAV_BLOCKS):
        q = f"{p}.feat_extractor.{i}"
        y = F.leaky_relu(_bn(sd, q + ".bn1", _conv(sd, q + ".conv1", x, stride, pad)), 0.01)
        y = _bn(sd, q + ".bn2", _conv(sd, q + ".conv2", y, 1, 7))
        sc = _bn(sd, q + ".downsample.1", _conv(sd, q + ".downsample.0", x, stride, pad)) if has_ds else x
        x = F.leaky_relu(y + sc, 0.01)
    return x.transpose(1, 2)


def _resblock(sd, p, x):
    """P.py:178-187."""
    return _conv(sd, p + ".model.2", F.leaky_relu(_conv(sd, p + ".model.0", x, 1, 1), 0.2), 1, 1) + x


def vq_encoder(sd, p, x, n_layer):
    """VQEncoderV5/V6, P.py:189-235.  x (bs,T,C_in) -> (bs,T,vae_length)."""
    x = x.permute(0, 2, 1)
    for i in range(n_layer):
        x = F.leaky_relu(_conv(sd, f"{p}.main.{3 * i}", x, 1, 1), 0.2)
        x = _resblock(sd, f"{p}.main.{3 * i + 2}", x)
    return x.permute(0, 2, 1)


def vq_decoder(sd, p, z, n_layer):
    """VQDecoderV5, P.py:237-261 (input_size == channels[0], so no stem conv)."""
    x = z.permute(0, 2, 1)
    x = _resblock(sd, p + ".main.0", x)
    x = _resblock(sd, p + ".main.1", x)
    for i in range(n_layer):
        x = F.leaky_relu(_conv(sd, f"{p}.main.{2 + 2 * i}", x, 1, 1), 0.2)
    x = _conv(sd, f"{p}.main.{2 + 2 * n_layer}", x, 1, 1)
    return x.permute(0, 2, 1)


def l2_argmin(z, codebook):
    """M.py:60-65 / P.py:158-164: d = |z|^2 + |e|^2 - 2 z.e^T, argmin over codes (first
    minimum wins, torch.argmin semantics)."""
    flat = z.reshape(-1, codebook.shape[1])
    d = (flat ** 2).sum(1, keepdim=True) + (codebook ** 2).sum(1) - 2 * flat @ codebook.t()
    return torch.argmin(d, dim=1).reshape(z.shape[:-1])


def logits_to_index(logits):
    """M.py:398-401: max(log_softmax(x)) index == first-max argmax of x."""
    return torch.max(F.log_softmax(logits, dim=2), dim=2)[1]


def pos_table(d_model, period, dtype=torch.float32):
    """P.py:328-340 (sin on even, cos on odd columns; rows repeat with `period`)."""
    pos = torch.arange(0, period, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe = torch.zeros(period, d_model)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.to(dtype)


def _mha(sd, p, q_in, kv_in, nhead=4):
    """nn.MultiheadAttention forward, batch-first view.  q_in (bs,Tq,E), kv_in (bs,Tk,E).
    Packed in_proj (3E,E): rows [0,E)=Q, [E,2E)=K, [2E,3E)=V; scores scaled by 1/sqrt(E/h);
    no masks, dropout off (eval)."""
    E = q_in.shape[-1]
    w, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    q = F.linear(q_in, w[:E], b[:E])
    k = F.linear(kv_in, w[E:2 * E], b[E:2 * E])
    v = F.linear(kv_in, w[2 * E:], b[2 * E:])
    bs, tq, _ = q.shape
    tk = k.shape[1]
    hd = E // nhead
    q = q.reshape(bs, tq, nhead, hd).transpose(1, 2)
    k = k.reshape(bs, tk, nhead, hd).transpose(1, 2)
    v = v.reshape(bs, tk, nhead, hd).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(hd)), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(bs, tq, E)
    return _lin(sd, p + ".out_proj", o)


def _ffn(sd, p, x):
    return _lin(sd, p + ".linear2", F.relu(_lin(sd, p + ".linear1", x)))


def enc_layer(sd, p, x):
    """nn.TransformerEncoderLayer, post-norm, ReLU (M.py:238-239)."""
    x = _ln(sd, p + ".norm1", x + _mha(sd, p + ".self_attn", x, x))
    return _ln(sd, p + ".norm2", x + _ffn(sd, p, x))


def dec_layer(sd, p, x, mem):
    """nn.TransformerDecoderLayer, post-norm, ReLU (M.py:241-250,261)."""
    x = _ln(sd, p + ".norm1", x + _mha(sd, p + ".self_attn", x, x))
    x = _ln(sd, p + ".norm2", x + _mha(sd, p + ".multihead_attn", x, mem))
    return _ln(sd, p + ".norm3", x + _ffn(sd, p, x))


def dec_stack(sd, p, x, mem, n):
    for i in range(n):
        x = dec_layer(sd, f"{p}.layers.{i}", x, mem)
    return x


# ----------------------------------------------------------------------------------------
# EmageAudioModel.forward (M.py:265-341) - one window
# ----------------------------------------------------------------------------------------

PARTS = ("face", "upper", "hands", "lower")


def emage_forward(sd, audio, speaker_id, masked_motion, mask, use_audio=True):
    motion = torch.where(mask == 1, sd["mask_embedding"].expand_as(masked_motion), masked_motion)  # M.py:267-268
    hint = vq_encoder(sd, "motion_encoder", motion, 3)                                             # M.py:271
    hint_body = mlp(sd, "bodyhints_body", hint)
    hint_face = mlp(sd, "bodyhints_face", hint)
    a_face = wav_encoder(sd, "audio_encoder_face", audio)                                          # M.py:275-276
    a_body = wav_encoder(sd, "audio_encoder_body", audio)
    T = hint_face.shape[1]
    if a_face.shape[1] > T:                     # M.py:278-281 (second `if` re-truncates face; body
        a_face = a_face[:, :T]                  # is never truncated - it is only attention memory)
    if a_body.shape[1] > T:
        a_face = a_face[:, :T]
    bs, t, _ = a_face.shape
    spk_body = sd["speaker_embedding_body.weight"][speaker_id].repeat(1, t, 1)                     # M.py:285-286
    spk_face = sd["speaker_embedding_face.weight"][speaker_id].repeat(1, t, 1)
    pe = sd["position_embeddings.pe"][:, :t]

    mem_face = _lin(sd, "audio_face_motion_proj", torch.cat([a_face, hint_face], dim=2))           # M.py:288
    dec_face = dec_stack(sd, "face_motion_decoder", spk_face + pe, mem_face, 4)                    # M.py:291-292
    face_latent = _lin(sd, "face_out_proj", dec_face)
    cls_face = mlp(sd, "face_cls", face_latent)

    x = spk_body + (_lin(sd, "moton_proj", hint_body) + pe)                                        # M.py:297-299
    fea = enc_layer(sd, "motion_self_encoder.layers.0", x)                                         # M.py:300
    mem_body = _lin(sd, "audio_body_motion_proj", a_body)                                          # M.py:304
    fea = (fea + spk_body) + pe                                                                    # M.py:307-308
    cross = dec_stack(sd, "audio_motion_cross_attn", fea, mem_body, 8)                             # M.py:309
    if not use_audio:
        cross = cross * 0.0
    fea = fea + cross                                                                              # M.py:312

    lat = {p: mlp(sd, "motion2latent_" + p, fea) for p in PARTS[1:]}                               # M.py:315-317
    others = {"upper": ("hands", "lower"), "hands": ("upper", "lower"), "lower": ("upper", "hands")}
    out = {"rec_face": face_latent, "cls_face": cls_face}
    for p in PARTS[1:]:
        a, b = others[p]
        refine = dec_stack(sd, "body_motion_decoder_" + p, lat[p] + spk_body, lat[a] + lat[b], 1)  # M.py:320-322
        rec = _lin(sd, "motion_out_proj_" + p, lat[p] + refine)                                    # M.py:323-325
        out["rec_" + p] = rec
        out["cls_" + p] = mlp(sd, "motion_cls_" + p, rec)                                          # M.py:328-330
    return out


# ----------------------------------------------------------------------------------------
# EmageVQModel.decode / get_global_motion (M.py:126-205)
# ----------------------------------------------------------------------------------------


def _scatter_joints(part_aa, joints, bs, t):
    """recover_from_mask_ts, P.py:118-132."""
    full = torch.zeros(bs, t, 55, 3, dtype=part_aa.dtype)
    full[:, :, list(joints)] = part_aa.reshape(bs, t, len(joints), 3)
    return full.reshape(bs, t, 165)


def vq_part_decode(vq, part, index=None, latent=None):
    """EmageVQVAEConv.decode / decode_from_latent, M.py:56-70."""
    sd, cfg = vq[part]
    cb = sd["quantizer.embedding.weight"]
    if index is None:
        index = l2_argmin(latent, cb)
    return vq_decoder(sd, "decoder", cb[index], cfg["vae_layer"]), index


def split_inputs(rot6d, expression, tar_contact=None, tar_trans=None):
    """EmageVQModel.spilt_inputs (sic), M.py:97-108: per-part encoder inputs from 55 x rot6d + expression."""
    bs, t, j6 = rot6d.shape
    r = rot6d.reshape(bs, t, j6 // 6, 6)
    contact = torch.zeros(bs, t, 4, dtype=rot6d.dtype) if tar_contact is None else tar_contact
    trans = torch.zeros(bs, t, 3, dtype=rot6d.dtype) if tar_trans is None else tar_trans
    return dict(face=torch.cat([r[:, :, 22], expression], dim=2), upper=r[:, :, list(UPPER_JOINTS)].reshape(bs, t, 78),
                hands=r[:, :, 25:55].reshape(bs, t, 180),
                lower=torch.cat([r[:, :, list(LOWER_JOINTS)].reshape(bs, t, 54), trans, contact], dim=2))


def vqvae_forward(vq, part, inputs):
    """EmageVQVAEConv.forward, M.py:42-46, with Quantizer.forward P.py:144-156: encoder, nearest code, the
    straight-through value z + (z_q - z) the decoder actually sees, commitment loss and code perplexity."""
    sd, cfg = vq[part]
    cb = sd["quantizer.embedding.weight"]
    z = vq_encoder(sd, "encoder", inputs, cfg["vae_layer"])
    index = l2_argmin(z, cb)
    z_q = cb[index]
    loss = torch.mean((z_q - z) ** 2) + cfg["vae_quantizer_lambda"] * torch.mean((z_q - z) ** 2)
    z_st = z + (z_q - z)
    e_mean = F.one_hot(index.reshape(-1), cb.shape[0]).to(z.dtype).mean(0)
    perplexity = torch.exp(-torch.sum(e_mean * torch.log(e_mean + 1e-10)))
    return {"poses_feat": z_st, "embedding_loss": loss, "perplexity": perplexity,
            "rec_pose": vq_decoder(sd, "decoder", z_st, cfg["vae_layer"]), "_index": index}


def vq_tokenise(vq, rot6d, expression, tar_contact=None, tar_trans=None):
    """EmageVQModel.map2index / map2latent, M.py:110-124: (indices, latents) per part."""
    parts = split_inputs(rot6d, expression, tar_contact, tar_trans)
    idx, lat = {}, {}
    for p in PARTS:
        sd, cfg = vq[p]
        idx[p] = l2_argmin(vq_encoder(sd, "encoder", parts[p], cfg["vae_layer"]), sd["quantizer.embedding.weight"])
        lat[p] = sd["quantizer.embedding.weight"][idx[p]]
    return idx, lat


def vq_decode(vq, face_index=None, upper_index=None, hands_index=None, lower_index=None,
              face_latent=None, upper_latent=None, hands_latent=None, lower_latent=None,
              get_global_motion=False, ref_trans=None):
    """vq: {part: (state_dict, cfg)} for face/upper/hands/lower/global.  M.py:126-193."""
    for ten in (face_index, upper_index, hands_index, lower_index, face_latent, upper_latent, hands_latent, lower_latent):
        if ten is not None:
            bs, t = ten.shape[:2]
            dt = vq["face"][0]["quantizer.embedding.weight"].dtype
            break
    used, raw = {}, {}
    if face_index is not None or face_latent is not None:
        mix, used["face"] = vq_part_decode(vq, "face", face_index, face_latent)
        raw["face"] = mix
        jaw, expression = rot6d_to_axis_angle(mix[:, :, :6]), mix[:, :, 6:]
    else:
        jaw, expression = torch.zeros(bs, t, 3, dtype=dt), torch.zeros(bs, t, 100, dtype=dt)
    if upper_index is not None or upper_latent is not None:
        u6, used["upper"] = vq_part_decode(vq, "upper", upper_index, upper_latent)
        raw["upper"] = u6
        upper = rot6d_to_axis_angle(u6.reshape(bs, t, -1, 6)).reshape(bs, t, -1)
    else:
        upper = torch.zeros(bs, t, 39, dtype=dt)
    if hands_index is not None or hands_latent is not None:
        h6, used["hands"] = vq_part_decode(vq, "hands", hands_index, hands_latent)
        raw["hands"] = h6
        hands = rot6d_to_axis_angle(h6.reshape(bs, t, -1, 6)).reshape(bs, t, -1)
    else:
        hands = torch.zeros(bs, t, 90, dtype=dt)
    if lower_index is not None or lower_latent is not None:
        lower_mix, used["lower"] = vq_part_decode(vq, "lower", lower_index, lower_latent)
        raw["lower"] = lower_mix
        l6, transfoot = lower_mix[:, :, :-7], lower_mix[:, :, -7:]
        lower = rot6d_to_axis_angle(l6.reshape(bs, t, -1, 6)).reshape(bs, t, -1)
    else:                                                                               # M.py:174-178
        lower = torch.zeros(bs, t, 27, dtype=dt)
        transfoot = torch.zeros(bs, t, 7, dtype=dt)
        lower_mix = torch.cat([axis_angle_to_rot6d(lower.reshape(bs, t, -1, 3)).reshape(bs, t, -1), transfoot], -1)

    aa = (_scatter_joints(upper, UPPER_JOINTS, bs, t) + _scatter_joints(hands, HANDS_JOINTS, bs, t)
          + _scatter_joints(lower, LOWER_JOINTS, bs, t))                                # M.py:180-184
    aa[:, :, 3 * JAW_JOINT:3 * JAW_JOINT + 3] = jaw                                     # M.py:185
    rot6d = axis_angle_to_rot6d(aa.reshape(bs, t, 55, 3)).reshape(bs, t, 330)
    out = dict(expression=expression, all_motion4inference=torch.cat([rot6d, transfoot], 2),
               motion_axis_angle=aa, trans=None, _index=used, _raw=raw)
    if get_global_motion:
        out["trans"] = global_motion(vq, lower_mix, ref_trans)
    return out


def global_motion(vq, lower_mix, ref_trans):
    """M.py:195-205 + EmageVAEConv.forward M.py:27-32 + velocity2position P.py:107-115."""
    sd, cfg = vq["global"]
    rec = vq_decoder(sd, "decoder", vq_encoder(sd, "encoder", lower_mix, cfg["vae_layer"]), cfg["vae_layer"])
    vel = rec[:, :, 54:57]
    if ref_trans.dim() == 2:
        ref_trans = ref_trans.unsqueeze(0).repeat(vel.shape[0], 1, 1)

    def integrate(v, x0):               # sequential sum, same order as the reference loop
        pos = [x0.unsqueeze(1)]
        for i in range(1, v.shape[1]):
            pos.append(v[:, i - 1:i] * (1 / 30) + pos[-1])
        return torch.cat(pos, dim=1)

    x = integrate(vel[:, :, 0:1], ref_trans[:, 0, 0:1])
    z = integrate(vel[:, :, 2:3], ref_trans[:, 0, 2:3])
    return torch.cat([x, vel[:, :, 1:2], z], dim=-1)


# ----------------------------------------------------------------------------------------
# EmageAudioModel.inference (M.py:343-490) and the caller plumbing of T.py:16-47
# ----------------------------------------------------------------------------------------


def window_plan(total_len, window=64, pre=4):
    """M.py:365-368,380-382,428-430: list of (start, end) frame ranges and the number of frames
    each window contributes to the output (its length minus `pre`, except the tail)."""
    rounds = (total_len - pre) // (window - pre)
    remain = (total_len - pre) % (window - pre)
    plan = [(i * (window - pre), i * (window - pre) + window, window - pre) for i in range(rounds)]
    if remain > pre:
        s = rounds * (window - pre)
        plan.append((s, s + pre + remain, pre + remain))
    return plan


def _select(cfg, out, idx):
    """M.py:403-410: latent for a part iff l?>0 and c?==0, class index iff c?>0."""
    kw = {}
    for p, lk, ck in (("face", "lf", "cf"), ("upper", "lu", "cu"), ("hands", "lh", "ch"), ("lower", "ll", "cl")):
        kw[p + "_latent"] = out["rec_" + p] if cfg[lk] > 0 and cfg[ck] == 0 else None
        kw[p + "_index"] = idx[p] if cfg[ck] > 0 else None
    return kw


def emage_inference(sd, cfg, vq, audio, speaker_id, masked_motion=None, mask=None, trace=None):
    """M.py:343-490.  Returns the dict of 8 concatenated tensors.  `trace`, if a list, receives one
    dict per window (inputs, raw outputs, indices, decoded seed) for teacher-forced tests."""
    dt = sd["mask_embedding"].dtype
    bs = audio.shape[0]
    length = audio.shape[1] * 30 // 16000                                              # M.py:345
    ident = axis_angle_to_rot6d(torch.zeros(bs, length, 55, 3, dtype=dt)).reshape(bs, length, -1)
    motion = torch.cat([ident, torch.zeros(bs, length, 7, dtype=dt)], dim=-1)          # M.py:348-351
    if masked_motion is not None:
        motion[:, :masked_motion.shape[1]] = masked_motion
    full_mask = torch.ones_like(motion)
    if mask is not None:
        full_mask[:, :mask.shape[1]] = mask
    window, pre = cfg["pose_length"], cfg["seed_frames"]
    spf = 16000 // 30                                                                  # 533, M.py:393
    plan = window_plan(length, window, pre)
    acc = {k + p: [] for k in ("rec_", "cls_") for p in PARTS}
    last = motion[:, :pre]                                                             # M.py:379
    for wi, (s, e, keep) in enumerate(plan):
        w_mask = full_mask[:, s:e].clone()
        w_motion = motion[:, s:e].clone()
        w_motion[:, :pre] = torch.where(w_mask[:, :pre] == 0, motion[:, s:s + pre], last)   # M.py:386-390
        w_mask[:, :pre] = 0
        a_slice = audio[:, s * spf: s * spf + (e - s) * spf]                           # M.py:393-394
        out = emage_forward(sd, a_slice, speaker_id, w_motion, w_mask)
        idx = {p: logits_to_index(out["cls_" + p]) for p in PARTS}
        dec = vq_decode(vq, **_select(cfg, out, idx))
        last = dec["all_motion4inference"][:, -pre:]                                   # M.py:418
        is_tail = keep == e - s
        for k in acc:
            acc[k].append(out[k] if is_tail else out[k][:, :-pre])                      # M.py:419-426,463-470
        if trace is not None:
            trace.append(dict(audio=a_slice, motion=w_motion, mask=w_mask, out=out, idx=idx,
                              used=dec["_index"], seed=last))
    return {k: torch.cat(v, dim=1) for k, v in acc.items()}


def emage_generate(sd, cfg, vq, audio, speaker_id, masked_motion=None, mask=None, ref_trans=None, trace=None):
    """The timed span of the reference demo, T.py:32-47: inference(), re-derive indices from
    the concatenated logits, final full-length decode with get_global_motion=True."""
    lat = emage_inference(sd, cfg, vq, audio, speaker_id, masked_motion, mask, trace)
    idx = {p: logits_to_index(lat["cls_" + p]) for p in PARTS}
    if ref_trans is None:
        ref_trans = torch.zeros(1, 3, dtype=audio.dtype)                               # trans[:,0], T.py:30,47
    pred = vq_decode(vq, **_select(cfg, lat, idx), get_global_motion=True, ref_trans=ref_trans)
    return lat, pred
