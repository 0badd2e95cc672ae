"""Pin the CPU oracle (oracle/emage_oracle.py) to the reference's own outputs.

The golden files were produced by the UNMODIFIED reference modules
(tests/golden/make_golden.py, run where /root/reference is visible).  These tests need
neither the reference tree nor a GPU.
"""
import os

import numpy as np
import pytest
import torch

from oracle import emage_oracle as O
from oracle.weights import make_checkpoint, synth_audio

CASES = ["tail11", "clip10s", "drop_tail", "short40", "seeded"]
PARTS = ("face", "upper", "hands", "lower")


@pytest.fixture(scope="module")
def ckpt():
    torch.manual_seed(0)
    return make_checkpoint(seed=0)


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference(case, ckpt, golden_dir):
    sd, cfg, vq = ckpt
    g = np.load(os.path.join(golden_dir, f"case_{case}.npz"))
    bs, n = int(g["bs"]), int(g["n_samples"])
    audio = torch.from_numpy(synth_audio(bs, n, int(g["audio_seed"])))
    mm = torch.from_numpy(g["masked_motion"]) if "masked_motion" in g else None
    mk = torch.from_numpy(g["mask"]) if "mask" in g else None
    with torch.no_grad():
        lat, pred = O.emage_generate(sd, cfg, vq, audio, torch.zeros(bs, 1, dtype=torch.long), mm, mk)
    for p in PARTS:
        # integer contract: emitted code indices are bit-exact
        assert np.array_equal(lat["cls_" + p].argmax(-1).numpy(), g["idx_cls_" + p]), p
        # float contract: same arithmetic, different op grouping -> fp32 rounding only
        np.testing.assert_allclose(lat["rec_" + p].numpy()[:, ::7], g["rec_" + p], atol=2e-4, rtol=0)
        np.testing.assert_allclose(lat["cls_" + p].numpy()[:, ::7], g["cls_" + p], atol=5e-4, rtol=0)
    assert np.array_equal(pred["_index"]["face"].numpy(), g["idx_l2_face"])
    for k in ("expression", "motion_axis_angle", "trans", "all_motion4inference"):
        assert pred[k].shape == g[k].shape, k
        np.testing.assert_allclose(pred[k].numpy(), g[k], atol=1e-4, rtol=0, err_msg=k)


def test_window_plan_edge_cases():
    # M.py:365-368: L=300 -> 4 full windows + tail of 60; L=124 -> tail dropped; L=40 -> one tail window
    assert O.window_plan(300) == [(0, 64, 60), (60, 124, 60), (120, 184, 60), (180, 244, 60), (240, 300, 60)]
    assert O.window_plan(124) == [(0, 64, 60), (60, 124, 60)]
    assert O.window_plan(40) == [(0, 40, 40)]
    assert O.window_plan(68) == [(0, 64, 60)]          # remain == 4 is not > pre -> dropped
    assert O.window_plan(4) == []


def test_rotation_round_trip():
    g = torch.Generator().manual_seed(5)
    aa = torch.randn(1000, 3, generator=g) * 0.8
    aa = aa * torch.clamp(2.0 / aa.norm(dim=-1, keepdim=True), max=1.0)      # stay away from the pi discontinuity
    back = O.rot6d_to_axis_angle(O.axis_angle_to_rot6d(aa))
    assert (back - aa).abs().max() < 1e-3   # fp32 sqrt(1+-trace) route loses components ~2e-4
    zero = O.rot6d_to_axis_angle(O.axis_angle_to_rot6d(torch.zeros(4, 3)))
    assert zero.abs().max() == 0


@pytest.mark.parametrize("kind", ["camn", "disco"])
def test_lstm_oracle_matches_reference(kind, golden_dir):
    """CaMN / DisCo (BASELINE configs[2], [3]): oracle/lstm_oracle.py vs outputs of the unmodified reference modules."""
    from oracle import lstm_oracle as L
    from oracle.weights import make_lstm_checkpoint
    sd, cfg = make_lstm_checkpoint(kind, seed=0)
    g = np.load(os.path.join(golden_dir, f"case_{kind}.npz"))
    bs, n = int(g["bs"]), int(g["n_samples"])
    audio = torch.from_numpy(synth_audio(bs, n, int(g["audio_seed"])))
    spk = torch.zeros(bs, 1, dtype=torch.long)
    fwd = L.camn_forward if kind == "camn" else L.disco_forward
    with torch.no_grad():
        a = fwd(sd, cfg, audio, spk, 4, None)
        b = fwd(sd, cfg, audio, spk, 4, torch.from_numpy(g["seed_motion"]))
    t = g["motion"].shape[1]
    np.testing.assert_allclose(a["motion"].reshape(bs, t, -1).numpy(), g["motion"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(b["motion"].reshape(bs, t, -1).numpy(), g["seeded_motion"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(a["motion_axis_angle"].numpy(), g["motion_axis_angle"], atol=1e-3, rtol=0)
    if kind == "disco":
        np.testing.assert_allclose(a["audio_fea_c"].numpy(), g["audio_fea_c"], atol=1e-5, rtol=0)
        np.testing.assert_allclose(a["audio_fea_r"].numpy(), g["audio_fea_r"], atol=1e-5, rtol=0)


def test_oracle_tokenisation_matches_reference(golden_dir):
    """Training-side tokenisation: EmageVQModel.map2index / map2latent and EmageVQVAEConv.forward
    of the real reference (tests/golden/make_golden_tokenise.py) vs the oracle restatement."""
    from oracle import emage_oracle as O
    from oracle.weights import make_checkpoint
    g = np.load(os.path.join(golden_dir, "case_tokenise.npz"))
    _, _, vq = make_checkpoint(seed=0)
    rot6d, expr = torch.from_numpy(g["rot6d"]), torch.from_numpy(g["expression"])
    contact, trans = torch.from_numpy(g["tar_contact"]), torch.from_numpy(g["tar_trans"])
    with torch.no_grad():
        idx, lat = O.vq_tokenise(vq, rot6d, expr, contact, trans)
        idx0, _ = O.vq_tokenise(vq, rot6d, expr)
        parts = O.split_inputs(rot6d, expr, contact, trans)
        for p in ("face", "upper", "hands", "lower"):
            assert np.array_equal(parts[p].numpy(), g["input_" + p]), p
            assert np.array_equal(idx[p].numpy(), g["idx_" + p]) and np.array_equal(idx0[p].numpy(), g["idx_default_" + p]), p
            assert np.array_equal(lat[p].numpy(), g["latent_" + p]), p
            fw = O.vqvae_forward(vq, p, parts[p])
            np.testing.assert_allclose(fw["rec_pose"].numpy(), g["rec_pose_" + p], atol=2e-5, rtol=0)
            np.testing.assert_allclose(fw["poses_feat"].numpy(), g["poses_feat_" + p], atol=1e-6, rtol=0)
            np.testing.assert_allclose(float(fw["embedding_loss"]), float(g["embedding_loss_" + p]), rtol=1e-5)
            np.testing.assert_allclose(float(fw["perplexity"]), float(g["perplexity_" + p]), rtol=1e-5)



@pytest.mark.parametrize("tag,flag", [("audio", True), ("noaudio", False)])
def test_oracle_forward_matches_reference(tag, flag, ckpt, golden_dir):
    """EmageAudioModel.forward on one window with user masked_motion / mask, with and without the audio cross-attention
    (use_audio=False, M.py:310-311), against the unmodified reference (tests/golden/make_golden_forward.py)."""
    import sys
    sys.path.insert(0, golden_dir)
    from make_golden_forward import BS, inputs
    sd, cfg, vq = ckpt
    g = np.load(os.path.join(golden_dir, "case_forward.npz"))
    audio, motion, mask = inputs()
    with torch.no_grad():
        got = O.emage_forward(sd, audio, torch.zeros(BS, 1, dtype=torch.long), motion, mask, use_audio=flag)
    for k, v in got.items():
        np.testing.assert_allclose(v.numpy()[:, ::3], g[f"{tag}_{k}"], atol=2e-4 if k.startswith("rec_") else 5e-4, rtol=0, err_msg=k)
        if k.startswith("cls_"):
            assert np.array_equal(v.argmax(-1).numpy(), g[f"{tag}_idx_{k}"]), k
