"""Shared test helpers (product-side model construction with the synthetic checkpoints)."""
import torch

from oracle.weights import EMAGE_CFG, VQ_CFGS, load_synthetic


def build_product(seed=0, device="cuda"):
    from pantomatrix_b200.emage_audio import (EmageAudioConfig, EmageAudioModel, EmageVAEConv, EmageVAEConvConfig,
                                              EmageVQModel, EmageVQVAEConv, EmageVQVAEConvConfig)
    model = load_synthetic(EmageAudioModel(EmageAudioConfig(**EMAGE_CFG)), seed, "emage").to(device).eval()
    vq = {p: load_synthetic(EmageVQVAEConv(EmageVQVAEConvConfig(**VQ_CFGS[p])), seed, "vq_" + p).to(device).eval()
          for p in ("face", "upper", "hands", "lower")}
    glob = load_synthetic(EmageVAEConv(EmageVAEConvConfig(**VQ_CFGS["global"])), seed, "vq_global").to(device).eval()
    vqm = EmageVQModel(face_model=vq["face"], upper_model=vq["upper"], lower_model=vq["lower"],
                       hands_model=vq["hands"], global_model=glob).to(device).eval()
    return model, vqm


def geodesic_deg(aa_a: torch.Tensor, aa_b: torch.Tensor) -> torch.Tensor:
    """Angle (degrees) of the relative rotation between two axis-angle tensors (..., 3): the error measure
    that is meaningful across the axis-angle discontinuity at pi."""
    from oracle.emage_oracle import axis_angle_to_quat, quat_to_matrix
    ra = quat_to_matrix(axis_angle_to_quat(aa_a.double()))
    rb = quat_to_matrix(axis_angle_to_quat(aa_b.double()))
    tr = (ra.transpose(-1, -2) @ rb).diagonal(dim1=-2, dim2=-1).sum(-1)
    return torch.rad2deg(torch.acos(torch.clamp((tr - 1) / 2, -1, 1)))


def build_lstm_product(kind, seed=0, device="cuda"):
    """CamnAudioModel ("camn") or DiscoAudioModel ("disco") with the synthetic checkpoint."""
    from oracle.weights import LSTM_CFG
    from pantomatrix_b200.lstm_audio import CamnAudioConfig, CamnAudioModel, DiscoAudioConfig, DiscoAudioModel
    cls, ccls = (CamnAudioModel, CamnAudioConfig) if kind == "camn" else (DiscoAudioModel, DiscoAudioConfig)
    return load_synthetic(cls(ccls(**LSTM_CFG)), seed, kind).to(device).eval()


# Opt-in GPU tests of code paths that have not been measured on hardware yet (fp16 operand planes, 96-column tiles):
# PM_TEST_EXPERIMENTAL=1 python -m pytest tests -m gpu
import os as _os  # noqa: E402

import pytest as _pytest  # noqa: E402

EXPERIMENTAL = _pytest.mark.skipif(_os.environ.get("PM_TEST_EXPERIMENTAL") != "1",
                                   reason="experimental path: set PM_TEST_EXPERIMENTAL=1")
