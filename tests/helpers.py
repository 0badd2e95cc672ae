"""Shared test helpers (product-side model construction with the synthetic checkpoints)."""
import torch

from synthetic_models import build_lstm_product, build_product  # noqa: F401  (construction lives at the repo root)


def geodesic_deg(aa_a: torch.Tensor, aa_b: torch.Tensor) -> torch.Tensor:
    """Angle (degrees) of the relative rotation between two axis-angle tensors (..., 3): the error measure
    that is meaningful across the axis-angle discontinuity at pi."""
    from oracle.emage_oracle import axis_angle_to_quat, quat_to_matrix
    ra = quat_to_matrix(axis_angle_to_quat(aa_a.double()))
    rb = quat_to_matrix(axis_angle_to_quat(aa_b.double()))
    tr = (ra.transpose(-1, -2) @ rb).diagonal(dim1=-2, dim2=-1).sum(-1)
    return torch.rad2deg(torch.acos(torch.clamp((tr - 1) / 2, -1, 1)))
