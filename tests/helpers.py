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


# Opt-in GPU tests of code paths that have not been measured on hardware yet (fp16 operand planes, 96-column tiles):
# PM_TEST_EXPERIMENTAL=1 python -m pytest tests -m gpu
import os as _os  # noqa: E402

import pytest as _pytest  # noqa: E402

EXPERIMENTAL = _pytest.mark.skipif(_os.environ.get("PM_TEST_EXPERIMENTAL") != "1",
                                   reason="experimental path: set PM_TEST_EXPERIMENTAL=1")
