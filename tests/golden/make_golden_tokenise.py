"""Golden fixture for the training-side tokenisation, from the REAL reference:
EmageVQModel.map2index / map2latent and EmageVQVAEConv.forward (rec_pose, embedding_loss, perplexity) on synthetic
weights and seeded inputs.  Run in the build container only (imports /root/reference):

    python tests/golden/make_golden_tokenise.py   ->  tests/golden/case_tokenise.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import ROOT, build_reference, import_reference  # noqa: E402


def main():
    ref = import_reference()
    _, vqm, vq, _ = build_reference(ref, seed=0)
    sys.path.insert(0, ROOT)
    from oracle.emage_oracle import axis_angle_to_rot6d
    g = np.random.Generator(np.random.PCG64(2024))
    bs, t = 2, 37
    aa = torch.from_numpy(g.standard_normal((bs, t, 55, 3)).astype(np.float32) * 0.4)
    rot6d = axis_angle_to_rot6d(aa).reshape(bs, t, 330)
    expr = torch.from_numpy(g.standard_normal((bs, t, 100)).astype(np.float32))
    contact = torch.from_numpy((g.random((bs, t, 4)) > 0.5).astype(np.float32))
    trans = torch.from_numpy(g.standard_normal((bs, t, 3)).astype(np.float32) * 0.1)
    out = {"rot6d": rot6d.numpy(), "expression": expr.numpy(), "tar_contact": contact.numpy(), "tar_trans": trans.numpy()}
    with torch.no_grad():
        idx = vqm.map2index(rot6d, expr, tar_contact=contact, tar_trans=trans)
        lat = vqm.map2latent(rot6d, expr, tar_contact=contact, tar_trans=trans)
        idx_default = vqm.map2index(rot6d, expr)                       # tar_contact / tar_trans default to zeros
        parts = vqm.spilt_inputs(rot6d, expr, tar_contact=contact, tar_trans=trans)
        for p in ("face", "upper", "hands", "lower"):
            out["idx_" + p] = idx[p].numpy().astype(np.int16)
            out["idx_default_" + p] = idx_default[p].numpy().astype(np.int16)
            out["latent_" + p] = lat[p].numpy()
            out["input_" + p] = parts[p].numpy()
            fw = vq[p](parts[p])
            out["rec_pose_" + p] = fw["rec_pose"].numpy()
            out["poses_feat_" + p] = fw["poses_feat"].numpy()
            out["embedding_loss_" + p] = np.float32(fw["embedding_loss"])
            out["perplexity_" + p] = np.float32(fw["perplexity"])
    np.savez_compressed(os.path.join(HERE, "case_tokenise.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") and v.ndim else float(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
