"""One-k-block, many-tap convs (the shapes the tap-GEMM's halo mode takes) against torch float64 convs: exit 0 when every
case is within fp16x3 accuracy.  PM_TC_HALO=0 (read once per process) runs the same cases through the per-tap path."""
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pantomatrix_b200 import ops  # noqa: E402

ops.set_plane_format("fp16")
CASES = [  # batch, rows, cin, cout, taps, pad
    (3, 300, 64, 64, 15, 7), (2, 1000, 64, 64, 15, 7), (1, 200, 32, 32, 15, 7), (2, 130, 64, 64, 3, 1),
    (4, 1258, 64, 64, 15, 7), (2, 257, 64, 48, 17, 8), (1, 128, 64, 64, 15, 0),
]
bad = 0
for i, (b, rows, cin, cout, taps, pad) in enumerate(CASES):
    g = torch.Generator().manual_seed(100 + i)
    x = torch.randn(b, rows, cin, generator=g).cuda()
    w = (torch.randn(taps, cout, cin, generator=g) / math.sqrt(cin * taps)).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    rows_out = rows + 2 * pad - taps + 1
    res = torch.randn(b, rows_out, cout, generator=g).cuda()
    a, pw = ops.split_bf16(x, 2), ops.PackedW(w, 2)
    out, pl = ops.tapgemm_tc(a, pw, bias, rows_out=rows_out, pad=pad, act=ops.ACT_LEAKY, slope=0.01, residual=res, out_nsplit=2)
    ref = F.conv1d(x.double().transpose(1, 2), w.double().permute(1, 2, 0), bias.double(), padding=pad).transpose(1, 2) + res.double()
    ref = F.leaky_relu(ref, 0.01)
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    perr = float((pl.t[..., :cout].double().sum(0) / ops.F16_ACT_SCALE - ref).abs().max() / ref.abs().max())
    ok = err < 1e-5 and perr < 1e-5
    bad += not ok
    print(f"case {i} {(b, rows, cin, cout, taps, pad)}: rel err f32 {err:.2e} planes {perr:.2e} {'ok' if ok else 'FAIL'}")
torch.cuda.synchronize()
sys.exit(1 if bad else 0)
