"""One warm-up + one EMAGE step (32 clips x 10 s) with eager launches, for ncu:

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python tools/profile_step.py bf16x3
  ncu --set full --clock-control none --import-source on -k regex:tapgemm_tc -s 300 -c 3 -o gpurun_out/prof \
      python tools/profile_step.py bf16x3
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from synthetic_models import build_product  # noqa: E402
from oracle.weights import synth_audio  # noqa: E402
from pantomatrix_b200.emage_audio import engine  # noqa: E402
from pantomatrix_b200.pipeline import generate  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
bs = int(sys.argv[3]) if len(sys.argv) > 3 else 32
engine.set_precision(precision)
model, vqm = build_product(0)
audio = torch.from_numpy(synth_audio(bs, 160000, 1234)).cuda()
generate(model, vqm, audio)            # warm-up: lazy weight packing, function attributes
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("step")
for _ in range(steps):
    generate(model, vqm, audio)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("done")
