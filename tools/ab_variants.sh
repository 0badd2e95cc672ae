#!/bin/bash
# A/B of tap-GEMM build variants on the GPU box (built in-tree beforehand with
#   python -m pantomatrix_b200.build --variant NAME -DMACRO ...   # lands in pantomatrix_b200/csrc/_build/variants/).
# For every variant given: the tap-GEMM parity tests, the GEMM microbenchmark and one bench line, all through
# PM_EMAGE_LIB so the default library is untouched.  Usage (under gpurun):
#   bash tools/ab_variants.sh base epi_prefetch ...        # "base" = the default libpm_emage.so
set -u
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then unset PM_EMAGE_LIB; else export PM_EMAGE_LIB=$PWD/pantomatrix_b200/csrc/_build/variants/libpm_emage_$v.so; fi
  echo "=== variant $v (${PM_EMAGE_LIB:-default library})"
  timeout 200 python -m pytest tests/test_tapgemm_tc_gpu.py -x -q -m gpu 2>&1 | tail -2
  timeout 120 python tools/bench_gemm.py "lin 2048x768x" > gpurun_out/gemm_$v.txt 2>&1; grep -E " 3 +f" gpurun_out/gemm_$v.txt
  timeout 120 python bench.py --cpu-baseline 0 > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{v}.json").read().strip().splitlines()[-1])
    print(v, "frames/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print(v, "bench failed:", e)
PY
done
