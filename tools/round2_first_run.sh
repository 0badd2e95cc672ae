#!/bin/bash
# First GPU call of round 2: everything that was written in round 1 without GPU minutes left.
# Build the variants on the build host first (they travel with the snapshot):
#   python -m pantomatrix_b200.build --variant epi_prefetch -DPM_TC_EPI_PREFETCH
#   python -m pantomatrix_b200.build --variant tma_store -DPM_TC_TMA_STORE
# then:  gpurun --timeout 1500 -- 'bash tools/round2_first_run.sh'        (about 12 minutes of box time)
set -u
mkdir -p gpurun_out
echo "=== 1. experimental parity tests (fp16 planes, fp16x3 end to end, 96-column tiles, mma attention)"
PM_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2_tests.txt
bench() {   # name, env assignments..., then bench args after --
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 150 python bench.py --cpu-baseline 0 "$@" > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2_bench_{n}.json").read().strip().splitlines()[-1])
    print(f"{n:28s} {d['value']:10.0f} frames/s  {d['ms_per_step']:6.2f} ms/step  e2e {d['e2e']['value']:10.0f}")
except Exception as e:
    print(n, "FAILED", e)
PY
}
echo "=== 2. bench lines"
bench default X=1 --
bench fp16x3 X=1 -- --precision fp16x3
bench bn96 PM_TC_BN=96 --
bench attn_mma PM_ATTN_MMA=1 --
bench fp16x3_bn96_attn PM_TC_BN=96 PM_ATTN_MMA=1 -- --precision fp16x3
for v in epi_prefetch tma_store; do
  lib=$PWD/pantomatrix_b200/csrc/_build/variants/libpm_emage_$v.so
  if [ -f "$lib" ]; then
    echo "=== 3. build variant $v"
    PM_EMAGE_LIB=$lib timeout 300 python -m pytest tests/test_tapgemm_tc_gpu.py tests/test_emage_gpu.py -x -q -m gpu 2>&1 | tail -2
    bench $v PM_EMAGE_LIB=$lib --
    bench ${v}_fp16x3 PM_EMAGE_LIB=$lib -- --precision fp16x3
  fi
done
echo "=== 4. free-running code agreement per precision mode (out of 38 400)"
PM_TEST_EXPERIMENTAL=1 timeout 400 python tests/diag_modes.py 32 > gpurun_out/r2_diag_modes.json 2>gpurun_out/r2_diag_modes.err; tail -c 1500 gpurun_out/r2_diag_modes.json
