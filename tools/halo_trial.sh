#!/bin/bash
# One-call trial of the halo-mode tap-GEMM (PM_TC_HALO): parity of the mode on its own, then the whole GPU suite and a
# bench with the mode on.  gpurun --timeout 400 -- 'bash tools/halo_trial.sh'
set -u
mkdir -p gpurun_out
good=0
for m in 1 2; do
  PM_TC_HALO=$m timeout 120 python tools/check_halo.py > gpurun_out/x4_halo_mode$m.log 2>&1
  rc=$?
  echo "halo mode $m rc=$rc"; tail -8 gpurun_out/x4_halo_mode$m.log
  if [ $rc -eq 0 ] && [ $good -eq 0 ]; then good=$m; fi
done
echo "good=$good"
[ $good -eq 0 ] && exit 0
export PM_TC_HALO=$good
timeout 200 python -m pytest tests -m gpu -q -x > gpurun_out/x4_tests_halo.log 2>&1; echo "tests=$?"; tail -3 gpurun_out/x4_tests_halo.log
timeout 150 python bench.py --cpu-baseline 0 --extra 0 > gpurun_out/x4_bench_halo.json 2> gpurun_out/x4_bench_halo.err; echo "bench=$?"
python -c "
import json
d=json.load(open('gpurun_out/x4_bench_halo.json')); print('halo', d['ms_per_step'], d['value'], d['e2e']['value'])"
timeout 60 python tools/bench_gemm.py "conv k15 128x7460" fp16 > gpurun_out/x4_conv_halo.txt 2>&1; cat gpurun_out/x4_conv_halo.txt
