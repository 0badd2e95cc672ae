#!/bin/bash
# ncu --set full captures of the four kernels the roofline claims rest on (one gpurun call, ~6 minutes of box time):
#   gpurun --timeout 1500 -- 'bash tools/ncu_captures.sh'
# Reports land in gpurun_out/r2_*.ncu-rep; tools/ncu_summary.py turns them into the tables under profiles/r2/.
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
# 1. trunk GEMM M = 2048, N = K = 768 in the default engine (two fp16 planes): fp32-only output, then fp32 + planes
timeout 400 $NCU -k regex:tapgemm_tc -s 3 -c 1 -o gpurun_out/r2_gemm python tools/bench_gemm.py "lin 2048x768x768" fp16 > gpurun_out/r2_ncu_gemm.log 2>&1
# 2. tcgen05 attention
timeout 400 $NCU -k regex:attention_tc -s 200 -c 2 -o gpurun_out/r2_attn python tools/profile_step.py fp16x3 > gpurun_out/r2_ncu_attn.log 2>&1
# 3. VQ lookup, 2^21 rows
timeout 300 $NCU -k regex:l2_argmin_tc -s 3 -c 1 -o gpurun_out/r2_vq python tools/bench_vq.py --engine tc --reps 1 > gpurun_out/r2_ncu_vq.log 2>&1
# 4. LSTM recurrence (CaMN layer)
timeout 400 $NCU -k regex:lstm_bidir -s 8 -c 1 -o gpurun_out/r2_lstm python tools/bench_lstm_models.py > gpurun_out/r2_ncu_lstm.log 2>&1
ls -la gpurun_out/*.ncu-rep
