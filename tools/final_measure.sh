set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/x2_bench_full.json 2> gpurun_out/x2_bench_full.err; echo bench=$?
python bench.py --impl reference > gpurun_out/x2_bench_ref.json 2> gpurun_out/x2_bench_ref.err; echo ref=$?
python tools/bench_gemm.py "conv k15" fp16 > gpurun_out/x2_conv_occ2.txt 2>&1
PM_TC_OCC2=0 python tools/bench_gemm.py "conv k15" fp16 > gpurun_out/x2_conv_occ1.txt 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x2_launches.csv python tools/profile_step.py fp16x3 1 > gpurun_out/x2_launches.log 2>&1; echo ll=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tapgemm_tc -s 3 -c 1 -o gpurun_out/x2_conv64 python tools/bench_gemm.py "conv k15 128x7460" fp16 > gpurun_out/x2_ncu_conv.log 2>&1; echo c1=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wav_stem -s 4 -c 1 -o gpurun_out/x2_stem python tools/profile_step.py fp16x3 1 > gpurun_out/x2_ncu_stem.log 2>&1; echo c2=$?
cat gpurun_out/x2_conv_occ2.txt gpurun_out/x2_conv_occ1.txt
python - <<PY
import json
d=json.load(open("gpurun_out/x2_bench_full.json")); print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["roofline"]["frac"], d["roofline_vq"]["frac"])
PY
