#!/bin/bash
# call Q (1 GPU): ncu full capture of the K1 recurrence kernel at the configs[2] shape; small-table latency of K2 (59 / 1184 / 4736
# families at the configs[1] shape) with a CTA-0 trace of the 1184-family case (one 8-family block per CTA)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_bd_matrix_rec -s 3 -c 1 -o gpurun_out/r2_k1_rec_cfg2 -f python bench.py --steps 1 --warmup 3 --no-sub --no-cpu-baseline > gpurun_out/r2_ncu_k1_rec.log 2>&1
ls -la gpurun_out/r2_k1_rec_cfg2.ncu-rep
for F in 59 1184 4736; do CAFE_BENCH_FAMILIES=$F K2_STEPS=20 python tools/k2_time.py 2>&1 | tail -1; done
CAFE_BENCH_FAMILIES=1184 K2_STEPS=2 CAFE_GPU_TRACE=gpurun_out/r2_trace_small.txt python tools/k2_time.py 2>&1 | tail -1
