#!/bin/bash
# call K (1 GPU): branch cutting timing; ncu launch list of the configs[4] pass; full captures of the windowed K2 instantiation and of K1
mkdir -p gpurun_out
timeout 600 python tools/time_branch_cutting.py 2>&1 | tail -1 | tee gpurun_out/r2_branch_cutting_timing.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_cd_pvalue.csv python tools/time_cd_pvalue.py 50 400 1000 25000 > gpurun_out/r2_cd_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_prune_fused2 -s 1 -c 1 -o gpurun_out/r2_k4_windowed -f python tools/time_cd_pvalue.py 50 400 1000 25000 > gpurun_out/r2_ncu_k4.log 2>&1
CAFE_BENCH_FAMILIES=25000 CAFE_BENCH_TAXA=50 CAFE_BENCH_MAXSIZE=400 CAFE_BENCH_MU=0.8 K2_STEPS=2 ncu --set full --clock-control none --import-source on -k regex:k_bd_matrix -s 3 -c 1 -o gpurun_out/r2_k1_cfg2 -f python tools/k2_time.py > gpurun_out/r2_ncu_k1.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/r2_cd_under_ncu.log
