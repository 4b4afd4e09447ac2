#!/bin/bash
# call G (1 GPU): CTA-0 warp profile + per-CTA times of K2 at the configs[2] table (200 k x 50 taxa, W = 481) and at configs[1]
mkdir -p gpurun_out
CAFE_BENCH_FAMILIES=200000 CAFE_BENCH_TAXA=50 CAFE_BENCH_MAXSIZE=400 CAFE_BENCH_MU=0.8 K2_STEPS=2 CAFE_GPU_TRACE=gpurun_out/r2_trace_cfg2.txt python tools/k2_time.py 2>&1 | tail -1
K2_STEPS=2 CAFE_GPU_TRACE=gpurun_out/r2_trace_cfg1.txt python tools/k2_time.py 2>&1 | tail -1
grep -c . gpurun_out/r2_trace_cfg2.txt
