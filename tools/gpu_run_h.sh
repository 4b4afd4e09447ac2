#!/bin/bash
# call H (1 GPU): K2 timing on three shapes after the lean stage boundary, quick parity
mkdir -p gpurun_out
python tools/k2_time.py 2>&1 | tail -1 | tee gpurun_out/r2_k1k2_cfg1_$1.json
CAFE_BENCH_FAMILIES=25000 CAFE_BENCH_TAXA=50 CAFE_BENCH_MAXSIZE=400 CAFE_BENCH_MU=0.8 python tools/k2_time.py 2>&1 | tail -1 | tee gpurun_out/r2_k1k2_cfg2shape_$1.json
CAFE_BENCH_FAMILIES=200000 CAFE_BENCH_TAXA=50 CAFE_BENCH_MAXSIZE=400 CAFE_BENCH_MU=0.8 K2_STEPS=5 python tools/k2_time.py 2>&1 | tail -1 | tee gpurun_out/r2_k1k2_cfg2_$1.json
python -m pytest tests/test_gpu_parity.py tests/test_gpu_pvalue.py -x -q -m gpu 2>&1 | tail -4
