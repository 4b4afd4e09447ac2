#!/bin/bash
# call P (1 GPU): per-group C halves - K2 timing against the one-time lag, then parity
mkdir -p gpurun_out
for lag in 0 1500 3000 6000 12000; do
  echo "lag $lag"
  CAFE_GPU_LAG=$lag python tools/k2_time.py 2>&1 | tail -1
  CAFE_GPU_LAG=$lag CAFE_BENCH_FAMILIES=25000 CAFE_BENCH_TAXA=50 CAFE_BENCH_MAXSIZE=400 CAFE_BENCH_MU=0.8 K2_STEPS=10 python tools/k2_time.py 2>&1 | tail -1
done
CAFE_BENCH_FAMILIES=200000 CAFE_BENCH_TAXA=50 CAFE_BENCH_MAXSIZE=400 CAFE_BENCH_MU=0.8 K2_STEPS=5 python tools/k2_time.py 2>&1 | tail -1
K2_STRESS_FRESH=1 python tools/k2_stress.py 12 | tail -1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_pvalue.py tests/test_gpu_lrt.py -x -q -m gpu 2>&1 | tail -4
