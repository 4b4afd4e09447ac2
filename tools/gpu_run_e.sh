#!/bin/bash
# call E (1 GPU): the windowed fused kernel (K4/K5) and the rebalanced last pass: parity, p-value tests, timings
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_pvalue.py tests/test_gpu_report.py tests/test_gpu_lrt.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_windowed_tests.log
cat gpurun_out/r2_windowed_tests.log
python tools/k2_time.py 2>&1 | tail -1 > gpurun_out/r2_k1k2_cfg1_b.json; cat gpurun_out/r2_k1k2_cfg1_b.json
CAFE_BENCH_FAMILIES=25000 CAFE_BENCH_TAXA=50 CAFE_BENCH_MAXSIZE=400 CAFE_BENCH_MU=0.8 python tools/k2_time.py 2>&1 | tail -1 > gpurun_out/r2_k1k2_cfg2shape_b.json; cat gpurun_out/r2_k1k2_cfg2shape_b.json
python bench.py --no-cpu-baseline > gpurun_out/r2_bench_full_n1_b.json 2> gpurun_out/r2_bench_full_n1_b.err; tail -c 2500 gpurun_out/r2_bench_full_n1_b.json; tail -3 gpurun_out/r2_bench_full_n1_b.err
