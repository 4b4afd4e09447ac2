#!/bin/bash
# call D (1 GPU): extra tests, the Ozaki throughput proxy, K1/K2 timings on two shapes, the full default bench line
mkdir -p gpurun_out
python -m pytest tests/test_gpu_extra.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_extra_tests.log
cat gpurun_out/r2_extra_tests.log
python tools/ozaki_gemm_proxy.py > gpurun_out/r2_ozaki_gemm_proxy.json 2> gpurun_out/r2_ozaki_gemm_proxy.err; cat gpurun_out/r2_ozaki_gemm_proxy.json; tail -3 gpurun_out/r2_ozaki_gemm_proxy.err
python tools/k2_time.py 2>&1 | tail -1 > gpurun_out/r2_k1k2_cfg1.json; cat gpurun_out/r2_k1k2_cfg1.json
CAFE_BENCH_FAMILIES=25000 CAFE_BENCH_TAXA=50 CAFE_BENCH_MAXSIZE=400 CAFE_BENCH_MU=0.8 python tools/k2_time.py 2>&1 | tail -1 > gpurun_out/r2_k1k2_cfg2shape.json; cat gpurun_out/r2_k1k2_cfg2shape.json
python bench.py > gpurun_out/r2_bench_full_n1.json 2> gpurun_out/r2_bench_full_n1.err; tail -c 4000 gpurun_out/r2_bench_full_n1.json; tail -3 gpurun_out/r2_bench_full_n1.err
