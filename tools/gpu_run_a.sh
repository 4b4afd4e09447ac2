#!/bin/bash
# call A (1 GPU): full GPU test suite, the default bench line, a CTA-0 trace of K2
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_gputests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
CAFE_GPU_TRACE=gpurun_out/r2_trace_v3.txt K2_STEPS=1 python tools/k2_time.py > gpurun_out/r2_trace_v3.json 2>&1
cat gpurun_out/r2_gputests.log; tail -c 3000 gpurun_out/r2_bench_n1.json; tail -5 gpurun_out/r2_bench_n1.err
