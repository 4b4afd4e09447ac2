#!/bin/bash
# call C (1 GPU): real-fixture tests, compiled bridge, ncu captures of K2 (launch list + full set for the DRAM traffic)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_integration.py tests/test_gpu_bridge.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_integration_tests.log
cat gpurun_out/r2_integration_tests.log
# launch list of the default bench command (headline only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-sub --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
# full capture of one K2 launch at configs[2] and one at configs[1]
ncu --set full --clock-control none --import-source on -k regex:k_prune_fused2 -s 3 -c 1 -o gpurun_out/r2_k2_cfg2 python bench.py --steps 1 --warmup 3 --no-sub --no-cpu-baseline > gpurun_out/r2_ncu_cfg2.log 2>&1
CAFE_BENCH_CONFIG="configs[1]" ncu --set full --clock-control none --import-source on -k regex:k_prune_fused2 -s 3 -c 1 -o gpurun_out/r2_k2_cfg1 python bench.py --steps 1 --warmup 3 --no-sub --no-cpu-baseline > gpurun_out/r2_ncu_cfg1.log 2>&1
ls -la gpurun_out/*.ncu-rep
