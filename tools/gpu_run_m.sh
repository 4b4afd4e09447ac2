#!/bin/bash
# call M (1 GPU): K1 product tables - timing on three shapes, K1 parity tests, bench scores (bit identity with the previous K1)
mkdir -p gpurun_out
python tools/k2_time.py 2>&1 | tail -1 | tee gpurun_out/r2_k1k2_cfg1_$1.json
CAFE_BENCH_FAMILIES=25000 CAFE_BENCH_TAXA=50 CAFE_BENCH_MAXSIZE=400 CAFE_BENCH_MU=0.8 python tools/k2_time.py 2>&1 | tail -1 | tee gpurun_out/r2_k1k2_cfg2shape_$1.json
python -m pytest tests/test_gpu_parity.py tests/test_gpu_extra.py tests/test_gpu_lrt.py -x -q -m gpu -k "k1 or lrt_example or lrt_two or S_1001 or lambda_mu or two_lambda" 2>&1 | tail -3
python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2_bench_n1_$1.json 2> gpurun_out/r2_bench_n1_$1.err; python - <<PY
import json
d = json.loads(open('gpurun_out/r2_bench_n1_$1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], repr(d['config']['last_score']), d['config']['ms_breakdown_rank0'])
for k,v in d['configs'].items(): print(k, v.get('value'), v.get('k1_ms'), v.get('k2_ms'), repr(v.get('last_score')))
PY
