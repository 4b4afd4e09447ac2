#!/bin/bash
# call F (1 GPU): K2 timing on the two shapes after the epilogue specialisation, p-value pass timing, quick parity
mkdir -p gpurun_out
python tools/k2_time.py 2>&1 | tail -1 > gpurun_out/r2_k1k2_cfg1_c.json; cat gpurun_out/r2_k1k2_cfg1_c.json
CAFE_BENCH_FAMILIES=25000 CAFE_BENCH_TAXA=50 CAFE_BENCH_MAXSIZE=400 CAFE_BENCH_MU=0.8 python tools/k2_time.py 2>&1 | tail -1 > gpurun_out/r2_k1k2_cfg2shape_c.json; cat gpurun_out/r2_k1k2_cfg2shape_c.json
python -m pytest tests/test_gpu_parity.py tests/test_gpu_pvalue.py -x -q -m gpu 2>&1 | tail -4
python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r2_bench_full_n1_c.json 2> gpurun_out/r2_bench_full_n1_c.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_full_n1_c.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['ms_breakdown_rank0'])
for k,v in d['configs'].items(): print(k, {a:b for a,b in v.items() if a in ('value','k2_ms','cd_s','pvalues_s','draws_per_s','cd_tflops')}, v.get('roofline',{}).get('frac'))
PY
tail -3 gpurun_out/r2_bench_full_n1_c.err
