#!/bin/bash
# call J (1 GPU): bench line (N=1, no CPU baseline leg), launch list, ncu full captures of K2 at configs[2] and configs[1]
mkdir -p gpurun_out
python bench.py --no-cpu-baseline > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; tail -c 600 gpurun_out/r2_bench_n1_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-sub --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_prune_fused2 -s 3 -c 1 -o gpurun_out/r2_k2_cfg2 -f python bench.py --steps 1 --warmup 3 --no-sub --no-cpu-baseline > gpurun_out/r2_ncu_cfg2.log 2>&1
CAFE_BENCH_CONFIG="configs[1]" ncu --set full --clock-control none --import-source on -k regex:k_prune_fused2 -s 3 -c 1 -o gpurun_out/r2_k2_cfg1 -f python bench.py --steps 1 --warmup 3 --no-sub --no-cpu-baseline > gpurun_out/r2_ncu_cfg1.log 2>&1
ls -la gpurun_out/*.ncu-rep
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n1_final.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['config']['ms_breakdown_rank0'])
for k,v in d.get('configs',{}).items(): print(k, {a:b for a,b in v.items() if a in ('value','k2_ms','cd_s','pvalues_s','draws_per_s','cd_tflops')}, v.get('roofline',{}).get('frac'))
PY
