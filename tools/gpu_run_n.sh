#!/bin/bash
# call N (1 GPU): what the driver runs at round end - smoke, the GPU test suite, the default bench (both arms), with wall times
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) 2>&1 | tail -4
( time python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r2_gputests_final.log 2>&1; tail -5 gpurun_out/r2_gputests_final.log
( time python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err ) 2>&1 | tail -3
( time python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err ) 2>&1 | tail -3
python - <<'PY'
import json
r = json.loads(open('gpurun_out/r2_bench_reference_arm.json').read().strip().splitlines()[-1])
d = json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
print('reference arm', r['value'], r['cpu_baseline']['cores'], r['config']['workload'] == d['config']['workload'])
print('ours', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], 'ratio e2e', d['e2e']['value'] / r['value'])
print('cpu_baseline', d.get('cpu_baseline', {}).get('value'), d['configs']['configs[4]'].get('cpu_baseline'), d['configs']['configs[4]'].get('cd_speedup_vs_cpu'))
PY
