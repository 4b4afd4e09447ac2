#!/bin/bash
# call O (1 GPU): the GPU suite on the final code, and a 200-step bench run (sustained clocks over a search-length run)
mkdir -p gpurun_out
( time python -m pytest tests/ -q -m gpu ) > gpurun_out/r2_gputests_final.log 2>&1; tail -4 gpurun_out/r2_gputests_final.log
python bench.py --steps 200 --no-cpu-baseline --no-sub > gpurun_out/r2_bench_n1_200steps.json 2> gpurun_out/r2_bench_n1_200steps.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n1_200steps.json').read().strip().splitlines()[-1])
print(d['steps'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])
PY
