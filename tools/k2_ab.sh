#!/bin/bash
# A/B timing of the K2 variants on the bench workload (BASELINE configs[1]); one JSON line per variant in gpurun_out/k2_ab.jsonl
# usage: tools/k2_ab.sh "NAME1:ENV=VAL,ENV=VAL" "NAME2:" ...
mkdir -p gpurun_out
: > gpurun_out/k2_ab.jsonl
for spec in "$@"; do
  name="${spec%%:*}"; envs="${spec#*:}"
  envs="${envs//,/ }"
  out=$(env $envs CAFE_BENCH_FAMILIES=50000 CAFE_BENCH_TAXA=20 CAFE_BENCH_MAXSIZE=200 python tools/k2_time.py 2>&1 | tail -1)
  echo "{\"variant\": \"$name\", \"env\": \"$envs\", \"result\": $out}" >> gpurun_out/k2_ab.jsonl
done
cat gpurun_out/k2_ab.jsonl
