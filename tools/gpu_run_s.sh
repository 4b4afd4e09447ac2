#!/bin/bash
# call S (1 GPU): ncu launch list of the default bench command's timed region on the final code (K1 recurrence kernel)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches_bench_final.csv python bench.py --steps 2 --warmup 3 --no-sub --no-cpu-baseline > gpurun_out/r2_bench_under_ncu_final.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2_launches_bench_final.csv")) if len(r) > 10 and r[0].isdigit()]
# the evaluations of the timed region: the last 2 steps = the last (k_bd_matrix_rec, k_transpose_keys, k_prune_fused2, k_score_reduce) groups
names = [r[4] for r in rows]; t = [float(r[-1]) / 1e3 for r in rows]
last = len(names) - 1 - names[::-1].index(next(n for n in names[::-1] if "k_prune_fused2" in n))
first = max(i for i in range(last) if "k_bd_matrix" in names[i])
tot = collections.OrderedDict()
for i in range(first, min(len(names), last + 2)):
    tot[names[i][:60]] = tot.get(names[i][:60], 0.0) + t[i]
s = sum(tot.values())
with open("gpurun_out/r2_launches_bench_final_summary.txt", "w") as f:
    f.write("# one objective evaluation of the default bench (BASELINE configs[2], N = 1) under ncu --metrics gpu__time_duration.sum --clock-control none\n")
    f.write("# (cold-cache, serialised launches: shares, not absolute times)\n# kernel, us, share\n")
    for k, v in tot.items():
        f.write(f"{k:62s} {v:12.1f} {100 * v / s:6.2f}%\n")
print(open("gpurun_out/r2_launches_bench_final_summary.txt").read())
PY
