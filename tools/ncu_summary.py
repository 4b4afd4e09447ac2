"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/: duration, DRAM traffic,
pipe utilisation, occupancy, stall mix, and the hottest SASS lines."""
import collections
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__ops_path_tensor_src_fp64.sum"]
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none summary of {rep}\n")
    for k in keys:
        if k in m:
            f.write(f"{k:85s} {m[k][0]} {m[k][1]}\n")
    f.write("\n# warp stall reasons (per issue-active cycle)\n")
    st = [(float(v[0]), h) for h, v in m.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for v, h in sorted(st, reverse=True)[:10]:
        f.write(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {v:.3f}\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    sh = srows[1]
    ix = {h: i for i, h in enumerate(sh)}
    data = [r for r in srows[2:] if len(r) >= len(sh)]
    tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
    ops = collections.Counter()
    for r in data:
        t = r[ix["Source"]].split()
        if t:
            op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
            ops[op] += int(r[ix["Instructions Executed"]])
    f.write("\n# executed warp-instructions by opcode (top 12)\n")
    for op, n in ops.most_common(12):
        f.write(f"{op:12s} {n}\n")
    f.write("\n# hottest SASS lines by stall samples\n")
    for i in sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:12]:
        r = data[i]
        f.write(f"{int(r[ix['# Samples']]) / tot:6.3f}  {r[ix['Source']].strip()[:80]}\n")
print(open(out).read())
