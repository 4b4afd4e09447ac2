#!/bin/bash
# call R (8 GPUs): strong-scaling lines of the headline workload at N = 4 and N = 8 (torchrun, as the driver launches them)
mkdir -p gpurun_out
for N in 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2b_bench_n$N.json 2> gpurun_out/r2b_bench_n$N.err
  tail -c 300 gpurun_out/r2b_bench_n$N.err | grep -v OMP
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2b_bench_n$N.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["config"]["ms_breakdown_rank0"], d["roofline"]["frac"], d["clocks"])
PY
done
