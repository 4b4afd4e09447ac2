#!/bin/bash
# call B (N GPUs): the multi-GPU tests and the strong-scaling bench line at N ranks
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_multi_tests.log
cat gpurun_out/r2_multi_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 2500 gpurun_out/r2_bench_n$N.json; tail -5 gpurun_out/r2_bench_n$N.err
