timeout 1200 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/gputests.log 2>&1; echo "rc=$?" >> gpurun_out/gputests.log; tail -16 gpurun_out/gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches_lrt.csv python tools/time_lrt.py 20 200 20000 > gpurun_out/lrt_under_ncu.log 2>&1; tail -2 gpurun_out/lrt_under_ncu.log
