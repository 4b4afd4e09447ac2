#!/bin/bash
# call L (1 GPU): compute-sanitizer on the fused kernels (memcheck, synccheck) + the tcgen05 prototype (memcheck), small cases
mkdir -p gpurun_out
T="tests/test_gpu_parity.py -k example_tree_or_two_lambda_or_error_model_band_or_kat_small"
T=$(echo $T | sed 's/_or_/ or /g')
( echo "## memcheck: fused K2 (score path), windowed K2 (K4/K5), branch cutting kernels"
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -k "example_tree or two_lambda or error_model_band or kat_small" tests/test_gpu_branchcut.py -x -q 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | tail -8
  echo "## memcheck: windowed path"
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pvalue.py -k "fused_windowed_path or k5_family" -x -q 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | tail -8
  echo "## synccheck: same score-path tests"
  timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -k "example_tree or two_lambda or kat_small" -x -q 2>&1 | grep -E "passed|failed|ERROR SUMMARY|error" | tail -6
) > gpurun_out/r2_memcheck.txt 2>&1
cat gpurun_out/r2_memcheck.txt
