#!/bin/bash
V=/root/repo/cafe_b200/build/variants
export K2_STRESS_FRESH=1
python tools/k2_stress.py 20 2>&1 | tail -1
CAFE_GPU_LIB=$V/libcafe_gpu_nofast.so python tools/k2_stress.py 20 2>&1 | tail -1
CAFE_GPU_LIB=$V/libcafe_gpu_head.so python tools/k2_stress.py 20 2>&1 | tail -1
