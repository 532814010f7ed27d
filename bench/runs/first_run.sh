#!/bin/bash
# First GPU contact: quick parity subset, then the variant sweep next to reference CUB.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not config" > gpurun_out/pytest_small.log 2>&1
echo "pytest small exit: $?" >> gpurun_out/pytest_small.log
tail -30 gpurun_out/pytest_small.log
B2S_LIB=cub_b200/libb2s_tune.so timeout 900 python bench/tune.py --log2n 27 --cases k4v4,k4v0,k8v4 --out gpurun_out/tune_r1a.jsonl > gpurun_out/tune.log 2>&1
echo "tune exit: $?" >> gpurun_out/tune.log
tail -60 gpurun_out/tune.log
