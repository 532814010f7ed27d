#!/bin/bash
# round 2, run k: tests of the constant-digit short circuit + evidence capture (r2j.sh) + skewed-key timings
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_frontend.py tests/test_abi.py -m gpu -q 2>&1 | tail -4
bash bench/runs/r2j.sh r2k
timeout 600 python bench/skew.py > gpurun_out/skew_r2k.log 2>&1; tail -14 gpurun_out/skew_r2k.log
rm -f gpurun_out/*.log.tmp
