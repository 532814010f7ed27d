#!/bin/bash
# GPU-box check script: parity tests, tuning sweep of the main cases, bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit: $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
B2S_LIB=cub_b200/libb2s_tune.so timeout 600 python bench/tune.py --log2n 27 --cases k4v0,k8v4,k8v4and3,k2v0,k8v0,k4v8 --out gpurun_out/tune_r1j.jsonl 2>&1 | python bench/tune_fmt.py | grep -E "cub| v (0|4|5|9|10) "
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err
echo "bench exit $?"; cat gpurun_out/bench_r1c.json; tail -3 gpurun_out/bench_r1c.err
