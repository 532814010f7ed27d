#!/bin/bash
# persistent / L2-prefetch variants of the digit pass
mkdir -p gpurun_out
export B2S_LIB=cub_b200/libb2s_tune.so
V=0,9,18,19,20,21,22,23,24,25,26,27,28,29,30,31,32,33,34,35
timeout 600 python bench/tune.py --log2n 28 --cases k4v4 --variants $V --iters 7 --out gpurun_out/tune_r1s.jsonl 2>&1 | python bench/tune_fmt.py
timeout 600 python bench/tune.py --log2n 27 --cases k4v0,k8v4 --variants $V --iters 5 --out gpurun_out/tune_r1s.jsonl 2>&1 | python bench/tune_fmt.py
