#!/bin/bash
mkdir -p gpurun_out
export B2S_LIB=cub_b200/libb2s_tune.so
V=85,88,89,90,91,92,93,94,95,96,97,98,99,100,101,102,103,104,105
timeout 600 python bench/tune.py --log2n 28 --cases k4v4 --variants $V --iters 7 --out gpurun_out/tune_r1x.jsonl 2>&1 | python bench/tune_fmt.py
timeout 600 python bench/tune.py --log2n 27 --cases k4v0,k8v4,k8v0,k4v8 --variants $V --iters 5 --out gpurun_out/tune_r1x.jsonl 2>&1 | python bench/tune_fmt.py
timeout 600 python bench/trace.py --variants 107,106 --out gpurun_out/trace_r1d.jsonl 2>&1
