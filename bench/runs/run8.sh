#!/bin/bash
mkdir -p gpurun_out
export B2S_LIB=cub_b200/libb2s_tune.so
V=0,108,109,110,111,112,113,114,115,116,117,118,119
timeout 600 python bench/tune.py --log2n 28 --cases k4v4 --variants $V --iters 7 --out gpurun_out/tune_r1y.jsonl 2>&1 | python bench/tune_fmt.py
timeout 600 python bench/tune.py --log2n 27 --cases k4v0,k8v4,k8v0,k4v8,k2v0 --variants $V --iters 5 --out gpurun_out/tune_r1y.jsonl 2>&1 | python bench/tune_fmt.py
timeout 600 python bench/trace.py --variants 106 --out gpurun_out/trace_r1e.jsonl 2>&1
