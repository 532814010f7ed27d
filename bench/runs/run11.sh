#!/bin/bash
# final tuning-build evidence for DESIGN.md: technique ladder, neighbours, all shapes at the production point, phase traces
mkdir -p gpurun_out
export B2S_LIB=cub_b200/libb2s_tune.so
timeout 600 python bench/tune.py --log2n 28 --cases k4v4 --variants 0,1,2,3,10,11,12,13,14,15,16,19,20,21,22,24 --iters 7 --out gpurun_out/tune_r1z.jsonl 2>&1 | python bench/tune_fmt.py
timeout 600 python bench/tune.py --log2n 27 --cases k4v0,k8v4,k8v0,k4v8,k2v0,k8v4and3,k1v0,k2v4 --variants 0,3,17,18 --iters 5 --out gpurun_out/tune_r1z.jsonl 2>&1 | python bench/tune_fmt.py
timeout 600 python bench/trace.py --variants 26,27 --out gpurun_out/trace_r1f.jsonl 2>&1 | tee gpurun_out/trace_r1f.txt
