#!/bin/bash
# round 2, run d: production kernel (b2s_pass.cuh) flows: split / pair / TMA write-out vs the lab kernel (variant 30)
mkdir -p gpurun_out
export B2S_LIB=cub_b200/libb2s_tune.so
timeout 600 python bench/tune.py --log2n 28 --cases k4v4 --variants 0,1,2,3,4,5,6,7,8,9,10,11,12,30 --iters 7 --out gpurun_out/tune_r2d.jsonl 2>&1 | python bench/tune_fmt.py
timeout 600 python bench/tune.py --log2n 27 --cases k4v0,k8v4,k8v0,k8v8 --variants 0,1,7,8,9,10,12,13,14,15,16,17,30 --iters 5 --out gpurun_out/tune_r2d.jsonl 2>&1 | python bench/tune_fmt.py
