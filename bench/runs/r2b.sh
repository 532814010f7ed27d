#!/bin/bash
# round 2, run b: (key,value) pair scatter variants 40-55 vs production (0), u32/u32 2^28
mkdir -p gpurun_out
export B2S_LIB=cub_b200/libb2s_tune.so
timeout 600 python bench/tune.py --log2n 28 --cases k4v4 --variants 0,40,41,42,43,44,45,46,47,48,49,50,51,52,53,54,55 --iters 7 --out gpurun_out/tune_r2b.jsonl 2>&1 | python bench/tune_fmt.py
