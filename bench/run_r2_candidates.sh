#!/bin/bash
# Round-2 starting point: measure the candidates that round 1 wrote but could not run (tuning variants 29-33:
# chunked key copy, early first look-back window) against the production point (variant 0).  ~1 GPU-minute.
#   gpurun --timeout 900 -- bash bench/run_r2_candidates.sh
mkdir -p gpurun_out
export B2S_LIB=cub_b200/libb2s_tune.so
V=0,29,30,31,32,33
timeout 600 python bench/tune.py --log2n 28 --cases k4v4 --variants $V --iters 7 --out gpurun_out/tune_r2a.jsonl 2>&1 | python bench/tune_fmt.py
timeout 600 python bench/tune.py --log2n 27 --cases k4v0,k8v4,k8v0,k2v0 --variants $V --iters 5 --out gpurun_out/tune_r2a.jsonl 2>&1 | python bench/tune_fmt.py
# keys alone with larger tiles (variants 34-39)
timeout 600 python bench/tune.py --log2n 27 --cases k4v0,k2v0,k1v0 --variants 0,34,35,36,37,38,39 --iters 5 --out gpurun_out/tune_r2a.jsonl 2>&1 | python bench/tune_fmt.py
# multi-GPU candidate (needs --gpus 2): partition kernel with 16-byte stores per destination run
#   gpurun --gpus 2 --timeout 1200 -- 'B2S_SPLIT_WIDE=1 bash bench/run_mgpu.sh 2'
