#!/bin/bash
mkdir -p gpurun_out
export B2S_LIB=cub_b200/libb2s_tune.so
timeout 600 python bench/trace.py --variants 36,37,38,39 --out gpurun_out/trace_r1.jsonl 2>&1
