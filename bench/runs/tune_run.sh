#!/bin/bash
# one tuning-build sweep on the GPU box:  bash bench/runs/tune_run.sh <cases> <variants> [log2n] [tag] [--ablate]
# e.g. gpurun -- 'bash bench/runs/tune_run.sh k4v4,k4v0 0,18,19 28 r2x'
mkdir -p gpurun_out
fmt='import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(d["case"], d.get("impl"), d.get("variant"), d.get("nt"), d.get("ipt"), d.get("minb"), d.get("flow"), d.get("match"), round(d["best_ms"], 3), round(d["gkeys_s"], 2), d.get("bit_exact_vs_ref"))'
B2S_LIB=cub_b200/libb2s_tune.so timeout 900 python bench/tune.py $5 --log2n ${3:-28} --cases ${1:-k4v4} --variants ${2:-0} --out gpurun_out/tune_${4:-run}.jsonl 2>&1 | python -c "$fmt"
