#!/bin/bash
# round 2, run s: depth-2 ranking pipeline (PF_DEEP) at 256 x {42,44,46} x 2
mkdir -p gpurun_out
fmt='import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(d["case"], d.get("impl"), d.get("variant"), d.get("nt"), d.get("ipt"), d.get("minb"), d.get("flow"), d.get("match"), round(d["best_ms"], 3), round(d["gkeys_s"], 2), d.get("bit_exact_vs_ref"))'
B2S_LIB=cub_b200/libb2s_tune.so timeout 600 python bench/tune.py --log2n 28 --cases k4v4 --variants ${1:-0,18,20,22,23} --out gpurun_out/tune_${2:-r2s}.jsonl 2>&1 | python -c "$fmt"
