#!/bin/bash
# round 2, run w: ranking atomic without the branch (PF_NOBR)
mkdir -p gpurun_out
fmt='import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(d["case"], d.get("impl"), d.get("variant"), d.get("nt"), d.get("ipt"), d.get("minb"), d.get("flow"), d.get("match"), round(d["best_ms"], 3), round(d["gkeys_s"], 2), d.get("bit_exact_vs_ref"))'
B2S_LIB=cub_b200/libb2s_tune.so timeout 600 python bench/tune.py --log2n 28 --cases k4v4 --variants 0,20 --out gpurun_out/tune_r2w.jsonl 2>&1 | python -c "$fmt"
B2S_LIB=cub_b200/libb2s_tune.so timeout 600 python bench/tune.py --log2n 28 --cases k4v0 --variants 0,27,25,26 --out gpurun_out/tune_r2w.jsonl 2>&1 | python -c "$fmt"
