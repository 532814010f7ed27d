#!/bin/bash
# round 2, run 3d: persistent pair kernel (variants 22-24) vs production
mkdir -p gpurun_out
fmt='import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    if d.get("impl") != "b2s": continue
    print(d["case"], d.get("variant"), d.get("nt"), d.get("ipt"), d.get("minb"), d.get("flow"), round(d["best_ms"], 3), round(d["gkeys_s"], 2), d.get("bit_exact_vs_ref"))'
B2S_LIB=cub_b200/libb2s_tune.so timeout 150 python bench/tune.py --ablate --log2n ${2:-28} --cases k4v4 --variants ${1:-0,22,23,24} --out gpurun_out/tune_r3d.jsonl 2>&1 | python -c "$fmt"
