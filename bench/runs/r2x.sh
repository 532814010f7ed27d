#!/bin/bash
# round 2, run x: PF_NOBR on the other item widths (18 = production shape + NOBR, 19 = 320 x 30 x 3 + NOBR)
mkdir -p gpurun_out
fmt='import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    if d.get("impl") != "b2s": continue
    print(d["case"], d.get("variant"), d.get("nt"), d.get("ipt"), d.get("minb"), d.get("flow"), round(d["best_ms"], 3), round(d["gkeys_s"], 2), d.get("bit_exact_vs_ref"))'
B2S_LIB=cub_b200/libb2s_tune.so timeout 900 python bench/tune.py --log2n 27 --cases k8v4,k8v0,k8v8,k4v8,k2v4,k2v0,k1v0,k4v0 --variants 0,18,19 --out gpurun_out/tune_r2x.jsonl 2>&1 | python -c "$fmt"
