#!/bin/bash
# round 2, run 3b: tile summary in shared memory instead of re-derived geometry
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
fmt='import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    if d.get("impl") != "b2s": continue
    print(d["case"], d.get("variant"), d.get("nt"), d.get("ipt"), d.get("minb"), d.get("flow"), round(d["best_ms"], 3), round(d["gkeys_s"], 2), d.get("bit_exact_vs_ref"))'
B2S_LIB=cub_b200/libb2s_tune.so timeout 900 python bench/tune.py --log2n 28 --cases k4v4,k4v0 --variants 0 --out gpurun_out/tune_r3b.jsonl 2>&1 | python -c "$fmt"
B2S_LIB=cub_b200/libb2s_tune.so timeout 900 python bench/tune.py --log2n 27 --cases k8v4,k8v0,k2v0 --variants 0 --out gpurun_out/tune_r3b.jsonl 2>&1 | python -c "$fmt"
