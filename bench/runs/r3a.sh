#!/bin/bash
# round 2, run 3a: production with 256 x 48 pairs; A/B of the constant-digit flag load
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for sk in 1 0; do
B2S_SKIP_CONSTANT=$sk timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r3a_$sk.json 2> gpurun_out/bench_r3a_$sk.err; echo "bench exit $?"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_r3a_$sk.json"))
print("skip_constant=$sk", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["e2e"]["value"])
PY
done
