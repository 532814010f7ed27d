#!/bin/bash
# round 2, run h: full GPU suite, bench N=1, reference arm, smoke
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err
echo "bench exit $?"; python -c "
import json; r=json.load(open('gpurun_out/bench_r2h.json')); print(r['value'], r['ms_per_step'], r['roofline']['frac'], r['roofline'].get('avg_launch_ms'), r['config'].get('parity'), r['e2e']['value'], r.get('reference_gpu',{}).get('value'), r['clocks'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_r2h.json 2> gpurun_out/bench_ref_r2h.err; echo "ref exit $?"; cat gpurun_out/bench_ref_r2h.json | head -c 1500
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
