#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit: $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err
echo "bench exit $?"; python - <<'PY'
import json
r=json.load(open('gpurun_out/bench_r1e.json'))
print(r['value'], r['ms_per_step'], r['roofline']['frac'], r['roofline']['avg_launch_ms'], r['roofline']['histogram_ms'], r['config']['parity'], r['clocks'], r['e2e']['value'])
PY
tail -3 gpurun_out/bench_r1e.err
