#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not config" > gpurun_out/pytest_small.log 2>&1
echo "pytest small exit: $?" >> gpurun_out/pytest_small.log
tail -5 gpurun_out/pytest_small.log
B2S_LIB=cub_b200/libb2s_tune.so timeout 900 python bench/tune.py --log2n 28 --cases k4v4 --out gpurun_out/tune_r1b.jsonl > gpurun_out/tune.log 2>&1
B2S_LIB=cub_b200/libb2s_tune.so timeout 900 python bench/tune.py --log2n 27 --cases k4v0,k8v4,k8v4and3,k2v0 --out gpurun_out/tune_r1b.jsonl >> gpurun_out/tune.log 2>&1
echo "tune exit: $?" >> gpurun_out/tune.log
grep -v '"variant": [0-9]*,' gpurun_out/tune.log | tail -3
python - <<'PY'
import json
best={}
for l in open('gpurun_out/tune_r1b.jsonl'):
    r=json.loads(l)
    k=(r['case'],r['impl'])
    if k not in best or r['gkeys_s']>best[k]['gkeys_s']: best[k]=r
for k,r in sorted(best.items()): print(k, round(r['gkeys_s'],2), r.get('variant'), r.get('nt'), r.get('ipt'), r.get('minb'), r.get('bit_exact_vs_ref'))
for l in open('gpurun_out/tune_r1b.jsonl'):
    r=json.loads(l)
    if r['impl']=='b2s': print(r['case'], r['variant'], r['nt'], r['ipt'], r['minb'], round(r['gkeys_s'],2), r['bit_exact_vs_ref'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1b.csv python bench/profile_target.py --reps 2 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:onesweep -s 4 -c 2 -o gpurun_out/prof_onesweep_r1b -f python bench/profile_target.py --reps 2 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:histogram -s 1 -c 1 -o gpurun_out/prof_hist_r1b -f python bench/profile_target.py --reps 2 >> gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err
echo "bench exit $?"; cat gpurun_out/bench_r1b.json; tail -5 gpurun_out/bench_r1b.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
