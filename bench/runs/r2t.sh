#!/bin/bash
# round 2, run t: N=1 bench with / without NUMA binding of the host side (e2e leg)
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -14
cat /sys/devices/system/node/node*/cpulist 2>/dev/null | head -4; nproc
for nb in 1 0; do
B2S_NUMA_BIND=$nb timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2t_numa$nb.json 2> gpurun_out/bench_r2t_numa$nb.err; echo "bench exit $?"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2t_numa$nb.json"))
print($nb, d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("host_numa"))
PY
done
