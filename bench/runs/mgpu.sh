#!/bin/bash
# multi-GPU tests (N >= 2) + the driver's bench command at N GPUs (native C++ host, both legs)
N=${1:-2}
TAG=${2:-r3}
mkdir -p gpurun_out
if [ "$N" -le 2 ]; then timeout 900 python -m pytest tests/test_multi_gpu_gpu.py -m gpu -x -q 2>&1 | tail -3; fi
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_n${N}.json 2> gpurun_out/bench_${TAG}_n${N}.err
echo "bench exit $?"
python - <<PY
import json
r = json.load(open("gpurun_out/bench_${TAG}_n${N}.json"))
nv = (r["roofline"].get("nvlink") or {})
print("value", round(r["value"], 1), "ms", round(r["ms_per_step"], 3), "verified", r["config"]["verified"], "e2e", round(r["e2e"]["value"], 2), r["e2e"].get("host_numa_rank0"),
      "nvlink GB/s", round(nv.get("achieved", 0), 1), "partition ms", round(nv.get("kernel_ms", 0), 2))
print("   phases", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r["config"]["phases_ms"].items()})
c5 = r.get("config5_u64_u32")
if c5:
    for k in ("uniform", "and3"):
        print("   config5", k, round(c5[k]["value"], 1), "GKeys/s", round(c5[k]["ms_per_step"], 1), "ms verified", c5[k]["verified"],
              "nvlink", round((c5[k]["nvlink"] or {}).get("achieved", 0), 1), {kk: (round(v, 2) if isinstance(v, float) else v) for kk, v in c5[k]["phases_ms"].items()})
PY
grep -v "Warning\|OMP_NUM\|\*\*\*\*" gpurun_out/bench_${TAG}_n${N}.err | tail -4
