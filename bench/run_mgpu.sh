#!/bin/bash
# usage: bash bench/run_mgpu.sh <ngpus> [tag]     multi-GPU tests + bench (native C++ host; torch-orchestrated host for A/B)
N=${1:-2}
TAG=${2:-r2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu_gpu.py -m gpu -x -q > gpurun_out/pytest_mgpu_$N.log 2>&1
echo "pytest mgpu exit: $?"; tail -30 gpurun_out/pytest_mgpu_$N.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_n${N}_$name.json 2> gpurun_out/bench_${TAG}_n${N}_$name.err
  echo "bench $name exit $?"; python - <<PY
import json
try:
    r = json.load(open("gpurun_out/bench_${TAG}_n${N}_$name.json"))
    nv = (r["roofline"].get("nvlink") or {})
    print("$name", "value", round(r["value"], 1), "ms", round(r["ms_per_step"], 2), "verified", r["config"]["verified"], "e2e", round(r["e2e"]["value"], 2),
          "nvlink GB/s", round(nv.get("achieved", 0), 1), "partition ms", round(nv.get("kernel_ms", 0), 2))
    print("   phases", r["config"]["phases_ms"])
    c5 = r.get("config5_u64_u32")
    if c5:
        for k in ("uniform", "and3"):
            print("   config5", k, round(c5[k]["value"], 1), "GKeys/s", round(c5[k]["ms_per_step"], 1), "ms verified", c5[k]["verified"],
                  "nvlink", round((c5[k]["nvlink"] or {}).get("achieved", 0), 1), c5[k]["phases_ms"])
except Exception as e:
    print("$name: no result", e)
PY
  grep -v Warning gpurun_out/bench_${TAG}_n${N}_$name.err | tail -8
}
run native B2S_MGPU_BACKEND=native
run native_itemstores B2S_MGPU_BACKEND=native B2S_SPLIT_BULK=0 B2S_CONFIG5_LOG2N=0
run torch_peer B2S_MGPU_BACKEND=torch B2S_EXCHANGE=peer B2S_CONFIG5_LOG2N=0
