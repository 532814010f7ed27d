#!/bin/bash
# round 2, run i: 8-GPU partition-shape sweep of the native multi-GPU sorter (threshold compares in SplitterOp)
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu_gpu.py -m gpu -x -q 2>&1 | tail -2
for shape in 0 1 2; do
B2S_SPLIT_SHAPE=$shape B2S_CONFIG5_LOG2N=$([ $shape = 0 ] && echo 30 || echo 0) timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r2i_n${N}_s$shape.json 2> gpurun_out/bench_r2i_n${N}_s$shape.err
echo "shape $shape exit $?"
python - <<PY
import json
r=json.loads(open("gpurun_out/bench_r2i_n${N}_s$shape.json").read())
nv=(r["roofline"].get("nvlink") or {})
print("value", round(r["value"],1), "ms", round(r["ms_per_step"],2), "verified", r["config"]["verified"], "e2e", round(r["e2e"]["value"],2), "nvlink", round(nv.get("achieved",0),1), "part ms", round(nv.get("kernel_ms",0),2))
print("   ", {k: (round(v,3) if isinstance(v,float) else v) for k,v in r["config"]["phases_ms"].items() if k not in ("exchange","timing")})
c5=r.get("config5_u64_u32")
if c5:
    for k in ("uniform","and3"):
        print("   c5",k, round(c5[k]["value"],1), round(c5[k]["ms_per_step"],1),"ms", c5[k]["verified"], "nvlink", round((c5[k]["nvlink"] or {}).get("achieved",0),1), {a: (round(b,2) if isinstance(b,float) else b) for a,b in c5[k]["phases_ms"].items() if a not in ("exchange","timing")})
print(r["clocks"])
PY
done
