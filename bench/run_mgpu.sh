#!/bin/bash
# usage: bash bench/run_mgpu.sh <ngpus>
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu_gpu.py -m gpu -x -q > gpurun_out/pytest_mgpu_$N.log 2>&1
echo "pytest mgpu exit: $?"; tail -30 gpurun_out/pytest_mgpu_$N.log
for ex in peer nccl; do
B2S_EXCHANGE=$ex timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_mgpu_${N}_$ex.json 2> gpurun_out/bench_mgpu_${N}_$ex.err
echo "bench $ex exit $?"; cat gpurun_out/bench_mgpu_${N}_$ex.json; grep -v Warning gpurun_out/bench_mgpu_${N}_$ex.err | tail -15
done
