"""bench/mgpu_config5.py -- BASELINE.json configs[4]: multi-GPU SortPairs u64 keys / u32 values, 2^LOG2N pairs per
GPU (2^30 per GPU x 8 GPUs = 2^33), uniform and AND-of-3 entropy-reduced keys; verified by local order +
cross-rank boundaries + global multiset checksums; device-timed, max over ranks.  A parity/scale case, not the
bench line.   torchrun --nproc-per-node N bench/mgpu_config5.py [--log2n 30] [--reps 3]"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cub_b200 import _lib, multi_gpu  # noqa: E402
from tests import harness as H  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=30)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--exchange", default="auto")
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
b2s = _lib.load()
n = 1 << a.log2n
sorter = multi_gpu.DistributedSorter(n, torch.uint64, torch.uint32, exchange=a.exchange)
vals = H.gen_device_iota(b2s, n, 4)
for name, rounds in (("uniform", 1), ("and3", 3)):
    keys = torch.empty(n, dtype=torch.int64, device="cuda")
    # distinct stream per rank: first_index offsets the counter-based generator
    assert b2s.b2s_fill_keys(keys.data_ptr(), n, 8, 42, rounds, rank * n, H.stream_handle()) == 0
    out = sorter.sort(keys, vals)  # warm-up
    ok = sorter.verify(keys, vals, out)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        out = sorter.sort(keys, vals)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.reps], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ph = sorter.last_phase_ms()
    if rank == 0:
        total = n * world
        print(json.dumps({"case": f"SortPairs u64/u32 {name}", "n_gpus": world, "n_per_gpu": n, "n_total": total,
                          "ms_per_sort": float(ms.item()), "gkeys_s": total / float(ms.item()) / 1e6, "verified": bool(ok),
                          "shard_sizes": out.counts_all, "phases_ms": ph}), flush=True)
    del keys
sorter.close()
dist.destroy_process_group()
