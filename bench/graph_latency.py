"""bench/graph_latency.py -- latency of small and mid-size sorts when the call is replayed from a CUDA graph (no host
enqueue cost), next to the plain call.  u32/u32 pairs, pointer form.   python bench/graph_latency.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cub_b200 as cb  # noqa: E402

out = open(os.path.join(ROOT, "gpurun_out", "graph_latency.jsonl"), "a")
for n in (4096, 8192, 16384, 65536, 262144, 1 << 20):
    keys = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), device="cuda", dtype=torch.int64).to(torch.int32)
    vals = torch.arange(n, device="cuda", dtype=torch.int32)
    ko, vo = torch.empty_like(keys), torch.empty_like(vals)
    err, nbytes = cb.DeviceRadixSort.SortPairs(None, 0, keys, ko, vals, vo, n)
    temp = torch.empty(nbytes, dtype=torch.uint8, device="cuda")

    def call():
        e, _ = cb.DeviceRadixSort.SortPairs(temp, nbytes, keys, ko, vals, vo, n)
        assert e == 0

    call()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        call()
    res = {}
    for name, fn in (("plain", call), ("graph", g.replay)):
        ts = []
        for _ in range(30):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res[name] = float(np.median(ts))
    ok = bool(torch.equal(ko, torch.sort(keys, stable=True).values))
    rec = {"n": n, "plain_us": res["plain"], "graph_us": res["graph"], "sorted": ok}
    print(json.dumps(rec), flush=True)
    out.write(json.dumps(rec) + "\n")
