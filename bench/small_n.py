"""bench/small_n.py -- latency of small sorts (u32/u32 pairs and u32 keys, DoubleBuffer form, CUDA events, best of 20):
ours vs reference CUB 2.2.0 (which switches to a single-tile kernel at n <= 4864).   python bench/small_n.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "bench"))
from cub_b200 import _lib  # noqa: E402
from tests import harness as H  # noqa: E402
from tune import time_sort  # noqa: E402
from oracle import pyoracle  # noqa: E402

b2s = _lib.load()
ref = pyoracle.load_gpu_reference("ref")
out = open(os.path.join(ROOT, "gpurun_out", "small_n.jsonl"), "a")
for vb in (4, 0):
    for n in (256, 1024, 4096, 4864, 4865, 16384, 65536, 262144, 1 << 20, 1 << 22):
        keys = H.gen_device_keys(b2s, n, 4, 42, 1)
        vals = H.gen_device_iota(b2s, n, 4) if vb else None
        r_ref = time_sort(ref.sort_db, keys, vals, 6, 20, 3)
        r = time_sort(b2s.b2s_radix_sort_db, keys, vals, 6, 20, 3)
        ok = bool(torch.equal(r[2], r_ref[2]) and (vals is None or torch.equal(r[3], r_ref[3])))
        rec = {"value_bytes": vb, "n": n, "b2s_us": r[0] * 1e3, "ref_us": r_ref[0] * 1e3, "bit_exact": ok}
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n")
