"""bench/skew.py -- digit-pass behaviour on skewed / low-entropy keys (u32/u32 pairs, DoubleBuffer form):
uniform keys masked to fewer bits (constant upper digits), all-equal keys, AND-of-k entropy reduction, sorted input.
    python bench/skew.py [--log2n 27]      (product library; reference CUB beside it)"""
import argparse
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

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=27)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "skew.jsonl"))
a = ap.parse_args()
b2s = _lib.load()
from oracle import pyoracle  # noqa: E402

ref = pyoracle.load_gpu_reference("ref")
n = 1 << a.log2n
base = H.gen_device_keys(b2s, n, 4, 42, 1)
vals = H.gen_device_iota(b2s, n, 4)
cases = {"uniform": base, "low 16 bits": base & 0xFFFF, "low 8 bits": base & 0xFF, "all equal": base & 0,
         "high 16 bits": base & 0xFFFF0000, "AND-of-3": H.gen_device_keys(b2s, n, 4, 42, 3),
         "AND-of-5": H.gen_device_keys(b2s, n, 4, 42, 5), "sorted": torch.sort(base.view(torch.int32) & 0x7FFFFFFF).values.view(torch.uint32),
         "90% zeros": torch.where((base & 0xF) < 14 + 0 * base, base & 0, base)}
out = open(a.out, "a")
for name, keys in cases.items():
    keys = keys.contiguous()
    r_ref = time_sort(ref.sort_db, keys, vals, 6, 5)
    r = time_sort(b2s.b2s_radix_sort_db, keys, vals, 6, 5)
    ok = bool(torch.equal(r[2], r_ref[2]) and torch.equal(r[3], r_ref[3]))
    rec = {"case": name, "n": n, "b2s_ms": r[0], "ref_ms": r_ref[0], "b2s_gkeys_s": n / r[0] / 1e6, "speedup": r_ref[0] / r[0], "bit_exact": ok}
    print(json.dumps(rec), flush=True)
    out.write(json.dumps(rec) + "\n")
