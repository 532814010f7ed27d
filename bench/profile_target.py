"""bench/profile_target.py -- minimal workload for ncu: BASELINE config 2 (SortPairs u32/u32) sorted a few times.
    ncu ... python bench/profile_target.py [--log2n 28] [--reps 3] [--case k4v4]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cub_b200 import _lib  # noqa: E402
from tests import harness as H  # noqa: E402

CASES = {"k4v4": (6, 4, 1), "k4v0": (6, 0, 1), "k8v4": (9, 4, 3), "f32desc": (8, 0, 1), "bf16desc": (5, 0, 1)}

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=28)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--case", default="k4v4")
a = ap.parse_args()
b2s = _lib.load()
kt, vb, rounds = CASES[a.case]
n = 1 << a.log2n
keys = H.gen_device_keys(b2s, n, H.KEY_BYTES[kt], 42, rounds)
if a.case.endswith("desc"):  # floating keys: +-0.0 at 1/256 each, as the reference's test generator forces them
    idx = torch.arange(n, device="cuda")
    keys[idx % 256 == 0] = 0
    keys[idx % 256 == 1] = torch.iinfo(keys.dtype).min
    del idx
vals = H.gen_device_iota(b2s, n, vb) if vb else None
ko = torch.empty_like(keys)
vo = torch.empty_like(vals) if vals is not None else None
for _ in range(a.reps):
    H.sort_ptr(b2s.b2s_radix_sort, keys, vals, kt, a.case.endswith("desc"), keys_out=ko, vals_out=vo, misalign=0)
torch.cuda.synchronize()
print("done", a.case, n)
