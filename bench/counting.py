"""bench/counting.py -- keys-only sorts of 1- and 2-byte keys: the counting path (b2s_narrow.cu) against our own digit passes
and the reference CUB on the same buffers, over sizes (to place the cut-over) and key distributions.  Every timed result is
compared bit for bit with the reference.  One JSON line per (key type, n, distribution); with --steps also the per-launch
times of the counting path (CUDA events between the launches, b2s_timing_*).  Not the bench line.
    python bench/counting.py [--out gpurun_out/counting.jsonl] [--max-log2 29]"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cub_b200 import _lib  # noqa: E402
from tests import harness as H  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "counting.jsonl"))
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--max-log2", type=int, default=29)
ap.add_argument("--min-log2", type=int, default=16)
ap.add_argument("--steps", action="store_true")
ap.add_argument("--keys", default="5,4,2,3,0,1", help="key type ids (tests/harness.py KEY_NAMES)")
ap.add_argument("--tag", default="")
a = ap.parse_args()
b2s = _lib.load()
from oracle import pyoracle as po  # noqa: E402  (comparator only)

ref = po.load_gpu_reference("ref")
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    peak = 6650.0


def run(fn, keys, out, kt, desc, iters):
    n = keys.numel()
    nbytes = ctypes.c_size_t(0)
    args = (H._p(keys), H._p(out), None, None, n, kt, 0, 4, int(desc), 0, H.KEY_BYTES[kt] * 8)
    assert fn(None, ctypes.byref(nbytes), *args, None) == 0
    temp = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    best = 1e30
    for it in range(iters + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        assert fn(ctypes.c_void_p(temp.data_ptr()), ctypes.byref(nbytes), *args, H.stream_handle()) == 0
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            best = min(best, e0.elapsed_time(e1))
    return best


def make(kt, n, dist):
    nb = H.KEY_BYTES[kt]
    keys = H.gen_device_keys(b2s, n, nb, seed=7, and_rounds=3 if dist == "and3" else 1)
    if dist == "const":
        keys.fill_(0x3C)
    elif dist == "few":  # 4 distinct keys
        keys &= 3
    elif kt in (4, 5):   # floating keys: +-0 as the reference's test generator forces them
        idx = torch.arange(n, device="cuda")
        keys[idx % 256 == 0] = 0
        keys[idx % 256 == 1] = torch.iinfo(H.CONTAINER[nb]).min
        if dist == "nozero":
            keys |= 1
    return keys


os.makedirs(os.path.dirname(a.out), exist_ok=True)
with open(a.out, "a") as f:
    for kt in [int(x) for x in a.keys.split(",")]:
        nb = H.KEY_BYTES[kt]
        sizes = [(lg, "uniform") for lg in range(a.min_log2, a.max_log2 + 1, 2)] + [(a.max_log2, "uniform")]
        sizes += [(min(28, a.max_log2), d) for d in ("and3", "const", "few")] + ([(min(28, a.max_log2), "nozero")] if kt in (4, 5) else [])
        seen = set()
        for lg, dist in sizes:
            if (lg, dist) in seen:
                continue
            seen.add((lg, dist))
            n = 1 << lg
            keys = make(kt, n, dist)
            out_c, out_d, out_r = (torch.empty_like(keys) for _ in range(3))
            desc = kt in (4, 5)
            old_min = b2s.b2s_set_counting_min_items(nb, 1)
            b2s.b2s_set_counting_sort(1)
            t_cnt = run(b2s.b2s_radix_sort, keys, out_c, kt, desc, a.iters)
            steps = None
            if a.steps:
                b2s.b2s_timing_enable(1)
                run(b2s.b2s_radix_sort, keys, out_c, kt, desc, 1)
                ms = (ctypes.c_float * 16)()
                k = b2s.b2s_timing_read(ms, 16)
                steps = [round(ms[i], 4) for i in range(k)]
                b2s.b2s_timing_enable(0)
            b2s.b2s_set_counting_sort(0)
            t_dig = run(b2s.b2s_radix_sort, keys, out_d, kt, desc, a.iters)
            b2s.b2s_set_counting_sort(1)
            b2s.b2s_set_counting_min_items(nb, old_min)
            t_ref = run(ref.sort, keys, out_r, kt, desc, a.iters) if ref is not None else None
            exact = bool(torch.equal(out_c, out_d)) and (ref is None or bool(torch.equal(out_c, out_r)))
            rec = {"tag": a.tag, "key": H.KEY_NAMES[kt], "log2n": lg, "dist": dist, "descending": desc, "counting_ms": round(t_cnt, 4),
                   "digit_passes_ms": round(t_dig, 4), "ref_cub_ms": None if t_ref is None else round(t_ref, 4),
                   "counting_gkeys": round(n / t_cnt / 1e6, 1), "digit_gkeys": round(n / t_dig / 1e6, 1),
                   "ref_gkeys": None if t_ref is None else round(n / t_ref / 1e6, 1),
                   "counting_frac_of_hbm_at_2K_bytes_per_key": round(2 * nb * n / t_cnt / 1e6 / peak, 3),
                   "bit_exact": exact}
            if steps is not None:
                rec["launch_ms"] = steps
            line = json.dumps(rec)
            print(line, flush=True)
            f.write(line + "\n")
            del keys, out_c, out_d, out_r
            torch.cuda.empty_cache()
