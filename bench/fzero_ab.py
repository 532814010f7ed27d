"""bench/fzero_ab.py -- full-range sorts of f32 / f64 keys (alone and with u32 values): zero recording (b2s_fzero.cu) on vs off vs
reference CUB on the same buffers; bit-exactness of every timed result.  One JSON line per case.  Not the bench line.
    python bench/fzero_ab.py [--out gpurun_out/fzero_ab.jsonl] [--log2n 28]"""
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
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "fzero_ab.jsonl"))
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--log2n", type=int, default=28)
a = ap.parse_args()
b2s = _lib.load()
from oracle import pyoracle as po  # noqa: E402  (comparator only)

ref = po.load_gpu_reference("ref")


def run(fn, keys, vals, ko, vo, kt, desc):
    n = keys.numel()
    vb = 4 if vals is not None else 0
    nbytes = ctypes.c_size_t(0)
    args = (H._p(keys), H._p(ko), H._p(vals), H._p(vo), n, kt, vb, 4, int(desc), 0, H.KEY_BYTES[kt] * 8)
    assert fn(None, ctypes.byref(nbytes), *args, None) == 0
    temp = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    best = 1e30
    for it in range(a.iters + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        assert fn(ctypes.c_void_p(temp.data_ptr()), ctypes.byref(nbytes), *args, H.stream_handle()) == 0
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            best = min(best, e0.elapsed_time(e1))
    return best


os.makedirs(os.path.dirname(a.out), exist_ok=True)
with open(a.out, "a") as f:
    for kt, lg in ((8, a.log2n), (8, a.log2n + 1), (11, a.log2n - 1)):
        nb = H.KEY_BYTES[kt]
        n = 1 << lg
        for zeros in ("spiked", "none"):
            for with_vals in (False, True):
                if with_vals and lg > a.log2n:
                    continue
                keys = H.gen_device_keys(b2s, n, nb, seed=9)
                if zeros == "spiked":
                    idx = torch.arange(n, device="cuda")
                    keys[idx % 256 == 0] = 0
                    keys[idx % 256 == 1] = torch.iinfo(H.CONTAINER[nb]).min
                    del idx
                else:
                    keys |= 1
                vals = H.gen_device_iota(b2s, n, 4) if with_vals else None
                outs = [(torch.empty_like(keys), torch.empty_like(vals) if with_vals else None) for _ in range(3)]
                desc = True
                b2s.b2s_set_float_zero_recording(1)
                t_on = run(b2s.b2s_radix_sort, keys, vals, *outs[0], kt, desc)
                b2s.b2s_set_float_zero_recording(0)
                t_off = run(b2s.b2s_radix_sort, keys, vals, *outs[1], kt, desc)
                b2s.b2s_set_float_zero_recording(1)
                t_ref = run(ref.sort, keys, vals, *outs[2], kt, desc) if ref is not None else None
                exact = all(torch.equal(outs[0][0], o[0]) and (not with_vals or torch.equal(outs[0][1], o[1])) for o in outs[1:3 if ref is not None else 2])
                rec = {"key": H.KEY_NAMES[kt], "values": "u32" if with_vals else None, "log2n": lg, "zeros": zeros, "descending": desc,
                       "recording_ms": round(t_on, 4), "collapse_every_digit_ms": round(t_off, 4), "ref_cub_ms": None if t_ref is None else round(t_ref, 4),
                       "recording_gkeys": round(n / t_on / 1e6, 2), "collapse_gkeys": round(n / t_off / 1e6, 2), "bit_exact": bool(exact)}
                line = json.dumps(rec)
                print(line, flush=True)
                f.write(line + "\n")
                del keys, vals, outs
                torch.cuda.empty_cache()
