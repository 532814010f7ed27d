"""bench/extras.py -- the SURVEY section 8(f) rows next to the reference's own entry points, same GPU, same buffers:
DeviceSegmentedRadixSort, struct keys (decomposer overloads) and 128-bit keys.  Times whole calls with CUDA events (best of
5, inputs > L2 where the shape allows), checks bit-exactness against the reference in the same run.
    python bench/extras.py [--out gpurun_out/extras.jsonl] [--log2n 26]
Development / evidence tool, not a bench line.  The reference (oracle/_ref/libref_cub.so) is the comparator only."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cub_b200 import _lib  # noqa: E402
from oracle import pyoracle as po  # noqa: E402  (comparator only)
from tests import harness as H  # noqa: E402
from tests.test_struct_gpu import CUSTOM, CUSTOM_FIELDS, _fields  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "extras.jsonl"))
ap.add_argument("--log2n", type=int, default=26)
a = ap.parse_args()
b2s = _lib.load()
ref = po.load_gpu_reference("ref")
out = open(a.out, "a")


def timed(call, iters=5, warm=2):
    best = 1e30
    for it in range(warm + iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        assert call() == 0
        e1.record()
        torch.cuda.synchronize()
        if it >= warm:
            best = min(best, e0.elapsed_time(e1))
    return best


def two_phase(fn, args):
    """size query + allocation; returns a closure that enqueues the sort"""
    nbytes = ctypes.c_size_t(0)
    assert fn(None, ctypes.byref(nbytes), *args, None) == 0
    temp = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device="cuda")
    return (lambda: fn(H._p(temp), ctypes.byref(nbytes), *args, H.stream_handle())), temp


def emit(rec):
    print(json.dumps(rec), flush=True)
    out.write(json.dumps(rec) + "\n")


n = 1 << a.log2n

# ---- segmented sort: u32 keys / u32 values, equal-sized segments of 64 .. 2^20 items (and ragged ones)
keys = H.gen_device_keys(b2s, n, 4, 42)
vals = H.gen_device_iota(b2s, n, 4)
rng = np.random.default_rng(5)
shapes = [("2^%d segments of %d" % (a.log2n - s, 1 << s), np.arange(0, n + 1, 1 << s, dtype=np.int64)) for s in (6, 10, 12, 14, 20)]
cuts = np.sort(rng.integers(0, n + 1, size=n // 1500))
shapes.append(("ragged, mean 1500 items", np.unique(np.concatenate(([0], cuts, [n]))).astype(np.int64)))
for name, offs in shapes:
    nseg = offs.shape[0] - 1
    d_offs = torch.from_numpy(offs.astype(np.int32)).cuda()
    res = {}
    for impl, fn, extra in (("ref_cub_2.2.0", getattr(ref, "segmented_sort", None), ()), ("b2s", b2s.b2s_segmented_radix_sort, (4,))):
        if fn is None:
            continue
        ko, vo = torch.empty_like(keys), torch.empty_like(vals)
        args = (H._p(keys), H._p(ko), H._p(vals), H._p(vo), n, nseg, H._p(d_offs[:-1]), H._p(d_offs[1:])) + extra + (6, 4, 0, 0, 32)
        call, temp = two_phase(fn, args)
        ms = timed(call)
        res[impl] = (ms, ko, vo)
        del temp
    exact = None
    if "ref_cub_2.2.0" in res:
        exact = bool(torch.equal(res["b2s"][1], res["ref_cub_2.2.0"][1]) and torch.equal(res["b2s"][2], res["ref_cub_2.2.0"][2]))
    for impl, (ms, _k, _v) in res.items():
        emit({"row": "DeviceSegmentedRadixSort::SortPairs u32/u32", "shape": name, "n": n, "segments": nseg, "impl": impl, "ms": ms,
              "gkeys_s": n / ms / 1e6, "hbm_gbs_one_read_one_write": n * 16 / ms / 1e6, "bit_exact_vs_ref": exact if impl == "b2s" else None})
    del res
del keys, vals
torch.cuda.empty_cache()

# ---- struct keys: struct { float f; long long lli; } (16 bytes, 96-bit image) with u32 values, all bits and 64 bits
ns = 1 << min(a.log2n, 25)
recs = np.zeros(ns, dtype=CUSTOM)
recs["f"] = rng.integers(0, 1 << 32, size=ns, dtype=np.uint64).astype(np.uint32)
recs["f"][(recs["f"] & 0x7FFFFFFF) == 0] = 1
recs["lli"] = rng.integers(0, 1 << 63, size=ns, dtype=np.uint64)
dk = torch.from_numpy(recs.view(np.uint8).reshape(-1).copy()).cuda()
dv = H.gen_device_iota(b2s, ns, 4)
f = _fields(CUSTOM_FIELDS)
for bb, eb, label in ((0, -1, "all 96 bits"), (32, 96, "bits [32,96)")):
    res = {}
    if hasattr(ref, "struct_sort"):
        ko, vo = torch.zeros_like(dk), torch.zeros_like(dv)
        call, temp = two_phase(ref.struct_sort, (H._p(dk), H._p(ko), H._p(dv), H._p(vo), ns, 4, 0, bb, eb))
        res["ref_cub_2.2.0"] = (timed(call), ko, vo)
    ko, vo = torch.zeros_like(dk), torch.zeros_like(dv)
    call, temp2 = two_phase(b2s.b2s_radix_sort_struct, (H._p(dk), H._p(ko), H._p(dv), H._p(vo), ns, 16, f, 2, 4, 0, bb, eb))
    res["b2s"] = (timed(call), ko, vo)
    exact = None
    if "ref_cub_2.2.0" in res:
        r_out = res["ref_cub_2.2.0"][1].cpu().numpy().view(CUSTOM)
        o_out = res["b2s"][1].cpu().numpy().view(CUSTOM)
        exact = bool(np.array_equal(r_out["f"], o_out["f"]) and np.array_equal(r_out["lli"], o_out["lli"])
                     and torch.equal(res["b2s"][2], res["ref_cub_2.2.0"][2]))
    for impl, (ms, _k, _v) in res.items():
        emit({"row": "SortPairs with a decomposer, struct {float, long long} / u32", "shape": label, "n": ns, "impl": impl, "ms": ms,
              "gkeys_s": ns / ms / 1e6, "bit_exact_vs_ref": exact if impl == "b2s" else None})
    del res
del dk, dv
torch.cuda.empty_cache()

# ---- 128-bit keys with u32 values
raw = np.stack([rng.integers(0, 1 << 63, size=ns, dtype=np.uint64) * np.uint64(2), rng.integers(0, 1 << 63, size=ns, dtype=np.uint64)], axis=1)
dk = torch.from_numpy(raw.view(np.int64).copy()).cuda()
dv = H.gen_device_iota(b2s, ns, 4)
res = {}
if hasattr(ref, "sort128"):
    ko, vo = torch.zeros_like(dk), torch.zeros_like(dv)
    call, temp = two_phase(ref.sort128, (H._p(dk), H._p(ko), H._p(dv), H._p(vo), ns, 0, 4, 0, 0, 128))
    res["ref_cub_2.2.0"] = (timed(call), ko, vo)
ko, vo = torch.zeros_like(dk), torch.zeros_like(dv)
call, temp2 = two_phase(b2s.b2s_radix_sort, (H._p(dk), H._p(ko), H._p(dv), H._p(vo), ns, 16, 4, 4, 0, 0, 128))
res["b2s"] = (timed(call), ko, vo)
exact = None
if "ref_cub_2.2.0" in res:
    exact = bool(torch.equal(res["b2s"][1], res["ref_cub_2.2.0"][1]) and torch.equal(res["b2s"][2], res["ref_cub_2.2.0"][2]))
for impl, (ms, _k, _v) in res.items():
    emit({"row": "SortPairs unsigned __int128 / u32", "shape": "all 128 bits", "n": ns, "impl": impl, "ms": ms, "gkeys_s": ns / ms / 1e6,
          "bit_exact_vs_ref": exact if impl == "b2s" else None})
out.close()
