"""bench/configs.py -- BASELINE.json configs[0..3] at full size on one B200: ours vs the unmodified reference CUB 2.2.0
(and toolkit CUB), device-resident, CUDA events, pointer form unless memory forces the DoubleBuffer form.  Prints one JSON
line per (config, impl) with GKeys/s and the fraction of the HBM roofline (SURVEY.md §8d bytes/key).  Not the bench line.
    python bench/configs.py [--out gpurun_out/configs.jsonl]"""
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
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.jsonl"))
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
b2s = _lib.load()
from oracle import pyoracle as po  # noqa: E402  (checker/comparator only)

impls = {"b2s": (b2s.b2s_radix_sort, b2s.b2s_radix_sort_db)}
tuned = None
for name, which in (("ref_cub_2.2.0", "ref"), ("toolkit_cub", "tk")):
    lib = po.load_gpu_reference(which)
    if lib is not None:
        impls[name] = (lib.sort, lib.sort_db)
        if which == "ref" and hasattr(lib, "tuned_sort"):
            tuned = lib.tuned_sort

# "best-known CUB on B200": the reference's dispatch with NVIDIA's B200 tuning points injected through SelectedPolicy
# (oracle/ref_shim_ext.cu, dispatch_radix_sort.cuh:1173): (key type, value bytes) -> shim selector
TUNED = {(6, 4): 0, (9, 4): 1, (6, 0): 2, (8, 0): 3}


def tuned_fn(which):
    def fn(tmp, nbytes, kin, kout, vin, vout, n, kt, vb, ob, desc, bb, eb, stream):
        return tuned(tmp, nbytes, kin, kout, vin, vout, n, which, bb, eb, stream)
    return fn
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    peak = 6650.0

# (label, key type, value bytes, log2 n, AND rounds, descending, begin_bit, end_bit, float spice, form)
CONFIGS = [
    ("c1 SortKeys u32 2^24 uniform", 6, 0, 24, 1, False, 0, 32, False, "ptr"),
    ("c2 SortPairs u32/u32 2^28 uniform", 6, 4, 28, 1, False, 0, 32, False, "ptr"),
    ("c2e SortPairs u32/u32 2^28 AND-of-3", 6, 4, 28, 3, False, 0, 32, False, "ptr"),
    ("c3s SortPairs u64/u32 2^28 uniform (tuned-CUB comparator size)", 9, 4, 28, 1, False, 0, 64, False, "ptr"),
    ("c3 SortPairs u64/u32 2^30 AND-of-3 bits[1,63)", 9, 4, 30, 3, False, 1, 63, False, "db"),
    ("c3b SortPairs u64/u32 2^30 AND-of-3 bits[24,56)", 9, 4, 30, 3, False, 24, 56, False, "db"),
    ("c4a SortKeys f32 2^28 ascending (tuned-CUB comparator)", 8, 0, 28, 1, False, 0, 32, True, "ptr"),
    ("c4 SortKeysDescending f32 2^29 (NaN, +-0, denormals)", 8, 0, 29, 1, True, 0, 32, True, "ptr"),
    ("c4b SortKeysDescending bf16 2^29", 5, 0, 29, 1, True, 0, 16, True, "ptr"),
]


def time_one(fns, keys, vals, kt, desc, bb, eb, form):
    n = keys.numel()
    vb = vals.element_size() if vals is not None else 0
    nbytes = ctypes.c_size_t(0)
    best = 1e30
    if form == "ptr":
        ko = torch.empty_like(keys)
        vo = torch.empty_like(vals) if vals is not None else None
        args = (H._p(keys), H._p(ko), H._p(vals), H._p(vo), n, kt, vb, 4, int(desc), bb, eb)
        assert fns[0](None, ctypes.byref(nbytes), *args, None) == 0
        temp = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
        for it in range(a.iters + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            assert fns[0](ctypes.c_void_p(temp.data_ptr()), ctypes.byref(nbytes), *args, H.stream_handle()) == 0
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                best = min(best, e0.elapsed_time(e1))
        return best, ko, vo
    kb = [torch.empty_like(keys), torch.empty_like(keys)]
    vbuf = [torch.empty_like(vals), torch.empty_like(vals)] if vals is not None else None
    kbp = (ctypes.c_void_p * 2)(kb[0].data_ptr(), kb[1].data_ptr())
    vbp = (ctypes.c_void_p * 2)(vbuf[0].data_ptr(), vbuf[1].data_ptr()) if vbuf else None
    ksel, vsel = ctypes.c_int(0), ctypes.c_int(0)
    rest = (n, kt, vb, 4, int(desc), bb, eb)
    assert fns[1](None, ctypes.byref(nbytes), kbp, ctypes.byref(ksel), vbp, ctypes.byref(vsel) if vbuf else None, *rest, None) == 0
    temp = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    for it in range(a.iters + 2):
        kb[0].copy_(keys)
        if vbuf:
            vbuf[0].copy_(vals)
        ksel.value = vsel.value = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        assert fns[1](ctypes.c_void_p(temp.data_ptr()), ctypes.byref(nbytes), kbp, ctypes.byref(ksel), vbp,
                      ctypes.byref(vsel) if vbuf else None, *rest, H.stream_handle()) == 0
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            best = min(best, e0.elapsed_time(e1))
    return best, kb[ksel.value], (vbuf[vsel.value] if vbuf else None)


os.makedirs(os.path.dirname(a.out), exist_ok=True)
with open(a.out, "a") as out:
    for label, kt, vb, lg, rounds, desc, bb, eb, spice, form in CONFIGS:
        n = 1 << lg
        kbytes = H.KEY_BYTES[kt]
        keys = H.gen_device_keys(b2s, n, kbytes, 42, rounds)
        if spice:
            idx = torch.arange(n, device="cuda")
            keys[idx % 256 == 0] = 0
            keys[idx % 256 == 1] = torch.iinfo(H.CONTAINER[kbytes]).min
            del idx
        vals = H.gen_device_iota(b2s, n, vb) if vb else None
        passes = (eb - bb + 7) // 8
        bytes_per_key = kbytes + passes * 2 * (kbytes + vb)
        golden = None
        run = dict(impls)
        if tuned is not None and (kt, vb) in TUNED and not desc and form == "ptr" and n < (1 << 31):
            run["ref_cub_2.2.0_b200_tuned"] = (tuned_fn(TUNED[(kt, vb)]), None)
        for name in ("ref_cub_2.2.0", "ref_cub_2.2.0_b200_tuned", "toolkit_cub", "b2s"):
            if name not in run:
                continue
            ms, ko, vo = time_one(run[name], keys, vals, kt, desc, bb, eb, form)
            exact = None
            if name == "ref_cub_2.2.0":
                golden = (ko.clone(), vo.clone() if vo is not None else None)
            elif name == "b2s" and golden is not None:
                exact = bool(torch.equal(ko, golden[0]) and (vo is None or torch.equal(vo, golden[1])))
            bpk, note = bytes_per_key, None
            if name == "b2s" and vb == 0 and kbytes <= 2 and bb == 0 and eb == 8 * kbytes:
                # counting path (b2s_narrow.cu): histogram read + expansion write; floating keys with both zeros present: one
                # more read of the keys and one bit per key written and read
                bpk = 2 * kbytes + ((kbytes + 0.25) if (spice and kbytes == 2) else 0)
                note = "counting path: its own algorithmic bytes/key; the LSD figure of the other impls is %d" % bytes_per_key
            rec = {"config": label, "impl": name, "n": n, "form": form, "ms": ms, "gkeys_s": n / ms / 1e6,
                   "bytes_per_key": bpk, "algo_gbs": n * bpk / ms / 1e6,
                   "hbm_roofline_frac": n * bpk / ms / 1e6 / peak, "bit_exact_vs_ref": exact}
            if note:
                rec["note"] = note
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n")
            del ko, vo
        del keys, vals, golden
        torch.cuda.empty_cache()
