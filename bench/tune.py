"""bench/tune.py -- sweep the tuning variants of the digit-pass kernel on one B200 and time the
reference CUB builds beside them.  Usage (GPU box):
    B2S_LIB=cub_b200/libb2s_tune.so python bench/tune.py [--log2n 26] [--cases k4v4,k4v0,k8v4] [--out gpurun_out/tune.jsonl]
Times the DoubleBuffer form (no alternate-buffer allocation) with CUDA events, inputs restored outside
the timed region, inputs >> L2.  Not a bench line: development tool.
"""
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

CASES = {"k4v4": (6, 4, 1), "k4v0": (6, 0, 1), "k8v4": (9, 4, 1), "k8v0": (9, 0, 1), "k4v8": (6, 8, 1),
         "k2v0": (2, 0, 1), "k8v8": (9, 8, 1), "k8v4and3": (9, 4, 3), "k1v0": (0, 0, 1), "k2v4": (2, 4, 1)}


def time_sort(fn_db, keys, vals, kt, iters=5, warm=2, bb=0, eb=None):
    n = keys.numel()
    kb = [torch.empty_like(keys), torch.empty_like(keys)]
    vb = [torch.empty_like(vals), torch.empty_like(vals)] if vals is not None else None
    eb = H.KEY_BYTES[kt] * 8 if eb is None else eb
    vbytes = vals.element_size() if vals is not None else 0
    kbp = (ctypes.c_void_p * 2)(kb[0].data_ptr(), kb[1].data_ptr())
    vbp = (ctypes.c_void_p * 2)(vb[0].data_ptr(), vb[1].data_ptr()) if vb else None
    ksel, vsel = ctypes.c_int(0), ctypes.c_int(0)
    nbytes = ctypes.c_size_t(0)
    rc = fn_db(None, ctypes.byref(nbytes), kbp, ctypes.byref(ksel), vbp, ctypes.byref(vsel) if vb else None, n, kt,
               vbytes, 4, 0, bb, eb, None)
    if rc != 0:
        return None
    temp = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    times = []
    for it in range(warm + iters):
        kb[0].copy_(keys)
        if vb:
            vb[0].copy_(vals)
        ksel.value = 0
        vsel.value = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn_db(ctypes.c_void_p(temp.data_ptr()), ctypes.byref(nbytes), kbp, ctypes.byref(ksel), vbp,
                   ctypes.byref(vsel) if vb else None, n, kt, vbytes, 4, 0, bb, eb, H.stream_handle())
        e1.record()
        torch.cuda.synchronize()
        if rc != 0:
            return None
        if it >= warm:
            times.append(e0.elapsed_time(e1))
    times.sort()
    return times[0], times[len(times) // 2], kb[ksel.value], (vb[vsel.value] if vb else None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=26)
    ap.add_argument("--cases", default="k4v4,k4v0,k8v4")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "tune.jsonl"))
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--variants", default="", help="comma-separated variant ids to time (default: all)")
    ap.add_argument("--ablate", action="store_true", help="also time the wrong-by-design ablation variants (k4 only)")
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    b2s = _lib.load()
    from oracle import pyoracle

    refs = {"ref_cub_2.2.0": pyoracle.load_gpu_reference("ref"), "toolkit_cub": pyoracle.load_gpu_reference("tk")}
    n = 1 << a.log2n
    out = open(a.out, "a")
    for case in a.cases.split(","):
        kt, vbytes, rounds = CASES[case]
        kbytes = H.KEY_BYTES[kt]
        keys = H.gen_device_keys(b2s, n, kbytes, 42, rounds)
        vals = H.gen_device_iota(b2s, n, vbytes) if vbytes else None
        passes = kbytes
        algo_bytes = n * (kbytes + passes * 2 * (kbytes + vbytes))
        golden = None
        for name, lib in refs.items():
            if lib is None:
                continue
            r = time_sort(lib.sort_db, keys, vals, kt, a.iters)
            if r is None:
                continue
            best, med, ko, vo = r
            golden = (ko.clone(), vo.clone() if vo is not None else None)
            rec = {"case": case, "n": n, "impl": name, "best_ms": best, "median_ms": med,
                   "gkeys_s": n / best / 1e6, "algo_gbs": algo_bytes / best / 1e6}
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n")
        nv = b2s.b2s_describe_variant(kbytes, vbytes, 0, None, None, None, None)
        only = [int(x) for x in a.variants.split(",") if x != ""]
        for v in range(nv):
            if only and v not in only:
                continue
            nt, ipt, minb, match = (ctypes.c_int() for _ in range(4))
            b2s.b2s_describe_variant(kbytes, vbytes, v, ctypes.byref(nt), ctypes.byref(ipt), ctypes.byref(minb),
                                     ctypes.byref(match))
            if (match.value >> 16) and not a.ablate:
                continue  # timing-only ablations write garbage; never run them in a normal sweep
            b2s.b2s_set_variant(v)
            try:
                r = time_sort(b2s.b2s_radix_sort_db, keys, vals, kt, a.iters)
            except Exception as e:  # noqa: BLE001
                print("variant", v, "failed:", e, flush=True)
                r = None
            if r is None:
                continue
            best, med, ko, vo = r
            ok = None
            if golden is not None:
                ok = bool(torch.equal(ko, golden[0]) and (vo is None or torch.equal(vo, golden[1])))
            rec = {"case": case, "n": n, "impl": "b2s", "variant": v, "nt": nt.value, "ipt": ipt.value,
                   "minb": minb.value, "match": match.value, "mode": b2s.b2s_variant_mode(kbytes, vbytes, v) & 3,
                   "flow": b2s.b2s_variant_flow(kbytes, vbytes, v), "pfd": b2s.b2s_variant_mode(kbytes, vbytes, v) >> 16, "modeflags": b2s.b2s_variant_mode(kbytes, vbytes, v) & 65535, "best_ms": best, "median_ms": med,
                   "gkeys_s": n / best / 1e6, "algo_gbs": algo_bytes / best / 1e6, "bit_exact_vs_ref": ok}
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n")
        b2s.b2s_set_variant(0)
        del keys, vals, golden
        torch.cuda.empty_cache()
    out.close()


if __name__ == "__main__":
    main()
