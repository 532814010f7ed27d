"""bench/trace.py -- per-tile phase timeline of one digit pass (tuning build, trace variants).
    B2S_LIB=cub_b200/libb2s_tune.so python bench/trace.py --variants 36,37 [--log2n 28] [--case k4v4] [--pass 1]
Thread 0 of every CTA stamps the SM clock at each phase boundary (MODE bit 4 in b2s_onesweep.cuh); this script turns the
stamps into mean / median / p90 phase durations (us at 1.965 GHz).  Development tool, not a bench line.
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
sys.path.insert(0, os.path.join(ROOT, "bench"))
from tune import CASES, time_sort  # noqa: E402

PHASES = [("key TMA wait", 1, 2), ("rank", 2, 3), ("digit scan P2", 3, 4),
          ("key scatter+val load", 4, 5), ("look-back (digit 0)", 5, 6), ("barrier S3b (all look-backs)", 6, 7),
          ("value scatter+S4", 7, 8), ("write-out issue (thread 0)", 8, 9), ("lifetime tile known->stores issued", 1, 9),
          ("end barrier+claim (persistent)", 9, 10)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--case", default="k4v4")
    ap.add_argument("--variants", default="36")
    ap.add_argument("--pass", dest="pas", type=int, default=1)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "trace.jsonl"))
    a = ap.parse_args()
    b2s = _lib.load()
    n = 1 << a.log2n
    kt, vbytes, rounds = CASES[a.case]
    kbytes = H.KEY_BYTES[kt]
    keys = H.gen_device_keys(b2s, n, kbytes, 42, rounds)
    vals = H.gen_device_iota(b2s, n, vbytes) if vbytes else None
    mhz = 1965.0
    out = open(a.out, "a")
    for v in [int(x) for x in a.variants.split(",")]:
        nt, ipt, minb, match = (ctypes.c_int() for _ in range(4))
        b2s.b2s_describe_variant(kbytes, vbytes, v, ctypes.byref(nt), ctypes.byref(ipt), ctypes.byref(minb), ctypes.byref(match))
        tile = nt.value * ipt.value
        tiles = (n + tile - 1) // tile
        trace = torch.zeros(tiles * 16, dtype=torch.int64, device="cuda")
        b2s.b2s_set_variant(v)
        b2s.b2s_set_trace(ctypes.c_void_p(trace.data_ptr()), a.pas)
        r = time_sort(b2s.b2s_radix_sort_db, keys, vals, kt, iters=3, warm=1)
        b2s.b2s_set_trace(None, 0)
        b2s.b2s_set_variant(0)
        t = trace.view(tiles, 16).cpu().double()
        t = t[: tiles - 1]  # drop the partial last tile
        rec = {"variant": v, "nt": nt.value, "ipt": ipt.value, "minb": minb.value, "mode": b2s.b2s_variant_mode(kbytes, vbytes, v),
               "case": a.case, "n": n, "sort_ms": r[0], "phases_us": {}}
        print(f"variant {v}: {nt.value}x{ipt.value}x{minb.value} mode {rec['mode']}  sort {r[0]:.3f} ms, tiles {tiles}")
        for name, s0, s1 in PHASES:
            if s0 == 9 and (rec["mode"] & 3) == 0:
                continue
            if (s0 == 6 or s1 == 7) and vbytes == 0:
                continue
            d = (t[:, s1] - t[:, s0]) / mhz
            d = d[(t[:, s1] > 0) & (t[:, s0] > 0)]
            if d.numel() == 0:
                continue
            q = torch.quantile(d, torch.tensor([0.1, 0.5, 0.9], dtype=torch.double))
            rec["phases_us"][name] = {"mean": d.mean().item(), "p10": q[0].item(), "p50": q[1].item(), "p90": q[2].item()}
            print(f"  {name:38s} mean {d.mean().item():7.2f}  p10 {q[0].item():7.2f}  p50 {q[1].item():7.2f}  p90 {q[2].item():7.2f} us")
        for name, slot in () if not (rec["mode"] & 512) else (("look-back round trips", 12), ("look-back spin reloads", 13), ("predecessors summed", 14)):
            d = t[1:, slot]
            q = torch.quantile(d, torch.tensor([0.1, 0.5, 0.9], dtype=torch.double))
            print(f"  {name:38s} mean {d.mean().item():7.2f}  p10 {q[0].item():7.2f}  p50 {q[1].item():7.2f}  p90 {q[2].item():7.2f}")
        d = (t[1:, 11] - t[1:, 5]) / mhz if (rec["mode"] & 512) else torch.zeros(1, dtype=torch.double)
        print(f"  first status word after look-back start mean {d.mean().item():.2f}  p10 {torch.quantile(d, 0.1).item():.2f}  p90 {torch.quantile(d, 0.9).item():.2f} us")
        # pass duration and tile start spacing from the global timer
        g = t[:, 0]
        span = (g.max() - g.min()).item() / 1e3
        rec["span_us"] = span
        print(f"  first->last tile start {span:.1f} us; mean spacing {span / tiles * 1e3:.1f} ns")
        # how far behind its predecessor does a tile reach the look-back?  (global timer is only taken at the start, so
        # compare tile starts: negative = started before its predecessor)
        dstart = (g[1:] - g[:-1])
        print(f"  start(t) - start(t-1): p1 {torch.quantile(dstart, 0.01).item():.0f}  p50 {dstart.median().item():.0f}  p99 {torch.quantile(dstart, 0.99).item():.0f} ns")
        out.write(json.dumps(rec) + "\n")
    out.close()


if __name__ == "__main__":
    main()
