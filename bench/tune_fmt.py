import sys, json
for l in sys.stdin:
    try:
        r = json.loads(l)
    except Exception:
        print(l.strip())
        continue
    m = r.get("match")
    kind = lbw = None
    if m is not None:
        kind, lbw, m = m >> 16, (m >> 8) & 255, m & 15
    print(r.get("case"), r.get("impl"), "v", r.get("variant"), r.get("nt"), r.get("ipt"), r.get("minb"), "abl", kind, "lbw", lbw, "flow", r.get("flow"),
          round(r["gkeys_s"], 2), "GK/s", round(r["best_ms"], 3), "ms", r.get("bit_exact_vs_ref"))
