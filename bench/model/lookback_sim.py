"""Discrete-event model of the decoupled look-back chain of the digit pass (development aid, CPU only).

Tiles reach their look-back every `delta` ns on average (plus jitter); every status read costs one L2 round trip `L`.
Compares the classic windowed walk with "push-forward" variants, where a tile that has its inclusive prefix also resolves
the next K tiles whose partial counts are already published.  Parameters default to what bench/trace.py measured on B200
for u32/u32 pairs, 256x20 tiles, 4 CTAs per SM (profiles/r1_trace_*.txt).
"""
import argparse
import heapq
import random


def simulate(n=6000, delta=24.0, sigma=800.0, L=450.0, proc=40.0, lbw=4, gap=1300.0, push=0, push_delay=0.0, own=True, seed=1):
    rnd = random.Random(seed)
    arrive = [t * delta + rnd.gauss(0, sigma) for t in range(n)]
    base = -min(arrive)
    arrive = [a + base + 5000 for a in arrive]
    partial = [a - gap for a in arrive]
    INF = float("inf")
    incl = [INF] * n          # time the inclusive word becomes visible in L2
    incl[0] = partial[0]
    done = [None] * n
    trips = [0] * n
    summed = [0] * n
    ev = []                   # (time, seq, kind, tile, pos)
    seq = 0
    for t in range(1, n):
        heapq.heappush(ev, (arrive[t], seq, "issue", t, t - 1)); seq += 1
    done[0] = arrive[0]
    if push:
        heapq.heappush(ev, (arrive[0] + push_delay, seq, "push", 0, 0)); seq += 1
    while ev:
        now, _, kind, t, pos = heapq.heappop(ev)
        if kind == "issue":
            if done[t] is not None:
                continue
            trips[t] += 1
            sample = now + L / 2
            heapq.heappush(ev, (sample, seq, "sample", t, pos)); seq += 1
        elif kind == "sample":
            if done[t] is not None:
                continue
            # own word pushed by a predecessor?
            if own and incl[t] <= now:
                done[t] = now + L / 2 + proc
                if push:
                    heapq.heappush(ev, (done[t] + push_delay, seq, "push", t, 0)); seq += 1
                continue
            p = pos
            finished = False
            cost = proc
            k = 0
            while k < lbw and p >= 0:
                if incl[p] <= now:
                    summed[t] += 1
                    finished = True
                    break
                if partial[p] <= now:
                    summed[t] += 1
                    p -= 1
                    k += 1
                    cost += proc / 4
                    continue
                break  # not published yet: re-poll this entry
            ret = now + L / 2 + cost
            if finished:
                done[t] = ret
                incl[t] = min(incl[t], ret + L / 2)
                if push:
                    heapq.heappush(ev, (ret + push_delay, seq, "push", t, 0)); seq += 1
            else:
                heapq.heappush(ev, (ret, seq, "issue", t, p)); seq += 1
        elif kind == "push":
            # read the next `push` status words (one round trip), resolve the leading run of published partials
            sample = now + L / 2
            heapq.heappush(ev, (sample, seq, "pushsample", t, 0)); seq += 1
        elif kind == "pushsample":
            vis = now + L / 2 + proc + L / 2
            for j in range(1, push + 1):
                q = t + j
                if q >= n:
                    break
                if incl[q] <= now:
                    continue
                if partial[q] <= now:
                    incl[q] = min(incl[q], vis)
                else:
                    break
    w = sorted(done[t] - arrive[t] for t in range(n // 4, n) if done[t] is not None)
    m = len(w)
    tr = [trips[t] for t in range(n // 4, n)]
    sm = [summed[t] for t in range(n // 4, n)]
    return {"mean_ns": sum(w) / m, "p10": w[m // 10], "p50": w[m // 2], "p90": w[9 * m // 10],
            "trips": sum(tr) / len(tr), "summed": sum(sm) / len(sm)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--delta", type=float, default=24.0)
    ap.add_argument("--sigma", type=float, default=800.0)
    ap.add_argument("--L", type=float, default=450.0)
    a = ap.parse_args()
    for name, kw in [("classic lbw4", dict(lbw=4)), ("classic lbw8", dict(lbw=8)), ("classic lbw16 (proc x3)", dict(lbw=16, proc=120)),
                     ("push4", dict(lbw=4, push=4)), ("push8", dict(lbw=4, push=8)), ("push16", dict(lbw=4, push=16)),
                     ("push8 delayed 1.4us", dict(lbw=4, push=8, push_delay=1400.0)),
                     ("push8 delayed 0.5us", dict(lbw=4, push=8, push_delay=500.0)),
                     ("push32", dict(lbw=4, push=32))]:
        r = simulate(delta=a.delta, sigma=a.sigma, L=a.L, **kw)
        print(f"{name:28s} look-back mean {r['mean_ns']:7.0f} ns  p10 {r['p10']:6.0f} p50 {r['p50']:6.0f} p90 {r['p90']:6.0f}  trips {r['trips']:.2f} summed {r['summed']:.1f}")
    print("-- efficient window processing (10 ns per entry)")
    for lbw in (4, 8, 16, 32):
        r = simulate(delta=a.delta, sigma=a.sigma, L=a.L, lbw=lbw, proc=40)
        print(f"lbw {lbw:3d} look-back mean {r['mean_ns']:7.0f} ns trips {r['trips']:.2f} summed {r['summed']:.1f}")
    for K in (48, 64, 96):
        r = simulate(delta=a.delta, sigma=a.sigma, L=a.L, lbw=4, push=K)
        print(f"push {K:3d} look-back mean {r['mean_ns']:7.0f} ns trips {r['trips']:.2f} summed {r['summed']:.1f}")
    for lbw, K in ((8, 16), (8, 32), (16, 16), (16, 32)):
        r = simulate(delta=a.delta, sigma=a.sigma, L=a.L, lbw=lbw, push=K)
        print(f"lbw {lbw} push {K:3d} look-back mean {r['mean_ns']:7.0f} ns trips {r['trips']:.2f} summed {r['summed']:.1f}")
