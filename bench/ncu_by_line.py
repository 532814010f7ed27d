"""bench/ncu_by_line.py -- per-source-line view of an ncu report taken with --import-source on: stall samples, warp
instructions and shared-memory wavefronts per row of 32 items, for the lines that matter.
    python bench/ncu_by_line.py <report.ncu-rep> <rows_per_launch> [min_pct]"""
import csv
import io
import subprocess
import sys

rep, rows_per_launch = sys.argv[1], float(sys.argv[2])
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
cur_file, hdr, lines = None, None, []
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] in ("", "Function Name") or not r[0].isdigit():
        continue
    d = dict(zip(hdr[4:], r[4:]))  # metric columns follow the two (line, source) + (address, sass) pairs

    def num(k):
        try:
            return float(d.get(k, "0").replace(",", ""))
        except ValueError:
            return 0.0

    lines.append((cur_file, int(r[0]), r[1].strip(), num("# Samples"), num("Instructions Executed"),
                  num("L1 Wavefronts Shared"), num("L1 Wavefronts Shared Ideal"), num("L2 Theoretical Sectors Global")))
tot_s = sum(x[3] for x in lines) or 1
tot_i = sum(x[4] for x in lines)
tot_w = sum(x[5] for x in lines)
print(f"{rep}: {tot_i / rows_per_launch:.1f} warp instructions, {tot_w / rows_per_launch:.1f} shared-memory wavefronts per row of 32 items")
print(f"{'file:line':26s} {'samples':>8s} {'instr/row':>9s} {'smem wf/row':>11s} {'ideal':>6s}  source")
for f, ln, src, s, i, w, wi, g in lines:
    if 100 * s / tot_s >= min_pct or w / rows_per_launch >= 0.3:
        print(f"{f + ':' + str(ln):26s} {100 * s / tot_s:7.2f}% {i / rows_per_launch:9.2f} {w / rows_per_launch:11.2f} {wi / rows_per_launch:6.2f}  {src[:110]}")
