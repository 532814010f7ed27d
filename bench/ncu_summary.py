"""bench/ncu_summary.py -- condensed text summary of an .ncu-rep (development + profiles/ evidence).
Usage: python bench/ncu_summary.py <report.ncu-rep> [rows_per_launch_divisor]
Prints the headline metrics, the stall breakdown, and an opcode / region histogram of the SASS."""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
        "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg", "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sector_op_read_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k} = {r[i]} {units[i]}")
    print("stalls (warps per issue-active cycle):")
    for i, k in enumerate(hdr):
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
            v = float(r[i] or 0)
            if v >= 0.05:
                print(f"   {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:22s} {v:6.2f}")
    print("---")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
if len(srows) > 2:
    h = srows[1]
    ix = {k: i for i, k in enumerate(h)}
    body = []
    for r in srows[2:]:
        if len(r) < 10 or r[0] in ("Kernel Name", "Address"):
            break
        body.append(r)
    tot = sum(int(r[ix["# Samples"]]) for r in body) or 1
    ex, sm = defaultdict(int), defaultdict(int)
    for r in body:
        t = r[ix["Source"]].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ex[op] += int(r[ix["Instructions Executed"]])
        sm[op] += int(r[ix["# Samples"]])
    te = sum(ex.values()) or 1
    div = float(sys.argv[2]) if len(sys.argv) > 2 else 8388608.0
    print(f"warp-instructions executed: {te}  ({te / div:.1f} per row of 32 items, divisor {div:.0f})")
    for k, v in sorted(ex.items(), key=lambda x: -x[1])[:22]:
        print(f"   {k:10s} {v / div:7.2f}/row  {100 * v / te:5.1f}% of instr   {100 * sm[k] / tot:5.1f}% of samples")
    print("hot instructions (>=0.8% of samples):")
    for n, r in enumerate(body):
        s = int(r[ix["# Samples"]])
        if s >= 0.008 * tot:
            st = {k[6:]: int(r[ix[k]]) for k in h if k.startswith("stall_") and "Not Issued" not in k and int(r[ix[k]] or 0) > 0.25 * s}
            print(f"   {n:5d} {100 * s / tot:5.2f}%  {r[ix['Source']].strip()[:64]:64s} {st}")
