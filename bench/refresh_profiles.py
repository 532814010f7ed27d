"""bench/refresh_profiles.py -- turn the artifacts of one bench/run*.sh GPU call (gpurun_out/*_<tag>.*) into the tracked
evidence under profiles/.  Usage: python bench/refresh_profiles.py r1i
Needs `ncu` (reads the .ncu-rep files) and `cuobjdump` (SASS listings of the production kernels); no GPU."""
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
PFX = sys.argv[2] if len(sys.argv) > 2 else tag[:2]  # file-name prefix under profiles/ (r1, r2, ...)
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def summary(rep, out, divisor):
    # bench/runs/r2l.sh extracts the summaries ON the GPU box (the reports are too large to bring back): use them when the
    # report itself is absent
    pre = {"prof_onesweep_": "prof_pass_pairs_", "prof_onesweep_keys_": "prof_pass_keys_", "prof_hist_": "prof_hist_"}
    base = os.path.basename(rep)[:-len(tag) - len(".ncu-rep")]
    txtfile = os.path.join(G, pre[base] + tag + ".ncu.txt")
    if not os.path.exists(rep) and os.path.exists(txtfile):
        txt = open(txtfile).read()
    else:
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "bench", "ncu_summary.py"), rep, str(divisor)],
                             capture_output=True, text=True).stdout
    open(out, "w").write(txt)
    return txt


def metric(txt, name):
    m = re.search(re.escape(name) + r" = ([0-9.]+) (\S*)", txt)
    v = float(m.group(1))
    u = m.group(2)
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}.get(u, 1)


n = 1 << 28
traffic = {}
t = summary(os.path.join(G, f"prof_onesweep_{tag}.ncu-rep"), os.path.join(P, f"{PFX}_digit_pass_production.ncu.txt"), n // 32)
name = re.search(r"Kernel Name = (.*)", t).group(1).strip()
rd, wr = metric(t, "dram__bytes_read.sum"), metric(t, "dram__bytes_write.sum")
traffic["onesweep_kernel_u32_u32_2p28"] = {
    "dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
    "algorithmic_bytes_per_launch": n * 16,
    "source": f"profiles/{PFX}_digit_pass_production.ncu.txt (ncu --set full --clock-control none, one launch of {name}, n = 2^28)"}
t = summary(os.path.join(G, f"prof_onesweep_keys_{tag}.ncu-rep"), os.path.join(P, f"{PFX}_digit_pass_production_keys_only.ncu.txt"), n // 32)
traffic["onesweep_kernel_u32_keys_2p28"] = {
    "dram_bytes_per_launch": int(metric(t, "dram__bytes_read.sum") + metric(t, "dram__bytes_write.sum")),
    "algorithmic_bytes_per_launch": n * 8, "source": f"profiles/{PFX}_digit_pass_production_keys_only.ncu.txt"}
t = summary(os.path.join(G, f"prof_hist_{tag}.ncu-rep"), os.path.join(P, f"{PFX}_histogram_production.ncu.txt"), n // 32)
traffic["histogram_kernel_u32_2p28"] = {
    "dram_bytes_per_launch": int(metric(t, "dram__bytes_read.sum") + metric(t, "dram__bytes_write.sum")),
    "algorithmic_bytes_per_launch": n * 4, "source": f"profiles/{PFX}_histogram_production.ncu.txt"}
json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)

# launch list of the bench command
rows = list(csv.reader(l for l in open(os.path.join(G, f"launches_{tag}.csv")) if not l.startswith("==")))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    k = re.sub(r"^void (b2s::)?", "", r[ki])
    k = re.sub(r"\(.*$", "", k).replace("b2s::", "").replace("cub_ref::", "").replace("cub::", "")
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, f"{PFX}_launches_bench.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 3   "
            "(per-launch times are cold-cache/serialised: compare shares)\n")
    f.write("The bench command also times reference CUB 2.2.0 (DeviceRadixSortPolicy) and toolkit CUB (policy_hub) on the same input.\n")
    f.write(f"{'kernel':92s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k[:90]:92s} {c:8d} {us:12.1f} {us / c:10.1f} {100 * us / tot:6.1f}%\n")
    ours = {k: v for k, v in agg.items() if k.startswith("digit_pass_kernel") or k.startswith("onesweep_kernel") or k.startswith("histogram_kernel")}
    ot = sum(v[1] for v in ours.values())
    h = sum(v[1] for k, v in ours.items() if k.startswith("histogram"))
    f.write(f"\nshares inside our sort: histogram_kernel {100 * h / ot:.1f}%, digit_pass_kernel {100 * (ot - h) / ot:.1f}%\n")

for src, dst in ((f"bench_{tag}.json", f"{PFX}_bench_n1.json"), (f"bench_ref_{tag}.json", f"{PFX}_bench_reference_arm_n1.json"),
                 (f"configs_{tag}.jsonl", f"{PFX}_configs_vs_reference.jsonl")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))

# SASS listings of the production kernels
def sass(obj, pattern, out):
    names = subprocess.run(["cuobjdump", "-elf", obj], capture_output=True, text=True).stdout
    syms = sorted(set(re.findall(r"\.text\.(\S*" + pattern + r"\S*)", names)))
    assert syms, pattern
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", syms[0], obj], capture_output=True, text=True).stdout
    open(out, "w").write(txt)
    return syms[0]

B = os.path.join(ROOT, "cub_b200", "csrc", "build")
print(sass(os.path.join(B, "k4.o"), r"digit_pass_kernelILi4ELi4ENS_7DigitOpILi4ELb0EEEjLi256ELi48ELi2ELi8ELi17E", os.path.join(P, f"{PFX}_digit_pass_u32_u32.sass")))
print(sass(os.path.join(B, "k4.o"), r"histogram_kernelILi4ELb0EjLb1", os.path.join(P, f"{PFX}_histogram_u32.sass")))
print(open(os.path.join(P, f"{PFX}_launches_bench.txt")).read())

for rep, out in ((f"prof_onesweep_{tag}.ncu-rep", f"{PFX}_digit_pass_production_by_line.txt"),
                 (f"prof_onesweep_keys_{tag}.ncu-rep", f"{PFX}_digit_pass_production_keys_only_by_line.txt")):
    pre = os.path.join(G, rep.replace("prof_onesweep_keys_", "prof_pass_keys_").replace("prof_onesweep_", "prof_pass_pairs_")
                       .replace(".ncu-rep", ".by_line.txt"))
    if not os.path.exists(os.path.join(G, rep)) and os.path.exists(pre):
        txt = open(pre).read()
    else:
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "bench", "ncu_by_line.py"), os.path.join(G, rep), str(n // 32)],
                             capture_output=True, text=True).stdout
    open(os.path.join(P, out), "w").write(txt)
