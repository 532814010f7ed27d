"""bench/regs.py -- registers / spills of the digit-pass kernels from the ptxas logs (development tool).
Usage: python bench/regs.py [t4|k4|...] [filter substring]"""
import re
import subprocess
import sys

which = sys.argv[1] if len(sys.argv) > 1 else "t4"
flt = sys.argv[2] if len(sys.argv) > 2 else "(int)4, (int)4, (bool)0, unsigned int"
log = open(f"/root/repo/cub_b200/csrc/build/{which}.ptxas.log").read()
ents = re.findall(r"Compiling entry function '([^']+)'.*?\n.*?Function properties.*?\n\s*(\d+) bytes stack frame, "
                  r"(\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", log)
names = subprocess.run(["cu++filt"] + [e[0] for e in ents], capture_output=True, text=True).stdout.splitlines()
for (name, stack, ss, sl, regs), dem in zip(ents, names):
    m = re.search(r"(onesweep\w*_kernel|histogram_kernel|split_count_kernel)<(.*)>\(", dem)
    if m and flt in m.group(2):
        print(f"{m.group(1)}<{m.group(2)}>  regs {regs} spill {ss}/{sl}")
