"""bench/cmp_sass.py -- are the kernels of two builds the same machine code?  Hashes the SASS instruction stream of every
function in k1/k2/k4/k8.o of a reference build (default /tmp/val, e.g. a `git worktree` of the last GPU-validated commit, built
with `make -C cub_b200/csrc all`) and of the current build, after normalising the MODE template argument.  Used at the end of
round 1, when experimental variants were added without GPU time left: 182 of 182 validated kernels unchanged."""
import subprocess, re, hashlib, sys
def funcs(obj):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    out = {}
    name = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            # normalise the MODE template argument (prefetch distance moved from bits 12+ to 16+)
            def norm(mm):
                v = int(mm.group(1))
                return "ELi%dEEEvNS_14OnesweepParams" % v
            name = re.sub(r"ELi(\d+)EEEvNS_14OnesweepParams", lambda mm: "ELi%dEEEvNS_14OnesweepParams" % ((int(mm.group(1)) & 4095) | ((int(mm.group(1)) >> 12) << 16) if int(mm.group(1)) < (1 << 22) and (int(mm.group(1)) >> 12) in (222,) else int(mm.group(1))), name)
            out[name] = hashlib.md5()
            continue
        if name and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
            out[name].update(re.sub(r"/\*[0-9a-f]+\*/", "", line).strip().encode())
    return {k: v.hexdigest() for k, v in out.items()}
same = diff = missing = 0
for k in ("k1", "k2", "k4", "k8"):
    a = funcs(f"/tmp/val/cub_b200/csrc/build/{k}.o")
    b = funcs(f"/root/repo/cub_b200/csrc/build/{k}.o")
    for name, h in a.items():
        if name not in b:
            missing += 1
            print("missing in current:", name[:150])
        elif b[name] != h:
            diff += 1
            print("DIFFERENT:", name[:170])
        else:
            same += 1
    extra = [n for n in b if n not in a]
    print(k, "functions:", len(a), "current:", len(b), "new in current:", len(extra))
    for n in extra[:4]:
        print("   new:", n[:150])
print("same", same, "different", diff, "missing", missing)
