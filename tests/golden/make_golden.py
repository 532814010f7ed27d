"""Regenerate tests/golden/golden_mt.npz (build container only: needs /root/reference + nvcc).

Compiles gen_golden.cu against the reference's own test headers (test/test_util.h, test/mersenne.h,
test/half.h, test/bfloat16.h) where they lie, runs it, and stores inputs + the harness-style
std::stable_sort solution (ascending and descending source ranks) as a compressed npz.
Usage: python tests/golden/make_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("B2S_REFERENCE", "/root/reference")
BITS = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}


def main():
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not present: fixtures can only be regenerated in the build container")
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "gen_golden")
        subprocess.check_call(["nvcc", "-std=c++17", "-O1", "-w", "-Wno-deprecated-gpu-targets",
                               "-DTHRUST_IGNORE_CUB_VERSION_CHECK", f"-I{REF}", f"-I{REF}/test", "-o", exe,
                               os.path.join(HERE, "gen_golden.cu")])
        blob = subprocess.check_output([exe])
    out, off = {}, 0
    names = []
    while off < len(blob):
        name = blob[off:off + 16].split(b"\0")[0].decode()
        n, kb, bb, eb = np.frombuffer(blob, dtype=np.int32, count=4, offset=off + 16)
        off += 32
        keys = np.frombuffer(blob, dtype=BITS[int(kb)], count=int(n), offset=off).copy()
        off += int(n) * int(kb)
        asc = np.frombuffer(blob, dtype=np.uint32, count=int(n), offset=off).copy()
        off += 4 * int(n)
        desc = np.frombuffer(blob, dtype=np.uint32, count=int(n), offset=off).copy()
        off += 4 * int(n)
        out[name + "__keys"] = keys
        out[name + "__meta"] = np.array([n, kb, bb, eb], dtype=np.int32)
        out[name + "__asc"] = asc
        out[name + "__desc"] = desc
        names.append(name)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "golden_mt.npz"), **out)
    print("wrote", len(names), "records:", ", ".join(names))


if __name__ == "__main__":
    main()
