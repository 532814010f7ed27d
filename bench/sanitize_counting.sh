#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over the counting-sort path (b2s_narrow.cu): every narrow key type, both
# directions, floating zeros present / absent, unaligned pointers, tiny and multi-tile sizes; results checked against the oracle.
mkdir -p gpurun_out
cat > /tmp/san_counting.py <<'PY'
import numpy as np, torch
from oracle import pyoracle as po
from cub_b200 import _lib
from tests import harness as H
b2s = _lib.load()
for kb in (1, 2):
    b2s.b2s_set_counting_min_items(kb, 1)
rng = np.random.default_rng(3)
for kt in range(6):
    nb = H.KEY_BYTES[kt]
    for n in (1, 9, 4097, 70_003, 200_001):
        raw = H.random_bits(rng, n, nb)
        if kt in (4, 5) and n > 1000:
            raw = H.spice_floats(raw, nb)
        for desc in (False, True):
            for off in (0, 1):
                big = H.to_dev(np.concatenate([np.zeros(off, dtype=raw.dtype), raw]))
                out = torch.zeros(n + off, dtype=big.dtype, device="cuda")
                H.sort_ptr(b2s.b2s_radix_sort, big[off:], None, kt, desc, keys_out=out[off:], n=n)
                ek, _ = po.radix_sort(raw, None, kt, desc)
                assert np.array_equal(H.to_np(out[off:], raw.dtype), ek), (kt, n, desc, off)
print("counting-path sanitizer workload ok")
PY
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  echo "== $tool"
  timeout 1200 compute-sanitizer --tool $tool --print-limit ${LIMIT:-5} env PYTHONPATH=$PWD python /tmp/san_counting.py 2>&1 | grep -v "^$" | tail -${TAIL:-8}
done > gpurun_out/sanitizer_counting_r2.txt 2>&1
cat gpurun_out/sanitizer_counting_r2.txt
