#!/bin/bash
# compute-sanitizer over the production digit-pass flow (small inputs): memcheck + racecheck + synccheck
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import numpy as np, torch, cub_b200 as cb
from oracle import pyoracle as po
rng = np.random.default_rng(1)
for n in (1, 33, 7680, 7681, 100_003):
    keys = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    vals = np.arange(n, dtype=np.uint32)
    dk = torch.from_numpy(keys.view(np.int32)).cuda().view(torch.uint32)
    dv = torch.from_numpy(vals.view(np.int32)).cuda().view(torch.uint32)
    ko, vo = cb.sort_pairs(dk, dv)
    k2 = cb.sort_keys(dk, descending=True)
    torch.cuda.synchronize()
    ek, ev = po.radix_sort(keys, vals, 6)
    assert np.array_equal(ko.view(torch.int32).cpu().numpy().view(np.uint32), ek)
    assert np.array_equal(vo.view(torch.int32).cpu().numpy().view(np.uint32), ev)
k64 = rng.integers(0, 1 << 63, size=50_001, dtype=np.uint64)
d64 = torch.from_numpy(k64.view(np.int64)).cuda()
v32 = torch.arange(50_001, dtype=torch.int32, device="cuda")
ko, vo = cb.sort_pairs(d64, v32)
torch.cuda.synchronize()
print("sanitizer workload ok")
PY
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --kernel-name kns=onesweep --print-limit ${LIMIT:-5} env PYTHONPATH=$PWD python /tmp/san.py 2>&1 | grep -v "^$" | tail -${TAIL:-6}
done > gpurun_out/sanitizer_r1.txt 2>&1
cat gpurun_out/sanitizer_r1.txt
