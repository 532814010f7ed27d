#!/bin/bash
# compute-sanitizer over the production flows (small inputs): memcheck + racecheck + synccheck
# pair flow (u32/u32), image-form floating keys, wide pairs (u64/u32), keys alone, constant-digit copies, segmented sort
# (all size classes + a multi-tile segment), struct keys.
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import ctypes, numpy as np, torch, cub_b200 as cb
from oracle import pyoracle as po
from cub_b200 import _lib
b2s = _lib.load()
rng = np.random.default_rng(1)
for n in (1, 33, 12288, 12289, 100_003):
    keys = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    vals = np.arange(n, dtype=np.uint32)
    dk = torch.from_numpy(keys.view(np.int32)).cuda().view(torch.uint32)
    dv = torch.from_numpy(vals.view(np.int32)).cuda().view(torch.uint32)
    ko, vo = cb.sort_pairs(dk, dv)
    k2 = cb.sort_keys(dk, descending=True)
    f = cb.sort_keys(dk.view(torch.float32), descending=True)
    fk, fv = cb.sort_pairs(dk.view(torch.float32), dv)
    torch.cuda.synchronize()
    ek, ev = po.radix_sort(keys, vals, 6)
    assert np.array_equal(ko.view(torch.int32).cpu().numpy().view(np.uint32), ek)
    assert np.array_equal(vo.view(torch.int32).cpu().numpy().view(np.uint32), ev)
    ef, _ = po.radix_sort(keys, None, 8, True)
    assert np.array_equal(f.view(torch.int32).cpu().numpy().view(np.uint32), ef)
# constant upper digits: two passes are copies
kc = (rng.integers(0, 1 << 16, size=60_001, dtype=np.uint64).astype(np.uint32) | np.uint32(0x5A5A0000))
dkc = torch.from_numpy(kc.view(np.int32)).cuda().view(torch.uint32)
ko, vo = cb.sort_pairs(dkc, torch.arange(60_001, dtype=torch.int32, device="cuda").view(torch.uint32))
k64 = rng.integers(0, 1 << 63, size=50_001, dtype=np.uint64)
d64 = torch.from_numpy(k64.view(np.int64)).cuda()
v32 = torch.arange(50_001, dtype=torch.int32, device="cuda")
ko, vo = cb.sort_pairs(d64, v32)
ko = cb.sort_keys(d64)
# segmented: empty, tiny, 256 / 1024 / 4096 classes, one multi-tile segment
sizes = [0, 5, 64, 256, 257, 1024, 1025, 4096, 4097, 20_000]
offs = np.concatenate(([0], np.cumsum(sizes))).astype(np.int32)
n = int(offs[-1])
sk = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
dsk = torch.from_numpy(sk.view(np.int32)).cuda()
dsv = torch.arange(n, dtype=torch.int32, device="cuda")
doffs = torch.from_numpy(offs).cuda()
out = cb.DeviceSegmentedRadixSort if hasattr(cb, "DeviceSegmentedRadixSort") else None
ko, vo = torch.empty_like(dsk), torch.empty_like(dsv)
nbytes = ctypes.c_size_t(0)
args = (ctypes.c_void_p(dsk.data_ptr()), ctypes.c_void_p(ko.data_ptr()), ctypes.c_void_p(dsv.data_ptr()), ctypes.c_void_p(vo.data_ptr()),
        n, len(sizes), ctypes.c_void_p(doffs.data_ptr()), ctypes.c_void_p(doffs.data_ptr() + 4), 4, 6, 4, 0, 0, 32)
assert b2s.b2s_segmented_radix_sort(None, ctypes.byref(nbytes), *args, None) == 0
temp = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device="cuda")
assert b2s.b2s_segmented_radix_sort(ctypes.c_void_p(temp.data_ptr()), ctypes.byref(nbytes), *args, None) == 0
torch.cuda.synchronize()
res = ko.cpu().numpy().view(np.uint32)
for b, e in zip(offs[:-1], offs[1:]):
    assert np.array_equal(res[b:e], np.sort(sk[b:e], kind="stable"))
print("sanitizer workload ok")
PY
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  echo "== $tool"
  timeout 1200 compute-sanitizer --tool $tool --print-limit ${LIMIT:-5} env PYTHONPATH=$PWD python /tmp/san.py 2>&1 | grep -v "^$" | tail -${TAIL:-8}
done > gpurun_out/sanitizer_r2.txt 2>&1
cat gpurun_out/sanitizer_r2.txt
