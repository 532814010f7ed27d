"""Pins the decomposer (custom struct key) restatement in oracle/pyoracle.py -- the checker for SURVEY.md §8(f)2 (product:
cub_b200/csrc/b2s_struct.cu, tests/test_struct_gpu.py) -- against the reference's own known-answer vectors
(test/catch2_test_device_radix_sort_custom.cu:555-1690, extracted by tests/golden/make_decomposer_kats.py), plus two
properties: a one-field decomposition equals the fundamental-type oracle, and the sort is stable."""
import json
import os

import numpy as np

from oracle import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "decomposer_kats.json")))["cases"]


def test_reference_kats():
    assert len(KATS) == 16
    for c in KATS:
        k = np.array(c["keys_in"], dtype=object)
        f = np.array([int(x) for x in k[:, 0]], dtype=np.uint32)
        lli = np.array([int(x) & 0xFFFFFFFFFFFFFFFF for x in k[:, 1]], dtype=np.uint64)
        b, e = c["bits"] if c["bits"] else (0, None)
        perm = po.decomposed_sort_permutation([(f, 8), (lli, 10)], c["descending"], b, e)
        got = [[int(f[i]), int(np.int64(lli[i]))] for i in perm]
        assert got == c["keys_expected"], (c["test"], c["section"], got)
        if c["values_in"] is not None:
            assert [c["values_in"][i] for i in perm] == c["values_expected"], (c["test"], c["section"])


def test_single_field_equals_fundamental_oracle():
    rng = np.random.default_rng(3)
    for kt, dt in ((6, np.uint32), (7, np.uint32), (8, np.uint32), (10, np.uint64), (5, np.uint16), (11, np.uint64)):
        raw = rng.integers(0, np.iinfo(dt).max, size=5000, dtype=dt, endpoint=True)
        if kt in (5, 8, 11):
            raw[::7] = 0
            raw[1::7] = dt(1) << dt(raw.dtype.itemsize * 8 - 1)  # -0.0
        vals = np.arange(raw.shape[0], dtype=np.uint32)
        for desc in (False, True):
            for bb, eb in ((0, None), (3, raw.dtype.itemsize * 8 - 2)):
                # floating keys with a partial range: both restatements follow the reference's onesweep zero handling
                perm = po.decomposed_sort_permutation([(raw, kt)], desc, bb, eb)
                ek, ev = po.radix_sort(raw, vals, kt, desc, bb, eb)
                assert np.array_equal(raw[perm], ek) and np.array_equal(vals[perm], ev), (kt, desc, bb, eb)


def test_lexicographic_and_stable():
    rng = np.random.default_rng(4)
    a = rng.integers(-3, 3, size=2000).astype(np.int16)
    b = rng.integers(0, 4, size=2000).astype(np.uint8)
    perm = po.decomposed_sort_permutation([(a.view(np.uint16), 3), (b, 0)])
    keys = list(zip(a[perm].tolist(), b[perm].tolist()))
    assert keys == sorted(keys)
    ref = sorted(range(2000), key=lambda i: (int(a[i]), int(b[i])))  # Python's sort is stable
    assert perm.tolist() == ref
