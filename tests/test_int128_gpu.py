"""GPU parity tests of 128-bit integer keys (B2S_U128 / B2S_I128 through b2s_radix_sort[_db]; cub/util_type.cuh:1225,1259,
test matrix test/test_device_radix_sort.cu:2238): against the UNMODIFIED reference's DeviceRadixSort<__uint128_t / __int128_t>
on the same device buffers, and against the oracle's two-member restatement.  Bar: bit-exact keys and values."""
import ctypes

import numpy as np
import pytest
import torch

from tests import harness as H

pytestmark = pytest.mark.gpu
U128, I128 = 16, 17


def _keys(rng, n):
    lo = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    hi = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    hi[::3] = hi[0]            # equal high words: the low word decides
    hi[1::7] = 0
    hi[2::7] = np.uint64(0xFFFFFFFFFFFFFFFF)
    lo[::5] = lo[min(1, n - 1)]  # fully equal keys: stability decides
    return np.stack([lo, hi], axis=1).copy()  # little endian: low word first


def _sort(fn, dk, dv, n, kt_or_signed, desc, bb, eb, ref):
    ko, vo = torch.zeros_like(dk), (torch.zeros_like(dv) if dv is not None else None)
    nbytes = ctypes.c_size_t(0)
    vb = 4 if dv is not None else 0
    if ref:
        args = (H._p(dk), H._p(ko), H._p(dv), H._p(vo), n, kt_or_signed, vb, int(desc), bb, eb)
    else:
        args = (H._p(dk), H._p(ko), H._p(dv), H._p(vo), n, kt_or_signed, vb, 4, int(desc), bb, eb)
    assert fn(None, ctypes.byref(nbytes), *args, None) == 0
    temp = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device="cuda")
    assert fn(H._p(temp), ctypes.byref(nbytes), *args, H.stream_handle()) == 0
    torch.cuda.synchronize()
    return ko, vo


@pytest.mark.parametrize("signed", [0, 1])
@pytest.mark.parametrize("n", [1, 777, 4865, 250_001])
def test_int128_vs_reference_and_oracle(b2s, refcub, oracle, signed, n):
    rng = np.random.default_rng(n + signed)
    raw = _keys(rng, n)
    vals = np.arange(n, dtype=np.uint32)
    dk = torch.from_numpy(raw.view(np.int64).copy()).cuda()
    dv = H.to_dev(vals)
    kt = I128 if signed else U128
    assert b2s.b2s_key_bytes(kt) == 16
    for desc in (False, True):
        for bb, eb in ((0, 128), (60, 68), (1, 127), (64, 128), (0, 64), (100, 101)):
            for pairs in (True, False):
                k_us, v_us = _sort(b2s.b2s_radix_sort, dk, dv if pairs else None, n, kt, desc, bb, eb, ref=False)
                perm = oracle.decomposed_sort_permutation([(raw[:, 1].copy(), 10 if signed else 9), (raw[:, 0].copy(), 9)], desc, bb, eb)
                assert np.array_equal(k_us.cpu().numpy().view(np.uint64), raw[perm]), f"vs oracle: n={n} signed={signed} desc={desc} [{bb},{eb})"
                if pairs:
                    assert np.array_equal(H.to_np(v_us, np.uint32), vals[perm])
                if hasattr(refcub, "sort128"):
                    k_ref, v_ref = _sort(refcub.sort128, dk, dv if pairs else None, n, signed, desc, bb, eb, ref=True)
                    assert torch.equal(k_us, k_ref), f"vs reference CUB: n={n} signed={signed} desc={desc} [{bb},{eb})"
                    if pairs:
                        assert torch.equal(v_us, v_ref)


def test_int128_double_buffer(b2s, oracle):
    rng = np.random.default_rng(9)
    n = 30_011
    raw = _keys(rng, n)
    for sel0 in (0, 1):
        bufs = [torch.zeros(n, 2, dtype=torch.int64, device="cuda") for _ in range(2)]
        bufs[sel0].copy_(torch.from_numpy(raw.view(np.int64).copy()))
        kb = (ctypes.c_void_p * 2)(bufs[0].data_ptr(), bufs[1].data_ptr())
        ks = ctypes.c_int(sel0)
        nbytes = ctypes.c_size_t(0)
        args = (kb, ctypes.byref(ks), None, None, n, U128, 0, 4, 1, 0, 128)
        assert b2s.b2s_radix_sort_db(None, ctypes.byref(nbytes), *args, None) == 0 and ks.value == sel0
        temp = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
        assert b2s.b2s_radix_sort_db(H._p(temp), ctypes.byref(nbytes), *args, H.stream_handle()) == 0
        torch.cuda.synchronize()
        perm = oracle.decomposed_sort_permutation([(raw[:, 1].copy(), 9), (raw[:, 0].copy(), 9)], True, 0, 128)
        assert np.array_equal(bufs[ks.value].cpu().numpy().view(np.uint64), raw[perm])
