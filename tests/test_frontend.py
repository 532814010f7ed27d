"""Front ends of SURVEY.md §8(f)4: Thrust-style in-place sort / sort_by_key and the torch custom operators.
CPU part: operator registration, fake-tensor kernels (what torch.compile traces), argument errors -- no compute.
GPU part: results against the oracle and against torch.sort(stable=True) on integer keys (bit-exact)."""
import numpy as np
import pytest
import torch

import cub_b200 as cb
from cub_b200 import frontend  # noqa: F401  (registers torch.ops.cub_b200.*)


def test_ops_registered_and_fake_kernels_trace():
    from torch._subclasses.fake_tensor import FakeTensorMode

    assert hasattr(torch.ops.cub_b200, "sort_pairs") and hasattr(torch.ops.cub_b200, "sort_keys")
    with FakeTensorMode():
        k = torch.empty(1000, dtype=torch.int32, device="cuda")
        v = torch.empty(1000, dtype=torch.int64, device="cuda")
        ko, vo = torch.ops.cub_b200.sort_pairs(k, v, True, 0, -1)
        assert ko.shape == k.shape and ko.dtype == k.dtype and vo.dtype == v.dtype and vo.device == v.device
        ks = torch.ops.cub_b200.sort_keys(k)
        assert ks.shape == k.shape


def test_no_cpu_path():
    k = torch.arange(10, dtype=torch.int32)
    with pytest.raises(ValueError):
        cb.sort(k)
    with pytest.raises(ValueError):
        cb.sort_by_key(k, k.clone())
    with pytest.raises(ValueError):
        torch.ops.cub_b200.sort_keys(k)
    with pytest.raises(ValueError):
        cb.sort_with_indices(k)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,vdtype", [(torch.int32, torch.int32), (torch.int64, torch.float32), (torch.int16, torch.int64),
                                          (torch.uint8, torch.int16), (torch.float32, torch.int32)])
@pytest.mark.parametrize("descending", [False, True])
def test_sort_by_key_in_place(dtype, vdtype, descending):
    from oracle import pyoracle as po

    n = 100_003
    g = torch.Generator(device="cuda").manual_seed(7)
    bits = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), device="cuda", generator=g, dtype=torch.int64)
    if dtype.is_floating_point:
        keys = bits.to(torch.int32).view(torch.float32).clone()
    else:
        keys = bits.to(dtype)
    vals = torch.arange(n, device="cuda").to(vdtype)
    k0, kp, vp = keys.clone(), keys.data_ptr(), vals.data_ptr()
    cb.sort_by_key(keys, vals, descending=descending)
    torch.cuda.synchronize()
    assert keys.data_ptr() == kp and vals.data_ptr() == vp  # in place
    kt = cb.key_type_of(dtype)
    raw = k0.cpu().numpy().view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[k0.element_size()])
    ek, ev = po.radix_sort(raw, np.arange(n, dtype=np.uint32), kt, descending)
    assert np.array_equal(keys.cpu().numpy().view(raw.dtype), ek)
    # the values were arange(n) converted to the value dtype (wraps for int16): convert the oracle's indices the same way
    assert torch.equal(vals.cpu(), torch.from_numpy(ev.astype(np.int64)).to(vdtype))


@pytest.mark.gpu
def test_sort_keys_in_place_and_stream():
    n = 70_001
    keys = torch.randint(0, 1 << 30, (n,), device="cuda", dtype=torch.int32)
    expect = torch.sort(keys).values
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    cb.stable_sort(keys, stream=s)
    s.synchronize()
    assert torch.equal(keys, expect)
    e = torch.empty(0, dtype=torch.int32, device="cuda")
    cb.sort(e)  # empty input: nothing launched


@pytest.mark.gpu
@pytest.mark.parametrize("descending", [False, True])
def test_sort_with_indices_matches_torch_stable_sort(descending):
    x = torch.randint(0, 1000, (200_000,), device="cuda", dtype=torch.int32)  # many duplicates: stability matters
    k, i = cb.sort_with_indices(x, descending=descending)
    rk, ri = torch.sort(x, descending=descending, stable=True)
    assert torch.equal(k, rk) and torch.equal(i, ri) and i.dtype == torch.int64


@pytest.mark.gpu
def test_custom_ops_are_functional():
    k = torch.randint(-1000, 1000, (50_000,), device="cuda", dtype=torch.int64)
    v = torch.arange(50_000, device="cuda", dtype=torch.int16 if False else torch.int32)
    k0 = k.clone()
    ko, vo = torch.ops.cub_b200.sort_pairs(k, v, False, 0, -1)
    assert torch.equal(k, k0)  # pointer form: the input is never written
    rk, ri = torch.sort(k, stable=True)
    assert torch.equal(ko, rk) and torch.equal(vo.to(torch.int64), ri)
    assert torch.equal(torch.ops.cub_b200.sort_keys(k, True, 0, -1), torch.sort(k, descending=True).values)
    torch.library.opcheck(torch.ops.cub_b200.sort_keys.default, (k,), test_utils=("test_schema", "test_faketensor"))
