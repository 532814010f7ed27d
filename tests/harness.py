"""Shared helpers for the GPU parity tests: drive the product C-ABI and the reference shim (same C
signatures) on torch device buffers holding RAW key bits (signed containers of the key width)."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

KEY_BYTES = [1, 1, 2, 2, 2, 2, 4, 4, 4, 8, 8, 8]
KEY_NAMES = ["u8", "i8", "u16", "i16", "f16", "bf16", "u32", "i32", "f32", "u64", "i64", "f64"]
CONTAINER = {1: torch.int8, 2: torch.int16, 4: torch.int32, 8: torch.int64}
NP_BITS = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}
NP_SIGNED = {1: np.int8, 2: np.int16, 4: np.int32, 8: np.int64}


def to_dev(a: np.ndarray) -> torch.Tensor:
    """numpy raw-bit array (any fixed width, possibly 2-D for 16-byte values) -> CUDA tensor, bit-exact."""
    a = np.ascontiguousarray(a)
    flat = a.view(NP_SIGNED[a.dtype.itemsize])
    return torch.from_numpy(flat.copy()).cuda()


def to_np(t: torch.Tensor, like: np.dtype) -> np.ndarray:
    return t.cpu().numpy().view(like)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_handle():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def sort_ptr(fn, keys: torch.Tensor, vals: torch.Tensor | None, kt: int, desc=False, bb=0, eb=None, vbytes=None,
             misalign=1, n=None, keys_out=None, vals_out=None):
    """Pointer form through `fn` (b2s_radix_sort or ref_cub_radix_sort). Temp storage is deliberately
    mis-aligned by `misalign` bytes like test/test_device_radix_sort.cu:1109-1110."""
    if eb is None:
        eb = KEY_BYTES[kt] * 8
    if n is None:
        n = keys.numel()
    if vbytes is None:
        vbytes = 0 if vals is None else vals.element_size() * (vals.shape[1] if vals.dim() == 2 else 1)
    if keys_out is None:
        keys_out = torch.empty_like(keys)
    if vals_out is None and vals is not None:
        vals_out = torch.empty_like(vals)
    nbytes = ctypes.c_size_t(0)
    ob = 4 if n < (1 << 32) else 8
    rc = fn(None, ctypes.byref(nbytes), _p(keys), _p(keys_out), _p(vals), _p(vals_out), n, kt, vbytes, ob, int(desc),
            bb, eb, None)
    assert rc == 0, f"size query failed rc={rc}"
    assert nbytes.value >= 1
    temp = torch.empty(nbytes.value + misalign, dtype=torch.uint8, device=keys.device)
    rc = fn(ctypes.c_void_p(temp.data_ptr() + misalign), ctypes.byref(nbytes), _p(keys), _p(keys_out), _p(vals),
            _p(vals_out), n, kt, vbytes, ob, int(desc), bb, eb, stream_handle())
    assert rc == 0, f"sort failed rc={rc}"
    torch.cuda.synchronize()
    return keys_out, vals_out


def sort_db(fn, kbufs, vbufs, kt: int, desc=False, bb=0, eb=None, vbytes=None, selector=0, n=None, temp=None):
    """DoubleBuffer form. kbufs/vbufs: lists of two tensors. Returns (key_selector, val_selector)."""
    if eb is None:
        eb = KEY_BYTES[kt] * 8
    if n is None:
        n = kbufs[0].numel()
    if vbytes is None:
        vbytes = 0 if vbufs is None else vbufs[0].element_size() * (vbufs[0].shape[1] if vbufs[0].dim() == 2 else 1)
    kb = (ctypes.c_void_p * 2)(kbufs[0].data_ptr(), kbufs[1].data_ptr())
    ksel = ctypes.c_int(selector)
    if vbufs is not None:
        vb = (ctypes.c_void_p * 2)(vbufs[0].data_ptr(), vbufs[1].data_ptr())
        vsel = ctypes.c_int(selector)
        vb_a, vsel_a = vb, ctypes.byref(vsel)
    else:
        vsel = ctypes.c_int(0)
        vb_a, vsel_a = None, None
    nbytes = ctypes.c_size_t(0)
    ob = 4 if n < (1 << 32) else 8
    rc = fn(None, ctypes.byref(nbytes), kb, ctypes.byref(ksel), vb_a, vsel_a, n, kt, vbytes, ob, int(desc), bb, eb,
            None)
    assert rc == 0 and nbytes.value >= 1
    assert ksel.value == selector, "size query must not touch the selector"
    if temp is None:
        temp = torch.empty(nbytes.value, dtype=torch.uint8, device=kbufs[0].device)
    rc = fn(ctypes.c_void_p(temp.data_ptr()), ctypes.byref(nbytes), kb, ctypes.byref(ksel), vb_a, vsel_a, n, kt,
            vbytes, ob, int(desc), bb, eb, stream_handle())
    assert rc == 0, f"sort failed rc={rc}"
    torch.cuda.synchronize()
    return ksel.value, vsel.value


def random_bits(rng: np.random.Generator, n: int, nbytes: int) -> np.ndarray:
    return rng.integers(0, 256, size=n * nbytes, dtype=np.uint8).view(NP_BITS[nbytes])


def spice_floats(raw: np.ndarray, nbytes: int) -> np.ndarray:
    """Force +0.0 / -0.0 (1/256 each, as test/test_util.h:576-598 does) plus NaNs, infinities and denormals."""
    raw = raw.copy()
    n = raw.shape[0]
    bits = 8 * nbytes
    high = 1 << (bits - 1)
    exp_mask = {2: 0x7C00, 4: 0x7F800000, 8: 0x7FF0000000000000}[nbytes]
    idx = np.arange(n)
    raw[idx % 256 == 0] = 0
    raw[idx % 256 == 1] = high
    raw[idx % 251 == 7] = exp_mask                      # +inf
    raw[idx % 251 == 8] = exp_mask | high               # -inf
    raw[idx % 241 == 3] = exp_mask | 1                  # signalling NaN
    raw[idx % 241 == 4] = (1 << bits) - 1 if bits < 64 else np.iinfo(np.uint64).max  # -NaN all ones
    raw[idx % 239 == 5] = 1                             # smallest denormal
    raw[idx % 239 == 6] = high | 3                      # negative denormal
    return raw


def gen_device_keys(b2s, n: int, kbytes: int, seed: int, and_rounds: int = 1) -> torch.Tensor:
    t = torch.empty(n, dtype=CONTAINER[kbytes], device="cuda")
    rc = b2s.b2s_fill_keys(ctypes.c_void_p(t.data_ptr()), n, kbytes, seed, and_rounds, 0, stream_handle())
    assert rc == 0
    return t


def gen_device_iota(b2s, n: int, vbytes: int) -> torch.Tensor:
    t = torch.empty(n, dtype=CONTAINER[vbytes], device="cuda")
    rc = b2s.b2s_fill_iota(ctypes.c_void_p(t.data_ptr()), n, vbytes, 0, stream_handle())
    assert rc == 0
    return t


def check_sorted(b2s, keys: torch.Tensor, vals, kt: int, desc=False, bb=0, eb=None):
    """(inversions, key checksum, pair checksum) computed on the device."""
    if eb is None:
        eb = KEY_BYTES[kt] * 8
    res = torch.zeros(3, dtype=torch.int64, device=keys.device)
    vb = 0 if vals is None else vals.element_size()
    rc = b2s.b2s_check_sorted(_p(keys), _p(vals), keys.numel(), kt, vb, int(desc), bb, eb,
                              ctypes.c_void_p(res.data_ptr()), stream_handle())
    assert rc == 0
    torch.cuda.synchronize()
    return [int(x) & 0xFFFFFFFFFFFFFFFF for x in res.cpu().tolist()]
