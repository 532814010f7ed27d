"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C-ABI, against
 (1) the golden fixtures made from the reference's own generators/harness,
 (2) the CPU oracle on seeded inputs for every key type / value width / direction / bit range / edge size,
 (3) the unmodified reference cub::DeviceRadixSort (oracle/_ref/libref_cub.so) on the same device buffers,
 (4) size-independent properties (sortedness, multiset checksum, permutation) at BASELINE.json's full sizes.
Bar: bit-exact keys AND values (integer/byte work).  Test structure follows test/test_device_radix_sort.cu:
sizes incl. 0/1 (:1649-1670), both back-ends (:1259-1376), mis-aligned temp storage (:1109-1110), pointer form
leaves the input untouched (:1141-1158), bit sub-ranges for unsigned keys (:1461-1494)."""
import ctypes

import numpy as np
import pytest
import torch

from tests import harness as H

pytestmark = pytest.mark.gpu

NAME_TO_TYPE = {n: i for i, n in enumerate(H.KEY_NAMES)}


def _expect(oracle, raw, vals_np, kt, desc, bb, eb):
    return oracle.radix_sort(raw, vals_np, kt, desc, bb, eb)


def _run_and_compare(b2s, oracle, raw, vals_np, kt, desc=False, bb=0, eb=None, label=""):
    dk = H.to_dev(raw)
    dv = H.to_dev(vals_np) if vals_np is not None else None
    before = dk.clone()
    ko, vo = H.sort_ptr(b2s.b2s_radix_sort, dk, dv, kt, desc, bb, eb)
    ek, ev = _expect(oracle, raw, vals_np, kt, desc, bb, eb)
    got_k = H.to_np(ko, raw.dtype)
    assert torch.equal(dk, before), f"{label}: pointer form modified its input"
    if not np.array_equal(got_k, ek):
        bad = np.nonzero(got_k != ek)[0]
        raise AssertionError(f"{label}: keys differ at {bad.size} of {raw.shape[0]} positions, first {bad[:5]}")
    if vals_np is not None:
        got_v = vo.cpu().numpy().view(vals_np.dtype).reshape(vals_np.shape)
        if not np.array_equal(got_v, ev):
            bad = np.nonzero((got_v != ev).reshape(raw.shape[0], -1).any(axis=1))[0]
            raise AssertionError(f"{label}: values differ at {bad.size} positions, first {bad[:5]}")


def test_golden_fixtures(b2s, golden):
    """Inputs from the reference's MT19937 stream; expected = reference harness solution."""
    for name in golden["names"]:
        name = str(name)
        if name == "mt_first8":
            continue
        n, kb, bb, eb = (int(x) for x in golden[name + "__meta"])
        kt = NAME_TO_TYPE[name.split("_")[0]]
        keys = golden[name + "__keys"]
        iota = np.arange(n, dtype=np.uint32)
        for desc, ranks in ((False, golden[name + "__asc"]), (True, golden[name + "__desc"])):
            ko, vo = H.sort_ptr(b2s.b2s_radix_sort, H.to_dev(keys), H.to_dev(iota), kt, desc, bb, eb)
            assert np.array_equal(H.to_np(vo, np.uint32), ranks), f"{name} desc={desc}: ranks"
            assert np.array_equal(H.to_np(ko, keys.dtype), keys[ranks]), f"{name} desc={desc}: keys"


@pytest.mark.parametrize("kt", range(12))
def test_all_key_types_vs_oracle(b2s, oracle, kt):
    rng = np.random.default_rng(100 + kt)
    nb = H.KEY_BYTES[kt]
    for n in (1, 2, 33, 4097, 70001):
        raw = H.random_bits(rng, n, nb)
        if kt in (4, 5, 8, 11):
            raw = H.spice_floats(raw, nb)
        for desc in (False, True):
            _run_and_compare(b2s, oracle, raw, None, kt, desc, label=f"{H.KEY_NAMES[kt]} n={n} keys-only desc={desc}")
            _run_and_compare(b2s, oracle, raw, np.arange(n, dtype=np.uint32), kt, desc,
                             label=f"{H.KEY_NAMES[kt]} n={n} pairs desc={desc}")


@pytest.mark.parametrize("kt", [6, 7, 8, 9, 5, 0])
def test_small_sizes_through_the_multi_kernel_path(b2s, oracle, kt):
    """Sorts of at most one tile take the single-tile kernel by default (covered by every small case in this file);
    with the hook off the same sizes must come out of the histogram + digit-pass kernels bit-exactly too."""
    rng = np.random.default_rng(300 + kt)
    nb = H.KEY_BYTES[kt]
    old = b2s.b2s_set_single_tile(0)
    try:
        for n in (1, 33, 4097, 8192):
            raw = H.random_bits(rng, n, nb)
            if kt in (5, 8):
                raw = H.spice_floats(raw, nb)
            _run_and_compare(b2s, oracle, raw, np.arange(n, dtype=np.uint32), kt, n % 2 == 1, label=f"multi-kernel kt={kt} n={n}")
        assert b2s.b2s_last_launch_count() > 1
    finally:
        b2s.b2s_set_single_tile(old)


def test_single_tile_kernel_shapes(b2s, oracle):
    """One launch for a sort of at most one tile: tile boundaries, bit sub-ranges, wide values, selector parity."""
    rng = np.random.default_rng(77)
    for kt, vb, n in ((6, 4, 8192), (6, 4, 8191), (9, 16, 4096), (9, 8, 8192), (2, 0, 8192), (0, 1, 5000), (11, 4, 8000)):
        raw = H.random_bits(rng, n, H.KEY_BYTES[kt])
        if kt == 11:
            raw = H.spice_floats(raw, 8)
        if vb == 16:
            vals = rng.integers(0, np.iinfo(np.int64).max, size=(n, 2), dtype=np.int64).view(np.uint64)
        elif vb:
            vals = H.random_bits(rng, n, vb)
        else:
            vals = None
        for desc in (False, True):
            _run_and_compare(b2s, oracle, raw, vals, kt, desc, label=f"single tile kt={kt} v={vb} n={n} desc={desc}")
            assert b2s.b2s_last_launch_count() == 1
    raw = H.random_bits(rng, 6000, 4)
    for bb, eb in ((3, 12), (0, 1), (31, 32), (7, 25)):
        _run_and_compare(b2s, oracle, raw, np.arange(6000, dtype=np.uint32), 6, False, bb, eb, label=f"single tile bits=[{bb},{eb})")
    # one item more than the tile: the multi-kernel path
    raw = H.random_bits(rng, 8193, 4)
    _run_and_compare(b2s, oracle, raw, np.arange(8193, dtype=np.uint32), 6, False, label="8193 items")
    assert b2s.b2s_last_launch_count() > 1


@pytest.mark.parametrize("vbytes", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("kt", [6, 9, 2, 0])
def test_value_widths(b2s, oracle, kt, vbytes):
    rng = np.random.default_rng(vbytes * 31 + kt)
    n = 50021
    raw = H.random_bits(rng, n, H.KEY_BYTES[kt])
    if vbytes == 16:
        vals = rng.integers(0, np.iinfo(np.int64).max, size=(n, 2), dtype=np.int64).view(np.uint64)
    else:
        vals = H.random_bits(rng, n, vbytes)
    _run_and_compare(b2s, oracle, raw, vals, kt, False, label=f"kt={kt} v={vbytes}")
    _run_and_compare(b2s, oracle, raw, vals, kt, True, label=f"kt={kt} v={vbytes} desc")


def test_edge_sizes_around_tiles(b2s, oracle):
    rng = np.random.default_rng(5)
    sizes = [0, 1, 31, 32, 255, 256, 257, 3071, 3072, 3073, 4095, 4096, 4097, 8191, 8192, 8193, 12288, 4096 * 5 + 17]
    for n in sizes:
        raw = H.random_bits(rng, n, 4)
        if n == 0:
            ko, vo = H.sort_ptr(b2s.b2s_radix_sort, torch.empty(0, dtype=torch.int32, device="cuda"),
                                torch.empty(0, dtype=torch.int32, device="cuda"), 6)
            assert ko.numel() == 0
            continue
        _run_and_compare(b2s, oracle, raw, np.arange(n, dtype=np.uint32), 6, False, label=f"u32/u32 n={n}")
        _run_and_compare(b2s, oracle, raw & 0x3, None, 6, True, label=f"u32 dup-heavy n={n}")


def test_bit_ranges_unsigned(b2s, oracle):
    rng = np.random.default_rng(11)
    n = 100003
    for kt in (6, 9, 2):
        bits = H.KEY_BYTES[kt] * 8
        raw = H.random_bits(rng, n, H.KEY_BYTES[kt])
        vals = np.arange(n, dtype=np.uint32)
        for bb, eb in ((1, bits - 1), (bits // 2 - 1, bits // 2 + 1), (0, 1), (bits - 1, bits), (3, 12), (5, 5)):
            for desc in (False, True):
                _run_and_compare(b2s, oracle, raw, vals, kt, desc, bb, eb, label=f"kt={kt} bits=[{bb},{eb}) d={desc}")


def test_double_buffer_semantics(b2s, oracle):
    rng = np.random.default_rng(21)
    n = 30011
    for kt, eb in ((6, 32), (6, 24), (9, 64), (9, 40), (2, 16), (0, 8)):
        raw = H.random_bits(rng, n, H.KEY_BYTES[kt])
        vals = np.arange(n, dtype=np.uint32)
        ek, ev = oracle.radix_sort(raw, vals, kt, False, 0, eb)
        for selector in (0, 1):
            kb = [torch.zeros(n, dtype=H.CONTAINER[H.KEY_BYTES[kt]], device="cuda") for _ in range(2)]
            vb = [torch.zeros(n, dtype=torch.int32, device="cuda") for _ in range(2)]
            kb[selector].copy_(H.to_dev(raw))
            vb[selector].copy_(H.to_dev(vals))
            ks, vs = H.sort_db(b2s.b2s_radix_sort_db, kb, vb, kt, False, 0, eb, selector=selector)
            passes = (eb + 7) // 8
            assert ks == vs == selector ^ (passes & 1)
            assert np.array_equal(H.to_np(kb[ks], raw.dtype), ek)
            assert np.array_equal(H.to_np(vb[vs], np.uint32), ev)
        # begin_bit == end_bit: no-op, selector unchanged (dispatch_radix_sort.cuh:1945)
        kb = [H.to_dev(raw), torch.zeros(n, dtype=H.CONTAINER[H.KEY_BYTES[kt]], device="cuda")]
        ks, _ = H.sort_db(b2s.b2s_radix_sort_db, kb, None, kt, False, 4, 4)
        assert ks == 0 and np.array_equal(H.to_np(kb[0], raw.dtype), raw)
    # keys-only DoubleBuffer
    raw = H.random_bits(rng, n, 4)
    kb = [H.to_dev(raw), torch.zeros(n, dtype=torch.int32, device="cuda")]
    ks, _ = H.sort_db(b2s.b2s_radix_sort_db, kb, None, 6, True)
    assert np.array_equal(H.to_np(kb[ks], np.uint32), oracle.radix_sort(raw, None, 6, True)[0])


def test_unaligned_pointers_and_temp(b2s, oracle):
    """Pointers need only element alignment (reference loads are scalar); temp may be at any byte offset."""
    rng = np.random.default_rng(31)
    n = 20000
    for kt, off in ((6, 1), (6, 3), (2, 1), (2, 5), (0, 7), (9, 1)):
        nb = H.KEY_BYTES[kt]
        raw = H.random_bits(rng, n + 8, nb)
        vals = np.arange(n + 8, dtype=np.uint32)
        dk, dv = H.to_dev(raw), H.to_dev(vals)
        ko, vo = H.sort_ptr(b2s.b2s_radix_sort, dk[off:off + n], dv[1:1 + n], kt, misalign=3)
        ek, ev = oracle.radix_sort(raw[off:off + n], vals[1:1 + n], kt)
        assert np.array_equal(H.to_np(ko, raw.dtype), ek), (kt, off)
        assert np.array_equal(H.to_np(vo, np.uint32), ev), (kt, off)


def test_temp_storage_protocol(b2s):
    n = 1 << 20
    nbytes = ctypes.c_size_t(0)
    # trivial problems report 1 byte (never 0) and launch nothing
    assert b2s.b2s_radix_sort(None, ctypes.byref(nbytes), None, None, None, None, 0, 6, 0, 4, 0, 0, 32, None) == 0
    assert nbytes.value == 1
    assert b2s.b2s_radix_sort(None, ctypes.byref(nbytes), None, None, None, None, n, 6, 4, 4, 0, 0, 32, None) == 0
    need = nbytes.value
    assert need > 2 * n * 4  # pointer form carries an alternate key + value buffer
    keys = torch.zeros(n, dtype=torch.int32, device="cuda")
    small = ctypes.c_size_t(need - 1)
    temp = torch.empty(need, dtype=torch.uint8, device="cuda")
    rc = b2s.b2s_radix_sort(ctypes.c_void_p(temp.data_ptr()), ctypes.byref(small), H._p(keys), H._p(keys.clone()),
                            H._p(keys), H._p(keys.clone()), n, 6, 4, 4, 0, 0, 32, H.stream_handle())
    assert rc != 0, "too-small temp storage must be rejected (cudaErrorInvalidValue)"
    assert b2s.b2s_radix_sort(None, ctypes.byref(nbytes), None, None, None, None, n, 99, 0, 4, 0, 0, 32, None) != 0


def test_non_default_stream(b2s, oracle):
    rng = np.random.default_rng(41)
    raw = H.random_bits(rng, 200000, 4)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ko, _ = H.sort_ptr(b2s.b2s_radix_sort, H.to_dev(raw), None, 6)
    assert np.array_equal(H.to_np(ko, np.uint32), oracle.radix_sort(raw, None, 6)[0])
    assert b2s.b2s_last_launch_count() == 1 + 1 + 4  # memset + histogram + 4 digit passes


# ---------------------------------------------------------------------------------------------
# against the unmodified reference CUB on the same device buffers
# ---------------------------------------------------------------------------------------------
REF_CASES = [
    # (key type, value bytes, log2 n, and_rounds, descending, begin_bit, end_bit)
    (6, 0, 24, 1, False, 0, 32),    # BASELINE config 1: SortKeys u32 2^24 uniform
    (6, 4, 22, 1, False, 0, 32),
    (6, 4, 22, 1, True, 0, 32),
    (9, 4, 22, 3, False, 1, 63),    # config 3 shape: u64/u32, AND-of-3, partial bits
    (9, 4, 21, 3, False, 24, 56),
    (9, 4, 21, 3, True, 31, 33),
    (8, 0, 22, 1, True, 0, 32),     # config 4 shape: f32 descending with NaN / +-0 / denormals
    (5, 0, 22, 1, True, 0, 16),     # bf16 descending
    (4, 4, 20, 1, False, 0, 16),
    (11, 8, 20, 1, True, 0, 64),
    (7, 4, 20, 1, False, 0, 32),
    (10, 0, 20, 1, True, 0, 64),
    (0, 4, 20, 1, False, 0, 8),
    (3, 0, 20, 1, True, 0, 16),
    (6, 8, 20, 2, False, 0, 32),
]


@pytest.mark.parametrize("case", REF_CASES, ids=lambda c: f"{H.KEY_NAMES[c[0]]}_v{c[1]}_2p{c[2]}_and{c[3]}_{'desc' if c[4] else 'asc'}_{c[5]}_{c[6]}")
def test_bit_exact_vs_reference_cub(b2s, refcub, case):
    kt, vb, lg, rounds, desc, bb, eb = case
    n = (1 << lg) + 12345
    nb = H.KEY_BYTES[kt]
    keys = H.gen_device_keys(b2s, n, nb, seed=42, and_rounds=rounds)
    if kt in (4, 5, 8, 11):  # uniform bit patterns already contain NaN/inf/denormals; force +-0.0 as test_util.h does
        idx = torch.arange(n, device="cuda")
        keys[idx % 256 == 0] = 0
        keys[idx % 256 == 1] = torch.iinfo(H.CONTAINER[nb]).min
    vals = H.gen_device_iota(b2s, n, vb) if vb else None
    k_ref, v_ref = H.sort_ptr(refcub.sort, keys, vals, kt, desc, bb, eb)
    k_us, v_us = H.sort_ptr(b2s.b2s_radix_sort, keys, vals, kt, desc, bb, eb)
    assert torch.equal(k_us, k_ref), "keys differ from reference CUB"
    if vb:
        assert torch.equal(v_us, v_ref), "values differ from reference CUB"
    # DoubleBuffer form against the same answer
    kb = [keys.clone(), torch.empty_like(keys)]
    vbuf = [vals.clone(), torch.empty_like(vals)] if vb else None
    ks, vs = H.sort_db(b2s.b2s_radix_sort_db, kb, vbuf, kt, desc, bb, eb)
    assert torch.equal(kb[ks], k_ref)
    if vb:
        assert torch.equal(vbuf[vs], v_ref)


# ---------------------------------------------------------------------------------------------
# BASELINE.json full sizes: properties (+ reference CUB where memory/time allow)
# ---------------------------------------------------------------------------------------------
def _property_check(b2s, keys, vals, k_out, v_out, kt, desc, bb, eb):
    inv, ksum, psum = H.check_sorted(b2s, k_out, v_out, kt, desc, bb, eb)
    inv0, ksum0, psum0 = H.check_sorted(b2s, keys, vals, kt, desc, bb, eb)
    assert inv == 0, f"{inv} adjacent inversions in the output"
    assert ksum == ksum0, "key multiset changed"
    if vals is not None:
        assert psum == psum0, "(key,value) pairing changed"


def test_config2_pairs_u32_2p28(b2s, refcub):
    n = 1 << 28
    keys = H.gen_device_keys(b2s, n, 4, seed=42)
    vals = H.gen_device_iota(b2s, n, 4)
    k_us, v_us = H.sort_ptr(b2s.b2s_radix_sort, keys, vals, 6)
    _property_check(b2s, keys, vals, k_us, v_us, 6, False, 0, 32)
    # values are source indices: gathering the input through them must reproduce the output (permutation +
    # pairing), and equal keys must keep increasing indices (stability)
    assert torch.equal(keys[v_us.long()], k_us)
    same = k_us[1:] == k_us[:-1]
    assert bool((v_us[1:][same] > v_us[:-1][same]).all()), "equal keys out of input order: not stable"
    k_ref, v_ref = H.sort_ptr(refcub.sort, keys, vals, 6)
    assert torch.equal(k_us, k_ref) and torch.equal(v_us, v_ref)


def test_config3_pairs_u64_2p30_partial_bits(b2s):
    n = 1 << 30  # exercises the 64-bit look-back words (n >= 2^30)
    keys = H.gen_device_keys(b2s, n, 8, seed=42, and_rounds=3)
    vals = H.gen_device_iota(b2s, n, 4)
    kb = [keys, torch.empty_like(keys)]
    vb = [vals, torch.empty_like(vals)]
    inv0 = H.check_sorted(b2s, keys, vals, 9, False, 1, 63)
    ks, vs = H.sort_db(b2s.b2s_radix_sort_db, kb, vb, 9, False, 1, 63)
    inv = H.check_sorted(b2s, kb[ks], vb[vs], 9, False, 1, 63)
    assert inv[0] == 0 and inv[1] == inv0[1] and inv[2] == inv0[2]
    k_out, v_out = kb[ks], vb[vs]
    # stability on the masked key: equal sort keys keep increasing source index
    mask = ((1 << 62) - 1) << 1
    mk = k_out & mask
    same = mk[1:] == mk[:-1]
    assert bool((v_out[1:][same] > v_out[:-1][same]).all())


@pytest.mark.parametrize("kt", [6, 8], ids=["u32", "f32"])
def test_pair_flow_with_64bit_offsets_2p30(b2s, refcub, kt):
    """4-byte keys with 4-byte values at n = 2^30: the pair flow (one 64-bit shared-memory store per item) with 64-bit
    look-back words and offsets (its own instantiation: fewer items per thread, half the window); for f32 keys also the
    image-form intermediate buffers.  DoubleBuffer form, bit-exact against the reference on the same input."""
    n = 1 << 30
    free, _ = torch.cuda.mem_get_info()
    if free < 48 * (1 << 30):
        pytest.skip("not enough device memory")
    keys = H.gen_device_keys(b2s, n, 4, seed=4242)
    if kt == 8:
        idx = torch.arange(n, device="cuda")
        keys[idx % 1024 == 0] = 0
        keys[idx % 1024 == 1] = torch.iinfo(torch.int32).min  # -0.0
        del idx
    vals = H.gen_device_iota(b2s, n, 4)
    inv0 = H.check_sorted(b2s, keys, vals, kt, True)
    kb = [keys.clone(), torch.empty_like(keys)]
    vb = [vals.clone(), torch.empty_like(vals)]
    ks, vs = H.sort_db(b2s.b2s_radix_sort_db, kb, vb, kt, True)
    inv = H.check_sorted(b2s, kb[ks], vb[vs], kt, True)
    assert inv[0] == 0 and inv[1] == inv0[1] and inv[2] == inv0[2]
    k_us, v_us = kb[ks], vb[vs]
    del kb, vb
    k_ref, v_ref = H.sort_ptr(refcub.sort, keys, vals, kt, True)
    assert torch.equal(k_us, k_ref) and torch.equal(v_us, v_ref)
    del k_us, v_us, k_ref, v_ref
    torch.cuda.empty_cache()


def test_config4_descending_float_2p29(b2s, refcub):
    n = 1 << 29
    for kt, nb in ((8, 4), (5, 2)):
        keys = H.gen_device_keys(b2s, n, nb, seed=7)
        idx = torch.arange(n, device="cuda")
        keys[idx % 256 == 0] = 0
        keys[idx % 256 == 1] = torch.iinfo(H.CONTAINER[nb]).min
        del idx
        k_us, _ = H.sort_ptr(b2s.b2s_radix_sort, keys, None, kt, True)
        _property_check(b2s, keys, None, k_us, None, kt, True, 0, nb * 8)
        k_ref, _ = H.sort_ptr(refcub.sort, keys, None, kt, True)
        assert torch.equal(k_us, k_ref)
        del k_us, k_ref, keys
        torch.cuda.empty_cache()


def test_more_than_2p32_items_u8(b2s):
    """n > 2^32 (the reference tests 4 350 000 007 items, test_device_radix_sort.cu:1694-1717): 64-bit num_items, 64-bit
    look-back words and offsets beyond 32 bits.  u8 keys keep it to one digit pass over 2 x 4.3 GB."""
    n = (1 << 32) + 100_003
    free, _ = torch.cuda.mem_get_info()
    if free < 3 * n + (2 << 30):
        pytest.skip("not enough device memory")
    keys = torch.empty(n, dtype=torch.int8, device="cuda")
    assert b2s.b2s_fill_keys(ctypes.c_void_p(keys.data_ptr()), n, 1, 99, 1, 0, H.stream_handle()) == 0
    out, _ = H.sort_ptr(b2s.b2s_radix_sort, keys, None, 0, n=n)
    inv, ksum, _ = H.check_sorted(b2s, out, None, 0)
    _inv0, ksum0, _ = H.check_sorted(b2s, keys, None, 0)
    assert inv == 0 and ksum == ksum0
    # exact multiset: per-value counts of input and output agree (with zero inversions this IS the sorted input)
    def hist(t):
        return torch.stack([t[i: i + (1 << 28)].view(torch.uint8).to(torch.int16).bincount(minlength=256)
                            for i in range(0, n, 1 << 28)]).sum(0)
    assert torch.equal(hist(keys), hist(out))
    assert int(out.view(torch.uint8)[0]) == 0 and int(out.view(torch.uint8)[n - 1]) == 255


@pytest.mark.parametrize("depth", [1, 2])
def test_host_sorter_pipeline(oracle, depth):
    """HostSorter (the e2e path of bench.py): host buffers in, host buffers out; with depth 2 consecutive calls are
    pipelined over three streams and two buffer sets -- every call must still return ITS OWN sorted data."""
    from cub_b200.device_radix_sort import HostSorter

    n = 300_007
    rng = np.random.default_rng(17)
    sorter = HostSorter(n, torch.uint32, torch.uint32, "cuda:0", depth=depth)
    inputs, outputs = [], []
    for i in range(5):
        k = H.random_bits(rng, n, 4)
        v = np.arange(n, dtype=np.uint32) + np.uint32(i)
        hk = torch.from_numpy(k.view(np.int32).copy()).pin_memory().view(torch.uint32)
        hv = torch.from_numpy(v.view(np.int32).copy()).pin_memory().view(torch.uint32)
        ok, ov = sorter(hk, hv)
        if depth == 1:
            sorter.synchronize()
            outputs.append((ok.view(torch.int32).numpy().view(np.uint32).copy(), ov.view(torch.int32).numpy().view(np.uint32).copy()))
        else:
            outputs.append((ok, ov))
        inputs.append((k, v, hk, hv))
        if depth == 2 and i >= 1:  # the slot of call i-1 is only reused by call i+1: read it after its own download
            pk, pv = outputs[i - 1]
            sorter.slots[(i - 1) % 2]["done"].synchronize()
            outputs[i - 1] = (pk.view(torch.int32).numpy().view(np.uint32).copy(), pv.view(torch.int32).numpy().view(np.uint32).copy())
    sorter.synchronize()
    if depth == 2:
        pk, pv = outputs[-1]
        outputs[-1] = (pk.view(torch.int32).numpy().view(np.uint32).copy(), pv.view(torch.int32).numpy().view(np.uint32).copy())
    for (k, v, _hk, _hv), (gk, gv) in zip(inputs, outputs):
        ek, ev = oracle.radix_sort(k, v, 6)
        assert np.array_equal(gk, ek) and np.array_equal(gv, ev)


# ---------------------------------------------------------------------------------------------
# CUDA graphs: a sort is stream-ordered work only (no allocation, no synchronisation), so it can be captured
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [5000, 300_000])
def test_cuda_graph_capture_and_replay(b2s, oracle, n):
    """Capture one pointer-form SortPairs (single-tile kernel for n = 5000, memset + histogram + 4 digit passes for
    n = 300000) in a CUDA graph and replay it on fresh inputs: the replays must sort whatever is in the input buffers."""
    import cub_b200 as cb

    rng = np.random.default_rng(n)
    keys = torch.empty(n, dtype=torch.int32, device="cuda")
    vals = torch.empty(n, dtype=torch.int32, device="cuda")
    keys_out, vals_out = torch.empty_like(keys), torch.empty_like(vals)
    err, nbytes = cb.DeviceRadixSort.SortPairs(None, 0, keys, keys_out, vals, vals_out, n)
    assert err == 0
    temp = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    raw0 = H.random_bits(rng, n, 4)
    keys.copy_(H.to_dev(raw0))
    vals.copy_(H.to_dev(np.arange(n, dtype=np.uint32)))
    # warm-up outside the capture (first use opts the kernels in to their shared-memory size)
    err, _ = cb.DeviceRadixSort.SortPairs(temp, nbytes, keys, keys_out, vals, vals_out, n)
    assert err == 0
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        err, _ = cb.DeviceRadixSort.SortPairs(temp, nbytes, keys, keys_out, vals, vals_out, n)
    assert err == 0
    for rep in range(3):
        raw = H.random_bits(rng, n, 4)
        v = rng.permutation(n).astype(np.uint32)
        keys.copy_(H.to_dev(raw))
        vals.copy_(H.to_dev(v))
        keys_out.zero_()
        g.replay()
        torch.cuda.synchronize()
        ek, ev = oracle.radix_sort(raw, v, 7)  # the tensors are int32: signed order
        assert np.array_equal(H.to_np(keys_out, np.uint32), ek), f"graph replay {rep}: keys differ"
        assert np.array_equal(H.to_np(vals_out, np.uint32), ev), f"graph replay {rep}: values differ"


# ---------------------------------------------------------------------------------------------
# round 2: the survey's "unpinned corner" -- floating keys x partial bit ranges x direction -- against the reference
# ---------------------------------------------------------------------------------------------
FLOAT_PARTIAL = [(8, 4), (5, 2), (11, 8), (4, 2)]  # (key type, bytes): f32, bf16, f64, f16


@pytest.mark.parametrize("kt,nb", FLOAT_PARTIAL, ids=[H.KEY_NAMES[k] for k, _ in FLOAT_PARTIAL])
@pytest.mark.parametrize("n", [1000, 4864, 4865, (1 << 20) + 77])
def test_float_partial_bit_ranges_vs_reference(b2s, refcub, oracle, kt, nb, n):
    """Floating keys with [begin_bit, end_bit) != the whole key, both directions, with +-0 / NaN / inf / denormals.
    The reference's two back-ends DISAGREE here when descending (radix_rank_sort_operations.cuh:55-66): its onesweep path
    (n > 4864) maps both zeros to the complemented image of -0.0, its single-tile path (n <= 4864) to the image of +0.0
    before reversing the order, so the zeros' digits of a PARTIAL range differ.  This library follows the onesweep
    behaviour at every size: bit-exact with the reference for n > 4864 in both directions and for
    n <= 4864 ascending; for n <= 4864 descending the checker is the oracle (onesweep semantics), and the reference is
    still matched on the same input with its zeros replaced (no +-0 => the two reference paths agree)."""
    bits = nb * 8
    rng = np.random.default_rng(1000 * kt + n % 997)
    raw = H.spice_floats(H.random_bits(rng, n, nb), nb)
    vals = np.arange(n, dtype=np.uint32)
    ranges = [(1, bits - 1), (0, bits - 1), (bits - 9, bits), (3, 12), (bits // 2 - 1, bits // 2 + 1)]
    for bb, eb in ranges:
        for desc in (False, True):
            dk, dv = H.to_dev(raw), H.to_dev(vals)
            k_us, v_us = H.sort_ptr(b2s.b2s_radix_sort, dk, dv, kt, desc, bb, eb)
            ek, ev = oracle.radix_sort(raw, vals, kt, desc, bb, eb)
            assert np.array_equal(H.to_np(k_us, raw.dtype), ek), f"vs oracle: {H.KEY_NAMES[kt]} n={n} [{bb},{eb}) desc={desc}"
            assert np.array_equal(H.to_np(v_us, np.uint32), ev)
            if n > 4864 or not desc:
                k_ref, v_ref = H.sort_ptr(refcub.sort, dk, dv, kt, desc, bb, eb)
                assert torch.equal(k_us, k_ref) and torch.equal(v_us, v_ref), \
                    f"vs reference CUB: {H.KEY_NAMES[kt]} n={n} [{bb},{eb}) desc={desc}"
            else:
                nz = raw.copy()
                high = np.array(1 << (bits - 1), dtype=np.uint64).astype(raw.dtype)
                nz[(raw == 0) | (raw == high)] = 1  # smallest denormal instead of +-0
                dk2 = H.to_dev(nz)
                k_us2, v_us2 = H.sort_ptr(b2s.b2s_radix_sort, dk2, dv, kt, desc, bb, eb)
                k_ref2, v_ref2 = H.sort_ptr(refcub.sort, dk2, dv, kt, desc, bb, eb)
                assert torch.equal(k_us2, k_ref2) and torch.equal(v_us2, v_ref2), \
                    f"vs reference CUB without zeros: {H.KEY_NAMES[kt]} n={n} [{bb},{eb}) desc"


@pytest.mark.parametrize("kt", [6, 7, 8, 9, 11, 2, 5, 0, 1])
def test_histogram_kernel_vs_oracle(b2s, oracle, kt):
    """The upfront histogram kernel alone (b2s_histogram.cuh) against oracle_histogram: exclusive digit offsets of every
    pass, whole keys and partial ranges, both directions, sizes around the vector / CTA granularity."""
    nb = H.KEY_BYTES[kt]
    bits = nb * 8
    rng = np.random.default_rng(4000 + kt)
    for n in (1, 15, 4097, 1_000_003):
        raw = H.random_bits(rng, n, nb)
        if kt in (5, 8, 11):
            raw = H.spice_floats(raw, nb)
        dk = H.to_dev(raw)
        for bb, eb in ((0, bits), (1, bits - 1), (3, min(bits, 14))):
            for desc in (False, True):
                passes = (eb - bb + 7) // 8
                out = torch.empty(passes * 256 + 1, dtype=torch.int64, device="cuda")
                rc = b2s.b2s_digit_histogram(H._p(dk), n, kt, int(desc), bb, eb, H._p(out), H.stream_handle())
                assert rc == 0
                torch.cuda.synchronize()
                got = out[: passes * 256].cpu().numpy().view(np.uint64).reshape(passes, 256)
                counts = oracle.histogram(raw, kt, desc, bb, eb).reshape(passes, 256)
                excl = np.cumsum(counts, axis=1) - counts
                assert np.array_equal(got, excl), f"histogram {H.KEY_NAMES[kt]} n={n} [{bb},{eb}) desc={desc}"


def test_lookback_forward_progress_under_concurrency(b2s, refcub):
    """Decoupled look-back stress: several multi-thousand-tile sorts in flight at once on different streams (their CTAs
    interleave on the SMs, so a tile's predecessors are NOT all resident when it starts), with block-index tile ids and
    with ticketed ids (b2s_set_tile_claim).  Every result must be bit-exact; a lost-progress bug would hang (the test
    runs under the suite's timeout)."""
    n = (1 << 24) + 4321
    keys = [H.gen_device_keys(b2s, n, 4, seed=50 + i) for i in range(4)]
    vals = H.gen_device_iota(b2s, n, 4)
    expect = [H.sort_ptr(refcub.sort, k, vals, 6) for k in keys]
    for claim in (0, 1):
        old = b2s.b2s_set_tile_claim(claim)
        try:
            streams = [torch.cuda.Stream() for _ in keys]
            outs = []
            nbytes = ctypes.c_size_t(0)
            assert b2s.b2s_radix_sort(None, ctypes.byref(nbytes), None, None, None, None, n, 6, 4, 4, 0, 0, 32, None) == 0
            temps = [torch.empty(nbytes.value, dtype=torch.uint8, device="cuda") for _ in keys]
            torch.cuda.synchronize()
            for rep in range(3):
                outs = []
                for k, s, t in zip(keys, streams, temps):
                    ko, vo = torch.empty_like(k), torch.empty_like(vals)
                    with torch.cuda.stream(s):
                        rc = b2s.b2s_radix_sort(H._p(t), ctypes.byref(nbytes), H._p(k), H._p(ko), H._p(vals), H._p(vo), n, 6, 4, 4,
                                                0, 0, 32, ctypes.c_void_p(s.cuda_stream))
                    assert rc == 0
                    outs.append((ko, vo))
                torch.cuda.synchronize()
                for (ko, vo), (ek, ev) in zip(outs, expect):
                    assert torch.equal(ko, ek) and torch.equal(vo, ev), f"claim={claim} rep={rep}"
        finally:
            b2s.b2s_set_tile_claim(old)


@pytest.mark.parametrize("kt,vb", [(6, 4), (7, 0), (9, 4), (8, 4), (10, 8), (2, 0)])
def test_constant_digit_passes_are_copies(b2s, refcub, kt, vb):
    """Keys whose upper (or lower, or all) digits are the same for every key: those passes run as plain copies (flag from the
    upfront histogram; the reference's single-bin short circuit, agent_radix_sort_onesweep.cuh:344-420).  Results must stay
    bit-exact against the reference -- pointer and DoubleBuffer forms, sizes with a partial last tile, unaligned pointers."""
    nb = H.KEY_BYTES[kt]
    bits = nb * 8
    for n, mask_lo, mask_hi in (((1 << 20) + 12345, 0, bits // 2), (300_007, bits // 4, bits), (70_001, 0, 0), (200_000, 8, bits - 8)):
        keys = H.gen_device_keys(b2s, n + 8, nb, seed=77)
        # keep only bits [mask_lo, mask_hi) random; everything else constant (also non-zero constants)
        width = mask_hi - mask_lo
        m = ((1 << width) - 1) << mask_lo if width > 0 else 0
        const = 0x5A5A5A5A5A5A5A5A & ((1 << bits) - 1) & ~m
        if m >= 1 << (bits - 1):
            m -= 1 << bits
        if const >= 1 << (bits - 1):
            const -= 1 << bits
        keys = (keys & m) | const
        vals = H.gen_device_iota(b2s, n + 8, vb) if vb else None
        for off in (0, 1):
            k_in = keys[off:off + n]
            v_in = vals[off:off + n] if vb else None
            for desc in (False, True):
                k_ref, v_ref = H.sort_ptr(refcub.sort, k_in, v_in, kt, desc)
                k_us, v_us = H.sort_ptr(b2s.b2s_radix_sort, k_in, v_in, kt, desc)
                assert torch.equal(k_us, k_ref), f"keys: kt={kt} n={n} bits[{mask_lo},{mask_hi}) desc={desc} off={off}"
                if vb:
                    assert torch.equal(v_us, v_ref), f"values: kt={kt} n={n} bits[{mask_lo},{mask_hi}) desc={desc} off={off}"
        kb = [keys[:n].clone(), torch.empty_like(keys[:n])]
        vbuf = [vals[:n].clone(), torch.empty_like(vals[:n])] if vb else None
        ks, vs = H.sort_db(b2s.b2s_radix_sort_db, kb, vbuf, kt)
        k_ref, v_ref = H.sort_ptr(refcub.sort, keys[:n], vals[:n] if vb else None, kt)
        assert torch.equal(kb[ks], k_ref) and (not vb or torch.equal(vbuf[vs], v_ref))
