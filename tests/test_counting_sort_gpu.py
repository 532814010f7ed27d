"""GPU parity tests of the counting-sort path (cub_b200/csrc/b2s_narrow.cu): keys-only sorts of 1- and 2-byte keys over all
their bits.  The path must be indistinguishable from the digit passes: bit-exact against the CPU oracle, against the unmodified
reference CUB on the same device buffers, and against our own digit passes (b2s_set_counting_sort(0)); special attention to
the one case where equal digits do not mean equal bits, -0.0 / +0.0 (cub/block/radix_rank_sort_operations.cuh:55-66, 79-89),
whose input order inside their common run has to survive."""
import numpy as np
import pytest
import torch

from tests import harness as H

pytestmark = pytest.mark.gpu

NARROW = [0, 1, 2, 3, 4, 5]  # u8 i8 u16 i16 f16 bf16


@pytest.fixture()
def counting_everywhere(b2s):
    """Cut-over lowered to one item, so that every eligible sort takes the counting path."""
    old = [b2s.b2s_set_counting_min_items(kb, 1) for kb in (1, 2)]
    on = b2s.b2s_set_counting_sort(1)
    yield b2s
    b2s.b2s_set_counting_sort(on)
    for kb, o in zip((1, 2), old):
        b2s.b2s_set_counting_min_items(kb, o)


def _check(b2s, oracle, raw, kt, desc, label):
    dk = H.to_dev(raw)
    before = dk.clone()
    ko, _ = H.sort_ptr(b2s.b2s_radix_sort, dk, None, kt, desc)
    launches = b2s.b2s_last_launch_count()
    ek, _ = oracle.radix_sort(raw, None, kt, desc)
    got = H.to_np(ko, raw.dtype)
    assert torch.equal(dk, before), f"{label}: pointer form modified its input"
    if not np.array_equal(got, ek):
        bad = np.nonzero(got != ek)[0]
        raise AssertionError(f"{label}: keys differ at {bad.size} of {raw.shape[0]} positions, first {bad[:5]}: "
                             f"got {got[bad[:5]]} expected {ek[bad[:5]]}")
    return launches


@pytest.mark.parametrize("kt", NARROW)
def test_counting_path_vs_oracle(counting_everywhere, oracle, kt):
    b2s = counting_everywhere
    nb = H.KEY_BYTES[kt]
    rng = np.random.default_rng(100 + kt)
    for n in (1, 2, 7, 8, 9, 31, 1000, 16384, 16385, 100_003, (1 << 20) + 11):
        raw = H.random_bits(rng, n, nb)
        if kt in (4, 5) and n >= 1000:
            raw = H.spice_floats(raw, nb)
        for desc in (False, True):
            launches = _check(b2s, oracle, raw, kt, desc, f"{H.KEY_NAMES[kt]} n={n} desc={desc}")
            # memset + histogram + expansion; 2-byte keys: + prefix; floating keys: + 2 zero kernels
            assert launches == (3 if nb == 1 else (6 if kt in (4, 5) else 4)), launches


@pytest.mark.parametrize("kt", [4, 5])
def test_floating_zeros_keep_their_input_order(counting_everywhere, oracle, kt):
    b2s = counting_everywhere
    rng = np.random.default_rng(7)
    n = 300_007
    pz, nz = np.uint16(0), np.uint16(0x8000)
    cases = {
        "only zeros, mixed": rng.choice([pz, nz], size=n),
        "only +0": np.full(n, pz),
        "only -0": np.full(n, nz),
        "one -0 among +0": np.where(np.arange(n) == n // 3, nz, pz).astype(np.uint16),
        "half zeros": np.where(rng.random(n) < 0.5, rng.choice([pz, nz], size=n), H.random_bits(rng, n, 2)).astype(np.uint16),
        "rare zeros": np.where(rng.random(n) < 1e-4, rng.choice([pz, nz], size=n), H.random_bits(rng, n, 2)).astype(np.uint16),
        "-0 only among random": np.where(rng.random(n) < 0.01, nz, H.random_bits(rng, n, 2) | np.uint16(1)).astype(np.uint16),
        "zeros in the last tile only": np.concatenate([H.random_bits(rng, n - 5, 2) | np.uint16(1),
                                                      np.array([nz, pz, nz, nz, pz], dtype=np.uint16)]),
    }
    for name, raw in cases.items():
        for desc in (False, True):
            _check(b2s, oracle, np.ascontiguousarray(raw), kt, desc, f"{H.KEY_NAMES[kt]} {name} desc={desc}")


@pytest.mark.parametrize("kt", NARROW)
def test_skewed_and_sparse_distributions(counting_everywhere, oracle, kt):
    b2s = counting_everywhere
    nb = H.KEY_BYTES[kt]
    ones = (1 << (8 * nb)) - 1
    dt = H.NP_BITS[nb]
    rng = np.random.default_rng(kt)
    n = 200_001
    cases = {
        "all equal": np.full(n, 0x5A & ones, dtype=dt),
        "two far apart": rng.choice(np.array([0, ones], dtype=dt), size=n),
        "extremes and middle": rng.choice(np.array([0, 1, ones >> 1, (ones >> 1) + 1, ones - 1, ones], dtype=dt), size=n),
        "multiples of 256": (H.random_bits(rng, n, nb) & dt(ones & ~0xFF)) if nb == 2 else H.random_bits(rng, n, nb) & dt(0xF0),
        "every image once or twice": np.concatenate([np.arange(ones + 1, dtype=dt), np.arange(0, ones + 1, 3, dtype=dt)]),
        "already sorted": np.sort(H.random_bits(rng, n, nb)),
    }
    for name, raw in cases.items():
        raw = np.ascontiguousarray(raw)
        if name == "every image once or twice":
            raw = rng.permutation(raw)
        for desc in (False, True):
            _check(b2s, oracle, raw, kt, desc, f"{H.KEY_NAMES[kt]} {name} desc={desc}")


@pytest.mark.parametrize("kt", [0, 3, 5])
def test_unaligned_input_and_output(counting_everywhere, oracle, kt):
    b2s = counting_everywhere
    nb = H.KEY_BYTES[kt]
    rng = np.random.default_rng(33)
    n = 70_001
    for off_in in (0, 1, 3, 7):
        for off_out in (0, 1, 5):
            raw = H.random_bits(rng, n, nb)
            if kt == 5:
                raw = H.spice_floats(raw, nb)
            big = H.to_dev(np.concatenate([np.zeros(off_in, dtype=raw.dtype), raw]))
            out = torch.zeros(n + off_out + 8, dtype=big.dtype, device="cuda")
            H.sort_ptr(b2s.b2s_radix_sort, big[off_in:], None, kt, False, keys_out=out[off_out:off_out + n], n=n)
            ek, _ = oracle.radix_sort(raw, None, kt, False)
            assert np.array_equal(H.to_np(out[off_out:off_out + n], raw.dtype), ek), (off_in, off_out)
            assert int(out[:off_out].abs().sum()) == 0 and int(out[off_out + n:].abs().sum()) == 0, "wrote outside the output"


@pytest.mark.parametrize("kt", NARROW)
def test_double_buffer_form_lands_in_the_alternate_buffer(counting_everywhere, oracle, kt):
    b2s = counting_everywhere
    nb = H.KEY_BYTES[kt]
    rng = np.random.default_rng(5)
    n = 50_000
    raw = H.random_bits(rng, n, nb)
    ek, _ = oracle.radix_sort(raw, None, kt, True)
    for selector in (0, 1):
        kb = [torch.zeros(n, dtype=H.CONTAINER[nb], device="cuda") for _ in range(2)]
        kb[selector].copy_(H.to_dev(raw))
        ks, _ = H.sort_db(b2s.b2s_radix_sort_db, kb, None, kt, True, selector=selector)
        assert ks == selector ^ 1
        assert np.array_equal(H.to_np(kb[ks], raw.dtype), ek)


def test_not_taken_with_values_or_partial_bits(counting_everywhere, oracle):
    """Values or a partial bit range make equal sort keys distinguishable: those sorts stay on the digit passes."""
    b2s = counting_everywhere
    rng = np.random.default_rng(9)
    n = 100_000
    raw = H.random_bits(rng, n, 2)
    vals = np.arange(n, dtype=np.uint32)
    ko, vo = H.sort_ptr(b2s.b2s_radix_sort, H.to_dev(raw), H.to_dev(vals), 2)
    ek, ev = oracle.radix_sort(raw, vals, 2)
    assert np.array_equal(H.to_np(ko, raw.dtype), ek) and np.array_equal(H.to_np(vo, np.uint32), ev)
    for bb, eb in ((0, 15), (1, 16), (4, 12)):
        ko, _ = H.sort_ptr(b2s.b2s_radix_sort, H.to_dev(raw), None, 2, False, bb, eb)
        ek, _ = oracle.radix_sort(raw, None, 2, False, bb, eb)
        assert np.array_equal(H.to_np(ko, raw.dtype), ek), (bb, eb)
        assert b2s.b2s_last_launch_count() <= 4  # memset + histogram + <= 2 digit passes


@pytest.mark.parametrize("kt,lg", [(5, 24), (4, 24), (2, 24), (3, 23), (0, 24), (1, 23)])
def test_default_cut_over_vs_reference_cub_and_digit_passes(b2s, refcub, kt, lg):
    """Default settings at a size above the cut-over: reference CUB, the counting path and our digit passes agree bit for bit."""
    nb = H.KEY_BYTES[kt]
    n = (1 << lg) + 4321
    keys = H.gen_device_keys(b2s, n, nb, seed=11)
    if kt in (4, 5):
        idx = torch.arange(n, device="cuda")
        keys[idx % 256 == 0] = 0
        keys[idx % 256 == 1] = torch.iinfo(H.CONTAINER[nb]).min
    for desc in (False, True):
        k_ref, _ = H.sort_ptr(refcub.sort, keys, None, kt, desc)
        k_cnt, _ = H.sort_ptr(b2s.b2s_radix_sort, keys, None, kt, desc)
        assert b2s.b2s_last_launch_count() == (3 if nb == 1 else (6 if kt in (4, 5) else 4)), "counting path not taken"
        old = b2s.b2s_set_counting_sort(0)
        try:
            k_dig, _ = H.sort_ptr(b2s.b2s_radix_sort, keys, None, kt, desc)
            assert b2s.b2s_last_launch_count() == 2 + nb, "digit passes not taken"
        finally:
            b2s.b2s_set_counting_sort(old)
        assert torch.equal(k_cnt, k_ref), "counting path differs from reference CUB"
        assert torch.equal(k_dig, k_ref), "digit passes differ from reference CUB"


@pytest.mark.parametrize("kt", [5, 2, 0])
def test_cuda_graph_capture_and_replay_of_the_counting_path(counting_everywhere, oracle, kt):
    """The counting path is stream-ordered work only (memset + kernels, device-side decisions about the zeros): captured once,
    the graph must sort whatever is in the input buffer at replay time -- including inputs whose zero flag differs."""
    b2s = counting_everywhere
    nb = H.KEY_BYTES[kt]
    n = 300_001
    rng = np.random.default_rng(kt)
    keys = torch.zeros(n, dtype=H.CONTAINER[nb], device="cuda")
    out = torch.empty_like(keys)
    import ctypes
    nbytes = ctypes.c_size_t(0)
    args = (H._p(keys), H._p(out), None, None, n, kt, 0, 4, 1, 0, 8 * nb)
    assert b2s.b2s_radix_sort(None, ctypes.byref(nbytes), *args, None) == 0
    temp = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        assert b2s.b2s_radix_sort(ctypes.c_void_p(temp.data_ptr()), ctypes.byref(nbytes), *args, ctypes.c_void_p(s.cuda_stream)) == 0
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        assert b2s.b2s_radix_sort(ctypes.c_void_p(temp.data_ptr()), ctypes.byref(nbytes), *args,
                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    for rep in range(4):
        raw = H.random_bits(rng, n, nb)
        if kt == 5:
            raw = H.spice_floats(raw, nb) if rep % 2 == 0 else (raw | np.uint16(1))  # both zeros / no zero at all
        keys.copy_(H.to_dev(raw))
        out.zero_()
        g.replay()
        torch.cuda.synchronize()
        ek, _ = oracle.radix_sort(raw, None, kt, True)
        assert np.array_equal(H.to_np(out, raw.dtype), ek), f"graph replay {rep}"


def test_concurrent_streams_share_nothing(counting_everywhere, oracle):
    """Several counting sorts in flight on different streams with their own temp storage."""
    b2s = counting_everywhere
    import ctypes
    rng = np.random.default_rng(77)
    jobs = []
    for i, kt in enumerate([5, 4, 2, 3, 0, 1, 5, 2]):
        nb = H.KEY_BYTES[kt]
        n = 150_000 + 1111 * i
        raw = H.random_bits(rng, n, nb)
        if kt in (4, 5):
            raw = H.spice_floats(raw, nb)
        keys, out = H.to_dev(raw), torch.empty(n, dtype=H.CONTAINER[nb], device="cuda")
        nbytes = ctypes.c_size_t(0)
        args = (H._p(keys), H._p(out), None, None, n, kt, 0, 4, i & 1, 0, 8 * nb)
        assert b2s.b2s_radix_sort(None, ctypes.byref(nbytes), *args, None) == 0
        temp = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
        jobs.append((kt, raw, keys, out, temp, nbytes, args, torch.cuda.Stream(), bool(i & 1)))
    torch.cuda.synchronize()
    for _ in range(3):
        for kt, raw, keys, out, temp, nbytes, args, st, desc in jobs:
            assert b2s.b2s_radix_sort(ctypes.c_void_p(temp.data_ptr()), ctypes.byref(nbytes), *args, ctypes.c_void_p(st.cuda_stream)) == 0
    torch.cuda.synchronize()
    for kt, raw, keys, out, temp, nbytes, args, st, desc in jobs:
        ek, _ = oracle.radix_sort(raw, None, kt, desc)
        assert np.array_equal(H.to_np(out, raw.dtype), ek), H.KEY_NAMES[kt]


def test_python_mirror_and_frontend_reach_the_counting_path(b2s, oracle):
    """cub_b200.sort_keys / DeviceRadixSort.SortKeys(DoubleBuffer) on torch bf16 / int16 / uint8 tensors above the default
    cut-over: same bits as the oracle; the DoubleBuffer's Current() names the result."""
    import cub_b200 as cb

    rng = np.random.default_rng(2026)
    for dtype, kt, n in ((torch.bfloat16, 5, (1 << 23) + 13), (torch.int16, 3, (1 << 22) + 7), (torch.uint8, 0, (1 << 17) + 3)):
        nb = H.KEY_BYTES[kt]
        raw = H.random_bits(rng, n, nb)
        if kt == 5:
            raw = H.spice_floats(raw, nb)
        t = H.to_dev(raw).view(dtype)
        for desc in (False, True):
            out = cb.sort_keys(t, descending=desc)
            torch.cuda.synchronize()
            ek, _ = oracle.radix_sort(raw, None, kt, desc)
            assert np.array_equal(H.to_np(out.view(H.CONTAINER[nb]), raw.dtype), ek), (dtype, desc)
        db = cb.DoubleBuffer(t.clone(), torch.empty_like(t))
        fn = cb.DeviceRadixSort.SortKeys
        err, nbytes = fn(None, 0, db, n)
        assert err == 0
        temp = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        err, _ = fn(temp, nbytes, db, n)
        assert err == 0
        torch.cuda.synchronize()
        ek, _ = oracle.radix_sort(raw, None, kt, False)
        assert db.selector == 1
        assert np.array_equal(H.to_np(db.Current().view(H.CONTAINER[nb]), raw.dtype), ek), dtype
