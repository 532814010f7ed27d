"""GPU parity tests of "zero recording" (cub_b200/csrc/b2s_fzero.cu, ImageFloatOp): full-range sorts of f32 / f64 keys give both
zeros one image in the first pass, record them, and restore their signs after the last pass.  Must be indistinguishable from
the reference's scheme (-0.0 / +0.0 equal in every digit, bits kept, stable: cub/block/radix_rank_sort_operations.cuh:55-66,
79-89): bit-exact keys and values against the CPU oracle, the reference CUB, and our own kernels with the recording switched
off."""
import ctypes

import numpy as np
import pytest
import torch

from tests import harness as H

pytestmark = pytest.mark.gpu

FLOATS = [(8, 4), (11, 8)]  # (key type, bytes): f32, f64


def _zero_cases(rng, n, nb):
    dt = H.NP_BITS[nb]
    pz, nz = dt(0), dt(1 << (8 * nb - 1))
    rnd = lambda: H.random_bits(rng, n, nb)  # noqa: E731
    nonzero = lambda: rnd() | dt(1)          # noqa: E731
    return {
        "uniform bits + spiked specials": H.spice_floats(rnd(), nb),
        "only zeros, mixed": rng.choice(np.array([pz, nz], dtype=dt), size=n),
        "only +0": np.full(n, pz, dtype=dt),
        "only -0": np.full(n, nz, dtype=dt),
        "no zero at all": nonzero(),
        "one -0 among +0 and noise": np.where(np.arange(n) == n // 3, nz, np.where(rng.random(n) < 0.3, pz, nonzero())).astype(dt),
        "half zeros": np.where(rng.random(n) < 0.5, rng.choice(np.array([pz, nz], dtype=dt), size=n), rnd()).astype(dt),
        "rare zeros": np.where(rng.random(n) < 1e-4, rng.choice(np.array([pz, nz], dtype=dt), size=n), nonzero()).astype(dt),
        "zeros in the last tile only": np.concatenate([nonzero()[: n - 5], np.array([nz, pz, nz, nz, pz], dtype=dt)]),
        "denormals and tiny keys around the zeros": (rnd() & dt((1 << (8 * nb - 1)) | 0x7)).astype(dt),
    }


@pytest.mark.parametrize("kt,nb", FLOATS, ids=["f32", "f64"])
@pytest.mark.parametrize("with_values", [False, True], ids=["keys", "pairs"])
def test_zero_recording_vs_oracle(b2s, oracle, kt, nb, with_values):
    rng = np.random.default_rng(kt * 10 + with_values)
    for n in (20_000, 100_003, (1 << 20) + 77):
        vals = np.arange(n, dtype=np.uint32)[::-1].copy() if with_values else None
        for name, raw in _zero_cases(rng, n, nb).items():
            raw = np.ascontiguousarray(raw)
            for desc in (False, True):
                dk = H.to_dev(raw)
                dv = H.to_dev(vals) if with_values else None
                before = dk.clone()
                ko, vo = H.sort_ptr(b2s.b2s_radix_sort, dk, dv, kt, desc)
                passes = nb
                assert b2s.b2s_last_launch_count() == 2 + passes + 2, "zero recording not taken"
                ek, ev = oracle.radix_sort(raw, vals, kt, desc)
                got = H.to_np(ko, raw.dtype)
                assert torch.equal(dk, before), "pointer form modified its input"
                if not np.array_equal(got, ek):
                    bad = np.nonzero(got != ek)[0]
                    raise AssertionError(f"{H.KEY_NAMES[kt]} n={n} {name} desc={desc}: keys differ at {bad.size} positions, "
                                         f"first {bad[:5]}: got {got[bad[:5]]} expected {ek[bad[:5]]}")
                if with_values:
                    assert np.array_equal(H.to_np(vo, np.uint32), ev), f"{H.KEY_NAMES[kt]} n={n} {name} desc={desc}: values"


@pytest.mark.parametrize("kt,nb", FLOATS, ids=["f32", "f64"])
def test_same_bits_with_the_recording_switched_off_and_as_reference_cub(b2s, refcub, kt, nb):
    n = (1 << 22) + 4321
    keys = H.gen_device_keys(b2s, n, nb, seed=5)
    idx = torch.arange(n, device="cuda")
    keys[idx % 256 == 0] = 0
    keys[idx % 256 == 1] = torch.iinfo(H.CONTAINER[nb]).min
    vals = H.gen_device_iota(b2s, n, 4)
    for desc in (False, True):
        for v in (None, vals):
            k_ref, v_ref = H.sort_ptr(refcub.sort, keys, v, kt, desc)
            k_on, v_on = H.sort_ptr(b2s.b2s_radix_sort, keys, v, kt, desc)
            assert b2s.b2s_last_launch_count() == 2 + nb + 2
            old = b2s.b2s_set_float_zero_recording(0)
            try:
                k_off, v_off = H.sort_ptr(b2s.b2s_radix_sort, keys, v, kt, desc)
                assert b2s.b2s_last_launch_count() == 2 + nb
            finally:
                b2s.b2s_set_float_zero_recording(old)
            assert torch.equal(k_on, k_ref) and torch.equal(k_off, k_ref)
            if v is not None:
                assert torch.equal(v_on, v_ref) and torch.equal(v_off, v_ref)


def test_not_taken_for_partial_ranges_wide_values_or_single_tiles(b2s, oracle):
    rng = np.random.default_rng(3)
    n = 50_000
    raw = H.spice_floats(H.random_bits(rng, n, 4), 4)
    # partial bit range: the zeros' run is not where digit 0x80 of the top pass starts
    ko, _ = H.sort_ptr(b2s.b2s_radix_sort, H.to_dev(raw), None, 8, True, 3, 29)
    assert b2s.b2s_last_launch_count() == 2 + 4
    assert np.array_equal(H.to_np(ko, raw.dtype), oracle.radix_sort(raw, None, 8, True, 3, 29)[0])
    # 8-byte values
    v8 = np.arange(n, dtype=np.uint64)
    ko, vo = H.sort_ptr(b2s.b2s_radix_sort, H.to_dev(raw), H.to_dev(v8), 8)
    assert b2s.b2s_last_launch_count() == 2 + 4
    ek, ev = oracle.radix_sort(raw, v8, 8)
    assert np.array_equal(H.to_np(ko, raw.dtype), ek) and np.array_equal(H.to_np(vo, np.uint64), ev)
    # one tile: single launch
    small = raw[:3000].copy()
    ko, _ = H.sort_ptr(b2s.b2s_radix_sort, H.to_dev(small), None, 8)
    assert b2s.b2s_last_launch_count() == 1
    assert np.array_equal(H.to_np(ko, raw.dtype), oracle.radix_sort(small, None, 8)[0])


@pytest.mark.parametrize("kt,nb", FLOATS, ids=["f32", "f64"])
def test_double_buffer_form_and_unaligned_pointers(b2s, oracle, kt, nb):
    rng = np.random.default_rng(17)
    n = 70_001
    raw = H.spice_floats(H.random_bits(rng, n, nb), nb)
    vals = rng.permutation(n).astype(np.uint32)
    ek, ev = oracle.radix_sort(raw, vals, kt, True)
    for selector in (0, 1):
        kb = [torch.zeros(n, dtype=H.CONTAINER[nb], device="cuda") for _ in range(2)]
        vb = [torch.zeros(n, dtype=torch.int32, device="cuda") for _ in range(2)]
        kb[selector].copy_(H.to_dev(raw))
        vb[selector].copy_(H.to_dev(vals))
        ks, vs = H.sort_db(b2s.b2s_radix_sort_db, kb, vb, kt, True, selector=selector)
        assert ks == vs == selector ^ (nb & 1)
        assert np.array_equal(H.to_np(kb[ks], raw.dtype), ek) and np.array_equal(H.to_np(vb[vs], np.uint32), ev)
    for off in (1, 3):  # element-aligned only
        big = H.to_dev(np.concatenate([np.zeros(off, dtype=raw.dtype), raw]))
        out = torch.zeros(n + off + 4, dtype=big.dtype, device="cuda")
        H.sort_ptr(b2s.b2s_radix_sort, big[off:], None, kt, False, keys_out=out[off:off + n], n=n)
        assert np.array_equal(H.to_np(out[off:off + n], raw.dtype), oracle.radix_sort(raw, None, kt, False)[0])
        assert int(out[:off].abs().sum()) == 0 and int(out[off + n:].abs().sum()) == 0


def test_cuda_graph_replay_with_and_without_zeros(b2s, oracle):
    n = 300_001
    rng = np.random.default_rng(8)
    keys = torch.zeros(n, dtype=torch.int32, device="cuda")
    out = torch.empty_like(keys)
    nbytes = ctypes.c_size_t(0)
    args = (H._p(keys), H._p(out), None, None, n, 8, 0, 4, 1, 0, 32)
    assert b2s.b2s_radix_sort(None, ctypes.byref(nbytes), *args, None) == 0
    temp = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    assert b2s.b2s_radix_sort(ctypes.c_void_p(temp.data_ptr()), ctypes.byref(nbytes), *args, H.stream_handle()) == 0
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        assert b2s.b2s_radix_sort(ctypes.c_void_p(temp.data_ptr()), ctypes.byref(nbytes), *args, H.stream_handle()) == 0
    for rep in range(4):
        raw = H.random_bits(rng, n, 4)
        raw = H.spice_floats(raw, 4) if rep % 2 == 0 else (raw | np.uint32(1))
        keys.copy_(H.to_dev(raw))
        out.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert np.array_equal(H.to_np(out, raw.dtype), oracle.radix_sort(raw, None, 8, True)[0]), f"graph replay {rep}"
