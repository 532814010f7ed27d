"""CPU-only: the C-ABI library loads, exports every symbol include/b2s_radix_sort.h declares, answers the
temp-storage size query without a GPU, and the Python mirror of cub::DeviceRadixSort forwards arguments
with the reference's meaning.  No compute is launched here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    """Every function declared by the C headers in include/ (b2s_radix_sort.h, b2s_mgpu.h, ...)."""
    import glob

    syms = set()
    for path in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
        syms |= set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", text))
    return sorted(syms)


def test_library_exports_every_declared_symbol(b2s):
    syms = _declared_symbols()
    assert {"b2s_radix_sort", "b2s_radix_sort_db"} <= set(syms)
    for s in syms:
        assert hasattr(b2s, s), f"libb2s.so does not export {s}"
    from cub_b200 import _lib

    assert set(_lib.EXPORTS) == set(syms), "ctypes prototypes and header disagree"
    assert b2s.b2s_version().decode().endswith("sm_100a")
    assert [b2s.b2s_key_bytes(i) for i in range(12)] == [1, 1, 2, 2, 2, 2, 4, 4, 4, 8, 8, 8]
    assert b2s.b2s_key_bytes(12) == 0


def test_sass_is_sm100a_with_bulk_copy():
    """The product is sm_100a-only and the digit pass stages tiles with the TMA engine (UBLKCP)."""
    import shutil
    import subprocess

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    from cub_b200 import _lib

    elf = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in elf and "sm_90" not in elf and "sm_80" not in elf
    syms = subprocess.run(["cuobjdump", "-symbols", _lib.LIB_PATH], capture_output=True, text=True).stdout
    names = sorted({t for line in syms.splitlines() for t in line.split()
                    if "digit_pass_kernelILi4ELi4ENS_7DigitOpILi4ELb0EEEj" in t and not t.startswith("$")})
    assert names, "production u32/u32 digit-pass kernel not found in libb2s.so"
    lab = [t for line in syms.splitlines() for t in line.split() if "onesweep_kernel" in t]
    assert not lab, "the laboratory kernel (ablation / trace branches) must not be in the product library"
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", names[0], _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass, "digit pass must stage its tile with the TMA engine"
    assert "VOTE" in sass and "ATOMS" in sass


def _query(b2s, n, kt, vb, bb=0, eb=None, db=False):
    nbytes = ctypes.c_size_t(0)
    eb = b2s.b2s_key_bytes(kt) * 8 if eb is None else eb
    if db:
        kb = (ctypes.c_void_p * 2)(None, None)
        sel = ctypes.c_int(0)
        vbufs = (ctypes.c_void_p * 2)(None, None)
        vsel = ctypes.c_int(0)
        rc = b2s.b2s_radix_sort_db(None, ctypes.byref(nbytes), kb, ctypes.byref(sel), vbufs, ctypes.byref(vsel), n, kt,
                                   vb, 4, 0, bb, eb, None)
        assert sel.value == 0
    else:
        rc = b2s.b2s_radix_sort(None, ctypes.byref(nbytes), None, None, None, None, n, kt, vb, 4, 0, bb, eb, None)
    assert rc == 0
    return nbytes.value


def test_temp_storage_query_on_cpu(b2s):
    # trivial cases: 1 byte (dispatch_radix_sort.cuh:1945-1952)
    assert _query(b2s, 0, 6, 4) == 1
    assert _query(b2s, 1000, 6, 4, 5, 5) == 1
    assert _query(b2s, 1000, 6, 4, 5, 5, db=True) == 1
    n = 1 << 28
    ptr = _query(b2s, n, 6, 4)
    dbl = _query(b2s, n, 6, 4, db=True)
    # pointer form carries one alternate key + value buffer; DoubleBuffer form only bookkeeping
    assert ptr - dbl >= 2 * n * 4 and ptr - dbl < 2 * n * 4 + 1024
    assert dbl < 256 << 20
    # a single digit pass needs no alternate buffers in either form
    assert _query(b2s, n, 6, 4, 0, 8) - _query(b2s, n, 6, 4, 0, 8, db=True) == 0
    # 64-bit look-back words from 2^30 items on
    assert _query(b2s, 1 << 30, 9, 4, db=True) > 2 * _query(b2s, (1 << 30) - 1, 9, 4, db=True) - (1 << 20)
    # bad arguments are rejected, not crashed on
    nbytes = ctypes.c_size_t(0)
    assert b2s.b2s_radix_sort(None, ctypes.byref(nbytes), None, None, None, None, 10, 12, 0, 4, 0, 0, 8, None) != 0
    assert b2s.b2s_radix_sort(None, ctypes.byref(nbytes), None, None, None, None, 10, 6, 3, 4, 0, 0, 8, None) != 0
    assert b2s.b2s_radix_sort(None, None, None, None, None, None, 10, 6, 0, 4, 0, 0, 8, None) != 0


def test_size_queries_of_the_counting_path_and_of_zero_recording(b2s):
    """Temp-storage sizes follow the path a sort will take (decided from its arguments alone, so that the query and the sort
    agree): the counting path needs no alternate buffer; zero recording adds two bits per key; the switches restore the plain
    digit-pass sizes.  No GPU work."""
    n = 1 << 26
    # u16 keys alone, all bits: 2 x 512 KB of counters / prefix instead of an n * 2-byte alternate buffer
    cnt = _query(b2s, n, 2, 0)
    assert cnt < (2 << 20)
    # bf16: + one bit per key for the zeros
    assert n // 8 <= _query(b2s, n, 5, 0) - cnt < n // 8 + (1 << 20)
    old = b2s.b2s_set_counting_sort(0)
    try:
        assert _query(b2s, n, 2, 0) >= n * 2
    finally:
        assert b2s.b2s_set_counting_sort(old) == 0
    # below the cut-over, with values or with a partial bit range: digit passes
    assert _query(b2s, 1 << 20, 2, 0) >= (1 << 20) * 2
    assert _query(b2s, n, 2, 4) >= n * 6
    assert _query(b2s, n, 2, 0, 0, 15) >= n * 2
    prev = b2s.b2s_set_counting_min_items(2, 1 << 30)
    try:
        assert prev == 1 << 22 and _query(b2s, n, 2, 0) >= n * 2
    finally:
        b2s.b2s_set_counting_min_items(2, prev)
        b2s.b2s_set_counting_min_items(1, b2s.b2s_set_counting_min_items(1, 1 << 16))
    # f32 keys, DoubleBuffer form: bookkeeping + two bits per key; without recording: bookkeeping only
    rec = _query(b2s, n, 8, 0, db=True)
    old = b2s.b2s_set_float_zero_recording(0)
    try:
        plain = _query(b2s, n, 8, 0, db=True)
    finally:
        assert b2s.b2s_set_float_zero_recording(old) == 0
    # (the recording sort runs on the integer tile shapes: fewer tiles, smaller status arrays than `plain`)
    assert rec >= n // 4 and plain < n // 4 and rec - plain < n // 4 + (4 << 20)
    # partial bit ranges and 8-byte values do not record
    assert _query(b2s, n, 8, 0, 1, 31, db=True) < (64 << 20)
    assert _query(b2s, n, 8, 8, db=True) < (64 << 20)


def test_variant_description(b2s):
    nt, ipt, minb, match = (ctypes.c_int() for _ in range(4))
    nv = b2s.b2s_describe_variant(4, 4, 0, ctypes.byref(nt), ctypes.byref(ipt), ctypes.byref(minb), ctypes.byref(match))
    assert nv >= 1 and nt.value % 32 == 0 and nt.value >= 256 and ipt.value >= 1
    assert b2s.b2s_describe_variant(3, 4, 0, None, None, None, None) == -1


def test_python_mirror_argument_handling():
    import torch

    import cub_b200 as cb

    assert cb.key_type_of(torch.float32) == 8 and cb.key_type_of(torch.bfloat16) == 5
    with pytest.raises(TypeError):
        cb.key_type_of(torch.bool)
    d = cb.DoubleBuffer("a", "b")
    assert d.Current() == "a" and d.Alternate() == "b"
    d.selector = 1
    assert d.Current() == "b"
    # size query through the mirror with raw (null) pointers: pointer and DoubleBuffer forms
    err, nbytes = cb.DeviceRadixSort.SortPairs(None, 0, 0, 0, 0, 0, 1 << 20, key_type=6, value_bytes=4)
    assert err == 0 and nbytes > 2 * 4 * (1 << 20)
    err, nb2 = cb.DeviceRadixSort.SortKeysDescending(None, 0, cb.DoubleBuffer(0, 0), 1 << 20, 0, 16, key_type=6)
    assert err == 0 and 1 < nb2 < nbytes
    with pytest.raises(ValueError):
        cb.DeviceRadixSort.SortKeys(None, 0, torch.zeros(4, dtype=torch.int32), torch.zeros(4, dtype=torch.int32), 4)


def test_no_cpu_fallback_in_product():
    """The product package must not import the oracle or fall back to torch.sort."""
    for root, _, files in os.walk(os.path.join(ROOT, "cub_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(root, f)).read()
                assert "pyoracle" not in src and "liboracle" not in src and "oracle/" not in src, f
                assert "torch.sort" not in src and "argsort" not in src, f


def _build_veneer_example():
    import shutil
    import subprocess

    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    exe = os.path.join(ROOT, "tests", "cpp", "veneer_example")
    subprocess.check_call(["nvcc", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "veneer_example.cu"), "-L" + os.path.join(ROOT, "cub_b200"), "-lb2s",
                           "-Xlinker", "-rpath", "-Xlinker", os.path.join(ROOT, "cub_b200"), "-o", exe])
    return exe


def test_cpp_veneer_call_site_compiles(b2s):
    """A reference-style C++ call site (namespace swapped) compiles and links against the veneer + libb2s.so."""
    assert os.path.exists(_build_veneer_example())


@pytest.mark.gpu
def test_cpp_veneer_call_site_runs(b2s):
    import subprocess

    exe = os.path.join(ROOT, "tests", "cpp", "veneer_example")
    if not os.path.exists(exe):
        exe = _build_veneer_example()
    out = subprocess.run([exe, "300007"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "veneer example: OK" in out.stdout, out.stdout + out.stderr
