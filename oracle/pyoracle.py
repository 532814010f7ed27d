"""oracle/pyoracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

ctypes loaders for the checkers:
  * liboracle.so            CPU restatement (radix_oracle.c) + host std::stable_sort harness port
  * _ref/libref_cub.so      the unmodified reference cub::DeviceRadixSort (GPU), built from /root/reference
  * _ref/libtk_cub.so       CUDA-toolkit CUB (informational comparator, GPU)
Only tests/, __graft_entry__.smoke() and bench.py may import this module.  cub_b200/ never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CPU_LIB = os.path.join(_HERE, "liboracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libref_cub.so")
TK_LIB = os.path.join(_HERE, "_ref", "libtk_cub.so")

# b2s_key_t order (include/b2s_radix_sort.h); numpy view dtype of the raw bits
KEY_NAMES = ["u8", "i8", "u16", "i16", "f16", "bf16", "u32", "i32", "f32", "u64", "i64", "f64"]
KEY_BYTES = [1, 1, 2, 2, 2, 2, 4, 4, 4, 8, 8, 8]
KEY_CATEGORY = [0, 1, 0, 1, 2, 2, 0, 1, 2, 0, 1, 2]  # 0 unsigned, 1 signed, 2 floating (b2s_key_t order)
BITS_DTYPE = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}

_cpu = None


def build_cpu():
    subprocess.check_call(["make", "-s", "-C", _HERE, "cpu"])


def cpu() -> ctypes.CDLL:
    global _cpu
    if _cpu is None:
        if not os.path.exists(CPU_LIB):
            build_cpu()
        lib = ctypes.CDLL(CPU_LIB)
        c = ctypes
        lib.oracle_radix_sort.restype = c.c_int
        lib.oracle_radix_sort.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint64, c.c_int,
                                          c.c_int, c.c_int, c.c_int, c.c_int]
        lib.oracle_histogram.restype = c.c_int
        lib.oracle_histogram.argtypes = [c.c_void_p, c.c_uint64, c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p]
        lib.oracle_sort_key.restype = c.c_uint64
        lib.oracle_sort_key.argtypes = [c.c_uint64, c.c_int, c.c_int, c.c_int, c.c_int]
        lib.oracle_twiddle_in.restype = c.c_uint64
        lib.oracle_twiddle_in.argtypes = [c.c_uint64, c.c_int, c.c_int]
        lib.oracle_twiddle_out.restype = c.c_uint64
        lib.oracle_twiddle_out.argtypes = [c.c_uint64, c.c_int, c.c_int]
        lib.oracle_digit_source.restype = c.c_uint64
        lib.oracle_digit_source.argtypes = [c.c_uint64, c.c_int]
        lib.host_stable_sort_solution.restype = c.c_int
        lib.host_stable_sort_solution.argtypes = [c.c_void_p, c.c_uint64, c.c_int, c.c_int, c.c_int, c.c_int,
                                                  c.c_void_p, c.c_void_p, c.c_int]
        lib.host_max_threads.restype = c.c_int
        _cpu = lib
    return _cpu


def _np_ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def radix_sort(keys_bits: np.ndarray, values: np.ndarray | None, key_type: int, descending=False, begin_bit=0,
               end_bit=None):
    """CPU oracle. keys_bits: raw key bits (uint8/16/32/64 view). values: any fixed-width array or None."""
    if end_bit is None:
        end_bit = KEY_BYTES[key_type] * 8
    keys_bits = np.ascontiguousarray(keys_bits)
    assert keys_bits.dtype.itemsize == KEY_BYTES[key_type]
    n = keys_bits.shape[0]
    kout = np.empty_like(keys_bits)
    vout = None
    vb = 0
    if values is not None:
        values = np.ascontiguousarray(values)
        vb = values.dtype.itemsize * (values.shape[1] if values.ndim == 2 else 1)
        vout = np.empty_like(values)
    rc = cpu().oracle_radix_sort(_np_ptr(keys_bits), _np_ptr(kout), _np_ptr(values), _np_ptr(vout), n, key_type, vb,
                                 int(descending), begin_bit, end_bit)
    assert rc == 0
    return kout, vout


def histogram(keys_bits: np.ndarray, key_type: int, descending=False, begin_bit=0, end_bit=None):
    if end_bit is None:
        end_bit = KEY_BYTES[key_type] * 8
    passes = (end_bit - begin_bit + 7) // 8
    out = np.zeros((max(passes, 0), 256), dtype=np.uint64)
    keys_bits = np.ascontiguousarray(keys_bits)
    rc = cpu().oracle_histogram(_np_ptr(keys_bits), keys_bits.shape[0], key_type, int(descending), begin_bit, end_bit,
                                _np_ptr(out))
    assert rc == 0
    return out


def host_stable_sort(keys_bits: np.ndarray, key_type: int, descending=False, begin_bit=0, end_bit=None, threads=1):
    """The reference test harness' host solution (typed compare + std::stable_sort)."""
    if end_bit is None:
        end_bit = KEY_BYTES[key_type] * 8
    keys_bits = np.ascontiguousarray(keys_bits)
    n = keys_bits.shape[0]
    kout = np.empty_like(keys_bits)
    ranks = np.empty(n, dtype=np.uint64)
    rc = cpu().host_stable_sort_solution(_np_ptr(keys_bits), n, key_type, int(descending), begin_bit, end_bit,
                                         _np_ptr(kout), _np_ptr(ranks), threads)
    assert rc == 0
    return kout, ranks


def host_max_threads() -> int:
    return cpu().host_max_threads()


def load_gpu_reference(which: str = "ref"):
    """Load the reference-CUB shim (GPU). Returns a ctypes lib exposing <prefix>_radix_sort[_db] with the
    signatures of include/b2s_radix_sort.h, or None when the prebuilt library is absent."""
    path, prefix = (REF_LIB, "ref_cub") if which == "ref" else (TK_LIB, "tk_cub")
    if not os.path.exists(path):
        return None
    from cub_b200 import _lib as product_binding  # prototypes only; no product code is executed

    lib = ctypes.CDLL(path)
    product_binding.bind(lib, prefix)
    lib.sort = getattr(lib, prefix + "_radix_sort")
    lib.sort_db = getattr(lib, prefix + "_radix_sort_db")
    c = ctypes
    if hasattr(lib, prefix + "_struct_sort"):  # oracle/ref_shim_ext.cu (reference build only)
        lib.struct_sort = getattr(lib, prefix + "_struct_sort")
        lib.struct_sort.restype = c.c_int
        lib.struct_sort.argtypes = [c.c_void_p, c.POINTER(c.c_size_t), c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint64,
                                    c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p]
        lib.segmented_sort = getattr(lib, prefix + "_segmented_sort")
        lib.segmented_sort.restype = c.c_int
        lib.segmented_sort.argtypes = [c.c_void_p, c.POINTER(c.c_size_t), c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint64,
                                       c.c_int, c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p]
        if hasattr(lib, prefix + "_tuned_sort"):
            lib.tuned_sort = getattr(lib, prefix + "_tuned_sort")
            lib.tuned_sort.restype = c.c_int
            lib.tuned_sort.argtypes = [c.c_void_p, c.POINTER(c.c_size_t), c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint64,
                                       c.c_int, c.c_int, c.c_int, c.c_void_p]
        lib.sort128 = getattr(lib, prefix + "_sort128")
        lib.sort128.restype = c.c_int
        lib.sort128.argtypes = [c.c_void_p, c.POINTER(c.c_size_t), c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint64,
                                c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p]
    return lib


# ---------------------------------------------------------------------------------------------------------------------
# Decomposer (custom struct key) semantics -- checker for SURVEY.md §8(f)2 (product: cub_b200/csrc/b2s_struct.cu).
# Reference: the decomposer overloads of cub::DeviceRadixSort (cub/device/device_radix_sort.cuh:486-530, 625-666, ...):
# a key is a tuple of arithmetic fields; the sort key is the concatenation of the fields' bit-ordered images, first tuple
# element most significant, LAST element holding bit 0 (test/catch2_test_device_radix_sort_custom.cu:1100-1115);
# [begin_bit, end_bit) indexes that concatenation; every field is transformed like a fundamental key (util_type.cuh:1031,
# 1078, 1179), -0.0 compares equal to +0.0 in a floating field (radix_rank_sort_operations.cuh:79-89), descending inverts
# the whole image (:592-599); the sort is stable.
# ---------------------------------------------------------------------------------------------------------------------
def _ordered_field(bits: np.ndarray, key_type: int, descending: bool = False) -> np.ndarray:
    """Bit-ordered image of one field as uint64 (only the low 8*bytes bits are used), complemented when descending.
    Floating fields: -0.0 and +0.0 share one image, HIGH in both directions, exactly as in the reference's onesweep path
    (radix_rank_sort_operations.cuh:55-66, 79-89: ascending -0.0 is mapped onto +0.0, descending the complemented +0.0 onto
    the complemented -0.0).  The reference's single-tile path (n <= 4864) maps onto +0.0 BEFORE reversing instead; the two
    agree unless a partial bit range cuts through a floating field of a descending sort, and none of the 16 known-answer
    vectors does (their zeros come with the full range)."""
    nbits = KEY_BYTES[key_type] * 8
    ones = (1 << nbits) - 1
    high = 1 << (nbits - 1)
    k = bits.astype(np.uint64)
    cat = KEY_CATEGORY[key_type]
    if cat == 1:
        k = k ^ np.uint64(high)
    elif cat == 2:
        sign = (k >> np.uint64(nbits - 1)) & np.uint64(1)
        k = k ^ np.where(sign == 1, np.uint64(ones), np.uint64(high))
    if descending:
        k = k ^ np.uint64(ones)
    if cat == 2:
        k = np.where(k == np.uint64(ones ^ high), np.uint64(high), k)
    return k


def decomposed_sort_permutation(fields, descending=False, begin_bit=0, end_bit=None):
    """fields: list of (raw bits array, key_type), most significant first.  Returns the stable permutation `perm` such that
    records[perm] is the sorted sequence over bits [begin_bit, end_bit) of the concatenated image."""
    widths = [KEY_BYTES[kt] * 8 for _, kt in fields]
    total = sum(widths)
    if end_bit is None:
        end_bit = total
    keys = []
    lo = 0  # bit offset of the current field inside the concatenation, starting from the LAST field
    for (bits, kt), w in zip(reversed(fields), reversed(widths)):
        img = _ordered_field(np.ascontiguousarray(bits), kt, descending)
        b, e = max(begin_bit, lo), min(end_bit, lo + w)
        if e > b:
            mask = ((1 << (e - lo)) - 1) ^ ((1 << (b - lo)) - 1)
            keys.append(img & np.uint64(mask))
        else:
            keys.append(np.zeros_like(img))
        lo += w
    # np.lexsort: last key is the primary one and the sort is stable; `keys` runs from least to most significant field
    return np.lexsort(tuple(keys))
