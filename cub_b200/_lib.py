"""ctypes binding of the product C-ABI (include/b2s_radix_sort.h -> cub_b200/libb2s.so).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C cub_b200/csrc``.
There is NO fallback: if the CUDA library is missing, importing a sort entry point raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2s.so")

_c = ctypes
_SORT_ARGS = [
    _c.c_void_p, _c.POINTER(_c.c_size_t),            # d_temp_storage, temp_storage_bytes
    _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,  # keys_in, keys_out, values_in, values_out
    _c.c_uint64, _c.c_int, _c.c_int, _c.c_int,        # num_items, key_type, value_bytes, offset_bytes
    _c.c_int, _c.c_int, _c.c_int, _c.c_void_p,        # descending, begin_bit, end_bit, stream
]
_SORT_DB_ARGS = [
    _c.c_void_p, _c.POINTER(_c.c_size_t),
    _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_int),    # key_bufs[2], key_selector
    _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_int),    # val_bufs[2], val_selector
    _c.c_uint64, _c.c_int, _c.c_int, _c.c_int,
    _c.c_int, _c.c_int, _c.c_int, _c.c_void_p,
]

# every symbol include/b2s_radix_sort.h declares: name -> (restype, argtypes)
EXPORTS = {
    "b2s_radix_sort": (_c.c_int, _SORT_ARGS),
    "b2s_radix_sort_db": (_c.c_int, _SORT_DB_ARGS),
    "b2s_radix_sort_struct": (_c.c_int, [_c.c_void_p, _c.POINTER(_c.c_size_t), _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                         _c.c_uint64, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                         _c.c_void_p]),
    "b2s_radix_sort_struct_db": (_c.c_int, [_c.c_void_p, _c.POINTER(_c.c_size_t), _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_int),
                                            _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_int), _c.c_uint64, _c.c_int, _c.c_void_p,
                                            _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p]),
    "b2s_segmented_radix_sort": (_c.c_int, [_c.c_void_p, _c.POINTER(_c.c_size_t), _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                            _c.c_uint64, _c.c_uint64, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int,
                                            _c.c_int, _c.c_int, _c.c_int, _c.c_void_p]),
    "b2s_segmented_radix_sort_db": (_c.c_int, [_c.c_void_p, _c.POINTER(_c.c_size_t), _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_int),
                                               _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_int), _c.c_uint64, _c.c_uint64, _c.c_void_p,
                                               _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p]),
    "b2s_key_bytes": (_c.c_int, [_c.c_int]),
    "b2s_version": (_c.c_char_p, []),
    "b2s_last_launch_count": (_c.c_int, []),
    "b2s_timing_enable": (_c.c_int, [_c.c_int]),
    "b2s_timing_read": (_c.c_int, [_c.POINTER(_c.c_float), _c.c_int]),
    "b2s_set_variant": (_c.c_int, [_c.c_int]),
    "b2s_describe_variant": (_c.c_int, [_c.c_int, _c.c_int, _c.c_int] + [_c.POINTER(_c.c_int)] * 4),
    "b2s_variant_mode": (_c.c_int, [_c.c_int, _c.c_int, _c.c_int]),
    "b2s_digit_histogram": (_c.c_int, [_c.c_void_p, _c.c_uint64, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p,
                                       _c.c_void_p]),
    "b2s_check_stable": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_uint64, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                    _c.c_void_p, _c.c_void_p]),
    "b2s_mgpu_unique_id": (_c.c_int, [_c.c_void_p]),
    "b2s_mgpu_create": (_c.c_int, [_c.POINTER(_c.c_void_p), _c.c_void_p, _c.c_int, _c.c_int, _c.c_uint64, _c.c_int, _c.c_int,
                                   _c.c_int, _c.c_int, _c.c_int, _c.c_double, _c.c_int]),
    "b2s_mgpu_sort": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_uint64, _c.POINTER(_c.c_void_p),
                                 _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_uint64), _c.POINTER(_c.c_uint64), _c.c_void_p]),
    "b2s_mgpu_last_phases": (_c.c_int, [_c.c_void_p, _c.POINTER(_c.c_float), _c.POINTER(_c.c_uint64)]),
    "b2s_mgpu_capacity": (_c.c_uint64, [_c.c_void_p]),
    "b2s_mgpu_last_error": (_c.c_char_p, [_c.c_void_p]),
    "b2s_mgpu_destroy": (_c.c_int, [_c.c_void_p]),
    "b2s_variant_flow": (_c.c_int, [_c.c_int, _c.c_int, _c.c_int]),
    "b2s_set_tile_claim": (_c.c_int, [_c.c_int]),
    "b2s_set_counting_sort": (_c.c_int, [_c.c_int]),
    "b2s_set_float_zero_recording": (_c.c_int, [_c.c_int]),
    "b2s_set_counting_min_items": (_c.c_uint64, [_c.c_int, _c.c_uint64]),
    "b2s_set_trace": (_c.c_int, [_c.c_void_p, _c.c_int]),
    "b2s_set_single_tile": (_c.c_int, [_c.c_int]),
    "b2s_enable_peer_access": (_c.c_int, [_c.c_int]),
    "b2s_ipc_open": (_c.c_int, [_c.c_char_p, _c.POINTER(_c.c_void_p)]),
    "b2s_ipc_close": (_c.c_int, [_c.c_void_p]),
    "b2s_split_count": (_c.c_int, [_c.c_void_p, _c.c_uint64, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p,
                                   _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p]),
    "b2s_split_scatter": (_c.c_int, [_c.c_void_p, _c.POINTER(_c.c_size_t), _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                     _c.c_void_p, _c.c_uint64, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                     _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p,
                                     _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_void_p), _c.c_uint64, _c.c_void_p]),
    "b2s_lower_bound": (_c.c_int, [_c.c_void_p, _c.c_uint64, _c.c_int, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p]),
    "b2s_fill_keys": (_c.c_int, [_c.c_void_p, _c.c_uint64, _c.c_int, _c.c_uint64, _c.c_int, _c.c_uint64, _c.c_void_p]),
    "b2s_fill_iota": (_c.c_int, [_c.c_void_p, _c.c_uint64, _c.c_int, _c.c_uint64, _c.c_void_p]),
    "b2s_check_sorted": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_uint64, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                    _c.c_int, _c.c_void_p, _c.c_void_p]),
}

_lib = None


def bind(lib: ctypes.CDLL, prefix: str = "b2s") -> ctypes.CDLL:
    """Attach prototypes. ``prefix`` lets the test-only reference shims (same signatures,
    ``ref_cub_*`` / ``tk_cub_*`` symbols) reuse the two sort prototypes."""
    for name, (res, args) in EXPORTS.items():
        sym = name if prefix == "b2s" else name.replace("b2s", prefix, 1)
        if prefix != "b2s" and name not in ("b2s_radix_sort", "b2s_radix_sort_db"):
            continue
        fn = getattr(lib, sym)
        fn.restype = res
        fn.argtypes = args
    return lib


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        path = os.environ.get("B2S_LIB", LIB_PATH)  # tuning builds (libb2s_tune.so) are selected explicitly
        if not os.path.exists(path):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C cub_b200/csrc`. cub_b200 has no CPU or PyTorch fallback."
            )
        _lib = bind(ctypes.CDLL(path))
    return _lib
