"""cub_b200 -- B200-native (sm_100a) drop-in for the cub::DeviceRadixSort hot path.

Layout:
  csrc/                 hand-written CUDA kernels + the C-ABI (include/b2s_radix_sort.h)
  device_radix_sort.py  host-side mirror of cub::DeviceRadixSort (same entry points / argument meaning)
  multi_gpu.py          single-box multi-GPU SortPairs: binding of the C++ host in libb2s.so (include/b2s_mgpu.h) + a
                        torch.distributed-orchestrated twin used by the CPU (gloo) tests
  frontend.py           Thrust-style in-place sort / sort_by_key and the torch.ops.cub_b200.* custom operators
"""
from .device_radix_sort import (  # noqa: F401
    DeviceRadixSort,
    DeviceSegmentedRadixSort,
    DoubleBuffer,
    KEY_TYPES,
    key_type_of,
    sort_keys,
    sort_pairs,
    sort_pairs_host,
    segmented_sort_pairs,
)

from .frontend import sort, sort_by_key, sort_with_indices, stable_sort, stable_sort_by_key  # noqa: E402,F401

__all__ = [
    "sort",
    "sort_by_key",
    "stable_sort",
    "stable_sort_by_key",
    "sort_with_indices",
    "DeviceRadixSort",
    "DeviceSegmentedRadixSort",
    "DoubleBuffer",
    "KEY_TYPES",
    "key_type_of",
    "sort_keys",
    "sort_pairs",
    "sort_pairs_host",
    "segmented_sort_pairs",
]
