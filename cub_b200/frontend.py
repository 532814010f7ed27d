"""cub_b200.frontend -- the two callers SURVEY.md §8(f)4 names, on top of the DeviceRadixSort mirror:

* Thrust-style in-place entry points ``sort`` / ``sort_by_key`` / ``stable_sort`` / ``stable_sort_by_key``: what
  ``thrust::sort(first, last[, thrust::greater<T>()])`` and ``thrust::sort_by_key`` reach for primitive keys, where Thrust
  dispatches to ``cub::DeviceRadixSort`` with a DoubleBuffer and copies back when the selector flipped
  (reference call shape: ``cub/device/device_radix_sort.cuh:781-790`` DoubleBuffer ``SortPairs``; Thrust itself is not
  vendored in /root/reference).  Radix sort is stable, so the ``stable_`` names are aliases.
* PyTorch custom operators ``torch.ops.cub_b200.sort_pairs`` / ``sort_keys`` (``torch.library.custom_op`` with fake-tensor
  kernels, so they trace under ``torch.compile`` / FakeTensorMode) and a helper shaped like PyTorch's own stable sort, ``sort_with_indices``.

There is no CPU path: every entry point raises on non-CUDA tensors, and the library loader raises if ``libb2s.so`` is
missing.  Ordering semantics are CUB's, not those of PyTorch's built-in sort: floating keys are ordered by their transformed bit patterns
(-NaN first and +NaN last when ascending, -0.0 == +0.0 and stable between them; ``device_radix_sort.cuh:70-105``).
"""
from __future__ import annotations

import contextlib
from typing import Optional, Tuple

import torch

from .device_radix_sort import DeviceRadixSort, DoubleBuffer, _check, key_type_of, sort_pairs as _sort_pairs

__all__ = ["sort", "sort_by_key", "stable_sort", "stable_sort_by_key", "sort_with_indices"]

_VALUE_BYTES = (1, 2, 4, 8, 16)


def _require_cuda_1d(t: torch.Tensor, what: str) -> None:
    if not isinstance(t, torch.Tensor) or t.device.type != "cuda":
        raise ValueError(f"{what} must be a CUDA tensor (cub_b200 has no CPU path)")
    if t.dim() != 1 or not t.is_contiguous():
        raise ValueError(f"{what} must be a contiguous 1-D tensor")


def _check_values(keys: torch.Tensor, values: torch.Tensor) -> None:
    _require_cuda_1d(values, "values")
    if values.numel() != keys.numel() or values.device != keys.device:
        raise ValueError("keys and values must have the same length and device")
    if values.element_size() not in _VALUE_BYTES:
        raise ValueError(f"unsupported value width {values.element_size()} bytes")


# ---------------------------------------------------------------------------------------------------------------------
# Thrust-style: in place, nothing returned
# ---------------------------------------------------------------------------------------------------------------------
def sort_by_key(keys: torch.Tensor, values: Optional[torch.Tensor], descending: bool = False, stream=None) -> None:
    """In-place stable sort of ``keys`` (and ``values`` along with them), ascending unless ``descending``
    (``thrust::sort_by_key(k, k + n, v[, thrust::greater<T>()])``).  One alternate buffer per array is allocated for
    the DoubleBuffer form; if the sorted data ends up in the alternate it is copied back (a u32 sort of 4 passes does
    not: an even number of passes returns to the caller's buffer)."""
    _require_cuda_1d(keys, "keys")
    key_type_of(keys.dtype)  # raises for unsupported key types
    if values is not None:
        _check_values(keys, values)
    n = keys.numel()
    if n == 0:
        return
    dk = DoubleBuffer(keys, torch.empty_like(keys))
    dv = DoubleBuffer(values, torch.empty_like(values)) if values is not None else None
    if dv is not None:
        fn = DeviceRadixSort.SortPairsDescending if descending else DeviceRadixSort.SortPairs
        args = (dk, dv, n)
    else:
        fn = DeviceRadixSort.SortKeysDescending if descending else DeviceRadixSort.SortKeys
        args = (dk, n)
    if stream is not None and not isinstance(stream, torch.cuda.Stream):
        raise TypeError("stream must be a torch.cuda.Stream (the copy-back runs on it too)")
    with torch.cuda.device(keys.device), (torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()):
        err, nbytes = fn(None, 0, *args)
        _check(err, "temp-storage query")
        temp = torch.empty(nbytes, dtype=torch.uint8, device=keys.device)
        err, _ = fn(temp, nbytes, *args)  # on the current stream, like the copy-back below
        _check(err, "radix sort")
        if dk.selector != 0:
            keys.copy_(dk.Current())
        if dv is not None and dv.selector != 0:
            values.copy_(dv.Current())


def sort(keys: torch.Tensor, descending: bool = False, stream=None) -> None:
    """In-place sort of ``keys`` (``thrust::sort``)."""
    sort_by_key(keys, None, descending, stream)


stable_sort = sort
stable_sort_by_key = sort_by_key


# ---------------------------------------------------------------------------------------------------------------------
# PyTorch custom operators (functional: inputs are never written -- the pointer form of the reference API)
# ---------------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("cub_b200::sort_pairs", mutates_args=())
def _op_sort_pairs(keys: torch.Tensor, values: torch.Tensor, descending: bool = False, begin_bit: int = 0,
                   end_bit: int = -1) -> Tuple[torch.Tensor, torch.Tensor]:
    _require_cuda_1d(keys, "keys")
    _check_values(keys, values)
    with torch.cuda.device(keys.device):
        k, v = _sort_pairs(keys, values, descending, begin_bit, None if end_bit < 0 else end_bit)
    return k, v


@_op_sort_pairs.register_fake
def _(keys, values, descending=False, begin_bit=0, end_bit=-1):
    return torch.empty_like(keys), torch.empty_like(values)


@torch.library.custom_op("cub_b200::sort_keys", mutates_args=())
def _op_sort_keys(keys: torch.Tensor, descending: bool = False, begin_bit: int = 0, end_bit: int = -1) -> torch.Tensor:
    _require_cuda_1d(keys, "keys")
    with torch.cuda.device(keys.device):
        k, _ = _sort_pairs(keys, None, descending, begin_bit, None if end_bit < 0 else end_bit)
    return k


@_op_sort_keys.register_fake
def _(keys, descending=False, begin_bit=0, end_bit=-1):
    return torch.empty_like(keys)


def sort_with_indices(x: torch.Tensor, descending: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Helper shaped like PyTorch's built-in stable sort, for 1-D CUDA tensors: returns (sorted keys, int64 source indices).
    The indices travel through the sort as 4-byte values when they fit (then widened), as 8-byte values otherwise.
    For integer keys the result equals that of PyTorch's built-in stable sort; for floating keys NaNs
    and signed zeros follow CUB's rule (module docstring)."""
    _require_cuda_1d(x, "x")
    n = x.numel()
    if n < (1 << 31):
        idx = torch.arange(n, dtype=torch.int32, device=x.device)
    else:
        idx = torch.arange(n, dtype=torch.int64, device=x.device)
    k, v = torch.ops.cub_b200.sort_pairs(x, idx, descending, 0, -1)
    return k, v.to(torch.int64)
