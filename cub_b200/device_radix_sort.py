"""Host-side mirror of ``cub::DeviceRadixSort`` over the B200-native C-ABI.

Reference interface being mirrored (same entry-point names, argument order and meaning,
two-phase temp-storage protocol, DoubleBuffer semantics):
  cub/device/device_radix_sort.cuh:312   SortPairs            (pointer form)
  cub/device/device_radix_sort.cuh:781   SortPairs            (DoubleBuffer form)
  cub/device/device_radix_sort.cuh:1214  SortPairsDescending  (pointer) / :1675 (DoubleBuffer)
  cub/device/device_radix_sort.cuh:2106  SortKeys             (pointer) / :2525 (DoubleBuffer)
  cub/device/device_radix_sort.cuh:2921  SortKeysDescending   (pointer) / :3330 (DoubleBuffer)
  cub/util_type.cuh:854-886              DoubleBuffer<T>

Differences forced by Python: ``temp_storage_bytes`` cannot be passed by reference, so every
entry point returns ``(cuda_error, temp_storage_bytes)``; device arrays are ``torch`` CUDA
tensors (or raw integer device pointers together with ``key_type=`` / ``value_bytes=``).
PyTorch is used for device memory and streams only -- all compute is in ``libb2s.so``; there
is no fallback path.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib

# b2s_key_t (include/b2s_radix_sort.h)
KEY_TYPES = {
    torch.uint8: 0, torch.int8: 1, torch.uint16: 2, torch.int16: 3, torch.float16: 4, torch.bfloat16: 5,
    torch.uint32: 6, torch.int32: 7, torch.float32: 8, torch.uint64: 9, torch.int64: 10, torch.float64: 11,
}
KEY_BYTES = [1, 1, 2, 2, 2, 2, 4, 4, 4, 8, 8, 8]
CUDA_SUCCESS = 0


def key_type_of(dtype: torch.dtype) -> int:
    try:
        return KEY_TYPES[dtype]
    except KeyError:
        raise TypeError(f"unsupported radix-sort key dtype {dtype}") from None


class DoubleBuffer:
    """Mirror of ``cub::DoubleBuffer<T>`` (cub/util_type.cuh:854-886): two device buffers and
    a selector naming the currently valid one."""

    def __init__(self, d_current=None, d_alternate=None):
        self.d_buffers = [d_current, d_alternate]
        self.selector = 0

    def Current(self):
        return self.d_buffers[self.selector]

    def Alternate(self):
        return self.d_buffers[self.selector ^ 1]


def _ptr(x) -> Optional[int]:
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            raise ValueError("device arrays must be CUDA tensors (no CPU path)")
        return x.data_ptr()
    return int(x)


def _stream_handle(stream, device=None) -> int:
    if stream is None:
        return torch.cuda.current_stream(device).cuda_stream
    if isinstance(stream, torch.cuda.Stream):
        return stream.cuda_stream
    return int(stream)


def _device_of(*objs):
    """The CUDA device the tensors among `objs` live on (None when only raw pointers are given); the C-ABI works on the
    CURRENT device, so callers run it under ``torch.cuda.device(dev)`` and take the stream from that device.  Tensors on
    different devices are an error (the kernels would fault or silently go through peer access)."""
    dev = None
    for o in objs:
        if isinstance(o, torch.Tensor) and o.is_cuda:
            if dev is not None and o.device != dev:
                raise ValueError(f"all device arrays of one sort must be on one device (got {dev} and {o.device})")
            dev = o.device
    return dev


class _on_device:
    def __init__(self, dev):
        self.ctx = torch.cuda.device(dev) if dev is not None else None

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def _infer(keys, values, key_type, value_bytes) -> Tuple[int, int]:
    if key_type is None:
        if not isinstance(keys, torch.Tensor):
            raise TypeError("key_type= is required with raw device pointers")
        key_type = key_type_of(keys.dtype)
    if value_bytes is None:
        if values is None:
            value_bytes = 0
        elif isinstance(values, torch.Tensor):
            value_bytes = values.element_size()
        else:
            raise TypeError("value_bytes= is required with raw device pointers")
    return key_type, value_bytes


def _offset_bytes(num_items: int) -> int:
    # mirrors detail::ChooseOffsetT (cub/detail/choose_offset.cuh:44-57) for a Python int
    return 4 if num_items < (1 << 32) else 8


def _call_ptr(descending, d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out,
              num_items, begin_bit, end_bit, stream, key_type, value_bytes):
    lib = _lib.load()
    key_type, value_bytes = _infer(d_keys_in if d_keys_in is not None else d_keys_out, d_values_in, key_type, value_bytes)
    if end_bit is None:
        end_bit = KEY_BYTES[key_type] * 8
    nbytes = ctypes.c_size_t(int(temp_storage_bytes or 0))
    dev = _device_of(d_temp_storage, d_keys_in, d_keys_out, d_values_in, d_values_out)
    with _on_device(dev):
        err = lib.b2s_radix_sort(_ptr(d_temp_storage), ctypes.byref(nbytes), _ptr(d_keys_in), _ptr(d_keys_out),
                                 _ptr(d_values_in), _ptr(d_values_out), int(num_items), key_type, value_bytes,
                                 _offset_bytes(int(num_items)), int(bool(descending)), int(begin_bit), int(end_bit),
                                 _stream_handle(stream, dev) if d_temp_storage is not None else None)
    return err, nbytes.value


def _call_db(descending, d_temp_storage, temp_storage_bytes, d_keys: DoubleBuffer, d_values: Optional[DoubleBuffer],
             num_items, begin_bit, end_bit, stream, key_type, value_bytes):
    lib = _lib.load()
    key_type, value_bytes = _infer(d_keys.d_buffers[0], d_values.d_buffers[0] if d_values is not None else None,
                                   key_type, value_bytes)
    if end_bit is None:
        end_bit = KEY_BYTES[key_type] * 8
    nbytes = ctypes.c_size_t(int(temp_storage_bytes or 0))
    kb = (ctypes.c_void_p * 2)(_ptr(d_keys.d_buffers[0]), _ptr(d_keys.d_buffers[1]))
    ksel = ctypes.c_int(d_keys.selector)
    if d_values is not None and value_bytes:
        vb = (ctypes.c_void_p * 2)(_ptr(d_values.d_buffers[0]), _ptr(d_values.d_buffers[1]))
        vsel = ctypes.c_int(d_values.selector)
        vb_arg, vsel_arg = vb, ctypes.byref(vsel)
    else:
        vsel = ctypes.c_int(0)
        vb_arg, vsel_arg = None, None
    dev = _device_of(d_temp_storage, *d_keys.d_buffers, *(d_values.d_buffers if d_values is not None else ()))
    with _on_device(dev):
        err = lib.b2s_radix_sort_db(_ptr(d_temp_storage), ctypes.byref(nbytes), kb, ctypes.byref(ksel), vb_arg, vsel_arg,
                                    int(num_items), key_type, value_bytes, _offset_bytes(int(num_items)),
                                    int(bool(descending)), int(begin_bit), int(end_bit),
                                    _stream_handle(stream, dev) if d_temp_storage is not None else None)
    if err == CUDA_SUCCESS and d_temp_storage is not None:
        d_keys.selector = ksel.value
        if d_values is not None and value_bytes:
            d_values.selector = vsel.value
    return err, nbytes.value


class DeviceRadixSort:
    """Static entry points with the reference's names.  Each accepts either the pointer form
    ``(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out[, d_values_in, d_values_out], num_items, ...)``
    or the DoubleBuffer form ``(d_temp_storage, temp_storage_bytes, d_keys[, d_values], num_items, ...)`` and
    returns ``(cudaError, temp_storage_bytes)``.  ``d_temp_storage=None`` only sizes the temp storage."""

    @staticmethod
    def _pairs(descending, d_temp_storage, temp_storage_bytes, *args, begin_bit=0, end_bit=None, stream=None,
               key_type=None, value_bytes=None):
        if isinstance(args[0], DoubleBuffer):
            d_keys, d_values, num_items = args[:3]
            rest = args[3:]
            begin_bit, end_bit, stream = DeviceRadixSort._opt(rest, begin_bit, end_bit, stream)
            return _call_db(descending, d_temp_storage, temp_storage_bytes, d_keys, d_values, num_items, begin_bit,
                            end_bit, stream, key_type, value_bytes)
        d_keys_in, d_keys_out, d_values_in, d_values_out, num_items = args[:5]
        begin_bit, end_bit, stream = DeviceRadixSort._opt(args[5:], begin_bit, end_bit, stream)
        return _call_ptr(descending, d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_values_in,
                         d_values_out, num_items, begin_bit, end_bit, stream, key_type, value_bytes)

    @staticmethod
    def _keys(descending, d_temp_storage, temp_storage_bytes, *args, begin_bit=0, end_bit=None, stream=None,
              key_type=None):
        if isinstance(args[0], DoubleBuffer):
            d_keys, num_items = args[:2]
            begin_bit, end_bit, stream = DeviceRadixSort._opt(args[2:], begin_bit, end_bit, stream)
            return _call_db(descending, d_temp_storage, temp_storage_bytes, d_keys, None, num_items, begin_bit,
                            end_bit, stream, key_type, 0)
        d_keys_in, d_keys_out, num_items = args[:3]
        begin_bit, end_bit, stream = DeviceRadixSort._opt(args[3:], begin_bit, end_bit, stream)
        return _call_ptr(descending, d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, None, None,
                         num_items, begin_bit, end_bit, stream, key_type, 0)

    @staticmethod
    def _opt(rest, begin_bit, end_bit, stream):
        # positional begin_bit, end_bit, stream exactly as in the reference signatures
        if len(rest) > 0:
            begin_bit = rest[0]
        if len(rest) > 1:
            end_bit = rest[1]
        if len(rest) > 2:
            stream = rest[2]
        return begin_bit, end_bit, stream

    @staticmethod
    def SortPairs(d_temp_storage, temp_storage_bytes, *args, **kw):
        return DeviceRadixSort._pairs(False, d_temp_storage, temp_storage_bytes, *args, **kw)

    @staticmethod
    def SortPairsDescending(d_temp_storage, temp_storage_bytes, *args, **kw):
        return DeviceRadixSort._pairs(True, d_temp_storage, temp_storage_bytes, *args, **kw)

    @staticmethod
    def SortKeys(d_temp_storage, temp_storage_bytes, *args, **kw):
        return DeviceRadixSort._keys(False, d_temp_storage, temp_storage_bytes, *args, **kw)

    @staticmethod
    def SortKeysDescending(d_temp_storage, temp_storage_bytes, *args, **kw):
        return DeviceRadixSort._keys(True, d_temp_storage, temp_storage_bytes, *args, **kw)


class DeviceSegmentedRadixSort:
    """Mirror of ``cub::DeviceSegmentedRadixSort`` (cub/device/device_segmented_radix_sort.cuh), pointer form:
    ``(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out[, d_values_in, d_values_out], num_items, num_segments,
    d_begin_offsets, d_end_offsets, begin_bit=0, end_bit=None, stream=None)`` -> ``(cudaError, temp_storage_bytes)``.
    The offsets are int32 / int64 CUDA tensors (begin and end may be views ``offsets[:-1]`` / ``offsets[1:]``)."""

    @staticmethod
    def _run(descending, d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out, num_items,
             num_segments, d_begin_offsets, d_end_offsets, begin_bit=0, end_bit=None, stream=None, key_type=None, value_bytes=None):
        lib = _lib.load()
        key_type, value_bytes = _infer(d_keys_in, d_values_in, key_type, value_bytes)
        if end_bit is None:
            end_bit = KEY_BYTES[key_type] * 8
        if d_begin_offsets.dtype != d_end_offsets.dtype or d_begin_offsets.dtype not in (torch.int32, torch.int64):
            raise TypeError("segment offsets must be int32 or int64 CUDA tensors of one dtype")
        nbytes = ctypes.c_size_t(int(temp_storage_bytes or 0))
        dev = _device_of(d_temp_storage, d_keys_in, d_keys_out, d_values_in, d_values_out, d_begin_offsets, d_end_offsets)
        with _on_device(dev):
            err = lib.b2s_segmented_radix_sort(_ptr(d_temp_storage), ctypes.byref(nbytes), _ptr(d_keys_in), _ptr(d_keys_out),
                                               _ptr(d_values_in), _ptr(d_values_out), int(num_items), int(num_segments),
                                               _ptr(d_begin_offsets), _ptr(d_end_offsets), d_begin_offsets.element_size(),
                                               key_type, value_bytes, int(bool(descending)), int(begin_bit), int(end_bit),
                                               _stream_handle(stream, dev) if d_temp_storage is not None else None)
        return err, nbytes.value

    @staticmethod
    def SortPairs(d_temp_storage, temp_storage_bytes, *args, **kw):
        return DeviceSegmentedRadixSort._run(False, d_temp_storage, temp_storage_bytes, *args, **kw)

    @staticmethod
    def SortPairsDescending(d_temp_storage, temp_storage_bytes, *args, **kw):
        return DeviceSegmentedRadixSort._run(True, d_temp_storage, temp_storage_bytes, *args, **kw)

    @staticmethod
    def SortKeys(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, *args, **kw):
        return DeviceSegmentedRadixSort._run(False, d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, None, None, *args, **kw)

    @staticmethod
    def SortKeysDescending(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, *args, **kw):
        return DeviceSegmentedRadixSort._run(True, d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, None, None, *args, **kw)


class CudaError(RuntimeError):
    pass


def _check(err: int, what: str):
    if err != CUDA_SUCCESS:
        raise CudaError(f"{what} failed with cudaError {err}")


def sort_pairs(keys: torch.Tensor, values: Optional[torch.Tensor], descending: bool = False, begin_bit: int = 0,
               end_bit: Optional[int] = None, stream=None):
    """Convenience wrapper: allocate outputs + temp storage and run the pointer form."""
    n = keys.numel()
    keys_out = torch.empty_like(keys)
    values_out = torch.empty_like(values) if values is not None else None
    fn = DeviceRadixSort.SortPairsDescending if descending else DeviceRadixSort.SortPairs
    if values is None:
        fn = DeviceRadixSort.SortKeysDescending if descending else DeviceRadixSort.SortKeys
        args = (keys, keys_out, n)
    else:
        args = (keys, keys_out, values, values_out, n)
    err, nbytes = fn(None, 0, *args, begin_bit=begin_bit, end_bit=end_bit, stream=stream)
    _check(err, "temp-storage query")
    temp = torch.empty(nbytes, dtype=torch.uint8, device=keys.device)
    err, _ = fn(temp, nbytes, *args, begin_bit=begin_bit, end_bit=end_bit, stream=stream)
    _check(err, "radix sort")
    return keys_out, values_out


def segmented_sort_pairs(keys: torch.Tensor, values: Optional[torch.Tensor], offsets: torch.Tensor, descending: bool = False,
                         begin_bit: int = 0, end_bit: Optional[int] = None, stream=None):
    """Convenience wrapper: segment s is [offsets[s], offsets[s + 1]); allocates outputs + temp storage."""
    n, nseg = keys.numel(), offsets.numel() - 1
    keys_out = keys.clone()  # items outside every segment keep their input value
    values_out = values.clone() if values is not None else None
    fn = DeviceSegmentedRadixSort.SortPairsDescending if descending else DeviceSegmentedRadixSort.SortPairs
    args = (keys, keys_out, values, values_out, n, nseg, offsets[:-1], offsets[1:])
    err, nbytes = fn(None, 0, *args, begin_bit=begin_bit, end_bit=end_bit)
    _check(err, "temp-storage query")
    temp = torch.empty(nbytes, dtype=torch.uint8, device=keys.device)
    err, _ = fn(temp, nbytes, *args, begin_bit=begin_bit, end_bit=end_bit, stream=stream)
    _check(err, "segmented radix sort")
    return keys_out, values_out


def sort_keys(keys: torch.Tensor, descending: bool = False, begin_bit: int = 0, end_bit: Optional[int] = None,
              stream=None) -> torch.Tensor:
    return sort_pairs(keys, None, descending, begin_bit, end_bit, stream)[0]


def bind_host_to_gpu_numa_node(device_index: int) -> dict:
    """Pin the calling process to the CPU cores of the NUMA node the GPU hangs off, BEFORE pinned host buffers are
    allocated (first touch then places them in that node's memory): host<->device copies of the end-to-end path then do
    not cross the socket interconnect.  Host-side deployment plumbing, no effect on results.  Returns what was done
    ({"numa_node": n, "cpus": k} or {"skipped": reason}); never raises."""
    import os

    try:
        prop = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return {"skipped": "no NUMA affinity reported for " + bdf}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return {"skipped": f"no allowed CPU on NUMA node {node}"}
        os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed), "pci": bdf}
    except Exception as e:  # noqa: BLE001
        return {"skipped": str(e)[:100]}


class HostSorter:
    """End-to-end path with HOST buffers: pinned host -> device, DoubleBuffer sort, device -> pinned host.
    Device buffers, temp storage and pinned result buffers are allocated once and reused (what a caller of the
    C-ABI does).  With ``depth > 1`` consecutive calls are pipelined over three streams (upload / sort / download) and
    ``depth`` buffer sets, so the upload of call i+1 overlaps the download of call i (PCIe is full duplex); every call
    still uploads its own inputs and downloads its own results.  A call returns its slot's pinned result tensors, valid
    after ``synchronize()`` (or after ``depth`` further calls have been waited for, see ``__call__``)."""

    def __init__(self, n: int, key_dtype: torch.dtype, value_dtype: Optional[torch.dtype], device="cuda:0",
                 descending: bool = False, begin_bit: int = 0, end_bit: Optional[int] = None, depth: int = 1):
        self.n, self.descending, self.begin_bit, self.end_bit, self.depth = n, descending, begin_bit, end_bit, depth
        dev = self.device = torch.device(device)
        self._fn = (DeviceRadixSort.SortPairsDescending if descending else DeviceRadixSort.SortPairs) \
            if value_dtype is not None else \
            (DeviceRadixSort.SortKeysDescending if descending else DeviceRadixSort.SortKeys)
        self.slots = []
        for _ in range(depth):
            k = [torch.empty(n, dtype=key_dtype, device=dev) for _ in range(2)]
            v = [torch.empty(n, dtype=value_dtype, device=dev) for _ in range(2)] if value_dtype is not None else None
            self.slots.append({"k": k, "v": v, "hk": torch.empty(n, dtype=key_dtype).pin_memory(),
                               "hv": torch.empty(n, dtype=value_dtype).pin_memory() if value_dtype is not None else None,
                               "done": torch.cuda.Event()})
        s0 = self.slots[0]
        dk, dv = DoubleBuffer(*s0["k"]), (DoubleBuffer(*s0["v"]) if s0["v"] else None)
        args = (dk, dv, n) if dv is not None else (dk, n)
        err, self.temp_bytes = self._fn(None, 0, *args, begin_bit=begin_bit, end_bit=end_bit)
        _check(err, "temp-storage query")
        self.temp = torch.empty(self.temp_bytes, dtype=torch.uint8, device=dev)  # sorts are serialised on one stream
        self.s_up, self.s_sort, self.s_down = (torch.cuda.Stream(dev) for _ in range(3)) if depth > 1 else (None,) * 3
        self._calls = 0
        # compatibility with the single-slot attribute names
        self.h_keys_out, self.h_vals_out = s0["hk"], s0["hv"]

    def _sort(self, slot, stream):
        dk = DoubleBuffer(*slot["k"])
        dv = DoubleBuffer(*slot["v"]) if slot["v"] else None
        args = (dk, dv, self.n) if dv is not None else (dk, self.n)
        err, _ = self._fn(self.temp, self.temp_bytes, *args, begin_bit=self.begin_bit, end_bit=self.end_bit, stream=stream)
        _check(err, "radix sort")
        return dk.Current(), (dv.Current() if dv is not None else None)

    def __call__(self, h_keys: torch.Tensor, h_values: Optional[torch.Tensor]):
        slot = self.slots[self._calls % self.depth]
        self._calls += 1
        if self.depth == 1:
            slot["k"][0].copy_(h_keys, non_blocking=True)
            if slot["v"]:
                slot["v"][0].copy_(h_values, non_blocking=True)
            ok, ov = self._sort(slot, None)
            slot["hk"].copy_(ok, non_blocking=True)
            if ov is not None:
                slot["hv"].copy_(ov, non_blocking=True)
            return slot["hk"], slot["hv"]
        slot["done"].synchronize()  # the slot's previous results have been downloaded (and may now be overwritten)
        with torch.cuda.stream(self.s_up):
            slot["k"][0].copy_(h_keys, non_blocking=True)
            if slot["v"]:
                slot["v"][0].copy_(h_values, non_blocking=True)
            up = self.s_up.record_event()
        self.s_sort.wait_event(up)
        with torch.cuda.stream(self.s_sort):
            ok, ov = self._sort(slot, self.s_sort)
            sorted_ev = self.s_sort.record_event()
        self.s_down.wait_event(sorted_ev)
        with torch.cuda.stream(self.s_down):
            slot["hk"].copy_(ok, non_blocking=True)
            if ov is not None:
                slot["hv"].copy_(ov, non_blocking=True)
            slot["done"].record(self.s_down)
        return slot["hk"], slot["hv"]

    def synchronize(self):
        if self.depth == 1:
            torch.cuda.current_stream(self.device).synchronize()
        else:
            for s in (self.s_up, self.s_sort, self.s_down):
                s.synchronize()


def sort_pairs_host(h_keys: torch.Tensor, h_values: Optional[torch.Tensor], descending: bool = False,
                    begin_bit: int = 0, end_bit: Optional[int] = None, device="cuda:0"):
    """One-shot host-buffer sort (allocates; use HostSorter to amortise allocations)."""
    s = HostSorter(h_keys.numel(), h_keys.dtype, h_values.dtype if h_values is not None else None, device,
                   descending, begin_bit, end_bit)
    k, v = s(h_keys, h_values)
    torch.cuda.synchronize()
    return k.clone(), (v.clone() if v is not None else None)
