"""Single-box multi-GPU SortPairs / SortKeys: one process per GPU, globally sorted across ranks.

New functionality (BASELINE.json configs[4]) -- the reference is a single-GPU library and has
nothing here.  The single-GPU building block is the DeviceRadixSort drop-in (cub/device/device_radix_sort.cuh).

Global order = the STABLE sort of the rank-order concatenation of the shards: (key, source rank, index on the
source rank).  Algorithm (sample sort around the local LSD sort):

  1. sample      every rank takes `s` regularly spaced keys of its (unsorted) shard; all_gather; every rank sorts
                 the same G*s (key, source rank) pairs with the local stable sort and picks the same G-1 splitters
                 (key, rank) -- ties on the key are broken by the source rank so that long runs of equal keys are
                 spread over several destinations without breaking stability;
  2. count       b2s_split_count: local keys per destination; all_gather -> G x G count matrix -> receive offsets;
  3. exchange    b2s_split_scatter: stable partition of the shard by destination.
                   exchange="peer": the partition kernel stores every item straight into the destination rank's
                                    receive buffer (CUDA-IPC mapped peer memory, NVLink stores) -- partition and
                                    all-to-all are ONE kernel, nothing is staged;
                   exchange="nccl": partition into a local bucketed buffer, then torch all_to_all_single (NCCL);
                 either way a rank receives G runs in source-rank order;
  4. local sort  one stable DeviceRadixSort (DoubleBuffer form) over what was received.  Runs arrive in rank order
                 and the sort is stable, so equal keys end up in (rank, index) order.

Two hosts drive the same kernels:
  * `NativeDistributedSorter` (the product path on GPUs): a thin binding of the C++ host inside libb2s.so
    (include/b2s_mgpu.h, cub_b200/csrc/b2s_mgpu.cu): NCCL called from C++ for the metadata, cudaIpc* peer mappings,
    everything enqueued on one stream; torch.distributed is used once, to hand the NCCL unique id to the ranks.
  * `DistributedSorter`: the same algorithm orchestrated from Python over torch.distributed (metadata, and the payload
    for exchange="nccl").  All device work goes through the C-ABI of libb2s.so (`LocalOps`); tests substitute a CPU
    `LocalOps` built on the oracle to exercise this host logic under gloo -- the product has no CPU path.
"""
from __future__ import annotations

import ctypes
import time
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .device_radix_sort import KEY_BYTES, DeviceRadixSort, DoubleBuffer, key_type_of

MAX_RANKS = 8
_CONTAINER = {1: torch.int8, 2: torch.int16, 4: torch.int32, 8: torch.int64}
_NP_BITS = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}
# category of b2s_key_t: 0 unsigned, 1 signed, 2 floating
_CATEGORY = [0, 1, 0, 1, 2, 2, 0, 1, 2, 0, 1, 2]


def sort_key(raw: int, key_type: int, descending: bool = False, begin_bit: int = 0, end_bit: Optional[int] = None) -> int:
    """Host restatement of the comparable form of a key: bits [begin_bit, end_bit) of the bit-ordered transform
    (cub/util_type.cuh:1031,1078,1179 TwiddleIn; radix_rank_sort_operations.cuh:592-599 descending; :79-89 -0.0).
    Used for splitter bookkeeping and the cross-rank boundary check; the device equivalent is SplitterOp."""
    bits = KEY_BYTES[key_type] * 8
    if end_bit is None:
        end_bit = bits
    ones = (1 << bits) - 1
    high = 1 << (bits - 1)
    k = int(raw) & ones
    cat = _CATEGORY[key_type]
    if cat == 2:
        zero_from, zero_to = (0, high) if descending else (high, 0)
        if k == zero_from:
            k = zero_to
        k ^= ones if (k & high) else high
        if descending:
            k ^= ones
    else:
        if cat == 1:
            k ^= high
        if descending:
            k ^= ones
    return (k >> begin_bit) & ((1 << (end_bit - begin_bit)) - 1)


def choose_splitters(sorted_keys: np.ndarray, sorted_ranks: np.ndarray, world: int, per_rank: int):
    """Splitters j = 1..world-1 are the samples at positions j * per_rank of the sorted (key, rank) sample list."""
    idx = [j * per_rank for j in range(1, world)]
    return sorted_keys[idx].copy(), sorted_ranks[idx].astype(np.int32).copy()


def exchange_plan(count_matrix: np.ndarray, rank: int):
    """count_matrix[src][dst] = items src sends to dst.  Returns (send_counts, send_offsets_local, recv_counts,
    recv_total, peer_offsets) where peer_offsets[dst] = position of THIS rank's segment in dst's receive buffer
    (segments are laid out in source-rank order, which is what keeps the final stable sort globally stable)."""
    c = np.asarray(count_matrix, dtype=np.int64)
    send_counts = c[rank].copy()
    send_offsets = np.concatenate(([0], np.cumsum(send_counts)[:-1]))
    recv_counts = c[:, rank].copy()
    peer_offsets = c[:rank, :].sum(axis=0) if rank > 0 else np.zeros(c.shape[1], dtype=np.int64)
    return send_counts, send_offsets, recv_counts, int(recv_counts.sum()), peer_offsets.astype(np.int64)


class LocalOps:
    """Device operations of one rank through the C-ABI (the product path: CUDA only, fails loudly otherwise)."""

    def __init__(self, device: torch.device):
        if device.type != "cuda":
            raise RuntimeError("cub_b200.multi_gpu needs CUDA devices; there is no CPU path")
        self.lib = _lib.load()
        self.device = device
        self.launches = 0

    # -- helpers
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def empty(self, n: int, dtype: torch.dtype) -> torch.Tensor:
        return torch.empty(n, dtype=dtype, device=self.device)

    def synchronize(self):
        torch.cuda.current_stream(self.device).synchronize()

    # -- local stable sort (DoubleBuffer form), returns the tensors holding the result
    def sort_db(self, kbufs, vbufs, n, key_type, descending, begin_bit, end_bit, temp_holder: dict):
        dk = DoubleBuffer(kbufs[0], kbufs[1])
        dv = DoubleBuffer(vbufs[0], vbufs[1]) if vbufs is not None else None
        fn = (DeviceRadixSort.SortPairsDescending if descending else DeviceRadixSort.SortPairs) if dv is not None else \
            (DeviceRadixSort.SortKeysDescending if descending else DeviceRadixSort.SortKeys)
        args = (dk, dv, n) if dv is not None else (dk, n)
        kw = dict(begin_bit=begin_bit, end_bit=end_bit, key_type=key_type)
        if dv is not None:
            kw["value_bytes"] = vbufs[0].element_size()
        err, nbytes = fn(None, 0, *args, **kw)
        if err:
            raise RuntimeError(f"temp-storage query failed: cudaError {err}")
        temp = temp_holder.get("sort")
        if temp is None or temp.numel() < nbytes:
            temp = temp_holder["sort"] = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        err, _ = fn(temp, temp.numel(), *args, **kw)
        if err:
            raise RuntimeError(f"radix sort failed: cudaError {err}")
        self.launches += self.lib.b2s_last_launch_count()
        return dk.Current(), (dv.Current() if dv is not None else None)

    @staticmethod
    def _dptr(t: Optional[torch.Tensor]):
        return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() else None

    def split_count(self, keys, n, key_type, descending, begin_bit, end_bit, sp_keys, sp_ranks, rank) -> torch.Tensor:
        """sp_keys (key container dtype) / sp_ranks (int32): DEVICE tensors.  Returns the int64 device tensor of counts."""
        ns = int(sp_keys.numel())
        counts = torch.zeros(ns + 1, dtype=torch.int64, device=self.device)
        err = self.lib.b2s_split_count(ctypes.c_void_p(keys.data_ptr()), n, key_type, int(descending), begin_bit, end_bit,
                                       self._dptr(sp_keys), self._dptr(sp_ranks), ns, rank,
                                       ctypes.c_void_p(counts.data_ptr()), self._stream())
        if err:
            raise RuntimeError(f"b2s_split_count failed: cudaError {err}")
        self.launches += 2
        return counts

    def split_scatter(self, keys, vals, out_keys, out_vals, n, key_type, descending, begin_bit, end_bit, sp_keys,
                      sp_ranks, rank, dest_offsets: torch.Tensor, peer_keys: Optional[Sequence[int]],
                      peer_vals: Optional[Sequence[int]], temp_holder: dict, peer_capacity: int = (1 << 64) - 1):
        """dest_offsets: int64 DEVICE tensor (items), one entry per destination."""
        ns = int(sp_keys.numel())
        vb = vals.element_size() if vals is not None else 0
        nb = ctypes.c_size_t(0)
        pkeys = pvals = None
        if peer_keys is not None:
            pkeys = (ctypes.c_void_p * MAX_RANKS)(*([int(p) for p in peer_keys] + [None] * (MAX_RANKS - len(peer_keys))))
            if vals is not None:
                pvals = (ctypes.c_void_p * MAX_RANKS)(*([int(p) for p in peer_vals] + [None] * (MAX_RANKS - len(peer_vals))))
        common = (ctypes.c_void_p(keys.data_ptr()), ctypes.c_void_p(out_keys.data_ptr()) if out_keys is not None else None,
                  ctypes.c_void_p(vals.data_ptr()) if vals is not None else None,
                  ctypes.c_void_p(out_vals.data_ptr()) if out_vals is not None else None,
                  n, key_type, vb, int(descending), begin_bit, end_bit, self._dptr(sp_keys), self._dptr(sp_ranks), ns, rank,
                  ctypes.c_void_p(dest_offsets.data_ptr()), pkeys, pvals, peer_capacity)
        err = self.lib.b2s_split_scatter(None, ctypes.byref(nb), *common, None)
        if err:
            raise RuntimeError(f"b2s_split_scatter size query failed: cudaError {err}")
        temp = temp_holder.get("split")
        if temp is None or temp.numel() < nb.value:
            temp = temp_holder["split"] = torch.empty(nb.value, dtype=torch.uint8, device=self.device)
        nb = ctypes.c_size_t(temp.numel())
        ev = temp_holder.get("split_events")
        if ev is None:
            ev = temp_holder["split_events"] = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        err = self.lib.b2s_split_scatter(ctypes.c_void_p(temp.data_ptr()), ctypes.byref(nb), *common, self._stream())
        ev[1].record()
        if err:
            raise RuntimeError(f"b2s_split_scatter failed: cudaError {err}")
        self.launches += 3

    def check_sorted(self, keys, vals, n, key_type, descending, begin_bit, end_bit):
        res = torch.zeros(3, dtype=torch.int64, device=self.device)
        err = self.lib.b2s_check_sorted(ctypes.c_void_p(keys.data_ptr()), ctypes.c_void_p(vals.data_ptr()) if vals is not None else None,
                                        n, key_type, vals.element_size() if vals is not None else 0, int(descending),
                                        begin_bit, end_bit, ctypes.c_void_p(res.data_ptr()), self._stream())
        if err:
            raise RuntimeError(f"b2s_check_sorted failed: cudaError {err}")
        return [int(x) for x in res.cpu().tolist()]


    def check_stable(self, keys, vals, n, key_type, descending, begin_bit, end_bit) -> int:
        """Adjacent positions with equal sort keys whose values do not increase (0 for a stable sort of (key, index))."""
        res = torch.zeros(1, dtype=torch.int64, device=self.device)
        err = self.lib.b2s_check_stable(ctypes.c_void_p(keys.data_ptr()), ctypes.c_void_p(vals.data_ptr()), n, key_type,
                                        vals.element_size(), int(descending), begin_bit, end_bit,
                                        ctypes.c_void_p(res.data_ptr()), self._stream())
        if err:
            raise RuntimeError(f"b2s_check_stable failed: cudaError {err}")
        return int(res.item())


@dataclass
class SortedShard:
    """NOTE: keys / values are VIEWS of the sorter's receive buffers: the next sort() on any rank overwrites them
    (peers store into these buffers).  Clone what must outlive the next call."""
    keys: torch.Tensor            # this rank's part of the globally sorted sequence (length = count)
    values: Optional[torch.Tensor]
    count: int
    counts_all: List[int]         # output shard sizes of all ranks


def _share_cuda(t: torch.Tensor):
    """(cudaIpcMemHandle bytes of the allocation holding `t`, byte offset of `t` inside it, exporting device)."""
    desc = t.untyped_storage()._share_cuda_()
    handle = bytes(desc[1])
    if len(handle) == 66:  # torch >= 2.5 prefixes {version, kind}: kind 'c' = cudaMalloc block, 'e' = expandable segment
        if handle[1:2] == b"e":
            raise RuntimeError("peer exchange needs plain cudaMalloc allocations: run without "
                               "PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True")
        handle = handle[2:]
    if len(handle) != 64:
        raise RuntimeError(f"unexpected CUDA IPC handle length {len(handle)}")
    return {"handle": handle, "offset": int(desc[3]) + t.storage_offset() * t.element_size(), "device": int(desc[0])}


class DistributedSorter:
    """Pre-allocates everything for repeated sorts of `n_local` items per rank (bench / serving loop)."""

    def __init__(self, n_local: int, key_dtype: torch.dtype, value_dtype: Optional[torch.dtype] = None, group=None,
                 descending: bool = False, begin_bit: int = 0, end_bit: Optional[int] = None,
                 samples_per_rank: int = 8192, slack: float = 1.10, exchange: str = "auto", ops: Optional[LocalOps] = None,
                 device: Optional[torch.device] = None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > MAX_RANKS:
            raise ValueError(f"at most {MAX_RANKS} ranks (one NVSwitch box)")
        self.key_type = key_type_of(key_dtype)
        self.kbytes = KEY_BYTES[self.key_type]
        if self.kbytes not in (4, 8):
            raise TypeError("multi-GPU sort supports 4- and 8-byte keys")
        self.key_dtype, self.value_dtype = key_dtype, value_dtype
        if value_dtype is not None and torch.empty(0, dtype=value_dtype).element_size() not in (4, 8):
            raise TypeError("multi-GPU sort supports 4- and 8-byte values (or keys only)")
        self.descending, self.begin_bit = descending, begin_bit
        self.end_bit = self.kbytes * 8 if end_bit is None else end_bit
        if ops is None:
            if device is None:
                device = torch.device("cuda", torch.cuda.current_device())
            ops = LocalOps(device)
        self.ops = ops
        self.device = self.ops.device
        self.n_local = n_local
        self.samples_per_rank = samples_per_rank
        self.capacity = int(n_local * slack) + 1024  # the same on every rank (n_local is the common maximum shard size)
        self.temp: dict = {}
        self._fence = self.ops.empty(1, torch.int32).zero_()
        self._phase_ms: dict = {}
        self._launches = 0
        kc, vc = _CONTAINER[self.kbytes], (value_dtype if value_dtype is not None else None)
        self.recv_k = [self.ops.empty(self.capacity, kc) for _ in range(2)]
        self.recv_v = [self.ops.empty(self.capacity, vc) for _ in range(2)] if vc is not None else None
        if exchange == "auto":
            exchange = "peer" if (self.device.type == "cuda" and self.world > 1) else "nccl"
        self.exchange = exchange
        self.part_k = self.part_v = None
        self.peer_k = self.peer_v = None
        if exchange == "peer":
            self._map_peers()
        else:
            self.part_k = self.ops.empty(n_local, kc)
            self.part_v = self.ops.empty(n_local, vc) if vc is not None else None

    # ---- CUDA-IPC mapping of every rank's receive buffers (peer exchange)
    def _map_peers(self):
        lib = self.ops.lib
        mine = {"k": _share_cuda(self.recv_k[0]), "v": _share_cuda(self.recv_v[0]) if self.recv_v is not None else None}
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        self._ipc_bases = {}

        def open_peer(d):
            base = self._ipc_bases.get(d["handle"])
            if base is None:
                err = lib.b2s_enable_peer_access(d["device"])
                if err:
                    raise RuntimeError(f"no peer access to device {d['device']}: cudaError {err}")
                p = ctypes.c_void_p()
                err = lib.b2s_ipc_open(d["handle"], ctypes.byref(p))
                if err:
                    raise RuntimeError(f"cudaIpcOpenMemHandle failed: cudaError {err}")
                base = self._ipc_bases[d["handle"]] = p.value
            return base + d["offset"]

        self.peer_k, self.peer_v = [], []
        for r, d in enumerate(everyone):
            if r == self.rank:
                self.peer_k.append(self.recv_k[0].data_ptr())
                self.peer_v.append(self.recv_v[0].data_ptr() if self.recv_v is not None else 0)
            else:
                self.peer_k.append(open_peer(d["k"]))
                self.peer_v.append(open_peer(d["v"]) if d["v"] is not None else 0)
        dist.barrier(group=self.group)

    def close(self):
        """Unmap the peer buffers (collective: every rank must call it before any rank frees its buffers)."""
        if getattr(self, "_ipc_bases", None):
            self.ops.synchronize()
            dist.barrier(group=self.group)
            for base in self._ipc_bases.values():
                self.ops.lib.b2s_ipc_close(ctypes.c_void_p(base))
            self._ipc_bases = {}
            dist.barrier(group=self.group)

    # ---- bookkeeping
    def last_phase_ms(self):
        return dict(self._phase_ms)

    def launches_per_sort(self):
        return self._launches

    def _container(self, t: torch.Tensor) -> torch.Tensor:
        return t.view(_CONTAINER[t.element_size()]) if t.dtype not in (torch.int8, torch.int16, torch.int32, torch.int64) else t

    # ---- the sort
    def sort(self, keys: torch.Tensor, values: Optional[torch.Tensor] = None) -> SortedShard:
        ops, G, me = self.ops, self.world, self.rank
        n = keys.numel()
        if n > self.n_local:
            raise ValueError("shard larger than the size this sorter was built for")
        if (values is None) != (self.value_dtype is None):
            raise ValueError("values must be given iff the sorter was built with a value dtype")
        kt, desc, bb, eb = self.key_type, self.descending, self.begin_bit, self.end_bit
        kin = self._container(keys)
        vin = values
        launches0 = ops.launches
        t0 = time.perf_counter()

        # 1. samples -> splitters (identical on every rank)
        # Every rank contributes exactly `samples_per_rank` regularly spaced samples whatever its shard size (indices
        # repeat when n < samples_per_rank), so no rank needs to know the others' sizes before sampling and every
        # call issues the same sequence of collectives -- shard sizes may change freely between calls.
        if n < 1:
            raise ValueError("every rank must hold at least one item")
        s = self.samples_per_rank
        pick = (torch.arange(s, device=self.device, dtype=torch.int64) * n) // s
        sample = kin[pick].contiguous()
        gathered = ops.empty(G * s, kin.dtype)
        dist.all_gather_into_tensor(gathered, sample, group=self.group)
        src = torch.arange(G, dtype=torch.int32, device=self.device).repeat_interleave(s)
        sk, sr = ops.sort_db([gathered, torch.empty_like(gathered)], [src, torch.empty_like(src)], G * s, kt, desc, bb, eb,
                             self.temp)
        idx = torch.arange(1, G, device=self.device, dtype=torch.int64) * s
        sp_keys, sp_ranks = sk[idx].contiguous(), sr[idx].contiguous()  # stay on the device: nobody waits for them
        t1 = time.perf_counter()

        # 2. counts -> exchange plan.  Offsets are computed on the device, so the partition kernel is enqueued before the
        #    host looks at the count matrix (which it needs only for the size of the final local sort).
        cm = ops.split_count(kin, n, kt, desc, bb, eb, sp_keys, sp_ranks, me)
        rows = ops.empty(G * G, torch.int64)
        dist.all_gather_into_tensor(rows, cm, group=self.group)
        dmatrix = rows.view(G, G)

        # 3. partition + exchange
        if self.exchange == "peer":
            # No barrier is needed BEFORE the stores: they are stream-ordered after the all_gathers above, which complete
            # only once every peer's stream has reached them, i.e. has finished reading its receive buffer (the local
            # sort of the previous call).  AFTER them a tiny stream-ordered all_reduce is the fence: it completes on this
            # rank only when every rank's partition kernel -- all stores into this rank's buffer -- has completed.
            peer_offsets_dev = dmatrix[:me].sum(dim=0) if me > 0 else torch.zeros(G, dtype=torch.int64, device=self.device)
            ops.split_scatter(kin, vin, None, None, n, kt, desc, bb, eb, sp_keys, sp_ranks, me, peer_offsets_dev,
                              self.peer_k, self.peer_v if vin is not None else None, self.temp, self.capacity)
            dist.all_reduce(self._fence, group=self.group)
            matrix = dmatrix.cpu().numpy()  # the host catches up while the partition kernel runs
            send_counts, send_offsets, recv_counts, total, _peer_offsets = exchange_plan(matrix, me)
        else:
            matrix = dmatrix.cpu().numpy()
            send_counts, send_offsets, recv_counts, total, _peer_offsets = exchange_plan(matrix, me)
            worst = int(matrix.sum(axis=0).max())
            if worst > self.capacity:  # the same matrix on every rank: every rank raises BEFORE the collective
                raise RuntimeError(f"receive capacity {self.capacity} too small for {worst} items; raise `slack`")
            send_offsets_dev = torch.from_numpy(np.ascontiguousarray(send_offsets, dtype=np.int64)).to(self.device)
            ops.split_scatter(kin, vin, self.part_k, self.part_v, n, kt, desc, bb, eb, sp_keys, sp_ranks, me,
                              send_offsets_dev, None, None, self.temp)
            dist.all_to_all_single(self.recv_k[0][:total], self.part_k[:n], [int(c) for c in recv_counts],
                                   [int(c) for c in send_counts], group=self.group)
            if vin is not None:
                dist.all_to_all_single(self._container(self.recv_v[0])[:total], self._container(self.part_v)[:n],
                                       [int(c) for c in recv_counts], [int(c) for c in send_counts], group=self.group)
        out_counts = [int(matrix[:, d].sum()) for d in range(G)]
        if max(out_counts) > self.capacity:
            raise RuntimeError(f"receive capacity {self.capacity} too small for {max(out_counts)} items; raise `slack` "
                               "(items beyond the capacity were dropped by the partition kernel, nothing was overrun)")
        t2 = t3 = time.perf_counter()

        # 4. final local stable sort over the G received runs
        if total > 0:
            ok, ov = ops.sort_db(self.recv_k, self.recv_v, total, kt, desc, bb, eb, self.temp)
        else:
            ok, ov = self.recv_k[0], (self.recv_v[0] if self.recv_v is not None else None)
        ops.synchronize()
        t4 = time.perf_counter()
        self._phase_ms = {"splitters": (t1 - t0) * 1e3, "count+plan": (t2 - t1) * 1e3, "partition+exchange": (t3 - t2) * 1e3,
                          "local_sort": (t4 - t3) * 1e3, "exchange": self.exchange}
        ev = self.temp.get("split_events")
        if ev is not None:
            self._phase_ms["partition_kernel_device_ms"] = ev[0].elapsed_time(ev[1])
            self._phase_ms["items_sent_to_peers"] = int(send_counts.sum() - send_counts[me])
        self._launches = ops.launches - launches0
        out_k = ok[:total].view(self.key_dtype) if ok.dtype != self.key_dtype else ok[:total]
        return SortedShard(out_k, ov[:total] if ov is not None else None, total, out_counts)

    def verify(self, keys_in: torch.Tensor, values_in: Optional[torch.Tensor], out: SortedShard,
               values_are_global_indices: bool = False) -> bool:
        return verify_global_sort(self.ops, self.group, self.world, self._container(keys_in), values_in,
                                  self._container(out.keys), out.values, out.count, self.key_type, self.descending,
                                  self.begin_bit, self.end_bit, values_are_global_indices)


def verify_global_sort(ops, group, world, keys_in, values_in, keys_out, values_out, count, key_type, descending, begin_bit,
                       end_bit, values_are_global_indices=False) -> bool:
    """Verification at any size (collective): local order, cross-rank boundaries, global multiset of keys and of
    (key, value) pairs.  With `values_are_global_indices` (values increase along the rank-order concatenation of the
    input, e.g. rank * n + i) also STABILITY: inside every run of equal sort keys the values increase, within a rank
    (device kernel) and across rank boundaries (last item of a rank against the first item of the next)."""
    G = world
    device = ops.device
    kt, desc, bb, eb = key_type, descending, begin_bit, end_bit
    inv_out, ksum_out, psum_out = ops.check_sorted(keys_out, values_out, count, kt, desc, bb, eb) if count else (0, 0, 0)
    _inv, ksum_in, psum_in = ops.check_sorted(keys_in, values_in, keys_in.numel(), kt, desc, bb, eb)
    unstable = 0
    stable_check = values_are_global_indices and values_out is not None and hasattr(ops, "check_stable")
    if stable_check and count > 1:
        unstable = ops.check_stable(keys_out, values_out, count, kt, desc, bb, eb)
    sums = torch.tensor([ksum_in, psum_in, ksum_out, psum_out, keys_in.numel(), count], dtype=torch.int64, device=device)
    dist.all_reduce(sums, group=group)  # int64 wraps mod 2^64 like the checksums
    sums = sums.cpu().tolist()
    ok = inv_out == 0 and unstable == 0 and sums[0] == sums[2] and sums[1] == sums[3] and sums[4] == sums[5]
    edge = torch.zeros(5, dtype=torch.int64, device=device)
    if count:
        edge[0] = 1
        edge[1] = keys_out[0].to(torch.int64)
        edge[2] = keys_out[count - 1].to(torch.int64)
        if values_out is not None:
            edge[3] = values_out[0].to(torch.int64)
            edge[4] = values_out[count - 1].to(torch.int64)
    edges = [torch.zeros_like(edge) for _ in range(G)]
    dist.all_gather(edges, edge, group=group)
    last = None
    vmask = (1 << (8 * values_out.element_size())) - 1 if values_out is not None else 0
    for e in edges:
        has, first_raw, last_raw, first_val, last_val = (int(x) for x in e.cpu().tolist())
        if not has:
            continue
        if last is not None:
            a, b = sort_key(last[0], kt, desc, bb, eb), sort_key(first_raw, kt, desc, bb, eb)
            if a > b or (stable_check and a == b and (last[1] & vmask) >= (first_val & vmask)):
                ok = False
        last = (last_raw, last_val)
    flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(flag.item())


class _DevicePointerView:
    """Zero-copy torch view of device memory owned by libb2s.so (through __cuda_array_interface__)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


_TYPESTR = {torch.int32: "<i4", torch.int64: "<i8"}


class NativeDistributedSorter:
    """The product multi-GPU sorter: binding of the C++ host in libb2s.so (include/b2s_mgpu.h).  One process per GPU;
    `group` (any torch.distributed backend) is used once, to broadcast the NCCL unique id, and by verify()."""

    def __init__(self, n_local: int, key_dtype: torch.dtype, value_dtype: Optional[torch.dtype] = None, group=None,
                 descending: bool = False, begin_bit: int = 0, end_bit: Optional[int] = None, samples_per_rank: int = 8192,
                 slack: float = 1.10, device: Optional[torch.device] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.key_type = key_type_of(key_dtype)
        self.kbytes = KEY_BYTES[self.key_type]
        self.key_dtype, self.value_dtype = key_dtype, value_dtype
        self.vbytes = torch.empty(0, dtype=value_dtype).element_size() if value_dtype is not None else 0
        self.descending, self.begin_bit = descending, begin_bit
        self.end_bit = self.kbytes * 8 if end_bit is None else end_bit
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.ops = LocalOps(device)
        self.device = device
        self.lib = self.ops.lib
        self.n_local = n_local
        self.exchange = "peer"
        ident = [None]
        if self.rank == 0:
            buf = ctypes.create_string_buffer(128)
            err = self.lib.b2s_mgpu_unique_id(buf)
            if err:
                raise RuntimeError(f"b2s_mgpu_unique_id failed: {err} (2001 = libnccl.so.2 not found)")
            ident[0] = bytes(buf.raw)
        if self.world > 1:
            dist.broadcast_object_list(ident, src=0, group=group)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(device):
            err = self.lib.b2s_mgpu_create(ctypes.byref(self._h), ident[0], self.rank, self.world, n_local, self.key_type,
                                           self.vbytes, int(descending), begin_bit, self.end_bit, float(slack), samples_per_rank)
        if err:
            msg = self.lib.b2s_mgpu_last_error(self._h).decode() if self._h else ""
            raise RuntimeError(f"b2s_mgpu_create failed: {err} {msg}")
        self.capacity = int(self.lib.b2s_mgpu_capacity(self._h))
        self._launches = 0

    def close(self):
        if self._h:
            with torch.cuda.device(self.device):
                self.lib.b2s_mgpu_destroy(self._h)
            self._h = ctypes.c_void_p()

    def _container(self, t: torch.Tensor) -> torch.Tensor:
        return t.view(_CONTAINER[t.element_size()]) if t.dtype not in (torch.int8, torch.int16, torch.int32, torch.int64) else t

    def sort(self, keys: torch.Tensor, values: Optional[torch.Tensor] = None) -> SortedShard:
        """Collective.  Returns views of the library's receive buffers (see SortedShard)."""
        n = keys.numel()
        if (values is None) != (self.value_dtype is None):
            raise ValueError("values must be given iff the sorter was built with a value dtype")
        if keys.device != self.device or (values is not None and values.device != self.device):
            raise ValueError("shards must live on the sorter's device")
        ko, vo = ctypes.c_void_p(), ctypes.c_void_p()
        cnt = ctypes.c_uint64(0)
        counts = (ctypes.c_uint64 * MAX_RANKS)()
        with torch.cuda.device(self.device):
            err = self.lib.b2s_mgpu_sort(self._h, ctypes.c_void_p(keys.data_ptr()),
                                         ctypes.c_void_p(values.data_ptr()) if values is not None else None, n,
                                         ctypes.byref(ko), ctypes.byref(vo), ctypes.byref(cnt), counts,
                                         ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if err:
            raise RuntimeError(f"b2s_mgpu_sort failed: {err} {self.lib.b2s_mgpu_last_error(self._h).decode()}")
        total = int(cnt.value)
        kc = _CONTAINER[self.kbytes]
        out_k = torch.as_tensor(_DevicePointerView(ko.value, max(total, 1), _TYPESTR[kc]), device=self.device)[:total]
        out_k = out_k.view(self.key_dtype) if out_k.dtype != self.key_dtype else out_k
        out_v = None
        if values is not None:
            vc = _CONTAINER[self.vbytes]
            out_v = torch.as_tensor(_DevicePointerView(vo.value, max(total, 1), _TYPESTR[vc]), device=self.device)[:total]
            out_v = out_v.view(self.value_dtype) if out_v.dtype != self.value_dtype else out_v
        self._launches = 1 + 1 + 8 + 2 + 3 + 1 + (1 + 1 + self.kbytes)  # kernels + memsets enqueued per call (nominal)
        return SortedShard(out_k, out_v, total, [int(counts[d]) for d in range(self.world)])

    def last_phase_ms(self):
        ms = (ctypes.c_float * 6)()
        sent = ctypes.c_uint64(0)
        err = self.lib.b2s_mgpu_last_phases(self._h, ms, ctypes.byref(sent))
        if err:
            return {}
        return {"splitters": ms[0], "count+plan": ms[1], "partition_kernel_device_ms": ms[2], "fence": ms[3],
                "local_sort": ms[4], "whole_call_device_ms": ms[5], "items_sent_to_peers": int(sent.value),
                "exchange": "peer (C++ host, bulk shared->peer copies)", "timing": "CUDA events on the sort's stream"}

    def launches_per_sort(self):
        return self._launches

    def verify(self, keys_in: torch.Tensor, values_in: Optional[torch.Tensor], out: SortedShard,
               values_are_global_indices: bool = False) -> bool:
        return verify_global_sort(self.ops, self.group, self.world, self._container(keys_in), values_in,
                                  self._container(out.keys), out.values, out.count, self.key_type, self.descending,
                                  self.begin_bit, self.end_bit, values_are_global_indices)
