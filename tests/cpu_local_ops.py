"""TEST INFRASTRUCTURE: a CPU stand-in for cub_b200.multi_gpu.LocalOps built on the oracle, so that the host
logic of the multi-GPU sort (splitter choice, count matrix, exchange plan, all-to-all plumbing, verification)
runs under gloo without a GPU.  The product never imports this."""
from __future__ import annotations

import numpy as np
import torch

from cub_b200 import multi_gpu as mg
from oracle import pyoracle as po

_NP_BITS = {4: np.uint32, 8: np.uint64}
_NP_SIGNED = {4: np.int32, 8: np.int64}


def _bits(t: torch.Tensor) -> np.ndarray:
    return t.numpy().view(_NP_BITS[t.element_size()])


def _mix(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


class CpuOps:
    def __init__(self):
        self.device = torch.device("cpu")
        self.launches = 0

    def empty(self, n, dtype):
        return torch.zeros(n, dtype=dtype)

    def synchronize(self):
        pass

    def sort_db(self, kbufs, vbufs, n, key_type, descending, begin_bit, end_bit, temp_holder):
        k = _bits(kbufs[0])[:n].copy()
        v = vbufs[0].numpy()[:n].copy() if vbufs is not None else None
        ko, vo = po.radix_sort(k, v, key_type, descending, begin_bit, end_bit)
        _bits(kbufs[1])[:n] = ko
        if vbufs is not None:
            vbufs[1].numpy()[:n] = vo
        self.launches += 1
        return kbufs[1], (vbufs[1] if vbufs is not None else None)

    def _dest(self, keys, n, key_type, descending, begin_bit, end_bit, sp_keys, sp_ranks, rank):
        sp_keys = _bits(sp_keys) if isinstance(sp_keys, torch.Tensor) else sp_keys
        sp_ranks = sp_ranks.numpy() if isinstance(sp_ranks, torch.Tensor) else sp_ranks
        sk = [mg.sort_key(int(k), key_type, descending, begin_bit, end_bit) for k in sp_keys]
        dest = np.zeros(n, dtype=np.int64)
        raw = _bits(keys)[:n]
        for i in range(n):
            o = mg.sort_key(int(raw[i]), key_type, descending, begin_bit, end_bit)
            dest[i] = sum(1 for j in range(len(sk)) if o > sk[j] or (o == sk[j] and int(sp_ranks[j]) <= rank))
        return dest

    def split_count(self, keys, n, key_type, descending, begin_bit, end_bit, sp_keys, sp_ranks, rank):
        dest = self._dest(keys, n, key_type, descending, begin_bit, end_bit, sp_keys, sp_ranks, rank)
        return torch.from_numpy(np.bincount(dest, minlength=len(sp_keys) + 1).astype(np.int64))

    def split_scatter(self, keys, vals, out_keys, out_vals, n, key_type, descending, begin_bit, end_bit, sp_keys,
                      sp_ranks, rank, dest_offsets, peer_keys, peer_vals, temp_holder, peer_capacity=None):
        assert peer_keys is None, "the CPU stand-in only does the bucketed (all_to_all) exchange"
        dest = self._dest(keys, n, key_type, descending, begin_bit, end_bit, sp_keys, sp_ranks, rank)
        order = np.argsort(dest, kind="stable")
        pos = dest_offsets.numpy().astype(np.int64) if isinstance(dest_offsets, torch.Tensor) else np.asarray(dest_offsets, dtype=np.int64)
        counts = np.bincount(dest, minlength=len(sp_keys) + 1)
        idx = 0
        for d in range(len(counts)):
            seg = order[idx: idx + counts[d]]
            out_keys.numpy()[pos[d]: pos[d] + counts[d]] = keys.numpy()[:n][seg]
            if vals is not None:
                out_vals.numpy()[pos[d]: pos[d] + counts[d]] = vals.numpy()[:n][seg]
            idx += counts[d]

    def check_sorted(self, keys, vals, n, key_type, descending, begin_bit, end_bit):
        raw = _bits(keys)[:n]
        sk = np.array([mg.sort_key(int(k), key_type, descending, begin_bit, end_bit) for k in raw], dtype=np.uint64)
        inv = int(np.count_nonzero(sk[:-1] > sk[1:])) if n > 1 else 0
        with np.errstate(over="ignore"):
            ksum = int(_mix(raw).sum(dtype=np.uint64).view(np.int64)) if n else 0
            psum = 0
            if vals is not None and n:
                v = vals.numpy()[:n].view(_NP_BITS[vals.element_size()]).astype(np.uint64)
                psum = int(_mix(raw.astype(np.uint64) ^ (v << np.uint64(32))).sum(dtype=np.uint64).view(np.int64))
        return [inv, ksum, psum]
