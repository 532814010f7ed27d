"""GPU tests of the multi-GPU SortPairs building blocks.
 * one GPU: b2s_split_count / b2s_split_scatter through the C-ABI against a numpy restatement of the destination rule
   (arbitrary splitters, tie bits, directions, bit ranges), bit-exact and stable;
 * >= 2 GPUs (skipped otherwise): the whole DistributedSorter over NCCL, both exchange modes, against the oracle's
   stable sort of the rank-order concatenation."""
import ctypes
import os
import tempfile

import numpy as np
import pytest
import torch

from cub_b200 import multi_gpu as mg
from tests import harness as H

pytestmark = pytest.mark.gpu


def _dest_np(raw, kt, desc, bb, eb, sp_keys, sp_ranks, rank):
    sk = [mg.sort_key(int(k), kt, desc, bb, eb) for k in sp_keys]
    o = np.array([mg.sort_key(int(k), kt, desc, bb, eb) for k in raw], dtype=np.uint64)
    dest = np.zeros(raw.shape[0], dtype=np.int64)
    for j, s in enumerate(sk):
        dest += (o > np.uint64(s)) | ((o == np.uint64(s)) & (int(sp_ranks[j]) <= rank))
    return dest


@pytest.mark.parametrize("kt,vb", [(6, 4), (9, 4), (8, 0), (10, 8), (11, 4), (7, 0)])
def test_split_kernels_vs_numpy(b2s, kt, vb):
    rng = np.random.default_rng(kt * 7 + vb)
    kb = H.KEY_BYTES[kt]
    ops = mg.LocalOps(torch.device("cuda", 0))
    for n, desc, bb, eb, nsp in ((100_003, False, 0, kb * 8, 7), (65_536, True, 0, kb * 8, 3), (30_001, False, 3, kb * 8 - 5, 1),
                                 (12_345, False, 0, kb * 8, 0)):
        raw = H.random_bits(rng, n, kb)
        raw[::7] = raw[3]                       # plenty of ties with a splitter
        if kt in (8, 11):
            raw = H.spice_floats(raw, kb)
        sp_idx = np.sort(rng.integers(0, n, nsp))
        cand = raw[sp_idx] if nsp else raw[:0]
        if nsp:
            cand[0] = raw[3]
        order = np.argsort([mg.sort_key(int(k), kt, desc, bb, eb) for k in cand], kind="stable")
        sp_keys = cand[order]
        # splitters are (key, source rank) pairs ascending in that order (they come out of a stable sort of the samples)
        sp_ranks = np.sort(rng.integers(0, 8, nsp)).astype(np.int32)
        rank = 3
        vals = np.arange(n, dtype=H.NP_BITS[vb]) if vb else None
        dk, dv = H.to_dev(raw), (H.to_dev(vals) if vb else None)
        d_sp_keys = H.to_dev(sp_keys) if nsp else torch.empty(0, dtype=H.CONTAINER[kb], device="cuda")
        d_sp_ranks = torch.from_numpy(sp_ranks).cuda()
        counts = ops.split_count(dk, n, kt, desc, bb, eb, d_sp_keys, d_sp_ranks, rank).cpu().numpy()
        dest = _dest_np(raw, kt, desc, bb, eb, sp_keys, sp_ranks, rank)
        exp_counts = np.bincount(dest, minlength=nsp + 1)
        assert counts.tolist() == exp_counts.tolist()
        offs = torch.from_numpy(np.concatenate(([0], np.cumsum(exp_counts)[:-1])).astype(np.int64)).cuda()
        ok, ov = torch.empty_like(dk), (torch.empty_like(dv) if vb else None)
        ops.split_scatter(dk, dv, ok, ov, n, kt, desc, bb, eb, d_sp_keys, d_sp_ranks, rank, offs, None, None, {})
        # destinations that are only element-aligned take the item-store write-out instead of the bulk copies
        okm = torch.empty(n + 4, dtype=dk.dtype, device="cuda")
        ovm = torch.empty(n + 4, dtype=dv.dtype, device="cuda") if vb else None
        ops.split_scatter(dk, dv, okm[1:], ovm[1:] if vb else None, n, kt, desc, bb, eb, d_sp_keys, d_sp_ranks, rank, offs, None,
                          None, {})
        torch.cuda.synchronize()
        assert torch.equal(okm[1:n + 1], ok) and (not vb or torch.equal(ovm[1:n + 1], ov)), "item-store and bulk write-out differ"
        # PEER write-out path with every "peer" pointing at the local buffer, and the capacity guard
        ok2, ov2 = torch.zeros_like(dk), (torch.zeros_like(dv) if vb else None)
        ops.split_scatter(dk, dv, None, None, n, kt, desc, bb, eb, d_sp_keys, d_sp_ranks, rank, offs, [ok2.data_ptr()] * 8,
                          [ov2.data_ptr() if vb else 0] * 8, {}, peer_capacity=n - 1000)
        torch.cuda.synchronize()
        assert torch.equal(ok2[: n - 1000], ok[: n - 1000]) and bool((ok2[n - 1000:] == 0).all())
        if vb:
            assert torch.equal(ov2[: n - 1000], ov[: n - 1000]) and bool((ov2[n - 1000:] == 0).all())
        torch.cuda.synchronize()
        order = np.argsort(dest, kind="stable")
        assert np.array_equal(H.to_np(ok, raw.dtype), raw[order]), "partition is not the stable one"
        if vb:
            assert np.array_equal(H.to_np(ov, vals.dtype), vals[order])


def _worker(rank, world, path, exchange, ret):
    import torch.distributed as dist

    from oracle import pyoracle as po

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"file://{path}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        for kt, kd, n_local, desc, mode in ((9, torch.uint64, 300_007, False, "random"), (6, torch.uint32, 200_000, True, "dups"),
                                            (9, torch.uint64, 100_000, False, "equal")):
            kb = H.KEY_BYTES[kt]
            rng = np.random.default_rng(99)
            shards, vals = [], []
            for r in range(world):
                if mode == "dups":
                    k = rng.integers(0, 5, size=n_local, dtype=np.uint64)
                elif mode == "equal":
                    k = np.full(n_local, 11, dtype=np.uint64)
                else:
                    k = H.random_bits(rng, n_local, 8)
                shards.append(k.astype(H.NP_BITS[kb]))
                vals.append(np.arange(n_local, dtype=np.uint32) + np.uint32(10_000_000 * r))
            if exchange == "native":  # the C++ host inside libb2s.so (include/b2s_mgpu.h)
                sorter = mg.NativeDistributedSorter(n_local, kd, torch.int32, descending=desc, samples_per_rank=1024, slack=1.6)
            else:
                sorter = mg.DistributedSorter(n_local, kd, torch.int32, descending=desc, samples_per_rank=1024, slack=1.6,
                                              exchange=exchange)
            tk = H.to_dev(shards[rank]).view(kd)
            tv = H.to_dev(vals[rank])
            # first a sort whose shard size differs on ONE rank only (sizes may change freely between calls), then the
            # full shards twice (buffer reuse across sorts)
            short = n_local - 12_345 if rank == 1 else n_local
            o1 = sorter.sort(tk[:short], tv[:short])
            assert sum(o1.counts_all) == n_local * world - 12_345
            assert sorter.verify(tk[:short], tv[:short], o1, values_are_global_indices=True)
            for _ in range(2):
                out = sorter.sort(tk, tv)
            assert sorter.verify(tk, tv, out, values_are_global_indices=True)
            ek, ev = po.radix_sort(np.concatenate(shards), np.concatenate(vals), kt, desc)
            lo = sum(out.counts_all[:rank])
            got_k = out.keys.view(H.CONTAINER[kb]).cpu().numpy().view(H.NP_BITS[kb])
            assert np.array_equal(got_k, ek[lo: lo + out.count]), f"{mode}: keys differ"
            assert np.array_equal(out.values.cpu().numpy().view(np.uint32), ev[lo: lo + out.count]), f"{mode}: values differ"
            if hasattr(sorter, "close"):
                sorter.close()
            del sorter
        ret[rank] = "ok"
    except Exception:  # noqa: BLE001
        import traceback

        ret[rank] = "FAIL: " + traceback.format_exc()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["nccl", "peer", "native"])
def test_distributed_sorter_nccl(exchange):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp

    with tempfile.TemporaryDirectory() as d:
        mgr = mp.Manager()
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, os.path.join(d, "rdzv"), exchange, ret), nprocs=world, join=True)
        for r in range(world):
            assert ret.get(r) == "ok", ret.get(r)
