"""CPU (gloo, world_size 2 and 3): host logic of the multi-GPU SortPairs -- splitter selection, count matrix,
exchange plan, all-to-all plumbing, stability across ranks, verification -- with the oracle standing in for the
device ops (tests/cpu_local_ops.py).  The CUDA path itself is covered by tests/test_multi_gpu_gpu.py (-m gpu)."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cub_b200 import multi_gpu as mg
from oracle import pyoracle as po


def test_sort_key_matches_oracle():
    lib = po.cpu()
    rng = np.random.default_rng(5)
    for kt in (6, 7, 8, 9, 10, 11):
        nb = po.KEY_BYTES[kt] * 8
        raws = [0, 1, (1 << nb) - 1, 1 << (nb - 1), (1 << (nb - 1)) - 1] + [int(x) for x in rng.integers(0, 1 << 63, 50)]
        for raw in raws:
            raw &= (1 << nb) - 1
            for desc in (False, True):
                for bb, eb in ((0, nb), (3, nb - 5), (nb // 2 - 1, nb // 2 + 1)):
                    assert mg.sort_key(raw, kt, desc, bb, eb) == lib.oracle_sort_key(raw, kt, int(desc), bb, eb)


def test_exchange_plan():
    m = np.array([[3, 1, 0], [2, 2, 2], [0, 5, 1]])
    send_counts, send_off, recv_counts, total, peer_off = mg.exchange_plan(m, 1)
    assert send_counts.tolist() == [2, 2, 2] and send_off.tolist() == [0, 2, 4]
    assert recv_counts.tolist() == [1, 2, 5] and total == 8
    assert peer_off.tolist() == [3, 1, 0]  # rank 1's segment starts after rank 0's in every destination
    assert mg.exchange_plan(m, 0)[4].tolist() == [0, 0, 0]


def _worker(rank, world, path, case, ret):
    from tests.cpu_local_ops import CpuOps

    dist.init_process_group("gloo", init_method=f"file://{path}", rank=rank, world_size=world)
    try:
        kt, vdtype, n_local, desc, bb, eb, mode = case
        kbytes = po.KEY_BYTES[kt]
        rng = np.random.default_rng(1234)  # same stream on every rank: everybody can build the global input
        sizes = [n_local + 17 * r for r in range(world)]
        shards = []
        for r in range(world):
            if mode == "dups":
                k = rng.integers(0, 4, size=sizes[r], dtype=np.uint64)
            elif mode == "equal":
                k = np.full(sizes[r], 7, dtype=np.uint64)
            else:
                k = rng.integers(0, 1 << 63, size=sizes[r], dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=sizes[r], dtype=np.uint64)
            shards.append(k.astype(mg._NP_BITS[kbytes]))
        vals = [np.arange(sizes[r], dtype=np.uint32) + np.uint32(1_000_000 * r) for r in range(world)]
        kd = {6: torch.uint32, 7: torch.int32, 8: torch.float32, 9: torch.uint64, 10: torch.int64, 11: torch.float64}[kt]
        sorter = mg.DistributedSorter(max(sizes), kd, vdtype, descending=desc, begin_bit=bb, end_bit=eb,
                                      samples_per_rank=64, slack=2.5, exchange="nccl", ops=CpuOps())
        cont = {4: np.int32, 8: np.int64}[kbytes]
        tk = torch.from_numpy(shards[rank].view(cont).copy())
        tv = torch.from_numpy(vals[rank].view(np.int32).copy()) if vdtype is not None else None
        # a first call with a shorter shard on ONE rank only: shard sizes may change between calls without any rank
        # taking a different sequence of collectives (round-1 advisor finding)
        short = sizes[rank] - 111 if rank == 0 else sizes[rank]
        o1 = sorter.sort(tk[:short], tv[:short] if tv is not None else None)
        assert sum(o1.counts_all) == sum(sizes) - 111 and sorter.verify(tk[:short], tv[:short] if tv is not None else None, o1)
        out = sorter.sort(tk, tv)
        assert sorter.verify(tk, tv, out, values_are_global_indices=mode == "random")
        # bit-exact against the oracle's stable sort of the rank-order concatenation
        ek, ev = po.radix_sort(np.concatenate(shards), np.concatenate(vals) if vdtype is not None else None, kt, desc, bb, eb)
        lo = sum(out.counts_all[:rank])
        got_k = out.keys.view({4: torch.int32, 8: torch.int64}[kbytes]).numpy().view(mg._NP_BITS[kbytes])
        assert out.count == out.counts_all[rank] and sum(out.counts_all) == sum(sizes)
        assert np.array_equal(got_k, ek[lo: lo + out.count]), "keys differ from the global stable sort"
        if vdtype is not None:
            assert np.array_equal(out.values.numpy().view(np.uint32), ev[lo: lo + out.count]), "values (stability) differ"
        if mode in ("dups", "equal"):  # ties are spread over the ranks instead of piling up on one
            assert max(out.counts_all) < 0.8 * sum(sizes)
        ret[rank] = "ok"
    except Exception as e:  # noqa: BLE001
        import traceback

        ret[rank] = "FAIL: " + traceback.format_exc()
    finally:
        dist.destroy_process_group()


CASES = [
    (9, torch.int32, 3001, False, 0, None, "random"),   # u64 keys / u32 values (BASELINE configs[4] shape)
    (6, torch.int32, 2500, True, 0, None, "random"),    # descending
    (9, torch.int32, 2000, False, 0, None, "dups"),     # heavy duplicates: tie-break on source rank
    (6, None, 1800, False, 0, None, "equal"),           # keys only, all equal
    (10, torch.int32, 2200, False, 5, 40, "random"),    # signed keys, partial bit range
    (11, torch.int32, 1500, True, 0, None, "random"),   # f64 descending (NaN patterns included by the raw bits)
]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"kt{c[0]}-{c[6]}-{'desc' if c[3] else 'asc'}")
def test_distributed_sort_gloo(world, case):
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "rdzv")
        mgr = mp.Manager()
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, path, case, ret), nprocs=world, join=True)
        for r in range(world):
            assert ret.get(r) == "ok", ret.get(r)
