"""2-GPU debug: (a) PEER split path with local pointers, (b) IPC + peer store with a trivial kernel."""
import ctypes, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from cub_b200 import _lib, multi_gpu as mg
from tests import harness as H

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
lib = _lib.load()
ops = mg.LocalOps(torch.device("cuda", rank))
# (a) PEER path, all destinations local
n = 100_000
rng = np.random.default_rng(1)
raw = H.random_bits(rng, n, 4)
sp = np.sort(raw[:3].copy()); spr = np.array([0, 1, 2], dtype=np.int32)
dk = H.to_dev(raw); dv = H.to_dev(np.arange(n, dtype=np.uint32))
counts = ops.split_count(dk, n, 6, False, 0, 32, sp, spr, 1)
offs = np.concatenate(([0], np.cumsum(counts)[:-1])).astype(np.uint64)
ok1, ov1 = torch.empty_like(dk), torch.empty_like(dv)
ops.split_scatter(dk, dv, ok1, ov1, n, 6, False, 0, 32, sp, spr, 1, offs, None, None, {})
ok2, ov2 = torch.zeros_like(dk), torch.zeros_like(dv)
ops.split_scatter(dk, dv, None, None, n, 6, False, 0, 32, sp, spr, 1, offs, [ok2.data_ptr()] * 4, [ov2.data_ptr()] * 4, {})
torch.cuda.synchronize()
print(rank, "PEER(local ptrs) == bucketed:", torch.equal(ok1, ok2), torch.equal(ov1, ov2), flush=True)
# (b) IPC + trivial kernel store into the peer
buf = torch.zeros(1 << 20, dtype=torch.int32, device="cuda")
desc = mg._share_cuda(buf)
every = [None] * world
dist.all_gather_object(every, desc)
peer = (rank + 1) % world
d = every[peer]
print(rank, "enable peer", lib.b2s_enable_peer_access(d["device"]), d["device"], d["offset"], flush=True)
p = ctypes.c_void_p()
print(rank, "ipc open rc", lib.b2s_ipc_open(d["handle"], ctypes.byref(p)), hex(p.value or 0), flush=True)
class PT:
    def __init__(s, a): s.a = a
    def data_ptr(s): return s.a
pt = PT(p.value + d["offset"])
dist.barrier()
rc = lib.b2s_fill_iota(ctypes.c_void_p(pt.data_ptr()), 1 << 20, 4, 1000 * (rank + 1), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
dist.barrier()
print(rank, "rc", rc, "my buf head (written by peer):", buf[:3].tolist(), flush=True)
dist.destroy_process_group()
