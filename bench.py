#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json: DeviceRadixSort::SortPairs throughput (GKeys/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

N = 1   workload = BASELINE.json configs[1]: SortPairs u32 keys / u32 values, 2^28 uniform random pairs, pointer form
        (input never modified, so every step sorts the same unsorted input; inputs 2 GiB >> 126 MB L2).
N > 1   the multi-GPU SortPairs (C++ host in libb2s.so, include/b2s_mgpu.h): 2^28 u32/u32 pairs PER GPU (weak scaling),
        one process per GPU, globally sorted across ranks: sampled splitters -> partition kernel whose per-destination
        runs are bulk-copied straight into the destination ranks' receive buffers over NVLink (CUDA-IPC peer memory)
        -> local sort.  A second leg reports BASELINE.json configs[4]: u64/u32, 2^30 pairs per GPU, uniform and AND-of-3
        (`config5_u64_u32`).  B2S_MGPU_BACKEND=torch selects the Python-orchestrated host (B2S_EXCHANGE=nccl|peer).
One "step" = one complete sort of the batch.  `value` = pairs sorted per second over all GPUs, device-timed
(CUDA events, max over ranks).  `e2e` = same metric through the Python mirror of cub::DeviceRadixSort with HOST
buffers (pinned H2D + sort + D2H inside the timed region).  `roofline` = dominant kernel (one digit pass of the
onesweep kernel): algorithmic bytes 2*(K+V)*n per launch over its CUDA-event duration, against the measured HBM
copy bandwidth in MEASURED_PEAKS.json.  `cpu_baseline` = the reference test harness' host std::stable_sort
(oracle/host_stable_sort.cpp, restated from test/test_device_radix_sort.cu:896-956), timed on this box's cores.
--impl reference: the reference's CPU implementation of the path (that same host stable_sort, on all host
threads); it also reports the unmodified reference CUB 2.2.0 kernels on this GPU (oracle/_ref/libref_cub.so)
as `reference_gpu`, the comparison north_star asks for.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KEY_TYPE_U32 = 6
KBYTES, VBYTES, PASSES = 4, 4, 4
ALGO_BYTES_PER_KEY = KBYTES + PASSES * 2 * (KBYTES + VBYTES)  # 68 B: SURVEY.md §8d
PASS_BYTES_PER_KEY = 2 * (KBYTES + VBYTES)                     # one digit pass reads+writes keys and values


def ncu_traffic():
    """DRAM bytes per digit-pass launch measured by ncu (dram__bytes_read.sum + dram__bytes_write.sum of one
    `ncu --set full` capture of this workload), recorded in profiles/ncu_traffic.json with its source report."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)["onesweep_kernel_u32_u32_2p28"]["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def time_reference_cpu(n: int, threads: int, steps: int, warmup: int):
    import numpy as np

    from oracle import pyoracle as po

    keys = np.random.default_rng(42).integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        po.host_stable_sort(keys, KEY_TYPE_U32, False, 0, 32, threads=threads)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def time_gpu_lib(fn, keys, vals, steps, warmup):
    """Pointer-form SortPairs through a C-ABI sort function; returns (ms_per_step, keys_out, vals_out)."""
    import torch

    from tests import harness as H

    n = keys.numel()
    ko, vo = torch.empty_like(keys), torch.empty_like(vals)
    nbytes = ctypes.c_size_t(0)
    args = (H._p(keys), H._p(ko), H._p(vals), H._p(vo), n, KEY_TYPE_U32, VBYTES, 4, 0, 0, 32)
    assert fn(None, ctypes.byref(nbytes), *args, None) == 0
    temp = torch.empty(nbytes.value, dtype=torch.uint8, device=keys.device)
    tp = ctypes.c_void_p(temp.data_ptr())
    s = H.stream_handle()
    for _ in range(warmup):
        assert fn(tp, ctypes.byref(nbytes), *args, s) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        assert fn(tp, ctypes.byref(nbytes), *args, s) == 0
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, ko, vo, temp


def host_threads() -> int:
    """Threads the CPU arm may use: the cores this process is allowed to run on (sched_getaffinity).  NOT OpenMP's
    default: torch.distributed.run exports OMP_NUM_THREADS=1, which made the round-1 arm single-threaded at N >= 2."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def run_reference(a):
    """Reference arm: the reference's own CPU implementation of the path (host std::stable_sort harness) on all
    host threads; plus the unmodified reference CUB kernels on this GPU as `reference_gpu`.  Loads nothing of the
    product (no libb2s.so): inputs come from numpy / torch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po

    threads = host_threads()
    n = 1 << int(os.environ.get("B2S_REF_LOG2N", "25"))
    sec = time_reference_cpu(n, threads, a.steps, a.warmup)
    value = n / sec / 1e9
    line = {
        "impl": "reference", "metric": "SortPairs GKeys/s (u32 keys / u32 values)", "value": value,
        "unit": "GKeys/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "SortPairs u32/u32 uniform random, host std::stable_sort harness of the reference "
                               "(test/test_device_radix_sort.cu:896-956)", "n_per_step": n},
        "cpu_baseline": {"value": value, "unit": "GKeys/s", "cores": threads, "kind": "port",
                         "sample": f"2^{n.bit_length() - 1} pairs per step, __gnu_parallel::stable_sort on {threads} "
                                   "threads = the cores this process may run on (the harness itself is single-threaded "
                                   "std::stable_sort)"},
        "e2e": {"value": value, "unit": "GKeys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    try:
        import torch

        if torch.cuda.is_available():
            ref = po.load_gpu_reference("ref")
            if ref is not None:
                ng = 1 << int(os.environ.get("B2S_BENCH_LOG2N", "28"))
                gen = torch.Generator(device="cuda").manual_seed(42)
                keys = torch.randint(-(1 << 31), (1 << 31) - 1, (ng,), dtype=torch.int32, device="cuda", generator=gen)
                vals = torch.arange(ng, dtype=torch.int32, device="cuda")
                ms, _, _, _ = time_gpu_lib(ref.sort, keys, vals, max(3, min(a.steps, 10)), 3)
                line["reference_gpu"] = {"impl": "reference CUB 2.2.0 (Policy900 onesweep) on this GPU, uniform random u32 keys",
                                         "value": ng / ms / 1e6, "unit": "GKeys/s", "ms_per_step": ms, "n": ng}
    except Exception as e:  # noqa: BLE001
        line["reference_gpu"] = {"unavailable": str(e)[:200]}
    emit(line)


def run_ours(a):
    import torch

    import cub_b200 as cb
    from cub_b200 import _lib
    from tests import harness as H

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        # NCCL's own log lines (e.g. "NCCL version ..." under NCCL_DEBUG=VERSION/INFO) go to stderr: stdout carries ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    b2s = _lib.load()
    n = 1 << int(os.environ.get("B2S_BENCH_LOG2N", "28"))
    peak, peak_src = measured_peak()
    # host side of the end-to-end path: run on (and allocate pinned memory from) the GPU's own NUMA node
    numa = cb.device_radix_sort.bind_host_to_gpu_numa_node(local) if os.environ.get("B2S_NUMA_BIND", "1") != "0" else {"skipped": "B2S_NUMA_BIND=0"}

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    keys = H.gen_device_keys(b2s, n, 4, 42 + 1000 * rank)
    vals = H.gen_device_iota(b2s, n, 4)
    line = {"metric": "SortPairs GKeys/s (u32 keys / u32 values)", "unit": "GKeys/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic"}

    if world == 1:
        sampler = ClockSampler(local)
        # ---- device-resident timing: K pointer-form sorts back to back
        barrier()
        sampler.start()
        time.sleep(0.3)  # the sampler's first rows arrive ~0.1-0.2 s after it starts; the timed region is ~0.1 s long
        ms, ko, vo, temp = time_gpu_lib(b2s.b2s_radix_sort, keys, vals, a.steps, a.warmup)
        launches_per_step = b2s.b2s_last_launch_count()
        value = n / ms / 1e6
        # ---- roofline leg: per-launch CUDA events inside the library, same workload
        b2s.b2s_timing_enable(1)
        nbytes = ctypes.c_size_t(temp.numel())
        args = (H._p(keys), H._p(ko), H._p(vals), H._p(vo), n, KEY_TYPE_U32, VBYTES, 4, 0, 0, 32)
        seg = (ctypes.c_float * 16)()
        pass_ms, hist_ms, memset_ms = [], [], []
        for _ in range(a.steps):
            # two sorts back to back, the events of the second one are read: reading them synchronises the host with the GPU, and a
            # sort that starts on an idle GPU is not what the timed region above runs (its steps follow each other without a gap)
            for _rep in range(2):
                assert b2s.b2s_radix_sort(ctypes.c_void_p(temp.data_ptr()), ctypes.byref(nbytes), *args,
                                          H.stream_handle()) == 0
            k = b2s.b2s_timing_read(seg, 16)
            assert k == 2 + PASSES, k
            memset_ms.append(seg[0])
            hist_ms.append(seg[1])
            pass_ms.extend(seg[2:2 + PASSES])
        b2s.b2s_timing_enable(0)
        # the same launch duration without the event records between the digit passes: the device-timed step above minus the
        # histogram and memset launches, over the passes.  Reported next to the per-launch figure, which carries the event
        # records' own gaps (~10 us each) and is the one the roofline fraction uses.
        pass_by_difference = (ms - sum(hist_ms) / len(hist_ms) - sum(memset_ms) / len(memset_ms)) / PASSES
        clocks = sampler.stop()  # sampled over the device-timed leg and the per-launch roofline leg (same kernels)
        avg_pass = sum(pass_ms) / len(pass_ms)
        achieved = n * PASS_BYTES_PER_KEY / avg_pass / 1e6  # GB/s
        whole = n * ALGO_BYTES_PER_KEY / ms / 1e6
        # ---- parity + comparator: unmodified reference CUB on the same buffers, same run
        ref_info, parity = None, "unchecked (oracle/_ref/libref_cub.so absent)"
        try:
            from oracle import pyoracle as po

            ref = po.load_gpu_reference("ref")
            if ref is not None:
                rms, rk, rv, rtemp = time_gpu_lib(ref.sort, keys, vals, max(3, min(a.steps, 10)), 3)
                parity = "bit-exact keys+values vs reference CUB 2.2.0" if (
                    torch.equal(rk, ko) and torch.equal(rv, vo)) else "MISMATCH vs reference CUB 2.2.0"
                ref_info = {"impl": "reference CUB 2.2.0 (Policy900 onesweep), same GPU, same input",
                            "value": n / rms / 1e6, "unit": "GKeys/s", "ms_per_step": rms}
                del rk, rv, rtemp
            tk = po.load_gpu_reference("tk")
            if tk is not None:
                tms, tkk, tkv, ttemp = time_gpu_lib(tk.sort, keys, vals, max(3, min(a.steps, 10)), 3)
                line["toolkit_cub_gpu"] = {"impl": "CUDA 12.9 toolkit CUB (SM100 policy)", "value": n / tms / 1e6,
                                           "unit": "GKeys/s", "ms_per_step": tms}
                del tkk, tkv, ttemp
        except Exception as e:  # noqa: BLE001
            parity = f"reference comparison failed: {str(e)[:120]}"
        del ko, vo, temp
        torch.cuda.empty_cache()
        # ---- e2e: host buffers through the Python mirror of cub::DeviceRadixSort (DoubleBuffer form)
        h_keys = torch.empty(n, dtype=torch.int32).pin_memory()
        h_vals = torch.empty(n, dtype=torch.int32).pin_memory()
        h_keys.copy_(keys.view(torch.int32))
        h_vals.copy_(vals.view(torch.int32))
        # three buffer sets, three streams: the upload of step i+1 overlaps the sort of step i and the download of step i-1
        # (PCIe is full duplex; with two sets the 3.9 ms sort sits between an upload and a download of the same set and
        # costs 2 ms per step); every step uploads its own 2 GiB of inputs and downloads its own 2 GiB of results
        sorter = cb.device_radix_sort.HostSorter(n, torch.uint32, torch.uint32, f"cuda:{local}", depth=3)
        hk, hv = h_keys.view(torch.uint32), h_vals.view(torch.uint32)
        for _ in range(max(2, min(a.warmup, 3))):
            sorter(hk, hv)
        sorter.synchronize()
        torch.cuda.synchronize()
        e2e_steps = max(2, min(a.steps, 20))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(sorter.s_up)        # device timestamps: before the first upload ...
        for _ in range(e2e_steps):
            out_k, out_v = sorter(hk, hv)
        e1.record(sorter.s_down)      # ... and after the last download
        sorter.synchronize()
        e2e_ms = e0.elapsed_time(e1) / e2e_steps
        u = lambda t, i: int(t.view(torch.int32)[i]) & 0xFFFFFFFF  # noqa: E731
        e2e_ok = u(out_k, 0) <= u(out_k, 1) <= u(out_k, n // 2) <= u(out_k, n - 1)
        # ---- CPU baseline: the reference harness' single-threaded std::stable_sort on a bounded sample
        cpu_n = 1 << int(os.environ.get("B2S_CPU_LOG2N", "25"))
        cpu_sec = time_reference_cpu(cpu_n, 1, 1, 0)
        line.update({
            "value": value, "ms_per_step": ms,
            "config": {"workload": "DeviceRadixSort::SortPairs u32 keys / u32 values, 2^%d uniform random pairs, "
                                   "pointer form (BASELINE.json configs[1])" % (n.bit_length() - 1),
                       "n": n, "l2": "inputs (2 GiB) larger than L2, no flush needed",
                       "launches_per_step": f"{launches_per_step} (1 memset + 1 histogram + {PASSES} digit passes)",
                       "parity": parity},
            "clocks": clocks,
            "e2e": {"value": n / e2e_ms / 1e6, "unit": "GKeys/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": n * (KBYTES + VBYTES), "d2h_bytes_per_step": n * (KBYTES + VBYTES),
                    "how": "HostSorter(depth=3): pinned host -> device, DoubleBuffer sort, device -> pinned host, "
                           "upload / sort / download on three streams, steps pipelined over three buffer sets; "
                           f"{e2e_steps} steps timed from before the first upload to after the last download",
                    "result_sane": e2e_ok, "host_numa": numa},
            "gpu_launches": (launches_per_step - 1) * a.steps,
            "roofline": {"bound": "hbm", "kernel": "digit_pass_kernel (one 8-bit digit pass, b2s_pass.cuh)", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": n * PASS_BYTES_PER_KEY,
                         "avg_launch_ms": avg_pass, "launch_ms_by_difference": pass_by_difference,
                         "frac_by_difference": n * PASS_BYTES_PER_KEY / pass_by_difference / 1e6 / peak,
                         "histogram_ms": sum(hist_ms) / len(hist_ms),
                         "memset_ms": sum(memset_ms) / len(memset_ms),
                         "whole_sort": {"bytes_per_key": ALGO_BYTES_PER_KEY, "achieved": whole,
                                        "frac": whole / peak}},
            "cpu_baseline": {"value": cpu_n / cpu_sec / 1e9, "unit": "GKeys/s", "cores": 1, "kind": "port",
                             "sample": f"2^{cpu_n.bit_length() - 1} u32 keys + index, std::stable_sort as in "
                                       "test/test_device_radix_sort.cu:896-956 (oracle/host_stable_sort.cpp)"},
        })
        if ref_info:
            line["reference_gpu"] = ref_info
        emit(line)
        return

    # ---- N > 1: multi-GPU SortPairs, weak scaling (2^28 pairs per GPU)
    from cub_b200 import multi_gpu

    backend = os.environ.get("B2S_MGPU_BACKEND", "native")

    def make_sorter(n_items, kdt, vdt):
        if backend == "native":  # C++ host inside libb2s.so (include/b2s_mgpu.h)
            return multi_gpu.NativeDistributedSorter(n_items, kdt, vdt)
        return multi_gpu.DistributedSorter(n_items, kdt, vdt, exchange=os.environ.get("B2S_EXCHANGE", "auto"))

    def timed_sorts(sorter, k, v, steps, warmup):
        for _ in range(warmup):
            sorter.sort(k, v)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            o = sorter.sort(k, v)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), o

    def nvlink_of(ph, item_bytes):
        if not ph.get("partition_kernel_device_ms"):
            return None
        sent = ph["items_sent_to_peers"] * item_bytes
        return {"kernel": "digit_pass_kernel<SplitterOp, PF_PEER | PF_TMAW> (partition fused with the all-to-all: bulk "
                          "shared->peer copies per destination run)",
                "bytes_sent_per_gpu": sent, "kernel_ms": ph["partition_kernel_device_ms"],
                "achieved": sent / ph["partition_kernel_device_ms"] / 1e6, "unit": "GB/s", "peak": 770.0,
                "frac": sent / ph["partition_kernel_device_ms"] / 1e6 / 770.0,
                "peak_source": "B200_PROFILING.md measured peer copy, per direction"}

    # values = global input index (rank * n + i): verify() can then check stability across rank boundaries too
    vals = H.gen_device_iota(b2s, n, 4)
    vals += rank * n if world * n < (1 << 32) else 0
    sorter = make_sorter(n, torch.uint32, torch.uint32)
    sampler = ClockSampler(local)  # started well before the (short) timed region: nvidia-smi needs ~0.3 s to deliver rows
    sampler.start()
    time.sleep(0.5)
    ms, out = timed_sorts(sorter, keys, vals, a.steps, a.warmup)
    t_roll = time.perf_counter()  # keep the same workload running (untimed) so that the 100 ms sampler sees it under load
    while time.perf_counter() - t_roll < 0.5:
        out = sorter.sort(keys, vals)
        torch.cuda.synchronize()
    clocks = sampler.stop()
    ok = sorter.verify(keys, vals, out, values_are_global_indices=world * n < (1 << 32))
    ph = sorter.last_phase_ms()

    # e2e: host shards in, host shards out, pipelined per rank: upload of step i+1 / sort of step i / download of step
    # i-1 on three streams; the sorted shard is copied out of the receive buffer (which the next exchange overwrites)
    h_keys = keys.view(torch.int32).cpu().pin_memory()
    h_vals = vals.view(torch.int32).cpu().pin_memory()
    e2e_steps = max(2, min(a.steps, 10))  # enough steps for the three-stage pipeline to reach its steady state
    cap = sorter.capacity
    hk_out = [torch.empty(cap, dtype=torch.int32).pin_memory() for _ in range(2)]
    hv_out = [torch.empty(cap, dtype=torch.int32).pin_memory() for _ in range(2)]
    dk = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(2)]
    dv = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(2)]
    ok_dev = [torch.empty(cap, dtype=torch.int32, device="cuda") for _ in range(2)]
    ov_dev = [torch.empty(cap, dtype=torch.int32, device="cuda") for _ in range(2)]
    s_up, s_down, s_main = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()
    up_done = [torch.cuda.Event() for _ in range(2)]
    sort_done = [torch.cuda.Event() for _ in range(2)]
    down_done = [torch.cuda.Event() for _ in range(2)]
    in_free = [torch.cuda.Event() for _ in range(2)]
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s_up)
    for i in range(e2e_steps):
        j = i & 1
        with torch.cuda.stream(s_up):
            if i >= 2:
                s_up.wait_event(in_free[j])       # the sort of step i-2 has consumed this input buffer
            dk[j].copy_(h_keys, non_blocking=True)
            dv[j].copy_(h_vals, non_blocking=True)
            up_done[j].record(s_up)
        s_main.wait_event(up_done[j])
        if i >= 2:
            s_main.wait_event(down_done[j])       # the download of step i-2 has drained this output buffer
        o = sorter.sort(dk[j].view(torch.uint32), dv[j].view(torch.uint32))
        in_free[j].record(s_main)
        ok_dev[j][:o.count].copy_(o.keys.view(torch.int32), non_blocking=True)
        ov_dev[j][:o.count].copy_(o.values.view(torch.int32), non_blocking=True)
        sort_done[j].record(s_main)
        with torch.cuda.stream(s_down):
            s_down.wait_event(sort_done[j])
            hk_out[j][:o.count].copy_(ok_dev[j][:o.count], non_blocking=True)
            hv_out[j][:o.count].copy_(ov_dev[j][:o.count], non_blocking=True)
            down_done[j].record(s_down)
    e1.record(s_down)
    torch.cuda.synchronize()
    barrier()
    e2e = torch.tensor([e0.elapsed_time(e1) / e2e_steps], device="cuda")
    dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    launches = sorter.launches_per_sort()
    if hasattr(sorter, "close"):
        sorter.close()
    del sorter, dk, dv, ok_dev, ov_dev, out, o
    torch.cuda.empty_cache()

    # ---- second leg: BASELINE.json configs[4] / north_star: u64 keys / u32 values, 2^30 pairs per GPU (2^33 at 8 GPUs),
    # uniform and AND-of-3 entropy-reduced keys.  B2S_CONFIG5_LOG2N overrides the per-GPU size (0 skips the leg).
    c5 = None
    lg5 = int(os.environ.get("B2S_CONFIG5_LOG2N", "30"))
    free_b, _tot = torch.cuda.mem_get_info()
    if lg5 and free_b > (1 << lg5) * 12 * 4.5:
        n5 = 1 << lg5
        del keys, vals
        torch.cuda.empty_cache()
        sorter5 = make_sorter(n5, torch.uint64, torch.uint32)
        v5 = H.gen_device_iota(b2s, n5, 4)
        c5 = {"workload": "multi-GPU SortPairs u64 keys / u32 values, 2^%d pairs per GPU (BASELINE.json configs[4])" % lg5,
              "n_per_gpu": n5, "n_total": n5 * world, "single_gpu_roofline_note":
              "one local sort of 2^%d u64/u32 pairs moves 8 + 8*2*12 = 200 B/key" % lg5}
        steps5 = max(2, min(a.steps, 5))
        for name, rounds in (("uniform", 1), ("and3", 3)):
            k5 = H.gen_device_keys(b2s, n5, 8, 42 + 1000 * rank, rounds).view(torch.uint64)
            ms5, o5 = timed_sorts(sorter5, k5, v5, steps5, 2)
            ok5 = sorter5.verify(k5, v5, o5)
            ph5 = sorter5.last_phase_ms()
            c5[name] = {"value": n5 * world / ms5 / 1e6, "unit": "GKeys/s", "ms_per_step": ms5, "steps": steps5, "verified": ok5,
                        "phases_ms": ph5, "nvlink": nvlink_of(ph5, 12)}
            del k5, o5
        if hasattr(sorter5, "close"):
            sorter5.close()
        del sorter5
    if rank == 0:
        total = n * world
        line.update({
            "value": total / ms / 1e6, "ms_per_step": ms,
            "config": {"workload": "multi-GPU SortPairs u32/u32, 2^%d uniform pairs per GPU, globally sorted across "
                                   "ranks (sampled splitters -> partition kernel whose run copies go straight into the peers' "
                                   "receive buffers over NVLink [host=%s] -> local LSD sort)" % (n.bit_length() - 1, backend),
                       "n_per_gpu": n, "n_total": total, "l2": "inputs larger than L2",
                       "verified": ok, "verifier": "local order + rank boundaries + multiset checksums + stability "
                                                   "(values = global input index must increase inside equal keys, across ranks too)",
                       "phases_ms": ph},
            "clocks": clocks,
            "e2e": {"value": total / float(e2e.item()) / 1e6, "unit": "GKeys/s", "ms_per_step": float(e2e.item()),
                    "h2d_bytes_per_step": total * 8, "d2h_bytes_per_step": total * 8,
                    "how": "per rank: pinned host -> device, global sort, device -> pinned host; upload / sort / download on "
                           "three streams, steps pipelined over two buffer sets", "host_numa_rank0": numa},
            "gpu_launches": launches * a.steps * world,
            "roofline": {"bound": "hbm", "kernel": "digit_pass_kernel (one 8-bit digit pass)", "achieved": None,
                         "peak": peak, "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src,
                         "note": "per-kernel HBM roofline is reported by the N=1 run; N>1 adds the NVLink exchange",
                         "nvlink": nvlink_of(ph, KBYTES + VBYTES)},
            "config5_u64_u32": c5,
        })
        emit(line)
    dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line of the contract goes to the process's real stdout; everything else any library prints to fd 1
    (e.g. NCCL's "NCCL version ..." banner, which is a plain printf) has been redirected to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
