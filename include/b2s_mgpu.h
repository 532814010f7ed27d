/*
 * b2s_mgpu.h -- C-ABI of the single-box multi-GPU SortPairs / SortKeys (part of libb2s.so).
 *
 * New functionality (BASELINE.json configs[4]): the reference (NVIDIA/cub) is a single-GPU
 * library and has no counterpart; the single-GPU building block is the cub::DeviceRadixSort drop-in of
 * b2s_radix_sort.h (cub/device/device_radix_sort.cuh:312,781,...).  This is the interface SURVEY.md section 8(b)
 * sketches as `b2s_mgpu_sort_pairs(per-device shards, ncclComm_t, streams ...)`.
 *
 * Model: ONE PROCESS PER GPU (ranks 0 .. world-1 of one NVSwitch box, world <= 8), every rank calls the same
 * functions collectively.  The host side is C++ inside the library: NCCL (loaded with dlopen("libnccl.so.2"), so the
 * single-GPU entry points have no NCCL dependency) carries metadata only -- samples, the count matrix, one fence word --
 * and CUDA IPC maps every rank's receive buffer into every other rank, so that the partition kernel's run copies ARE the
 * all-to-all (bulk shared->peer copies over NVLink).  Python (cub_b200/multi_gpu.py) is a thin binding.
 *
 * Global order = the STABLE sort of the rank-order concatenation of the shards: (key, source rank, index on the
 * source rank), on bits [begin_bit, end_bit) of the bit-ordered key, ascending or descending -- the result a single
 * DeviceRadixSort over the concatenation would give.  Algorithm: regular samples -> ncclAllGather -> every rank sorts
 * the same (key, rank) samples and takes the same world-1 splitters (ties on the key broken by source rank, which
 * spreads runs of equal keys over destinations without breaking stability) -> b2s_split_count -> ncclAllGather of the
 * count matrix -> b2s_split_scatter straight into the peers' receive buffers -> one-word ncclAllReduce as the fence ->
 * one local stable DeviceRadixSort (DoubleBuffer form) over the world received runs (they arrive in rank order).
 *
 * All functions return 0 on success, a cudaError_t (< 1000), 1000 + ncclResult_t, or B2S_MGPU_E_* below.
 */
#ifndef B2S_MGPU_H_
#define B2S_MGPU_H_

#include <stddef.h>
#include <stdint.h>

#include "b2s_radix_sort.h"

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_MGPU_ID_BYTES 128      /* sizeof(ncclUniqueId) */
#define B2S_MGPU_MAX_RANKS 8
#define B2S_MGPU_E_NCCL_MISSING 2001  /* libnccl.so.2 could not be loaded */
#define B2S_MGPU_E_ARGUMENT 2002
#define B2S_MGPU_E_CAPACITY 2003      /* a rank would receive more than its receive capacity: raise `slack` */

typedef struct b2s_mgpu_sorter b2s_mgpu_sorter_t;

/* Rank 0 creates the id and hands the 128 bytes to the other ranks by any means (MPI, a file, torch.distributed ...). */
int b2s_mgpu_unique_id(void *id128);

/* Collective.  Binds to the CURRENT device.  Allocates (cudaMalloc) two receive buffers of
 * capacity = max_items_per_rank * slack + 1024 items per array, the sample / metadata buffers and the temp storage of
 * the local sort, joins the NCCL communicator and maps the peers' receive buffers (cudaIpcOpenMemHandle).
 * key_type: a 4- or 8-byte b2s_key_t; value_bytes: 0 (keys only), 4 or 8.  slack >= 1 (1.10 is a good default: shard
 * sizes stay within a few % of n/world with 8192 samples per rank).  samples_per_rank <= 65536. */
int b2s_mgpu_create(b2s_mgpu_sorter_t **sorter, const void *id128, int rank, int world, uint64_t max_items_per_rank,
                    int key_type, int value_bytes, int descending, int begin_bit, int end_bit, double slack,
                    int samples_per_rank);

/* Collective; everything is enqueued on `stream` (the one host wait is for the 8 x 8 count matrix, which arrives while
 * the partition kernel runs).  d_keys / d_values: this rank's shard (num_items >= 1, <= max_items_per_rank), never
 * written.  On return *d_keys_out / *d_values_out point INTO the sorter's receive buffers: this rank's slice of the
 * globally sorted sequence, *out_count items, valid (in stream order) until the next call on any rank overwrites them.
 * counts_all (may be NULL): world entries, the slice sizes of all ranks. */
int b2s_mgpu_sort(b2s_mgpu_sorter_t *sorter, const void *d_keys, const void *d_values, uint64_t num_items,
                  void **d_keys_out, void **d_values_out, uint64_t *out_count, uint64_t *counts_all, b2s_stream_t stream);

/* Device-timed phases of the last sort on this rank (synchronises on its last event), milliseconds:
 * [0] samples + splitters, [1] count + count-matrix exchange, [2] partition kernel (== the all-to-all), [3] fence,
 * [4] local sort, [5] whole call; items_sent = items this rank stored into OTHER ranks' buffers. */
int b2s_mgpu_last_phases(b2s_mgpu_sorter_t *sorter, float *ms6, uint64_t *items_sent);

uint64_t b2s_mgpu_capacity(const b2s_mgpu_sorter_t *sorter);
const char *b2s_mgpu_last_error(const b2s_mgpu_sorter_t *sorter);

/* Collective: unmaps the peers' buffers (after a fence so that nobody is still storing), frees everything. */
int b2s_mgpu_destroy(b2s_mgpu_sorter_t *sorter);

/* Checker helper for (key, index) pairs whose values were increasing in input order: *d_result (uint64) = number of
 * adjacent positions with EQUAL sort keys whose values do not increase -- 0 for a stable sort. */
int b2s_check_stable(const void *d_keys, const void *d_values, uint64_t num_items, int key_type, int value_bytes,
                     int descending, int begin_bit, int end_bit, uint64_t *d_result, b2s_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B2S_MGPU_H_ */
