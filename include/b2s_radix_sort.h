/*
 * b2s_radix_sort.h -- C-ABI of the B200-native device-wide radix sort.
 *
 * This is the drop-in boundary for the one hot path of NVIDIA/cub that this
 * repository replaces: cub::DeviceRadixSort (reference: cub/device/device_radix_sort.cuh).
 * The reference boundary is a header-only C++ template API; the entry points
 * below are what a type-erased FFI for that API binds (one symbol family,
 * key/value types erased into enums and byte sizes).  A C++ template veneer
 * with the exact cub::DeviceRadixSort signatures lives in
 * include/b200/device_radix_sort.cuh and forwards here.
 *
 * Conventions kept from the reference (device_radix_sort.cuh:312-350, 781-810;
 * dispatch_radix_sort.cuh:1939-1978):
 *   - returns a cudaError_t value as int (0 == cudaSuccess);
 *   - d_temp_storage == NULL  =>  only *temp_storage_bytes is written (never 0),
 *     nothing is launched;
 *   - no allocation, no synchronisation, no ownership transfer; all work is
 *     ordered on `stream`; uses the current device;
 *   - d_temp_storage may have ANY alignment (test_device_radix_sort.cu:1109-1110);
 *   - key/value pointers only need element alignment;
 *   - num_items == 0 => success, nothing launched;
 *   - begin_bit == end_bit => copy (pointer form) / no-op (DoubleBuffer form);
 *   - pointer form never writes keys_in / values_in;
 *   - DoubleBuffer form may clobber both buffers and updates the selectors so
 *     that bufs[selector] holds the result.
 * What a call enqueues: 1 memset + 1 histogram kernel + one digit-pass kernel per 8 key bits; for at most one tile
 * (8192 items of up to 16 bytes, 4096 beyond) ONE single-CTA kernel that needs no temp storage (the size query and
 * the "too small" check stay the same, as in the reference's single-tile path, dispatch_radix_sort.cuh:1272).  Only
 * stream-ordered work with arguments fixed at call time: a call can be captured into a CUDA graph and replayed.
 *
 * No torch types, no C++ types: plain pointers and sizes only.
 */
#ifndef B2S_RADIX_SORT_H_
#define B2S_RADIX_SORT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Key types of the fundamental-type overloads (util_type.cuh:1209-1305). */
typedef enum {
  B2S_U8 = 0,
  B2S_I8 = 1,
  B2S_U16 = 2,
  B2S_I16 = 3,
  B2S_F16 = 4,  /* __half          */
  B2S_BF16 = 5, /* __nv_bfloat16   */
  B2S_U32 = 6,
  B2S_I32 = 7,
  B2S_F32 = 8,
  B2S_U64 = 9,
  B2S_I64 = 10,
  B2S_F64 = 11,
  B2S_KEY_TYPE_COUNT = 12,
  /* 128-bit integer keys (util_type.cuh:1225,1259; test matrix test_device_radix_sort.cu:2238): accepted by
   * b2s_radix_sort / b2s_radix_sort_db (values of any size up to 64 bytes); sorted as a (high word, low word) composite
   * through the pack -> SortPairs<u64, index> -> gather path of b2s_radix_sort_struct. */
  B2S_U128 = 16,
  B2S_I128 = 17
} b2s_key_t;

/* cudaStream_t is passed as an opaque pointer so that C callers need no CUDA headers. */
typedef void *b2s_stream_t;

/*
 * Pointer form. Replaces
 *   cub::DeviceRadixSort::SortKeys            (device_radix_sort.cuh:2106)  value_bytes=0, descending=0
 *   cub::DeviceRadixSort::SortKeysDescending  (device_radix_sort.cuh:2921)  value_bytes=0, descending=1
 *   cub::DeviceRadixSort::SortPairs           (device_radix_sort.cuh:312)   descending=0
 *   cub::DeviceRadixSort::SortPairsDescending (device_radix_sort.cuh:1214)  descending=1
 *
 * value_bytes: 0 (keys only), 1, 2, 4, 8 or 16 -- values are moved as opaque words.
 * offset_bytes: 4 or 8, mirrors detail::ChooseOffsetT<NumItemsT> (choose_offset.cuh:44-57);
 *               accepted for interface fidelity -- this implementation sizes its
 *               offsets from num_items itself.
 */
int b2s_radix_sort(void *d_temp_storage, size_t *temp_storage_bytes,
                   const void *d_keys_in, void *d_keys_out,
                   const void *d_values_in, void *d_values_out,
                   uint64_t num_items, int key_type, int value_bytes, int offset_bytes,
                   int descending, int begin_bit, int end_bit, b2s_stream_t stream);

/*
 * DoubleBuffer form. Replaces the cub::DoubleBuffer overloads
 *   SortPairs (device_radix_sort.cuh:781), SortPairsDescending (:1675),
 *   SortKeys (:2525), SortKeysDescending (:3330).
 * key_bufs / val_bufs + selector mirror cub::DoubleBuffer<T>{d_buffers[2], selector}
 * (util_type.cuh:854-886).  On return *key_selector / *val_selector name the
 * buffer that holds the sorted output.  For keys-only pass value_bytes = 0
 * (val_bufs / val_selector may then be NULL).
 */
int b2s_radix_sort_db(void *d_temp_storage, size_t *temp_storage_bytes,
                      void *key_bufs[2], int *key_selector,
                      void *val_bufs[2], int *val_selector,
                      uint64_t num_items, int key_type, int value_bytes, int offset_bytes,
                      int descending, int begin_bit, int end_bit, b2s_stream_t stream);

/*
 * User-defined key structs ("decomposer" overloads).  Replaces the 16 overloads
 *   cub::DeviceRadixSort::{SortKeys,SortPairs}[Descending](..., decomposer [, begin_bit, end_bit], stream)
 *   (device_radix_sort.cuh:486-530, 625-666, 922-962, 1055-1105, 1368-1426, 1515-1563, 1816-1856, 1949-...),
 * pointer and DoubleBuffer forms.  The decomposer -- a callable returning a tuple of references to the arithmetic
 * members of the key, MOST significant first -- is type-erased into `fields`: byte offset and fundamental type of every
 * tuple element, in tuple order (the C++ veneer derives them by applying the decomposer to a probe object).
 * The sort is the stable sort on bits [begin_bit, end_bit) of the concatenated bit-ordered image (bit 0 is in the LAST
 * field); end_bit < 0 means "all bits" (the overloads without a bit range).  Keys are moved as opaque
 * key_struct_bytes-byte records, values as opaque value_bytes-byte records (0 = keys only, up to 64).
 * Same conventions as b2s_radix_sort; the temp storage holds one (u64 word, index) pair per item, twice.
 */
#define B2S_MAX_STRUCT_FIELDS 8
typedef struct {
  int32_t offset;   /* byte offset of the member inside the key struct */
  int32_t key_type; /* b2s_key_t of the member */
} b2s_field_t;
int b2s_radix_sort_struct(void *d_temp_storage, size_t *temp_storage_bytes,
                          const void *d_keys_in, void *d_keys_out, const void *d_values_in, void *d_values_out,
                          uint64_t num_items, int key_struct_bytes, const b2s_field_t *fields, int num_fields,
                          int value_bytes, int descending, int begin_bit, int end_bit, b2s_stream_t stream);
int b2s_radix_sort_struct_db(void *d_temp_storage, size_t *temp_storage_bytes,
                             void *key_bufs[2], int *key_selector, void *val_bufs[2], int *val_selector,
                             uint64_t num_items, int key_struct_bytes, const b2s_field_t *fields, int num_fields,
                             int value_bytes, int descending, int begin_bit, int end_bit, b2s_stream_t stream);

/*
 * Segmented sort: num_segments independent stable sorts of the contiguous ranges
 * [d_begin_offsets[s], d_end_offsets[s]) of one array.  Replaces cub::DeviceSegmentedRadixSort::{SortKeys,SortPairs}
 * [Descending], pointer and DoubleBuffer forms (cub/device/device_segmented_radix_sort.cuh; kernel
 * dispatch_radix_sort.cuh:383, dispatch :2076).  The offset arrays are DEVICE arrays of offset_bytes-byte signed integers
 * (4 or 8; the reference takes iterators -- int* in its tests and documentation); begin and end may alias
 * (d_offsets, d_offsets + 1); segments with end <= begin are empty; items outside every segment are not written.
 * One launch, one CTA per segment: segments of at most 4096 items are sorted entirely in shared memory.
 * Conventions as b2s_radix_sort / b2s_radix_sort_db (begin_bit == end_bit copies every segment in the pointer form).
 */
int b2s_segmented_radix_sort(void *d_temp_storage, size_t *temp_storage_bytes,
                             const void *d_keys_in, void *d_keys_out, const void *d_values_in, void *d_values_out,
                             uint64_t num_items, uint64_t num_segments, const void *d_begin_offsets,
                             const void *d_end_offsets, int offset_bytes, int key_type, int value_bytes, int descending,
                             int begin_bit, int end_bit, b2s_stream_t stream);
int b2s_segmented_radix_sort_db(void *d_temp_storage, size_t *temp_storage_bytes,
                                void *key_bufs[2], int *key_selector, void *val_bufs[2], int *val_selector,
                                uint64_t num_items, uint64_t num_segments, const void *d_begin_offsets,
                                const void *d_end_offsets, int offset_bytes, int key_type, int value_bytes, int descending,
                                int begin_bit, int end_bit, b2s_stream_t stream);

/* Size in bytes of a key of the given type (0 for an invalid type). */
int b2s_key_bytes(int key_type);

/* Library / build identification: "b2s <version> sm_100a". */
const char *b2s_version(void);

/*
 * Introspection used by bench.py and the tests (not part of the reference API):
 * number of kernels / memsets the last b2s_radix_sort* call on this thread enqueued.
 */
int b2s_last_launch_count(void);

/* Per-launch timing for bench.py's roofline leg (not part of the reference API).  While enabled,
 * every sort call records CUDA events on its stream around each operation it enqueues.
 * b2s_timing_read synchronises on the last event of the most recent call and returns the number of
 * segments written to ms[]: [memset, histogram, digit pass 0, digit pass 1, ...] in milliseconds. */
int b2s_timing_enable(int on);
int b2s_timing_read(float *ms, int capacity);

/* Tuning hooks (not part of the reference API).  Variant 0 is the production tuning of the
 * digit-pass kernel; a tuning build (-DB2S_TUNING) carries more points.  b2s_set_variant returns
 * the previous variant; b2s_describe_variant returns the number of variants (or -1) and the
 * {threads, items/thread, min CTAs/SM, match mode} of one of them. */
int b2s_set_variant(int variant);
int b2s_describe_variant(int key_bytes, int value_bytes, int variant, int *threads, int *items_per_thread,
                         int *min_ctas_per_sm, int *match_mode);
/* Test hook: sorts of at most one tile (8192 items of up to 16 bytes, 4096 beyond) normally run as ONE kernel that keeps
 * every digit pass in shared memory (the reference's single-tile path, dispatch_radix_sort.cuh:1272).  0 sends them through the
 * multi-kernel path instead, so that tests can cover it at small sizes; returns the previous setting.  Also B2S_SINGLE_TILE=0. */
int b2s_set_single_tile(int enable);
/* Scheduling mode of a variant: bits 0-1 = 0 one tile per CTA, 1/2 persistent CTAs (next tile claimed after/before the
 * write-out); bits 16+ = L2 prefetch distance in tiles.  -1 for an unknown variant. */
int b2s_variant_mode(int key_bytes, int value_bytes, int variant);
/* The upfront histogram kernel alone (replaces DeviceRadixSortHistogramKernel + DeviceRadixSortExclusiveSumKernel,
 * dispatch_radix_sort.cuh:556,603): d_offsets[p * 256 + d] = number of keys whose digit of pass p (bits
 * [begin_bit + 8p, min(begin_bit + 8p + 8, end_bit)) of the bit-ordered key) is < d.  d_offsets holds
 * ceil((end_bit - begin_bit) / 8) * 256 + 1 uint64 (the last word is scratch).  Exposed so that tests can check the
 * kernel directly against the oracle; a sort runs it internally. */
int b2s_digit_histogram(const void *d_keys, uint64_t num_items, int key_type, int descending, int begin_bit, int end_bit,
                        uint64_t *d_offsets, b2s_stream_t stream);
/* Kernel and flow of a variant: >= 0 production kernel (b2s_pass.cuh) with these flag bits (1 = (key,value) scattered as
 * one 64-bit store, 2 = bulk-copy write-out, 4 = ticketed tile ids); -1 laboratory kernel (tuning builds); -2 unknown. */
int b2s_variant_flow(int key_bytes, int value_bytes, int variant);
/* Tile ids of the digit pass.  0 (default): tile id = block index -- relies on CTAs being dispatched in index order, as
 * CUB's decoupled look-back scan does (cub/agent/agent_scan.cuh).  1: every CTA takes an atomic ticket, as the reference's
 * onesweep agent does (cub/agent/agent_radix_sort_onesweep.cuh:650-687): forward progress of the look-back then holds
 * under ANY dispatch order, at ~1.5 % of throughput.  Also B2S_TILE_CLAIM=1.  Returns the previous setting. */
int b2s_set_tile_claim(int enable);
/* Keys-only sorts of 1- and 2-byte keys over ALL their bits (begin_bit == 0, end_bit == key bits) run as a counting sort from
 * a cut-over size on: joint histogram of the keys, exclusive prefix, expansion with 128-bit stores -- 2*K bytes of HBM traffic
 * per key instead of K + 2*K*K, bit-identical results (the input order of -0.0 / +0.0 inside their common run is re-created by
 * a stable compaction).  In the DoubleBuffer form the result is then in the ALTERNATE buffer (selector flips) whatever the
 * number of digit passes would have been.  b2s_set_counting_sort(0) / B2S_COUNTING_SORT=0 sends those sorts through the digit
 * passes instead; b2s_set_counting_min_items(key_bytes = 1 | 2, n) sets the cut-over (defaults: 2^16 items for 1-byte keys, 2^22 for 2-byte integers, 2^23 for f16 / bf16 -- where the counting path overtakes the digit passes on a B200; setting key_bytes = 2 sets both 2-byte cut-overs).  Both
 * return the previous setting. */
int b2s_set_counting_sort(int enable);
/* Sorts of 4- / 8-byte floating keys over ALL their bits (no values or 4-byte values, more than one tile) give -0.0 and +0.0 one
 * image in the first digit pass, record which keys were zeros and their signs (two bits per key of temporary storage), and
 * restore the signs in the run of zeros after the last pass: digits then cost what integer digits cost in every pass, results
 * are bit-identical.  b2s_set_float_zero_recording(0) / B2S_FLOAT_ZERO_RECORD=0 keeps the reference's scheme (zeros collapsed
 * in every digit extraction).  Returns the previous setting. */
int b2s_set_float_zero_recording(int enable);
uint64_t b2s_set_counting_min_items(int key_bytes, uint64_t min_items);
/* Tuning builds: digit pass number `pass` of every following sort writes per-tile phase timestamps (u64[tiles][16], SM clock
 * cycles; slot 0 = global timer in ns, slot 15 = SM id) to d_trace when the active variant is a trace variant.  NULL disables. */
int b2s_set_trace(void *d_trace, int pass);

/*
 * Multi-GPU SortPairs building blocks (new functionality, SURVEY.md §8e -- the reference has no multi-GPU
 * path).  One process per GPU; cub_b200/multi_gpu.py drives them with torch.distributed for the metadata.
 * The global order is (key, source rank, index on the source rank), i.e. the stable sort of the rank-order
 * concatenation of the shards.  Keys: 4- or 8-byte types; values: 0, 4 or 8 bytes.
 *
 * Splitters are (raw key, source rank) pairs in DEVICE memory (they come out of a device-side sort of the samples, and
 * the host never has to wait for them), ascending in (sort key, source rank) order -- what a stable sort of the
 * (key, rank) samples produces; at most 7 (8 ranks).  The per-destination offsets
 * of b2s_split_scatter are in device memory as well: the whole exchange is enqueued without a host round trip.
 * A local key goes to destination d = number of splitters (k*, r*) with (k*, r*) <= (key, my_rank), comparing
 * keys on bits [begin_bit, end_bit) of their bit-ordered transform.
 *
 * b2s_split_count:   d_counts[d] (uint64, num_splitters + 1 entries) = local keys destined for rank d.
 * b2s_split_scatter: stable partition of the local shard by destination.  Destination d's items are written
 *                    to  peer_keys[d] + d_dest_offsets[d]  (items) when peer_keys != NULL -- receive buffers
 *                    of the other ranks mapped into this process, so the all-to-all exchange is fused into the
 *                    partition pass as NVLink stores -- and to  d_keys_out + d_dest_offsets[d]  otherwise
 *                    (then exchanged by an NCCL all-to-all).  Peer stores at item positions >= peer_capacity are
 *                    dropped, so an under-sized receive buffer is never overrun (the caller sees the overflow in
 *                    the count matrix).  Two-phase temp-storage query like the sort.
 */
/* Let kernels of the CURRENT device store into memory of `peer_device` (needed once per peer before
 * b2s_split_scatter with peer buffers; idempotent). */
int b2s_enable_peer_access(int peer_device);
/* Map a buffer of another process (cudaIpcGetMemHandle bytes, e.g. from torch's UntypedStorage._share_cuda_) for
 * kernels of the CURRENT device; *d_ptr is the base of the exporting allocation.  One open per handle per process. */
int b2s_ipc_open(const void *handle64, void **d_ptr);
int b2s_ipc_close(void *d_ptr);
int b2s_split_count(const void *d_keys_in, uint64_t num_items, int key_type, int descending,
                    int begin_bit, int end_bit, const void *d_splitter_keys, const int *d_splitter_ranks,
                    int num_splitters, int my_rank, uint64_t *d_counts, b2s_stream_t stream);
int b2s_split_scatter(void *d_temp_storage, size_t *temp_storage_bytes,
                      const void *d_keys_in, void *d_keys_out, const void *d_values_in, void *d_values_out,
                      uint64_t num_items, int key_type, int value_bytes, int descending, int begin_bit, int end_bit,
                      const void *d_splitter_keys, const int *d_splitter_ranks, int num_splitters, int my_rank,
                      const uint64_t *d_dest_offsets, void *const *peer_keys, void *const *peer_vals,
                      uint64_t peer_capacity, b2s_stream_t stream);

/*
 * Device helpers for the test/bench harness.  All enqueue on `stream`, no sync.
 */

/* out[i] = number of keys in sorted d_keys[0..n) (ascending, unsigned compare of
 * the bit-ordered transform of key_type) that are < splitters[i] (lower bound), for
 * i in [0, num_splitters).  Splitters are raw keys of key_type. out is uint64_t[]. */
int b2s_lower_bound(const void *d_sorted_keys, uint64_t num_items, int key_type,
                    const void *d_splitters, int num_splitters, uint64_t *d_out,
                    b2s_stream_t stream);

/* Fill keys with the counter-based generator of SURVEY.md §8d:
 *   key_i = splitmix64(seed * 0x100000001B3 + first_index + i), AND-ed over (and_rounds) consecutive
 * seeds (seed, seed+1, ...), truncated to key_bytes. */
int b2s_fill_keys(void *d_keys, uint64_t num_items, int key_bytes, uint64_t seed,
                  int and_rounds, uint64_t first_index, b2s_stream_t stream);

/* values_i = (uint32/uint64)(first_index + i) */
int b2s_fill_iota(void *d_values, uint64_t num_items, int value_bytes, uint64_t first_index,
                  b2s_stream_t stream);

/* Order + multiset check of a sorted run, for sizes beyond the oracle:
 *  d_result[0] = number of adjacent inversions under the bit-ordered transform restricted
 *                to [begin_bit,end_bit) (0 for a correctly sorted array),
 *  d_result[1] = sum over i of mix64(key_i) (order-independent checksum, wraps mod 2^64),
 *  d_result[2] = sum over i of mix64(key_i ^ rotl(value_i, 32)) if values given. */
int b2s_check_sorted(const void *d_keys, const void *d_values, uint64_t num_items, int key_type,
                     int value_bytes, int descending, int begin_bit, int end_bit,
                     uint64_t *d_result, b2s_stream_t stream);

#ifdef __cplusplus
}
#endif

#endif /* B2S_RADIX_SORT_H_ */
