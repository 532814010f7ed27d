// oracle/ref_shim_ext.cu -- TEST INFRASTRUCTURE, NOT PRODUCT.
//
// More extern "C" wrappers around the UNMODIFIED reference (compiled from /root/reference where it lies, CUB 2.2.0), for
// the SURVEY.md section 8(f) rows: decomposer (user-defined struct key) overloads, cub::DeviceSegmentedRadixSort and
// 128-bit keys.  Only *calls* the reference's public API (cub/device/device_radix_sort.cuh:486-530 ...,
// cub/device/device_segmented_radix_sort.cuh); no reference source is copied.  Part of oracle/_ref/libref_cub.so.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_segmented_radix_sort.cuh>
#include <cuda/std/tuple>
#include <cstdint>

#ifndef REF_NS
#define REF_NS refcub
#endif
namespace rc = REF_NS::cub;

#ifndef REF_PREFIX
#define REF_PREFIX ref_cub
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

namespace {
// the key struct of the reference's own decomposer tests (test/catch2_test_device_radix_sort_custom.cu) and documentation
struct custom_t {
  float f;
  long long lli;
};
struct decomposer_t {
  __host__ __device__ ::cuda::std::tuple<float&, long long&> operator()(custom_t& key) const { return {key.f, key.lli}; }
};
}  // namespace

// vbytes: 0 (keys only) or 4; eb < 0: the overloads without a bit range
extern "C" int CAT(REF_PREFIX, _struct_sort)(void* tmp, size_t* bytes, const void* kin, void* kout, const void* vin, void* vout,
                                             uint64_t n, int vbytes, int desc, int bb, int eb, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const custom_t* ki = (const custom_t*)kin;
  custom_t* ko = (custom_t*)kout;
  const uint32_t* vi = (const uint32_t*)vin;
  uint32_t* vo = (uint32_t*)vout;
  const int num = (int)n;
  decomposer_t d;
  if (vbytes == 0) {
    if (eb < 0)
      return (int)(desc ? rc::DeviceRadixSort::SortKeysDescending(tmp, *bytes, ki, ko, num, d, s)
                        : rc::DeviceRadixSort::SortKeys(tmp, *bytes, ki, ko, num, d, s));
    return (int)(desc ? rc::DeviceRadixSort::SortKeysDescending(tmp, *bytes, ki, ko, num, d, bb, eb, s)
                      : rc::DeviceRadixSort::SortKeys(tmp, *bytes, ki, ko, num, d, bb, eb, s));
  }
  if (vbytes != 4) return -1;
  if (eb < 0)
    return (int)(desc ? rc::DeviceRadixSort::SortPairsDescending(tmp, *bytes, ki, ko, vi, vo, num, d, s)
                      : rc::DeviceRadixSort::SortPairs(tmp, *bytes, ki, ko, vi, vo, num, d, s));
  return (int)(desc ? rc::DeviceRadixSort::SortPairsDescending(tmp, *bytes, ki, ko, vi, vo, num, d, bb, eb, s)
                    : rc::DeviceRadixSort::SortPairs(tmp, *bytes, ki, ko, vi, vo, num, d, bb, eb, s));
}

namespace {
template <typename KeyT>
int seg_sort(void* tmp, size_t* bytes, const void* kin, void* kout, const void* vin, void* vout, int n, int segs,
             const int* begin, const int* end, int vbytes, int desc, int bb, int eb, cudaStream_t s) {
  const KeyT* ki = (const KeyT*)kin;
  KeyT* ko = (KeyT*)kout;
  if (vbytes == 0)
    return (int)(desc ? rc::DeviceSegmentedRadixSort::SortKeysDescending(tmp, *bytes, ki, ko, n, segs, begin, end, bb, eb, s)
                      : rc::DeviceSegmentedRadixSort::SortKeys(tmp, *bytes, ki, ko, n, segs, begin, end, bb, eb, s));
  if (vbytes != 4) return -1;
  const uint32_t* vi = (const uint32_t*)vin;
  uint32_t* vo = (uint32_t*)vout;
  return (int)(desc ? rc::DeviceSegmentedRadixSort::SortPairsDescending(tmp, *bytes, ki, ko, vi, vo, n, segs, begin, end, bb, eb, s)
                    : rc::DeviceSegmentedRadixSort::SortPairs(tmp, *bytes, ki, ko, vi, vo, n, segs, begin, end, bb, eb, s));
}
}  // namespace

// key_type: b2s_key_t 6 (u32), 7 (i32), 8 (f32), 9 (u64), 2 (u16); offsets are 32-bit ints as in the reference's tests
extern "C" int CAT(REF_PREFIX, _segmented_sort)(void* tmp, size_t* bytes, const void* kin, void* kout, const void* vin, void* vout,
                                                uint64_t n, int num_segments, const void* begin_offsets, const void* end_offsets,
                                                int key_type, int vbytes, int desc, int bb, int eb, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int* b = (const int*)begin_offsets;
  const int* e = (const int*)end_offsets;
  switch (key_type) {
    case 2: return seg_sort<uint16_t>(tmp, bytes, kin, kout, vin, vout, (int)n, num_segments, b, e, vbytes, desc, bb, eb, s);
    case 6: return seg_sort<uint32_t>(tmp, bytes, kin, kout, vin, vout, (int)n, num_segments, b, e, vbytes, desc, bb, eb, s);
    case 7: return seg_sort<int32_t>(tmp, bytes, kin, kout, vin, vout, (int)n, num_segments, b, e, vbytes, desc, bb, eb, s);
    case 8: return seg_sort<float>(tmp, bytes, kin, kout, vin, vout, (int)n, num_segments, b, e, vbytes, desc, bb, eb, s);
    case 9: return seg_sort<unsigned long long>(tmp, bytes, kin, kout, vin, vout, (int)n, num_segments, b, e, vbytes, desc, bb, eb, s);
    default: return -1;
  }
}

// 128-bit integer keys (cub/util_type.cuh:1225,1259); vbytes 0 or 4
extern "C" int CAT(REF_PREFIX, _sort128)(void* tmp, size_t* bytes, const void* kin, void* kout, const void* vin, void* vout,
                                         uint64_t n, int is_signed, int vbytes, int desc, int bb, int eb, void* stream) {
#if CUB_IS_INT128_ENABLED
  cudaStream_t s = (cudaStream_t)stream;
  const uint32_t* vi = (const uint32_t*)vin;
  uint32_t* vo = (uint32_t*)vout;
  const int num = (int)n;
#define GO128(T)                                                                                                          \
  {                                                                                                                       \
    const T* ki = (const T*)kin;                                                                                          \
    T* ko = (T*)kout;                                                                                                     \
    if (vbytes == 0)                                                                                                      \
      return (int)(desc ? rc::DeviceRadixSort::SortKeysDescending(tmp, *bytes, ki, ko, num, bb, eb, s)                    \
                        : rc::DeviceRadixSort::SortKeys(tmp, *bytes, ki, ko, num, bb, eb, s));                            \
    if (vbytes != 4) return -1;                                                                                           \
    return (int)(desc ? rc::DeviceRadixSort::SortPairsDescending(tmp, *bytes, ki, ko, vi, vo, num, bb, eb, s)             \
                      : rc::DeviceRadixSort::SortPairs(tmp, *bytes, ki, ko, vi, vo, num, bb, eb, s));                     \
  }
  if (is_signed) GO128(__int128_t) else GO128(__uint128_t)
#undef GO128
#else
  return -2;
#endif
}

// ---- "best-known CUB on B200": the reference's own dispatch with NVIDIA's B200 tuning points injected through its
// SelectedPolicy hook (cub/device/dispatch/dispatch_radix_sort.cuh:1173), the way benchmarks/bench/radix_sort/pairs.cu:42-92
// builds its tuning variants.  Points (threads x items per thread) from the CUDA 12.9 toolkit's
// cub/device/dispatch/tuning/tuning_radix_sort.cuh (unused by that release's selector): u32/u32 pairs 448 x 20,
// u64/u32 pairs 352 x 12, u32 keys 512 x 21, f32 keys 512 x 20.  Speed comparator only (results equal the stock build's).
namespace {
template <typename KeyT, typename ValueT, typename OffsetT, int THREADS, int ITEMS>
struct tuned_policy_hub {
  using DominantT = rc::detail::conditional_t<(sizeof(ValueT) > sizeof(KeyT)), ValueT, KeyT>;
  struct policy_t : rc::ChainedPolicy<300, policy_t, policy_t> {
    static constexpr int ONESWEEP_RADIX_BITS = 8;
    static constexpr bool ONESWEEP = true;
    static constexpr bool OFFSET_64BIT = sizeof(OffsetT) == 8;
    using OnesweepPolicy = rc::AgentRadixSortOnesweepPolicy<THREADS, ITEMS, DominantT, 1, rc::RADIX_RANK_MATCH_EARLY_COUNTS_ANY,
                                                            rc::BLOCK_SCAN_RAKING_MEMOIZE, rc::RADIX_SORT_STORE_DIRECT,
                                                            ONESWEEP_RADIX_BITS>;
    using HistogramPolicy = rc::AgentRadixSortHistogramPolicy<128, 16, 1, KeyT, ONESWEEP_RADIX_BITS>;
    using ExclusiveSumPolicy = rc::AgentRadixSortExclusiveSumPolicy<256, ONESWEEP_RADIX_BITS>;
    using ScanPolicy = rc::AgentScanPolicy<512, 23, OffsetT, rc::BLOCK_LOAD_WARP_TRANSPOSE, rc::LOAD_DEFAULT,
                                           rc::BLOCK_STORE_WARP_TRANSPOSE, rc::BLOCK_SCAN_RAKING_MEMOIZE>;
    static constexpr int SINGLE_TILE_RADIX_BITS = (sizeof(KeyT) > 1) ? 6 : 5;
    using SingleTilePolicy = rc::AgentRadixSortDownsweepPolicy<256, 19, DominantT, rc::BLOCK_LOAD_DIRECT, rc::LOAD_LDG,
                                                               rc::RADIX_RANK_MEMOIZE, rc::BLOCK_SCAN_WARP_SCANS,
                                                               SINGLE_TILE_RADIX_BITS>;
  };
  using MaxPolicy = policy_t;
};

template <typename KeyT, typename ValueT, int THREADS, int ITEMS>
int tuned_sort(void* tmp, size_t* bytes, const void* kin, void* kout, const void* vin, void* vout, uint64_t n, int bb, int eb,
               cudaStream_t s) {
  using OffsetT = int;
  rc::DoubleBuffer<KeyT> dk(const_cast<KeyT*>((const KeyT*)kin), (KeyT*)kout);
  rc::DoubleBuffer<ValueT> dv(const_cast<ValueT*>((const ValueT*)vin), (ValueT*)vout);
  using Dispatch = rc::DispatchRadixSort<false, KeyT, ValueT, OffsetT, tuned_policy_hub<KeyT, ValueT, OffsetT, THREADS, ITEMS>>;
  return (int)Dispatch::Dispatch(tmp, *bytes, dk, dv, (OffsetT)n, bb, eb, /*is_overwrite_okay=*/false, s);
}
}  // namespace

// which: 0 = u32/u32 pairs (448 x 20), 1 = u64/u32 pairs (352 x 12), 2 = u32 keys (512 x 21), 3 = f32 keys (512 x 20);
// pointer form, ascending, n < 2^31
extern "C" int CAT(REF_PREFIX, _tuned_sort)(void* tmp, size_t* bytes, const void* kin, void* kout, const void* vin, void* vout,
                                            uint64_t n, int which, int bb, int eb, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  switch (which) {
    case 0: return tuned_sort<uint32_t, uint32_t, 448, 20>(tmp, bytes, kin, kout, vin, vout, n, bb, eb, s);
    case 1: return tuned_sort<unsigned long long, uint32_t, 352, 12>(tmp, bytes, kin, kout, vin, vout, n, bb, eb, s);
    case 2: return tuned_sort<uint32_t, rc::NullType, 512, 21>(tmp, bytes, kin, kout, nullptr, nullptr, n, bb, eb, s);
    case 3: return tuned_sort<float, rc::NullType, 512, 20>(tmp, bytes, kin, kout, nullptr, nullptr, n, bb, eb, s);
    default: return -1;
  }
}
