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
