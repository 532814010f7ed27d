// include/b200/device_segmented_radix_sort.cuh -- C++ veneer with the signatures of cub::DeviceSegmentedRadixSort,
// forwarding to the C-ABI of libb2s.so (b2s_segmented_radix_sort[_db], include/b2s_radix_sort.h).
//
// A reference call site recompiles by swapping the namespace:
//     cub::DeviceSegmentedRadixSort::SortPairs(...)  ->  b200::DeviceSegmentedRadixSort::SortPairs(...)
// Mirrors (signatures only; no reference code): cub/device/device_segmented_radix_sort.cuh -- SortPairs / SortPairsDescending /
// SortKeys / SortKeysDescending, pointer and DoubleBuffer forms, `int num_items, int num_segments`, begin / end offset
// "iterators".  The reference accepts any random-access iterator for the offsets; a C-ABI cannot, so this veneer takes
// POINTERS to 32- or 64-bit integers (what the reference's tests and documentation use: `int* d_offsets`,
// `d_offsets + 1`); anything else fails to compile with a clear message.
#pragma once
#include "device_radix_sort.cuh"

namespace b200 {
namespace detail {
template <typename It>
struct offset_pointer {
  static_assert(std::is_pointer<It>::value, "segment offsets must be pointers to 32- or 64-bit integers in device memory");
  using T = typename std::remove_cv<typename std::remove_pointer<It>::type>::type;
  static_assert(std::is_integral<T>::value && (sizeof(T) == 4 || sizeof(T) == 8), "segment offsets must be 32- or 64-bit integers");
  static constexpr int bytes = (int)sizeof(T);
};

template <bool DESC, typename KeyT, typename ValueT, typename BeginIt, typename EndIt>
cudaError_t seg_ptr(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in, KeyT* d_keys_out,
                    const ValueT* d_values_in, ValueT* d_values_out, long long num_items, long long num_segments,
                    BeginIt d_begin_offsets, EndIt d_end_offsets, int begin_bit, int end_bit, cudaStream_t stream) {
  static_assert(offset_pointer<BeginIt>::bytes == offset_pointer<EndIt>::bytes, "begin and end offsets must have one type");
  return (cudaError_t)b2s_segmented_radix_sort(d_temp_storage, &temp_storage_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out,
                                               (uint64_t)num_items, (uint64_t)num_segments, d_begin_offsets, d_end_offsets,
                                               offset_pointer<BeginIt>::bytes, key_enum<KeyT>::value, value_bytes<ValueT>(),
                                               DESC ? 1 : 0, begin_bit, end_bit, (b2s_stream_t)stream);
}
template <bool DESC, typename KeyT, typename ValueT, typename BeginIt, typename EndIt>
cudaError_t seg_db(void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys, DoubleBuffer<ValueT>* d_values,
                   long long num_items, long long num_segments, BeginIt d_begin_offsets, EndIt d_end_offsets, int begin_bit,
                   int end_bit, cudaStream_t stream) {
  static_assert(offset_pointer<BeginIt>::bytes == offset_pointer<EndIt>::bytes, "begin and end offsets must have one type");
  void* kb[2] = {d_keys.d_buffers[0], d_keys.d_buffers[1]};
  void* vb[2] = {d_values ? (void*)d_values->d_buffers[0] : nullptr, d_values ? (void*)d_values->d_buffers[1] : nullptr};
  int vsel = d_values ? d_values->selector : 0;
  cudaError_t e = (cudaError_t)b2s_segmented_radix_sort_db(
      d_temp_storage, &temp_storage_bytes, kb, &d_keys.selector, d_values ? vb : nullptr, d_values ? &vsel : nullptr, (uint64_t)num_items,
      (uint64_t)num_segments, d_begin_offsets, d_end_offsets, offset_pointer<BeginIt>::bytes, key_enum<KeyT>::value,
      value_bytes<ValueT>(), DESC ? 1 : 0, begin_bit, end_bit, (b2s_stream_t)stream);
  if (d_values) d_values->selector = vsel;
  return e;
}
}  // namespace detail

struct DeviceSegmentedRadixSort {
#define B200_SEG_PAIRS(NAME, DESC)                                                                                              \
  template <typename KeyT, typename ValueT, typename BeginOffsetIteratorT, typename EndOffsetIteratorT>                         \
  static cudaError_t NAME(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in, KeyT* d_keys_out,            \
                          const ValueT* d_values_in, ValueT* d_values_out, int num_items, int num_segments,                     \
                          BeginOffsetIteratorT d_begin_offsets, EndOffsetIteratorT d_end_offsets, int begin_bit = 0,            \
                          int end_bit = sizeof(KeyT) * 8, cudaStream_t stream = 0) {                                            \
    return detail::seg_ptr<DESC>(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out, num_items, \
                                 num_segments, d_begin_offsets, d_end_offsets, begin_bit, end_bit, stream);                     \
  }                                                                                                                             \
  template <typename KeyT, typename ValueT, typename BeginOffsetIteratorT, typename EndOffsetIteratorT>                         \
  static cudaError_t NAME(void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys,                         \
                          DoubleBuffer<ValueT>& d_values, int num_items, int num_segments, BeginOffsetIteratorT d_begin_offsets, \
                          EndOffsetIteratorT d_end_offsets, int begin_bit = 0, int end_bit = sizeof(KeyT) * 8,                  \
                          cudaStream_t stream = 0) {                                                                            \
    return detail::seg_db<DESC>(d_temp_storage, temp_storage_bytes, d_keys, &d_values, num_items, num_segments, d_begin_offsets, \
                                d_end_offsets, begin_bit, end_bit, stream);                                                     \
  }
#define B200_SEG_KEYS(NAME, DESC)                                                                                               \
  template <typename KeyT, typename BeginOffsetIteratorT, typename EndOffsetIteratorT>                                          \
  static cudaError_t NAME(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in, KeyT* d_keys_out,            \
                          int num_items, int num_segments, BeginOffsetIteratorT d_begin_offsets,                                \
                          EndOffsetIteratorT d_end_offsets, int begin_bit = 0, int end_bit = sizeof(KeyT) * 8,                  \
                          cudaStream_t stream = 0) {                                                                            \
    return detail::seg_ptr<DESC, KeyT, NullType>(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, nullptr, nullptr,   \
                                                 num_items, num_segments, d_begin_offsets, d_end_offsets, begin_bit, end_bit,   \
                                                 stream);                                                                       \
  }                                                                                                                             \
  template <typename KeyT, typename BeginOffsetIteratorT, typename EndOffsetIteratorT>                                          \
  static cudaError_t NAME(void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys, int num_items,          \
                          int num_segments, BeginOffsetIteratorT d_begin_offsets, EndOffsetIteratorT d_end_offsets,             \
                          int begin_bit = 0, int end_bit = sizeof(KeyT) * 8, cudaStream_t stream = 0) {                         \
    return detail::seg_db<DESC, KeyT, NullType>(d_temp_storage, temp_storage_bytes, d_keys, nullptr, num_items, num_segments,   \
                                                d_begin_offsets, d_end_offsets, begin_bit, end_bit, stream);                    \
  }
  B200_SEG_PAIRS(SortPairs, false)
  B200_SEG_PAIRS(SortPairsDescending, true)
  B200_SEG_KEYS(SortKeys, false)
  B200_SEG_KEYS(SortKeysDescending, true)
#undef B200_SEG_PAIRS
#undef B200_SEG_KEYS
};

}  // namespace b200
