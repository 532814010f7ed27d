// include/b200/device_radix_sort.cuh -- C++ veneer with the EXACT signatures of cub::DeviceRadixSort for
// fundamental key types, forwarding to the C-ABI of libb2s.so (include/b2s_radix_sort.h).
//
// A reference call site recompiles by swapping the namespace:
//     cub::DeviceRadixSort::SortPairs(...)   ->   b200::DeviceRadixSort::SortPairs(...)
//     cub::DoubleBuffer<T>                   ->   b200::DoubleBuffer<T>   (layout-compatible: T* d_buffers[2]; int selector)
//
// Mirrors (signatures only; no reference code):
//   cub/device/device_radix_sort.cuh:312 (SortPairs), :781 (SortPairs DoubleBuffer), :1214/:1675 (SortPairsDescending),
//   :2106/:2525 (SortKeys), :2921/:3330 (SortKeysDescending); cub/util_type.cuh:854-886 (DoubleBuffer);
//   cub/detail/choose_offset.cuh:44-57 (NumItemsT -> offset width).
// Out of scope here: decomposer overloads for user-defined key structs, 128-bit keys.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <type_traits>

#if __has_include(<cuda_fp16.h>)
#include <cuda_fp16.h>
#define B200_HAS_HALF 1
#endif
#if __has_include(<cuda_bf16.h>)
#include <cuda_bf16.h>
#define B200_HAS_BF16 1
#endif

#include "../b2s_radix_sort.h"

namespace b200 {

template <typename T>
struct DoubleBuffer {
  T* d_buffers[2];
  int selector;
  DoubleBuffer() : d_buffers{nullptr, nullptr}, selector(0) {}
  DoubleBuffer(T* d_current, T* d_alternate) : d_buffers{d_current, d_alternate}, selector(0) {}
  T* Current() { return d_buffers[selector]; }
  T* Alternate() { return d_buffers[selector ^ 1]; }
};

struct NullType {};

namespace detail {
template <typename K> struct key_enum;  // undefined for unsupported key types -> compile error, like the reference's traits
#define B200_KEY(T, E) template <> struct key_enum<T> { static constexpr int value = E; }
B200_KEY(unsigned char, B2S_U8);
B200_KEY(signed char, B2S_I8);
B200_KEY(char, (std::is_signed<char>::value ? B2S_I8 : B2S_U8));
B200_KEY(unsigned short, B2S_U16);
B200_KEY(short, B2S_I16);
B200_KEY(unsigned int, B2S_U32);
B200_KEY(int, B2S_I32);
B200_KEY(float, B2S_F32);
B200_KEY(unsigned long, (sizeof(unsigned long) == 8 ? B2S_U64 : B2S_U32));
B200_KEY(long, (sizeof(long) == 8 ? B2S_I64 : B2S_I32));
B200_KEY(unsigned long long, B2S_U64);
B200_KEY(long long, B2S_I64);
B200_KEY(double, B2S_F64);
#ifdef B200_HAS_HALF
B200_KEY(__half, B2S_F16);
#endif
#ifdef B200_HAS_BF16
B200_KEY(__nv_bfloat16, B2S_BF16);
#endif
#undef B200_KEY

template <typename V>
constexpr int value_bytes() {
  if constexpr (std::is_same<V, NullType>::value) {
    return 0;
  } else {
    static_assert(sizeof(V) == 1 || sizeof(V) == 2 || sizeof(V) == 4 || sizeof(V) == 8 || sizeof(V) == 16,
                  "value types of 1, 2, 4, 8 or 16 bytes are supported");
    return (int)sizeof(V);
  }
}
template <typename N>
constexpr int offset_bytes() {
  static_assert(std::is_integral<N>::value && !std::is_same<typename std::remove_cv<N>::type, bool>::value,
                "NumItemsT must be an integral type, but not bool");
  return sizeof(N) <= 4 ? 4 : 8;
}

template <bool DESC, typename KeyT, typename ValueT, typename NumItemsT>
cudaError_t sort_ptr(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in, KeyT* d_keys_out,
                     const ValueT* d_values_in, ValueT* d_values_out, NumItemsT num_items, int begin_bit, int end_bit,
                     cudaStream_t stream) {
  return (cudaError_t)b2s_radix_sort(d_temp_storage, &temp_storage_bytes, d_keys_in, d_keys_out, d_values_in,
                                     d_values_out, (uint64_t)num_items, key_enum<KeyT>::value, value_bytes<ValueT>(),
                                     offset_bytes<NumItemsT>(), DESC ? 1 : 0, begin_bit, end_bit, (b2s_stream_t)stream);
}

template <bool DESC, typename KeyT, typename ValueT, typename NumItemsT>
cudaError_t sort_db(void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys,
                    DoubleBuffer<ValueT>* d_values, NumItemsT num_items, int begin_bit, int end_bit,
                    cudaStream_t stream) {
  void* kb[2] = {d_keys.d_buffers[0], d_keys.d_buffers[1]};
  void* vb[2] = {d_values ? (void*)d_values->d_buffers[0] : nullptr, d_values ? (void*)d_values->d_buffers[1] : nullptr};
  int vsel = d_values ? d_values->selector : 0;
  cudaError_t e = (cudaError_t)b2s_radix_sort_db(d_temp_storage, &temp_storage_bytes, kb, &d_keys.selector,
                                                 d_values ? vb : nullptr, d_values ? &vsel : nullptr,
                                                 (uint64_t)num_items, key_enum<KeyT>::value, value_bytes<ValueT>(),
                                                 offset_bytes<NumItemsT>(), DESC ? 1 : 0, begin_bit, end_bit,
                                                 (b2s_stream_t)stream);
  if (d_values) d_values->selector = vsel;
  return e;
}
}  // namespace detail

struct DeviceRadixSort {
  // ---- SortPairs ---------------------------------------------------------------------------
  template <typename KeyT, typename ValueT, typename NumItemsT>
  static cudaError_t SortPairs(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in,
                               KeyT* d_keys_out, const ValueT* d_values_in, ValueT* d_values_out, NumItemsT num_items,
                               int begin_bit = 0, int end_bit = sizeof(KeyT) * 8, cudaStream_t stream = 0) {
    return detail::sort_ptr<false>(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_values_in,
                                   d_values_out, num_items, begin_bit, end_bit, stream);
  }
  template <typename KeyT, typename ValueT, typename NumItemsT>
  static cudaError_t SortPairs(void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys,
                               DoubleBuffer<ValueT>& d_values, NumItemsT num_items, int begin_bit = 0,
                               int end_bit = sizeof(KeyT) * 8, cudaStream_t stream = 0) {
    return detail::sort_db<false>(d_temp_storage, temp_storage_bytes, d_keys, &d_values, num_items, begin_bit, end_bit,
                                  stream);
  }
  // ---- SortPairsDescending -----------------------------------------------------------------
  template <typename KeyT, typename ValueT, typename NumItemsT>
  static cudaError_t SortPairsDescending(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in,
                                         KeyT* d_keys_out, const ValueT* d_values_in, ValueT* d_values_out,
                                         NumItemsT num_items, int begin_bit = 0, int end_bit = sizeof(KeyT) * 8,
                                         cudaStream_t stream = 0) {
    return detail::sort_ptr<true>(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_values_in,
                                  d_values_out, num_items, begin_bit, end_bit, stream);
  }
  template <typename KeyT, typename ValueT, typename NumItemsT>
  static cudaError_t SortPairsDescending(void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys,
                                         DoubleBuffer<ValueT>& d_values, NumItemsT num_items, int begin_bit = 0,
                                         int end_bit = sizeof(KeyT) * 8, cudaStream_t stream = 0) {
    return detail::sort_db<true>(d_temp_storage, temp_storage_bytes, d_keys, &d_values, num_items, begin_bit, end_bit,
                                 stream);
  }
  // ---- SortKeys ----------------------------------------------------------------------------
  template <typename KeyT, typename NumItemsT>
  static cudaError_t SortKeys(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in,
                              KeyT* d_keys_out, NumItemsT num_items, int begin_bit = 0,
                              int end_bit = sizeof(KeyT) * 8, cudaStream_t stream = 0) {
    return detail::sort_ptr<false, KeyT, NullType>(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, nullptr,
                                                   nullptr, num_items, begin_bit, end_bit, stream);
  }
  template <typename KeyT, typename NumItemsT>
  static cudaError_t SortKeys(void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys,
                              NumItemsT num_items, int begin_bit = 0, int end_bit = sizeof(KeyT) * 8,
                              cudaStream_t stream = 0) {
    return detail::sort_db<false, KeyT, NullType>(d_temp_storage, temp_storage_bytes, d_keys, nullptr, num_items,
                                                  begin_bit, end_bit, stream);
  }
  // ---- SortKeysDescending ------------------------------------------------------------------
  template <typename KeyT, typename NumItemsT>
  static cudaError_t SortKeysDescending(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in,
                                        KeyT* d_keys_out, NumItemsT num_items, int begin_bit = 0,
                                        int end_bit = sizeof(KeyT) * 8, cudaStream_t stream = 0) {
    return detail::sort_ptr<true, KeyT, NullType>(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, nullptr,
                                                  nullptr, num_items, begin_bit, end_bit, stream);
  }
  template <typename KeyT, typename NumItemsT>
  static cudaError_t SortKeysDescending(void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys,
                                        NumItemsT num_items, int begin_bit = 0, int end_bit = sizeof(KeyT) * 8,
                                        cudaStream_t stream = 0) {
    return detail::sort_db<true, KeyT, NullType>(d_temp_storage, temp_storage_bytes, d_keys, nullptr, num_items,
                                                 begin_bit, end_bit, stream);
  }
};

}  // namespace b200
