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
//   decomposer overloads for user-defined key structs: :486-530, 625-666, 922-962, 1055-1105, 1368-1426, 1515-1563,
//   1816-1856, 1949-... (b2s_radix_sort_struct); the deprecated overloads with a trailing `bool debug_synchronous`
//   (:366, 821, 1265, 1715, ...) are accepted and the flag ignored, as the reference itself does since 2.0.
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

#if __has_include(<cuda/std/tuple>)
#include <cuda/std/tuple>
#define B200_HAS_CUDA_STD_TUPLE 1
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
#if defined(__SIZEOF_INT128__)
B200_KEY(__uint128_t, B2S_U128);  // cub/util_type.cuh:1225,1259
B200_KEY(__int128_t, B2S_I128);
#endif
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
                  "value types of 1, 2, 4, 8 or 16 bytes are supported (any size up to 64 bytes with struct / 128-bit keys)");
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

#ifdef B200_HAS_CUDA_STD_TUPLE
// Type erasure of a decomposer: apply it to a probe object and record, for every element of the returned tuple of
// references (most significant first), its byte offset inside the key and its fundamental type.
template <typename KeyT, typename TupleT, size_t... I>
void record_fields(const unsigned char* base, TupleT& tup, b2s_field_t* out, std::index_sequence<I...>) {
  ((out[I].offset = (int32_t)(reinterpret_cast<const unsigned char*>(&::cuda::std::get<I>(tup)) - base),
    out[I].key_type = key_enum<typename std::remove_cv<typename std::remove_reference<
        typename ::cuda::std::tuple_element<I, TupleT>::type>::type>::type>::value),
   ...);
}
template <typename KeyT, typename DecomposerT>
int decompose(DecomposerT& decomposer, b2s_field_t (&out)[B2S_MAX_STRUCT_FIELDS]) {
  alignas(KeyT) unsigned char raw[sizeof(KeyT)] = {};
  KeyT& probe = *reinterpret_cast<KeyT*>(raw);
  auto tup = decomposer(probe);
  constexpr size_t N = ::cuda::std::tuple_size<decltype(tup)>::value;
  static_assert(N >= 1 && N <= B2S_MAX_STRUCT_FIELDS, "a decomposer may name 1..8 arithmetic members");
  record_fields<KeyT>(raw, tup, out, std::make_index_sequence<N>{});
  return (int)N;
}
template <typename D>
using enable_if_decomposer = typename std::enable_if<!std::is_convertible<D, int>::value, cudaError_t>::type;

template <bool DESC, typename KeyT, typename ValueT, typename NumItemsT, typename DecomposerT>
cudaError_t struct_ptr(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in, KeyT* d_keys_out,
                       const ValueT* d_values_in, ValueT* d_values_out, NumItemsT num_items, DecomposerT decomposer,
                       int begin_bit, int end_bit, cudaStream_t stream) {
  b2s_field_t f[B2S_MAX_STRUCT_FIELDS];
  const int nf = decompose<KeyT>(decomposer, f);
  constexpr int vb = std::is_same<ValueT, NullType>::value ? 0 : (int)sizeof(ValueT);
  return (cudaError_t)b2s_radix_sort_struct(d_temp_storage, &temp_storage_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out,
                                            (uint64_t)num_items, (int)sizeof(KeyT), f, nf, vb, DESC ? 1 : 0, begin_bit, end_bit,
                                            (b2s_stream_t)stream);
}
template <bool DESC, typename KeyT, typename ValueT, typename NumItemsT, typename DecomposerT>
cudaError_t struct_db(void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys, DoubleBuffer<ValueT>* d_values,
                      NumItemsT num_items, DecomposerT decomposer, int begin_bit, int end_bit, cudaStream_t stream) {
  b2s_field_t f[B2S_MAX_STRUCT_FIELDS];
  const int nf = decompose<KeyT>(decomposer, f);
  constexpr int vb = std::is_same<ValueT, NullType>::value ? 0 : (int)sizeof(ValueT);
  void* kb[2] = {d_keys.d_buffers[0], d_keys.d_buffers[1]};
  void* vbuf[2] = {d_values ? (void*)d_values->d_buffers[0] : nullptr, d_values ? (void*)d_values->d_buffers[1] : nullptr};
  int vsel = d_values ? d_values->selector : 0;
  cudaError_t e = (cudaError_t)b2s_radix_sort_struct_db(d_temp_storage, &temp_storage_bytes, kb, &d_keys.selector,
                                                        d_values ? vbuf : nullptr, d_values ? &vsel : nullptr, (uint64_t)num_items,
                                                        (int)sizeof(KeyT), f, nf, vb, DESC ? 1 : 0, begin_bit, end_bit,
                                                        (b2s_stream_t)stream);
  if (d_values) d_values->selector = vsel;
  return e;
}
#endif  // B200_HAS_CUDA_STD_TUPLE
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

  // ---- deprecated overloads with a trailing `bool debug_synchronous` (device_radix_sort.cuh:366, 821, 1265, 1715, 2155,
  //      2565, 2970, 3370): the flag is ignored, as in the reference since CUB 2.0 ----------------------------------------
#define B200_DEBUG_SYNC_OVERLOADS(NAME)                                                                                        \
  template <typename KeyT, typename ValueT, typename NumItemsT>                                                                \
  static cudaError_t NAME(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in, KeyT* d_keys_out,           \
                          const ValueT* d_values_in, ValueT* d_values_out, NumItemsT num_items, int begin_bit, int end_bit,    \
                          cudaStream_t stream, bool /*debug_synchronous*/) {                                                   \
    return NAME(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out, num_items, begin_bit,    \
                end_bit, stream);                                                                                              \
  }                                                                                                                            \
  template <typename KeyT, typename ValueT, typename NumItemsT>                                                                \
  static cudaError_t NAME(void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys,                        \
                          DoubleBuffer<ValueT>& d_values, NumItemsT num_items, int begin_bit, int end_bit, cudaStream_t stream,\
                          bool /*debug_synchronous*/) {                                                                        \
    return NAME(d_temp_storage, temp_storage_bytes, d_keys, d_values, num_items, begin_bit, end_bit, stream);                  \
  }
#define B200_DEBUG_SYNC_KEYS_OVERLOADS(NAME)                                                                                   \
  template <typename KeyT, typename NumItemsT>                                                                                 \
  static cudaError_t NAME(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in, KeyT* d_keys_out,           \
                          NumItemsT num_items, int begin_bit, int end_bit, cudaStream_t stream, bool /*debug_synchronous*/) {  \
    return NAME(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, num_items, begin_bit, end_bit, stream);             \
  }                                                                                                                            \
  template <typename KeyT, typename NumItemsT>                                                                                 \
  static cudaError_t NAME(void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys, NumItemsT num_items,   \
                          int begin_bit, int end_bit, cudaStream_t stream, bool /*debug_synchronous*/) {                       \
    return NAME(d_temp_storage, temp_storage_bytes, d_keys, num_items, begin_bit, end_bit, stream);                            \
  }
  B200_DEBUG_SYNC_OVERLOADS(SortPairs)
  B200_DEBUG_SYNC_OVERLOADS(SortPairsDescending)
  B200_DEBUG_SYNC_KEYS_OVERLOADS(SortKeys)
  B200_DEBUG_SYNC_KEYS_OVERLOADS(SortKeysDescending)
#undef B200_DEBUG_SYNC_OVERLOADS
#undef B200_DEBUG_SYNC_KEYS_OVERLOADS

#ifdef B200_HAS_CUDA_STD_TUPLE
  // ---- decomposer overloads for user-defined key structs: with [begin_bit, end_bit) and without (all bits) --------------
#define B200_DECOMPOSER_PAIRS(NAME, DESC)                                                                                      \
  template <typename KeyT, typename ValueT, typename NumItemsT, typename DecomposerT>                                          \
  static detail::enable_if_decomposer<DecomposerT> NAME(void* d_temp_storage, size_t& temp_storage_bytes,                      \
      const KeyT* d_keys_in, KeyT* d_keys_out, const ValueT* d_values_in, ValueT* d_values_out, NumItemsT num_items,           \
      DecomposerT decomposer, int begin_bit, int end_bit, cudaStream_t stream = 0) {                                           \
    return detail::struct_ptr<DESC>(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out,      \
                                    num_items, decomposer, begin_bit, end_bit, stream);                                        \
  }                                                                                                                            \
  template <typename KeyT, typename ValueT, typename NumItemsT, typename DecomposerT>                                          \
  static detail::enable_if_decomposer<DecomposerT> NAME(void* d_temp_storage, size_t& temp_storage_bytes,                      \
      const KeyT* d_keys_in, KeyT* d_keys_out, const ValueT* d_values_in, ValueT* d_values_out, NumItemsT num_items,           \
      DecomposerT decomposer, cudaStream_t stream = 0) {                                                                       \
    return detail::struct_ptr<DESC>(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out,      \
                                    num_items, decomposer, 0, -1, stream);                                                     \
  }                                                                                                                            \
  template <typename KeyT, typename ValueT, typename NumItemsT, typename DecomposerT>                                          \
  static detail::enable_if_decomposer<DecomposerT> NAME(void* d_temp_storage, size_t& temp_storage_bytes,                      \
      DoubleBuffer<KeyT>& d_keys, DoubleBuffer<ValueT>& d_values, NumItemsT num_items, DecomposerT decomposer, int begin_bit,  \
      int end_bit, cudaStream_t stream = 0) {                                                                                  \
    return detail::struct_db<DESC>(d_temp_storage, temp_storage_bytes, d_keys, &d_values, num_items, decomposer, begin_bit,    \
                                   end_bit, stream);                                                                           \
  }                                                                                                                            \
  template <typename KeyT, typename ValueT, typename NumItemsT, typename DecomposerT>                                          \
  static detail::enable_if_decomposer<DecomposerT> NAME(void* d_temp_storage, size_t& temp_storage_bytes,                      \
      DoubleBuffer<KeyT>& d_keys, DoubleBuffer<ValueT>& d_values, NumItemsT num_items, DecomposerT decomposer,                 \
      cudaStream_t stream = 0) {                                                                                               \
    return detail::struct_db<DESC>(d_temp_storage, temp_storage_bytes, d_keys, &d_values, num_items, decomposer, 0, -1,        \
                                   stream);                                                                                    \
  }
#define B200_DECOMPOSER_KEYS(NAME, DESC)                                                                                       \
  template <typename KeyT, typename NumItemsT, typename DecomposerT>                                                           \
  static detail::enable_if_decomposer<DecomposerT> NAME(void* d_temp_storage, size_t& temp_storage_bytes,                      \
      const KeyT* d_keys_in, KeyT* d_keys_out, NumItemsT num_items, DecomposerT decomposer, int begin_bit, int end_bit,        \
      cudaStream_t stream = 0) {                                                                                               \
    return detail::struct_ptr<DESC, KeyT, NullType>(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, nullptr,        \
                                                    nullptr, num_items, decomposer, begin_bit, end_bit, stream);               \
  }                                                                                                                            \
  template <typename KeyT, typename NumItemsT, typename DecomposerT>                                                           \
  static detail::enable_if_decomposer<DecomposerT> NAME(void* d_temp_storage, size_t& temp_storage_bytes,                      \
      const KeyT* d_keys_in, KeyT* d_keys_out, NumItemsT num_items, DecomposerT decomposer, cudaStream_t stream = 0) {         \
    return detail::struct_ptr<DESC, KeyT, NullType>(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, nullptr,        \
                                                    nullptr, num_items, decomposer, 0, -1, stream);                            \
  }                                                                                                                            \
  template <typename KeyT, typename NumItemsT, typename DecomposerT>                                                           \
  static detail::enable_if_decomposer<DecomposerT> NAME(void* d_temp_storage, size_t& temp_storage_bytes,                      \
      DoubleBuffer<KeyT>& d_keys, NumItemsT num_items, DecomposerT decomposer, int begin_bit, int end_bit,                     \
      cudaStream_t stream = 0) {                                                                                               \
    return detail::struct_db<DESC, KeyT, NullType>(d_temp_storage, temp_storage_bytes, d_keys, nullptr, num_items, decomposer, \
                                                   begin_bit, end_bit, stream);                                                \
  }                                                                                                                            \
  template <typename KeyT, typename NumItemsT, typename DecomposerT>                                                           \
  static detail::enable_if_decomposer<DecomposerT> NAME(void* d_temp_storage, size_t& temp_storage_bytes,                      \
      DoubleBuffer<KeyT>& d_keys, NumItemsT num_items, DecomposerT decomposer, cudaStream_t stream = 0) {                      \
    return detail::struct_db<DESC, KeyT, NullType>(d_temp_storage, temp_storage_bytes, d_keys, nullptr, num_items, decomposer, \
                                                   0, -1, stream);                                                             \
  }
  B200_DECOMPOSER_PAIRS(SortPairs, false)
  B200_DECOMPOSER_PAIRS(SortPairsDescending, true)
  B200_DECOMPOSER_KEYS(SortKeys, false)
  B200_DECOMPOSER_KEYS(SortKeysDescending, true)
#undef B200_DECOMPOSER_PAIRS
#undef B200_DECOMPOSER_KEYS
#endif  // B200_HAS_CUDA_STD_TUPLE
};

}  // namespace b200
