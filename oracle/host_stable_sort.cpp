// oracle/host_stable_sort.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
//
// Restatement of the reference test harness' host solution, the "test harness's host
// std::stable_sort reference" that BASELINE.json names as the CPU baseline:
//   test/test_device_radix_sort.cu:668-680   Pair<KeyT, index> with operator< on the TYPED key
//   test/test_device_radix_sort.cu:898-936   mask the raw bits to [begin_bit,end_bit) when the range is partial,
//                                            descending = reverse -> std::stable_sort -> reverse
//   test/test_device_radix_sort.cu:945-953   expected keys = ORIGINAL keys permuted by the ranks
// Half / bfloat16 keys compare through float like test/half.h and test/bfloat16.h do.
// threads == 1 is exactly the harness (std::stable_sort); threads > 1 uses
// __gnu_parallel::stable_sort on that many OpenMP threads (reported, not part of the harness).
//
// Like the harness, this is only meaningful where the harness uses it: full bit range for
// every key type, partial ranges for unsigned integer keys (test_device_radix_sort.cu:1470-1473),
// and inputs without NaN (test/test_util.h:518-520).  The bit-level oracle is oracle/radix_oracle.c.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <parallel/algorithm>
#include <omp.h>
#include <vector>

namespace {

float half_to_float(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1f;
  uint32_t man = h & 0x3ffu;
  uint32_t f;
  if (exp == 0) {
    if (man == 0) {
      f = sign;
    } else {  // subnormal: normalise
      int e = -1;
      do { ++e; man <<= 1; } while (!(man & 0x400u));
      f = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
    }
  } else if (exp == 31) {
    f = sign | 0x7f800000u | (man << 13);
  } else {
    f = sign | ((exp + 127 - 15) << 23) | (man << 13);
  }
  float r;
  std::memcpy(&r, &f, 4);
  return r;
}

struct half_key {
  uint16_t bits;
  bool operator<(const half_key& o) const { return half_to_float(bits) < half_to_float(o.bits); }
};
struct bf16_key {
  uint16_t bits;
  static float cvt(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; std::memcpy(&f, &u, 4); return f; }
  bool operator<(const bf16_key& o) const { return cvt(bits) < cvt(o.bits); }
};

template <typename KeyT, typename IndexT>
struct Pair {
  KeyT key;
  IndexT value;
  bool operator<(const Pair& b) const { return key < b.key; }
};

template <typename KeyT, typename UnsignedBits, typename IndexT>
int solve(const void* keys_v, uint64_t n, int descending, int begin_bit, int end_bit, void* out_keys_v,
          uint64_t* out_ranks, int threads) {
  const KeyT* h_keys = static_cast<const KeyT*>(keys_v);
  KeyT* out_keys = static_cast<KeyT*>(out_keys_v);
  using PairT = Pair<KeyT, IndexT>;
  std::vector<PairT> pairs(n);
  const int num_bits = end_bit - begin_bit;
  for (uint64_t i = 0; i < n; ++i) {
    if (num_bits < static_cast<int>(sizeof(KeyT) * 8)) {
      UnsignedBits base = 0;
      std::memcpy(&base, &h_keys[i], sizeof(KeyT));
      base &= (num_bits <= 0 ? UnsignedBits(0) : UnsignedBits(((UnsignedBits{1} << num_bits) - 1) << begin_bit));
      std::memcpy(&pairs[i].key, &base, sizeof(KeyT));
    } else {
      pairs[i].key = h_keys[i];
    }
    pairs[i].value = static_cast<IndexT>(i);
  }
  if (descending) std::reverse(pairs.begin(), pairs.end());
  if (threads <= 1) {
    std::stable_sort(pairs.begin(), pairs.end());
  } else {
    omp_set_num_threads(threads);
    __gnu_parallel::stable_sort(pairs.begin(), pairs.end());
  }
  if (descending) std::reverse(pairs.begin(), pairs.end());
  for (uint64_t i = 0; i < n; ++i) {
    if (out_ranks) out_ranks[i] = pairs[i].value;
    if (out_keys) out_keys[i] = h_keys[pairs[i].value];
  }
  return 0;
}

template <typename KeyT, typename UnsignedBits>
int solve_idx(const void* k, uint64_t n, int d, int bb, int eb, void* ok, uint64_t* r, int t) {
  return n < (1ull << 32) ? solve<KeyT, UnsignedBits, uint32_t>(k, n, d, bb, eb, ok, r, t)
                          : solve<KeyT, UnsignedBits, uint64_t>(k, n, d, bb, eb, ok, r, t);
}

}  // namespace

extern "C" int host_max_threads(void) { return omp_get_max_threads(); }

// out_keys: n keys (nullable); out_ranks: n uint64 source indices (nullable).
extern "C" int host_stable_sort_solution(const void* keys, uint64_t n, int key_type, int descending, int begin_bit,
                                         int end_bit, void* out_keys, uint64_t* out_ranks, int threads) {
  switch (key_type) {
    case 0: return solve_idx<uint8_t, uint8_t>(keys, n, descending, begin_bit, end_bit, out_keys, out_ranks, threads);
    case 1: return solve_idx<int8_t, uint8_t>(keys, n, descending, begin_bit, end_bit, out_keys, out_ranks, threads);
    case 2: return solve_idx<uint16_t, uint16_t>(keys, n, descending, begin_bit, end_bit, out_keys, out_ranks, threads);
    case 3: return solve_idx<int16_t, uint16_t>(keys, n, descending, begin_bit, end_bit, out_keys, out_ranks, threads);
    case 4: return solve_idx<half_key, uint16_t>(keys, n, descending, begin_bit, end_bit, out_keys, out_ranks, threads);
    case 5: return solve_idx<bf16_key, uint16_t>(keys, n, descending, begin_bit, end_bit, out_keys, out_ranks, threads);
    case 6: return solve_idx<uint32_t, uint32_t>(keys, n, descending, begin_bit, end_bit, out_keys, out_ranks, threads);
    case 7: return solve_idx<int32_t, uint32_t>(keys, n, descending, begin_bit, end_bit, out_keys, out_ranks, threads);
    case 8: return solve_idx<float, uint32_t>(keys, n, descending, begin_bit, end_bit, out_keys, out_ranks, threads);
    case 9: return solve_idx<uint64_t, uint64_t>(keys, n, descending, begin_bit, end_bit, out_keys, out_ranks, threads);
    case 10: return solve_idx<int64_t, uint64_t>(keys, n, descending, begin_bit, end_bit, out_keys, out_ranks, threads);
    case 11: return solve_idx<double, uint64_t>(keys, n, descending, begin_bit, end_bit, out_keys, out_ranks, threads);
    default: return 1;
  }
}
