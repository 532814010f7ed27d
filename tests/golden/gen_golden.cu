// tests/golden/gen_golden.cu -- fixture generator (run in the build container only).
// Draws test inputs with the REFERENCE's own generators, included from where they lie
// (/root/reference/test/test_util.h: RandomBits :473-522, InitValue :545-626, MT19937 seeding
// :136-138 via test/mersenne.h), and solves them the way the reference harness does
// (test/test_device_radix_sort.cu:896-956: Pair<key,index>, std::stable_sort, reverse for
// descending).  Output: raw little-endian records on stdout, parsed by make_golden.py.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "test_util.h"
#include "half.h"
#include "bfloat16.h"

template <typename KeyT>
struct Pair {
  KeyT key;
  uint32_t value;
  bool operator<(const Pair& b) const { return key < b.key; }
};

template <typename KeyT>
void emit(const char* name, GenMode mode, int n, int begin_bit, int end_bit) {
  using UnsignedBits = typename cub::Traits<KeyT>::UnsignedBits;
  std::vector<KeyT> keys(n);
  for (int i = 0; i < n; ++i) InitValue(mode, keys[i], i);
  for (int desc = 0; desc < 2; ++desc) {
    std::vector<Pair<KeyT>> p(n);
    const int num_bits = end_bit - begin_bit;
    for (int i = 0; i < n; ++i) {
      if (num_bits < (int)sizeof(KeyT) * 8) {
        UnsignedBits base = 0;
        memcpy(&base, &keys[i], sizeof(KeyT));
        base &= ((UnsignedBits{1} << num_bits) - 1) << begin_bit;
        memcpy(&p[i].key, &base, sizeof(KeyT));
      } else {
        p[i].key = keys[i];
      }
      p[i].value = i;
    }
    if (desc) std::reverse(p.begin(), p.end());
    std::stable_sort(p.begin(), p.end());
    if (desc) std::reverse(p.begin(), p.end());
    if (desc == 0) {
      // header: name(16) n(4) key_bytes(4) begin(4) end(4)
      char nm[16] = {0};
      strncpy(nm, name, 15);
      fwrite(nm, 1, 16, stdout);
      int32_t h[4] = {n, (int32_t)sizeof(KeyT), begin_bit, end_bit};
      fwrite(h, 4, 4, stdout);
      fwrite(keys.data(), sizeof(KeyT), n, stdout);
    }
    std::vector<uint32_t> ranks(n);
    for (int i = 0; i < n; ++i) ranks[i] = p[i].value;
    fwrite(ranks.data(), 4, n, stdout);
  }
}

int main(int argc, char** argv) {
  CommandLineArgs args(argc, argv);  // seeds MT19937 exactly like every reference test binary
  // first raw draws of the stream (pins the generator itself)
  {
    unsigned int first[8];
    for (int i = 0; i < 8; ++i) first[i] = mersenne::genrand_int32();
    char nm[16] = "mt_first8";
    fwrite(nm, 1, 16, stdout);
    int32_t h[4] = {8, 4, 0, 32};
    fwrite(h, 4, 4, stdout);
    fwrite(first, 4, 8, stdout);
    unsigned int z[16] = {0};
    fwrite(z, 4, 16, stdout);  // two dummy rank arrays keep the record shape uniform
    unsigned int seed4[4] = {0x123, 0x234, 0x345, 0x456};
    mersenne::init_by_array(seed4, 4);  // rewind the stream
  }
  emit<unsigned int>("u32_rand", RANDOM, 5000, 0, 32);
  emit<unsigned int>("u32_bits_1_31", RANDOM, 3000, 1, 31);
  emit<unsigned int>("u32_bits_15_17", RANDOM, 3000, 15, 17);
  emit<int>("i32_rand", RANDOM, 3000, 0, 32);
  emit<float>("f32_pmzero", RANDOM_MINUS_PLUS_ZERO, 5000, 0, 32);
  emit<unsigned long long>("u64_rand", RANDOM, 3000, 0, 64);
  emit<unsigned long long>("u64_bits_1_63", RANDOM, 2000, 1, 63);
  emit<long long>("i64_rand", RANDOM, 2000, 0, 64);
  emit<double>("f64_pmzero", RANDOM_MINUS_PLUS_ZERO, 3000, 0, 64);
  emit<unsigned short>("u16_rand", RANDOM, 3000, 0, 16);
  emit<short>("i16_rand", RANDOM, 3000, 0, 16);
  emit<half_t>("f16_pmzero", RANDOM_MINUS_PLUS_ZERO, 3000, 0, 16);
  emit<bfloat16_t>("bf16_pmzero", RANDOM_MINUS_PLUS_ZERO, 3000, 0, 16);
  emit<unsigned char>("u8_rand", RANDOM, 2000, 0, 8);
  emit<signed char>("i8_rand", RANDOM, 2000, 0, 8);
  emit<unsigned int>("u32_uniform", UNIFORM, 1000, 0, 32);
  emit<unsigned int>("u32_iota", INTEGER_SEED, 1000, 0, 32);
  return 0;
}
