/*
 * oracle/radix_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement (plain C) of the semantics of the reference cub::DeviceRadixSort
 * onesweep path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product path (cub_b200/) never does.
 *
 * Parity pin: checked in tests/test_oracle.py against
 *   - the reference test harness' host std::stable_sort solution (restated in
 *     oracle/host_stable_sort.cpp from test/test_device_radix_sort.cu:896-956) on the
 *     reference's own MT19937-seeded generator stream (tests/golden/, generated from the
 *     reference's test/mersenne.h by tests/golden/make_golden.py),
 *   - the +-0.0 known-answer vectors of test/catch2_test_device_radix_sort_custom.cu:593-700,
 *   - and, on the GPU box, against the unmodified reference itself (oracle/_ref/libref_cub.so).
 *
 * What is restated (reference file:line):
 *   - Traits<T>::TwiddleIn for unsigned / signed / floating keys     cub/util_type.cuh:1031, 1078, 1179
 *   - descending = bitwise complement of the bit-ordered key          cub/block/radix_rank_sort_operations.cuh:592-599
 *   - -0.0 collapse for DIGIT EXTRACTION ONLY (stored bits keep -0.0) cub/block/radix_rank_sort_operations.cuh:79-89
 *   - digit = (collapsed >> bit_start) & mask                         cub/block/radix_rank_sort_operations.cuh:118-135
 *   - LSD pass structure: RADIX_BITS = 8, num_passes = ceil(bits/8), last pass narrower;
 *     per pass: histogram -> exclusive sum -> stable scatter          cub/device/dispatch/dispatch_radix_sort.cuh:1533-1537, 1653-1704
 *     (agent_radix_sort_histogram.cuh:202-241, agent_radix_sort_onesweep.cuh:616-647)
 *   - begin_bit == end_bit: output = input (copy)                     cub/device/dispatch/dispatch_radix_sort.cuh:1955-1963
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { K_U8, K_I8, K_U16, K_I16, K_F16, K_BF16, K_U32, K_I32, K_F32, K_U64, K_I64, K_F64, K_COUNT };

static const int k_bytes[K_COUNT] = {1, 1, 2, 2, 2, 2, 4, 4, 4, 8, 8, 8};
/* 0 unsigned, 1 signed, 2 floating */
static const int k_cat[K_COUNT] = {0, 1, 0, 1, 2, 2, 0, 1, 2, 0, 1, 2};

int oracle_key_bytes(int key_type) { return (key_type >= 0 && key_type < K_COUNT) ? k_bytes[key_type] : 0; }

static uint64_t all_ones(int bytes) { return bytes == 8 ? ~0ull : ((1ull << (8 * bytes)) - 1); }

/* util_type.cuh:1031 (unsigned), :1078 (signed), :1179-1183 (floating); then
 * radix_rank_sort_operations.cuh:592-599 (descending inverts every bit). */
uint64_t oracle_twiddle_in(uint64_t bits, int key_type, int descending) {
  const int nb = k_bytes[key_type];
  const uint64_t ones = all_ones(nb);
  const uint64_t high = 1ull << (8 * nb - 1);
  uint64_t k = bits & ones;
  switch (k_cat[key_type]) {
    case 1: k ^= high; break;
    case 2: k ^= (k & high) ? ones : high; break;
    default: break;
  }
  if (descending) k = (~k) & ones;
  return k;
}

/* util_type.cuh:1036 / :1083 / :1185-1189 and the inverse inversion. */
uint64_t oracle_twiddle_out(uint64_t ordered, int key_type, int descending) {
  const int nb = k_bytes[key_type];
  const uint64_t ones = all_ones(nb);
  const uint64_t high = 1ull << (8 * nb - 1);
  uint64_t k = ordered & ones;
  if (descending) k = (~k) & ones;
  switch (k_cat[key_type]) {
    case 1: k ^= high; break;
    case 2: k ^= (k & high) ? high : ones; break;
    default: break;
  }
  return k;
}

/* radix_rank_sort_operations.cuh:79-89: a bit-ordered key equal to TwiddleIn(-0.0) is
 * replaced by TwiddleIn(+0.0) before the digit is taken.  Applied AFTER the descending
 * inversion, exactly as the onesweep agent does (the inverted +0.0 has that pattern then). */
uint64_t oracle_digit_source(uint64_t ordered, int key_type) {
  if (k_cat[key_type] != 2) return ordered;
  const int nb = k_bytes[key_type];
  const uint64_t high = 1ull << (8 * nb - 1);
  const uint64_t tw_minus_zero = high - 1; /* TwiddleIn(HIGH_BIT) = HIGH ^ ones = 0x7f..f */
  const uint64_t tw_zero = high;           /* TwiddleIn(0)        = 0x80..0              */
  return ordered == tw_minus_zero ? tw_zero : ordered;
}

/* The value the sort is stable with respect to: bits [begin_bit, end_bit) of the
 * collapsed bit-ordered key. */
uint64_t oracle_sort_key(uint64_t bits, int key_type, int descending, int begin_bit, int end_bit) {
  int nbits = end_bit - begin_bit;
  if (nbits <= 0) return 0;
  uint64_t s = oracle_digit_source(oracle_twiddle_in(bits, key_type, descending), key_type);
  s >>= begin_bit;
  if (nbits < 64) s &= (1ull << nbits) - 1;
  return s;
}

static uint64_t load_key(const uint8_t *p, int nb) {
  uint64_t v = 0;
  memcpy(&v, p, (size_t)nb); /* little endian host */
  return v;
}

/*
 * All digit histograms in one read of the keys, as the upfront histogram kernel does
 * (agent_radix_sort_histogram.cuh:202-241).  hist is [num_passes][256] uint64.
 */
int oracle_histogram(const void *keys, uint64_t n, int key_type, int descending, int begin_bit, int end_bit,
                     uint64_t *hist) {
  if (key_type < 0 || key_type >= K_COUNT) return 1;
  const int nb = k_bytes[key_type];
  const int passes = (end_bit - begin_bit + 7) / 8;
  memset(hist, 0, sizeof(uint64_t) * 256 * (size_t)(passes > 0 ? passes : 0));
  const uint8_t *kp = (const uint8_t *)keys;
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t s = oracle_digit_source(oracle_twiddle_in(load_key(kp + i * nb, nb), key_type, descending), key_type);
    for (int p = 0; p < passes; ++p) {
      int bit = begin_bit + 8 * p;
      int w = end_bit - bit < 8 ? end_bit - bit : 8;
      hist[p * 256 + ((s >> bit) & ((1u << w) - 1))]++;
    }
  }
  return 0;
}

/*
 * Stable LSD radix sort, one 8-bit digit per pass (narrower last pass), restating
 * DispatchRadixSort::InvokeOnesweep's pass loop.  keys_in/vals_in are never written.
 * value_bytes may be 0 (keys only).  Returns 0 on success.
 */
int oracle_radix_sort(const void *keys_in, void *keys_out, const void *vals_in, void *vals_out, uint64_t n,
                      int key_type, int value_bytes, int descending, int begin_bit, int end_bit) {
  if (key_type < 0 || key_type >= K_COUNT) return 1;
  const int nb = k_bytes[key_type];
  const size_t vb = (size_t)value_bytes;
  if (n == 0) return 0;
  const int passes = (end_bit - begin_bit + 7) / 8;
  if (passes <= 0) { /* dispatch_radix_sort.cuh:1955-1963 (InvokeCopy) */
    memmove(keys_out, keys_in, n * (size_t)nb);
    if (vb) memmove(vals_out, vals_in, n * vb);
    return 0;
  }
  uint8_t *kbuf[2], *vbuf[2] = {0, 0};
  kbuf[0] = (uint8_t *)malloc(n * (size_t)nb);
  kbuf[1] = (uint8_t *)malloc(n * (size_t)nb);
  if (vb) {
    vbuf[0] = (uint8_t *)malloc(n * vb);
    vbuf[1] = (uint8_t *)malloc(n * vb);
  }
  uint64_t *offs = (uint64_t *)malloc(sizeof(uint64_t) * 256);
  if (!kbuf[0] || !kbuf[1] || (vb && (!vbuf[0] || !vbuf[1])) || !offs) return 2;

  const uint8_t *ksrc = (const uint8_t *)keys_in;
  const uint8_t *vsrc = (const uint8_t *)vals_in;
  int cur = 0;
  for (int p = 0; p < passes; ++p) {
    const int bit = begin_bit + 8 * p;
    const int w = end_bit - bit < 8 ? end_bit - bit : 8;
    const uint32_t mask = (1u << w) - 1;
    uint8_t *kdst = kbuf[cur], *vdst = vbuf[cur];
    /* histogram (agent_radix_sort_histogram.cuh) */
    uint64_t cnt[256];
    memset(cnt, 0, sizeof cnt);
    for (uint64_t i = 0; i < n; ++i) {
      uint64_t s = oracle_digit_source(oracle_twiddle_in(load_key(ksrc + i * nb, nb), key_type, descending), key_type);
      cnt[(s >> bit) & mask]++;
    }
    /* exclusive sum (dispatch_radix_sort.cuh:603-635) */
    uint64_t run = 0;
    for (int d = 0; d < 256; ++d) { offs[d] = run; run += cnt[d]; }
    /* stable scatter: position = bins_in[d] + #(earlier items with digit d)
     * (agent_radix_sort_onesweep.cuh:476-492, 593-610) -- keys stay in their original encoding */
    for (uint64_t i = 0; i < n; ++i) {
      uint64_t s = oracle_digit_source(oracle_twiddle_in(load_key(ksrc + i * nb, nb), key_type, descending), key_type);
      uint64_t dst = offs[(s >> bit) & mask]++;
      memcpy(kdst + dst * nb, ksrc + i * nb, (size_t)nb);
      if (vb) memcpy(vdst + dst * vb, vsrc + i * vb, vb);
    }
    ksrc = kdst;
    vsrc = vdst;
    cur ^= 1;
  }
  memcpy(keys_out, ksrc, n * (size_t)nb);
  if (vb) memcpy(vals_out, vsrc, n * vb);
  free(kbuf[0]); free(kbuf[1]); free(vbuf[0]); free(vbuf[1]); free(offs);
  return 0;
}
