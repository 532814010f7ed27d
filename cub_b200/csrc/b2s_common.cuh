// b2s_common.cuh -- shared device helpers of the B200 radix sort (sm_100a only).
//
// Key bit-ordering arithmetic mirrors the reference semantics (NOT its code):
//   cub/util_type.cuh:1031,1078,1179        Traits<T>::TwiddleIn (unsigned / signed / floating)
//   cub/block/radix_rank_sort_operations.cuh:592-599  descending = complement of the bit-ordered key
//   cub/block/radix_rank_sort_operations.cuh:79-89    -0.0 collapses onto +0.0 for DIGIT extraction only
// Here all of that is folded into one per-pass "digit functor" driven by three runtime
// constants so that one kernel instantiation serves unsigned, signed, ascending and
// descending keys of a given width (floating keys use a second instantiation).
#pragma once
#include <type_traits>
#include <utility>
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef __CUDA_ARCH__
#define B2S_DEVICE_ARCH_OK 1
#elif __CUDA_ARCH__ >= 1000
#define B2S_DEVICE_ARCH_OK 1
#else
#error "b2s kernels are written for sm_100a only"
#endif

namespace b2s {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;

template <int BYTES> struct UIntOf;
template <> struct UIntOf<1> { using type = uint8_t; };
template <> struct UIntOf<2> { using type = uint16_t; };
template <> struct UIntOf<4> { using type = uint32_t; };
template <> struct UIntOf<8> { using type = unsigned long long; };
// 16-byte values are opaque and only guaranteed element-aligned by the caller; 8-byte alignment keeps
// every access legal for e.g. a struct of two doubles
struct alignas(8) U128 { unsigned long long lo, hi; };
template <> struct UIntOf<16> { using type = U128; };

// Register-width type a key is widened to for arithmetic (8/16-bit keys compute in 32 bits).
template <int BYTES> struct WideOf { using type = uint32_t; };
template <> struct WideOf<8> { using type = unsigned long long; };

// ---------------------------------------------------------------------------------------
// Digit functor.  digit(k) = (ordered(k) >> bit) & mask
//   integer keys : ordered(k) = k ^ xor_mask            (sign bit for signed, all ones more for descending)
//   floating keys: t = k ^ (sign(k) ? ~0 : HIGH) ^ xor_mask   (Traits<fp>::TwiddleIn; xor_mask = ONES if descending)
//                  ordered(k) = (t == zero_img) ? HIGH : t
//     -0.0 and +0.0 must share their digits (never their stored bits).  Ascending, t(+0.0) = HIGH and t(-0.0) = ~HIGH;
//     descending it is the other way round; the reference collapses -0.0 onto +0.0 ascending and +0.0 onto -0.0
//     descending (radix_rank_sort_operations.cuh:55-66, 79-89), i.e. in BOTH directions the image ~HIGH is replaced by
//     HIGH.  zero_img is that image as this functor computes it (for 16-bit keys held in 32-bit registers the bits
//     above the key are ones for negative keys; they never reach a digit).
// ---------------------------------------------------------------------------------------
template <int KBYTES, bool IS_FLOAT>
struct DigitOp {
  using W = typename WideOf<KBYTES>::type;
  W xor_mask;    // integer: HIGH (signed) ^ ONES (descending); float: ONES if descending else 0
  W zero_img;    // float only: see above
  uint32_t xor_digit;  // integer: (xor_mask >> bit) & mask, so that a digit is one shift + one 3-input logic op
  uint32_t bit;  // first bit of this pass' digit
  uint32_t mask; // (1 << digit_bits) - 1

  static constexpr int kMaxDigit = 255;  // largest value operator() can return
  static constexpr int KBITS = KBYTES * 8;
  static constexpr int WBITS = sizeof(W) * 8;
  static constexpr W ONES = KBYTES == 8 ? ~W(0) : (W)((1ull << (KBITS % 64)) - 1);
  static constexpr W HIGH = W(1) << (KBITS - 1);

  __device__ __forceinline__ void prepare() {}
  __host__ __device__ __forceinline__ W ordered(W k) const {
    if (IS_FLOAT) {
      // sign of the KBITS-wide key replicated over the register: 0 or ~0
      const W m = KBYTES == 8 ? (W)((long long)k >> 63) : (W)((int)((unsigned int)k << (WBITS - KBITS)) >> 31);
      const W t = k ^ (m | HIGH) ^ xor_mask;
      return t == zero_img ? HIGH : t;
    } else {
      return k ^ xor_mask;
    }
  }
  __device__ __forceinline__ uint32_t operator()(W k) const {
    if (IS_FLOAT) return (uint32_t)(ordered(k) >> bit) & mask;
    return ((uint32_t)(k >> bit) ^ xor_digit) & mask;  // == ((k ^ xor_mask) >> bit) & mask
  }
};

// Digit functor of the multi-pass sort for FLOATING keys.  Between the first and the last digit pass the keys live in
// the intermediate buffers as their bit-ordered image t = k ^ (sign(k) ? ~0 : HIGH) ^ xor_mask (a bijection: -0.0 and +0.0
// keep distinct images), so the sign-dependent transform is paid once per sort instead of three times per key per pass;
// a digit is then the collapse of the zero image (DigitOp above, same rule) + shift + mask.  The first pass converts on
// the way in (raw_in), the last one on the way out (raw_out); results are bit-identical to converting in every pass, which
// is what the reference does (Traits<fp>::TwiddleIn / TwiddleOut, cub/util_type.cuh:1078-1100).
template <int KBYTES>
struct OrderedFloatOp {
  using W = typename WideOf<KBYTES>::type;
  static constexpr bool kConverts = true;
  static constexpr int kMaxDigit = 255;
  static constexpr int KBITS = KBYTES * 8;
  static constexpr int WBITS = sizeof(W) * 8;
  static constexpr W ONES = KBYTES == 8 ? ~W(0) : (W)((1ull << (KBITS % 64)) - 1);
  static constexpr W HIGH = W(1) << (KBITS - 1);
  static constexpr W ZERO_IMG = ONES ^ HIGH;  // the image that shares its digits with HIGH (DigitOp: zero_img, without register garbage)
  W xor_mask;     // ONES if descending else 0
  uint32_t bit, mask;
  int raw_in, raw_out;  // keys arrive / leave in their raw encoding (first / last pass of a sort)

  __device__ __forceinline__ void prepare() {}
  // raw key (zero-extended in the register) -> image, bits above the key cleared
  __host__ __device__ __forceinline__ W to_image(W k) const {
    const W m = KBYTES == 8 ? (W)((long long)k >> 63) : (W)((int)((unsigned int)k << (WBITS - KBITS)) >> 31);
    return (k ^ (m | HIGH) ^ xor_mask) & ONES;
  }
  __host__ __device__ __forceinline__ W to_raw(W t) const {
    const W o = t ^ xor_mask;  // ascending image: top bit set <=> the key was not negative
    const W m = KBYTES == 8 ? (W)((long long)o >> 63) : (W)((int)((unsigned int)o << (WBITS - KBITS)) >> 31);
    return (o ^ (~m | HIGH)) & ONES;
  }
  // digit of an IMAGE
  __device__ __forceinline__ uint32_t operator()(W t) const {
    const W c = t == ZERO_IMG ? HIGH : t;
    return (uint32_t)(c >> bit) & mask;
  }
};

// Digit functor of full-range sorts of 4- and 8-byte FLOATING keys ("zero recording", b2s_fzero.cu).  The first pass maps BOTH
// zeros onto one image (HIGH -- the image the reference's collapse rule gives them in every digit, see DigitOp) and records, per
// row of 32 input keys, which keys were zeros and their signs; since the passes are stable the zeros then travel as equal keys in
// input order, every later pass extracts digits like an integer pass (the middle passes ARE the integer kernels), and after the
// last pass the run of zeros gets its recorded sign bits back.  Saves the collapse compare (2 of 4 instructions) in three digit
// extractions per key and pass, and the integer shapes apply.
template <int KBYTES>
struct ImageFloatOp {
  using W = typename WideOf<KBYTES>::type;
  static constexpr bool kConverts = true;
  static constexpr bool kRecordsZeros = true;
  static constexpr int kMaxDigit = 255;
  static constexpr int KBITS = KBYTES * 8;
  static constexpr int WBITS = sizeof(W) * 8;
  static constexpr W ONES = KBYTES == 8 ? ~W(0) : (W)((1ull << (KBITS % 64)) - 1);
  static constexpr W HIGH = W(1) << (KBITS - 1);
  W xor_mask;     // ONES if descending else 0
  uint32_t bit, mask;
  int raw_in, raw_out;

  __device__ __forceinline__ void prepare() {}
  __host__ __device__ __forceinline__ static bool is_zero(W k) { return (k & (ONES ^ HIGH)) == 0; }  // +0.0 or -0.0 (raw)
  __host__ __device__ __forceinline__ W to_image_nz(W k) const {  // image of a key that is not a zero
    const W m = KBYTES == 8 ? (W)((long long)k >> 63) : (W)((int)((unsigned int)k << (WBITS - KBITS)) >> 31);
    return (k ^ (m | HIGH) ^ xor_mask) & ONES;
  }
  __host__ __device__ __forceinline__ W to_image(W k) const { return is_zero(k) ? HIGH : to_image_nz(k); }
  __host__ __device__ __forceinline__ W to_raw(W t) const {
    const W o = t ^ xor_mask;
    const W m = KBYTES == 8 ? (W)((long long)o >> 63) : (W)((int)((unsigned int)o << (WBITS - KBITS)) >> 31);
    return (o ^ (~m | HIGH)) & ONES;
  }
  __device__ __forceinline__ uint32_t operator()(W t) const { return (uint32_t)(t >> bit) & mask; }
};

template <typename OpT, typename = void>
struct OpRecordsZeros { static constexpr bool value = false; };
template <typename OpT>
struct OpRecordsZeros<OpT, std::enable_if_t<OpT::kRecordsZeros>> { static constexpr bool value = true; };

template <typename OpT, typename = void>
struct OpConverts { static constexpr bool value = false; };
template <typename OpT>
struct OpConverts<OpT, std::enable_if_t<OpT::kConverts>> { static constexpr bool value = true; };

template <typename OpT, typename W>
__device__ __forceinline__ W image_to_raw(const OpT& op, W t) {
  if constexpr (OpConverts<OpT>::value) return op.to_raw(t);
  else return t;
}

// Destination functor of the multi-GPU partition pass: "digit" = number of splitters that order at or before
// this key, i.e. the rank the key is sent to.  Splitters are (key, source rank) pairs; a key equal to splitter
// j goes right of it iff the splitter was sampled on a rank <= this one (`tie` bit j), which spreads long runs
// of equal keys over several destinations without breaking stability (equal keys stay in rank order).
template <int KBYTES, bool IS_FLOAT>
struct SplitterOp {
  using W = typename WideOf<KBYTES>::type;
  using KeyU = typename UIntOf<KBYTES>::type;
  static constexpr int MAX_SPLITTERS = 7;
  static constexpr int kMaxDigit = MAX_SPLITTERS;  // destinations 0 .. count
  DigitOp<KBYTES, IS_FLOAT> base;  // .bit = begin_bit of the sort; .mask unused
  W range_mask;                    // ones over (end_bit - begin_bit) bits
  const void* d_keys;              // DEVICE: `count` raw splitter keys, ascending in sort order
  const int* d_ranks;              // DEVICE: source rank of every splitter
  int my_rank;
  int count;
  W s[MAX_SPLITTERS];              // filled by prepare(): THRESHOLDS: sort key of splitter j, + 1 when its tie bit is clear
  uint32_t tie;                    // filled by prepare(): bit j = splitter j's source rank <= this rank

  __device__ __forceinline__ W sort_key(W k) const { return (W)(base.ordered(k) >> base.bit) & range_mask; }
  // Splitters live in device memory (they come out of a device-side sort of the samples), so that the host never
  // has to wait for them; every thread reads the <= 7 of them once (L2 hits).
  // "key orders at or after splitter j"  <=>  o > s_j || (o == s_j && tie_j)  <=>  o >= s_j + (tie_j ? 0 : 1): one compare
  // against a precomputed threshold.  A splitter whose threshold would overflow (largest sort key, tie bit clear) can
  // never match, nor can any later one (splitters ascend in (key, rank) order): `count` is cut there.
  __device__ __forceinline__ void prepare() {
    tie = 0;
    int live = count;
#pragma unroll
    for (int j = 0; j < MAX_SPLITTERS; ++j) {
      s[j] = ~W(0);
      if (j < count) {
        const W sj = sort_key((W)reinterpret_cast<const KeyU*>(d_keys)[j]);
        const bool t = d_ranks[j] <= my_rank;
        tie |= (t ? 1u : 0u) << j;
        if (!t && sj == range_mask && j < live) live = j;  // threshold sj + 1 does not exist
        s[j] = t ? sj : sj + 1;
      }
    }
    count = live;
  }
  __device__ __forceinline__ uint32_t operator()(W k) const {
    const W o = sort_key(k);
    uint32_t d = 0;
#pragma unroll
    for (int j = 0; j < MAX_SPLITTERS; ++j)
      if (j < count) d += (o >= s[j]) ? 1u : 0u;
    return d;
  }
  // ge[j] += "k orders at or after splitter j" (the summands of operator()); monotone in j
  __device__ __forceinline__ void add_ge(W k, uint32_t (&ge)[MAX_SPLITTERS]) const {
    const W o = sort_key(k);
#pragma unroll
    for (int j = 0; j < MAX_SPLITTERS; ++j)
      if (j < count) ge[j] += (o >= s[j]) ? 1u : 0u;
  }
};

// Host-side description of a key type, resolved once per call.
struct KeyDesc {
  int bytes;       // 1,2,4,8
  int category;    // 0 unsigned, 1 signed, 2 floating
};

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "B2S_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra B2S_DONE_%=;\n"
      "bra B2S_WAIT_%=;\n"
      "B2S_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP); 16-byte aligned src/dst/size.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// L2 prefetch of a contiguous global window (16-byte aligned address and size); no shared memory, no completion.
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
// Orders earlier generic-proxy accesses to shared memory (LDS/STS) before later async-proxy ones (TMA writes).
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// gpu-scope relaxed loads / stores for the look-back status words (the word carries flag AND
// value, so no acquire/release pairing is needed; L1 is bypassed by the scope).
__device__ __forceinline__ uint32_t ld_status(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Same, at a compile-time byte offset from `p` (the offset goes into the instruction: no address arithmetic per load).
template <int OFF>
__device__ __forceinline__ uint32_t ld_status_at(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1+%2];" : "=r"(v) : "l"(p), "n"(OFF) : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ unsigned long long ld_status_at(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1+%2];" : "=l"(v) : "l"(p), "n"(OFF) : "memory");
  return v;
}
// win[j] = status word of the j-th predecessor row (rows are ROW_BYTES apart), j = 0 .. sizeof...(J) - 1
template <int ROW_BYTES, typename OffT, int... J>
__device__ __forceinline__ void load_status_window(const OffT* p, OffT* win, std::integer_sequence<int, J...>) {
  ((win[J] = ld_status_at<-J * ROW_BYTES>(p)), ...);
}
__device__ __forceinline__ void st_status(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// streaming (evict-first) global load: the value is used exactly once
template <typename T>
__device__ __forceinline__ T ld_stream(const T* p) {
  return __ldcs(p);
}
__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
__device__ __forceinline__ uint32_t lanemask_le() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
  return m;
}

// Peer mask of lanes holding the same BITS-bit digit.  Per digit bit: x_i = ballot(bit_i), complemented in the lanes
// whose bit is clear (the lanes that agree with me on bit i); the peer mask is the AND of the x_i, combined three at a
// time with 3-input LOP3s (depth 2-3 instead of a chain of 8).  SASS per row of 32 keys: 1-2 R2P/LOP3.P for the
// predicates, BITS x (VOTE + predicated LOP3) and BITS/2 LOP3 -- 21 instructions for 8 bits.
template <int I, bool NOT_ON_FMA>
__device__ __forceinline__ uint32_t agree_bit(uint32_t d, uint32_t ones) {
  uint32_t b;
  if (NOT_ON_FMA) {
    // ~b == b * -1 + -1: the complement as an IMAD, i.e. on the FMA pipe, which the ranking loop leaves idle, instead of one
    // more LOP3 on the half-rate ALU pipe that bounds it
    asm volatile("{\n"
        ".reg .pred p;\n"
        ".reg .b32 t;\n"
        "and.b32 t, %1, %2;\n"
        "setp.ne.u32 p, t, 0;\n"
        "vote.sync.ballot.b32 %0, p, 0xffffffff;\n"
        "@!p mad.lo.u32 %0, %0, %3, %3;\n"
        "}\n"
        : "=r"(b)
        : "r"(d), "n"(1u << I), "r"(ones));  // `ones` must not be a compile-time constant, or ptxas turns this into IADD3
  } else {
    asm volatile("{\n"
        ".reg .pred p;\n"
        ".reg .b32 t;\n"
        "and.b32 t, %1, %2;\n"
        "setp.ne.u32 p, t, 0;\n"
        "vote.sync.ballot.b32 %0, p, 0xffffffff;\n"
        "@!p not.b32 %0, %0;\n"
        "}\n"
        : "=r"(b)
        : "r"(d), "n"(1u << I));
  }
  return b;
}
__device__ __forceinline__ uint32_t and3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0x80;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
template <int BITS, bool NOT_ON_FMA = false>
__device__ __forceinline__ uint32_t match_ballot(uint32_t d, uint32_t ones = 0xffffffffu) {
  static_assert(BITS >= 1 && BITS <= 11, "digit width");
  uint32_t x[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) x[i] = 0xffffffffu;
  x[0] = agree_bit<0, NOT_ON_FMA>(d, ones);
  if (BITS > 1) x[1] = agree_bit<1, NOT_ON_FMA>(d, ones);
  if (BITS > 2) x[2] = agree_bit<2, NOT_ON_FMA>(d, ones);
  if (BITS > 3) x[3] = agree_bit<3, NOT_ON_FMA>(d, ones);
  if (BITS > 4) x[4] = agree_bit<4, NOT_ON_FMA>(d, ones);
  if (BITS > 5) x[5] = agree_bit<5, NOT_ON_FMA>(d, ones);
  if (BITS > 6) x[6] = agree_bit<6, NOT_ON_FMA>(d, ones);
  if (BITS > 7) x[7] = agree_bit<7, NOT_ON_FMA>(d, ones);
  if (BITS > 8) x[8] = agree_bit<8, NOT_ON_FMA>(d, ones);
  if (BITS > 9) x[9] = agree_bit<9, NOT_ON_FMA>(d, ones);
  if (BITS > 10) x[10] = agree_bit<10, NOT_ON_FMA>(d, ones);
  if (BITS <= 2) return x[0] & x[1];
  uint32_t m = and3(x[0], x[1], x[2]);
  if (BITS > 3) m = BITS > 4 ? and3(m, x[3], x[4]) : (m & x[3]);
  if (BITS > 5) m = BITS > 6 ? and3(m, x[5], x[6]) : (m & x[5]);
  if (BITS > 7) m = BITS > 8 ? and3(m, x[7], x[8]) : (m & x[7]);
  if (BITS > 9) m = BITS > 10 ? and3(m, x[9], x[10]) : (m & x[9]);
  return m;
}
// (The hardware MATCH.ANY instruction is not an option: on B200 its cost grows with the number of distinct
// values in the warp, ~45-60 cycles per warp instruction per SM for random 8-bit digits against ~7 for the
// eight ballot rounds above -- bench/micro/prim.cu, profiles/r1_prim_microbench.txt.)

// index of the most significant set bit (SASS FLO)
__device__ __forceinline__ uint32_t bfind(uint32_t x) {
  uint32_t r;
  asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}

// Optimisation barrier: the compiler may not assume anything about the value afterwards.  Used to make it RECOMPUTE
// a digit from its key (2 instructions) instead of keeping one more register per item alive across block barriers.
__device__ __forceinline__ uint32_t opaque(uint32_t x) {
  asm volatile("" : "+r"(x));
  return x;
}
__device__ __forceinline__ unsigned long long opaque(unsigned long long x) {
  asm volatile("" : "+l"(x));
  return x;
}

// Shared-memory add without a result (reduction form).
__device__ __forceinline__ void red_shared_add(uint32_t smem_addr, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(smem_addr), "r"(v) : "memory");
}

// Leader-only shared-memory fetch-add: returns the old value on the lanes where `pred` holds; the result is
// UNSPECIFIED on the other lanes (callers only ever read the leader's copy through a shuffle), which saves
// initialising it.
__device__ __forceinline__ uint32_t atoms_add_if(bool pred, uint32_t smem_addr, uint32_t v) {
  uint32_t old;
  asm volatile("{\n"
               ".reg .pred p;\n"
               "setp.ne.u32 p, %3, 0;\n"
               "@p atom.shared.add.u32 %0, [%1], %2;\n"
               "}\n"
               : "=r"(old)
               : "r"(smem_addr), "r"(v), "r"((uint32_t)pred)
               : "memory");
  return old;
}

__device__ __forceinline__ uint32_t atoms_add(uint32_t smem_addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_addr), "r"(v) : "memory");
  return old;
}

}  // namespace b2s
