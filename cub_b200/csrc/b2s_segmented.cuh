// b2s_segmented.cuh -- cub::DeviceSegmentedRadixSort: many independent sorts of contiguous segments of one array.
//
// Replaces (reference, for parity of RESULT only):
//   cub::DeviceSegmentedRadixSort::{SortKeys,SortPairs}[Descending], pointer + DoubleBuffer forms
//       cub/device/device_segmented_radix_sort.cuh
//   DeviceSegmentedRadixSortKernel (one CTA per segment per pass: upsweep / scan / downsweep)
//       cub/device/dispatch/dispatch_radix_sort.cuh:383-540, dispatch :2076-2420
//   tests  test/test_device_radix_sort.cu:293-470 (segmented back-ends), :1385-1455
//
// B200-first design: ONE launch for all passes, one CTA per segment (segment sizes live on the device, so nothing the
// host plans may depend on them):
//   * a segment of at most one tile (4096 items) is loaded once, sorted entirely in shared memory -- counting sweep, 256-wide
//     scan, ballot ranking fused with an in-place scatter per digit (the tile pass of the single-tile kernel) -- and stored
//     once: 2 x (K + V) bytes of HBM traffic per item for the whole sort instead of per pass, no temp storage touched;
//   * a larger segment is processed pass by pass by its CTA: digit histogram of the segment in shared memory, scan, then
//     tile after tile through the same tile pass with running per-digit output offsets (tiles are taken in order, so no
//     look-back is needed), ping-ponging between the output and the alternate buffer so that the last pass lands in the
//     output.  Like the reference's kernel this gives a huge segment one SM only; the device-wide sort is the tool for that.
// Stability: the tile pass ranks in tile order, tiles are taken in order.
#pragma once
#include "b2s_common.cuh"

namespace b2s {

constexpr int SEG_NT = 256;
constexpr int SEG_IPT = 16;

template <int KBYTES, int VBYTES, int NT, int IPT>
struct TileSmem {
  static constexpr int TILE = NT * IPT;
  static constexpr int NW = NT / 32;
  static constexpr int OFF_KEYS = 0;
  static constexpr int OFF_VALS = (TILE * KBYTES + 127) / 128 * 128;
  static constexpr int OFF_WHIST = OFF_VALS + (TILE * VBYTES + 127) / 128 * 128;
  static constexpr int OFF_WTOT = OFF_WHIST + NW * RADIX * 4;  // uint32[8]
  static constexpr int OFF_DSTART = OFF_WTOT + 64;             // uint32[256] digit starts inside the sorted tile
  static constexpr int OFF_BASE = OFF_DSTART + RADIX * 4;      // uint32[256] running output offsets of the segment
  static constexpr int TOTAL = OFF_BASE + RADIX * 4;
};

// One digit pass over a tile held in shared memory (sk / sv, TILE slots, padded with a key that orders last): stable
// partition by op's digit, in place.  Every thread of the CTA calls it.  s_dstart (may be null): receives the position of
// every digit's first item inside the sorted tile.
template <int KBYTES, int VBYTES, int NT, int IPT, typename OpT>
__device__ __forceinline__ void tile_digit_pass(typename UIntOf<KBYTES>::type* sk, typename UIntOf<VBYTES ? VBYTES : 1>::type* sv,
                                                unsigned int* whist, unsigned int* s_wtot, const OpT& op, unsigned int ones,
                                                unsigned int* s_dstart) {
  using KeyU = typename UIntOf<KBYTES>::type;
  using W = typename WideOf<KBYTES>::type;
  using ValU = typename UIntOf<VBYTES ? VBYTES : 1>::type;
  constexpr bool HAS_VALUES = VBYTES != 0;
  constexpr int NW = NT / 32;
  static_assert(NT >= RADIX, "one thread per digit needed");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_base = warp * 32 * IPT;
  unsigned int* myhist = whist + warp * RADIX;
  const unsigned int myhist_s = smem_u32(myhist);
  const unsigned int lt = lanemask_lt();

  // ---- items -> registers; counting sweep on this warp's counters
  W key[IPT];
  ValU val[HAS_VALUES ? IPT : 1];
#pragma unroll
  for (int u = 0; u < IPT; ++u) key[u] = (W)sk[warp_base + u * 32 + lane];
  if (HAS_VALUES) {
#pragma unroll
    for (int u = 0; u < IPT; ++u) val[u] = sv[warp_base + u * 32 + lane];
  }
#pragma unroll
  for (int i = lane; i < RADIX; i += 32) myhist[i] = 0;
  __syncwarp();
#pragma unroll
  for (int u = 0; u < IPT; ++u) red_shared_add(myhist_s + op(key[u]) * 4, 1u);
  __syncthreads();  // every count is in; every item of the tile is in a register (the scatter below is in place)

  // ---- 256-wide exclusive scan of the tile's digit counts; counters become absolute positions
  unsigned int total = 0;
  if (tid < RADIX) {
#pragma unroll
    for (int w = 0; w < NW; ++w) total += whist[w * RADIX + tid];
  }
  unsigned int incl = total;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (tid < RADIX && lane == 31) s_wtot[warp] = incl;
  __syncthreads();
  if (tid < RADIX) {
    unsigned int run = incl - total;
#pragma unroll
    for (int w = 0; w < RADIX / 32; ++w)
      if (w < warp) run += s_wtot[w];
    if (s_dstart) s_dstart[tid] = run;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const unsigned int c = whist[w * RADIX + tid];
      whist[w * RADIX + tid] = run;
      run += c;
    }
  }
  __syncthreads();

  // ---- ranking sweep fused with the scatter (same software pipeline as the digit-pass kernel)
  unsigned int d = op(key[0]);
  unsigned int m = match_ballot<RADIX_BITS, true>(d, ones);
  unsigned int bcast_prev = 0, below_prev = 0;
#pragma unroll
  for (int u = 0; u < IPT; ++u) {
    const unsigned int leader = bfind(m);
    const unsigned int below = __popc(m & lt);
    const unsigned int raw = atoms_add_if(lane == leader, myhist_s + d * 4, (unsigned int)__popc(m));
    unsigned int d_next = 0, m_next = 0;
    if (u + 1 < IPT) {
      d_next = op(key[u + 1]);
      m_next = match_ballot<RADIX_BITS, true>(d_next, ones);
    }
    if (u > 0) {
      const unsigned int r = bcast_prev + below_prev;
      sk[r] = (KeyU)key[u - 1];
      if (HAS_VALUES) sv[r] = val[u - 1];
    }
    bcast_prev = __shfl_sync(0xffffffffu, raw, leader);
    below_prev = below;
    d = d_next;
    m = m_next;
  }
  {
    const unsigned int r = bcast_prev + below_prev;
    sk[r] = (KeyU)key[IPT - 1];
    if (HAS_VALUES) sv[r] = val[IPT - 1];
  }
  __syncthreads();
}

template <int KBYTES, bool F>
struct SegmentedParams {
  const void* keys_src;  // pass 0 reads here (never written in the pointer form)
  void* keys_a;          // the last pass lands here
  void* keys_b;          // the other ping-pong buffer (may be null when there is a single pass)
  const void* vals_src;
  void* vals_a;
  void* vals_b;
  const void* begin_offsets;
  const void* end_offsets;
  unsigned long long pad_key;
  DigitOp<KBYTES, F> op;  // xor_mask / zero_img set by the host; bit, mask, xor_digit per pass here
  int begin_bit, end_bit, passes;
  unsigned int ones;
};

template <int KBYTES, int VBYTES, bool F, typename SegOffT>
__global__ void __launch_bounds__(SEG_NT, 2) segmented_sort_kernel(const SegmentedParams<KBYTES, F> P) {
  using KeyU = typename UIntOf<KBYTES>::type;
  using W = typename WideOf<KBYTES>::type;
  using ValU = typename UIntOf<VBYTES ? VBYTES : 1>::type;
  using S = TileSmem<KBYTES, VBYTES, SEG_NT, SEG_IPT>;
  constexpr int NT = SEG_NT, TILE = S::TILE;
  constexpr bool HAS_VALUES = VBYTES != 0;

  extern __shared__ __align__(128) unsigned char smem[];
  KeyU* sk = reinterpret_cast<KeyU*>(smem + S::OFF_KEYS);
  ValU* sv = reinterpret_cast<ValU*>(smem + S::OFF_VALS);
  unsigned int* whist = reinterpret_cast<unsigned int*>(smem + S::OFF_WHIST);
  unsigned int* s_wtot = reinterpret_cast<unsigned int*>(smem + S::OFF_WTOT);
  unsigned int* s_dstart = reinterpret_cast<unsigned int*>(smem + S::OFF_DSTART);
  unsigned int* s_base = reinterpret_cast<unsigned int*>(smem + S::OFF_BASE);

  const int tid = threadIdx.x;
  const unsigned long long seg = blockIdx.x;
  const long long begin = (long long)reinterpret_cast<const SegOffT*>(P.begin_offsets)[seg];
  const long long end = (long long)reinterpret_cast<const SegOffT*>(P.end_offsets)[seg];
  if (end <= begin) return;
  const unsigned long long len = (unsigned long long)(end - begin);
  auto op = P.op;
  auto set_pass = [&](int p) {
    const int bit = P.begin_bit + RADIX_BITS * p;
    int nbits = P.end_bit - bit;
    nbits = nbits < 0 ? 0 : (nbits > RADIX_BITS ? RADIX_BITS : nbits);
    op.bit = (uint32_t)bit;
    op.mask = (1u << nbits) - 1u;
    op.xor_digit = (uint32_t)(op.xor_mask >> bit) & op.mask;
  };
  auto load_tile = [&](const KeyU* gk, const ValU* gv, int valid, int slots) {
#pragma unroll 4
    for (int i = tid; i < slots; i += NT) sk[i] = i < valid ? gk[i] : (KeyU)P.pad_key;
    if (HAS_VALUES) {
#pragma unroll 4
      for (int i = tid; i < valid; i += NT) sv[i] = gv[i];
    }
    __syncthreads();
  };

  if (len <= (unsigned long long)TILE) {
    // ---- the whole segment is one tile: every pass in shared memory
    // The tile pass is instantiated for 1, 4 and 16 items per thread: a segment of <= 256 / <= 1024 items is padded to and
    // worked on as a 256- / 1024-slot tile instead of the full 4096 (a 64-item segment: 8 rows of 32 instead of 128).
    const int n = (int)len;
    const int slots = n <= NT ? NT : (n <= NT * 4 ? NT * 4 : TILE);
    load_tile(reinterpret_cast<const KeyU*>(P.keys_src) + begin, reinterpret_cast<const ValU*>(P.vals_src) + begin, n, slots);
    for (int p = 0; p < P.passes; ++p) {
      set_pass(p);
      if (slots == NT) tile_digit_pass<KBYTES, VBYTES, NT, 1>(sk, sv, whist, s_wtot, op, P.ones, nullptr);
      else if (slots == NT * 4) tile_digit_pass<KBYTES, VBYTES, NT, 4>(sk, sv, whist, s_wtot, op, P.ones, nullptr);
      else tile_digit_pass<KBYTES, VBYTES, NT, SEG_IPT>(sk, sv, whist, s_wtot, op, P.ones, nullptr);
    }
    KeyU* gk = reinterpret_cast<KeyU*>(P.keys_a) + begin;
    for (int i = tid; i < n; i += NT) gk[i] = sk[i];
    if (HAS_VALUES) {
      ValU* gv = reinterpret_cast<ValU*>(P.vals_a) + begin;
      for (int i = tid; i < n; i += NT) gv[i] = sv[i];
    }
    return;
  }

  // ---- a segment of several tiles: pass by pass through global memory, this CTA alone
  const KeyU* src_k = reinterpret_cast<const KeyU*>(P.keys_src) + begin;
  const ValU* src_v = reinterpret_cast<const ValU*>(P.vals_src) + begin;
  for (int p = 0; p < P.passes; ++p) {
    set_pass(p);
    const bool to_a = ((P.passes - 1 - p) & 1) == 0;
    KeyU* dst_k = reinterpret_cast<KeyU*>(to_a ? P.keys_a : P.keys_b) + begin;
    ValU* dst_v = reinterpret_cast<ValU*>(to_a ? P.vals_a : P.vals_b) + begin;
    // digit histogram of the segment -> exclusive offsets
    s_base[tid] = 0;  // NT == RADIX threads
    __syncthreads();
    for (unsigned long long i = tid; i < len; i += NT) atomicAdd(&s_base[op((W)src_k[i])], 1u);
    __syncthreads();
    {
      const unsigned int c = s_base[tid];
      unsigned int incl = c;
      const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) s_wtot[warp] = incl;
      __syncthreads();
      unsigned int run = incl - c;
#pragma unroll
      for (int w = 0; w < RADIX / 32; ++w)
        if (w < warp) run += s_wtot[w];
      s_base[tid] = run;
    }
    __syncthreads();
    for (unsigned long long t0 = 0; t0 < len; t0 += TILE) {
      const int valid = (len - t0) < (unsigned long long)TILE ? (int)(len - t0) : TILE;
      load_tile(src_k + t0, src_v + t0, valid, TILE);
      tile_digit_pass<KBYTES, VBYTES, NT, SEG_IPT>(sk, sv, whist, s_wtot, op, P.ones, s_dstart);
      for (int pos = tid; pos < valid; pos += NT) {
        const KeyU k = sk[pos];
        const unsigned int d = op((W)k);
        const unsigned int o = s_base[d] + ((unsigned int)pos - s_dstart[d]);
        dst_k[o] = k;
        if (HAS_VALUES) dst_v[o] = sv[pos];
      }
      __syncthreads();
      {  // advance the running offsets by this tile's digit counts (padding sits behind position `valid`: not counted)
        const unsigned int s0 = s_dstart[tid] < (unsigned int)valid ? s_dstart[tid] : (unsigned int)valid;
        unsigned int s1 = tid + 1 < RADIX ? s_dstart[tid + 1] : (unsigned int)TILE;
        if (s1 > (unsigned int)valid) s1 = (unsigned int)valid;
        __syncthreads();
        s_base[tid] += s1 - s0;
      }
      __syncthreads();
    }
    src_k = dst_k;
    src_v = dst_v;
  }
}

}  // namespace b2s
