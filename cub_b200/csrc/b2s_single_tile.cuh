// b2s_single_tile.cuh -- a whole sort of at most one tile in ONE launch of ONE CTA: every digit pass runs in shared
// memory, nothing touches global memory between the load and the store, no histogram kernel, no look-back, no memset.
//
// Replaces (reference, for parity of RESULT only):
//   DeviceRadixSortSingleTileKernel   cub/device/dispatch/dispatch_radix_sort.cuh:258-364
//   BlockRadixSort::SortBlockedToStriped  cub/block/block_radix_sort.cuh:431-490
// The reference takes this route for n <= 4864 (256 threads x 19 items); small sorts are latency-bound (launch count), so
// the cut-over here is the tile this kernel can hold (8192 items of up to 16 bytes, 4096 beyond).
//
// Per pass: keys (and values) -> registers in warp-striped rows; counting sweep on warp-private counters; 256-wide scan
// turns them into absolute positions; the ranking sweep of the digit-pass kernel (ballot match, leader atomic, fused
// scatter -- b2s_onesweep.cuh, EARLY flow) puts every item into its slot, in place.  Stable for the same reason.
#pragma once
#include "b2s_common.cuh"

namespace b2s {

template <int KBYTES, bool F>
struct SingleTileParams {
  const void* keys_in;
  void* keys_out;
  const void* vals_in;
  void* vals_out;
  unsigned int n;
  unsigned long long pad_key;  // raw key that orders last in every pass
  DigitOp<KBYTES, F> op;       // xor_mask / zero_img set by the host; bit, mask, xor_digit are set per pass here
  int begin_bit, end_bit;
  unsigned int ones;           // 0xffffffff as a launch parameter (see agree_bit)
};

template <int KBYTES, int VBYTES>
struct SingleTileShape {
  static constexpr int NT = 512;
  static constexpr int IPT = (KBYTES + VBYTES <= 16) ? 16 : 8;
  static constexpr int TILE = NT * IPT;
  static constexpr int NW = NT / 32;
  static constexpr int OFF_KEYS = 0;
  static constexpr int OFF_VALS = (TILE * KBYTES + 127) / 128 * 128;
  static constexpr int OFF_WHIST = OFF_VALS + (TILE * VBYTES + 127) / 128 * 128;
  static constexpr int OFF_MISC = OFF_WHIST + NW * RADIX * 4;
  static constexpr int TOTAL = OFF_MISC + 64;
};

template <int KBYTES, int VBYTES, bool F>
__global__ void __launch_bounds__(SingleTileShape<KBYTES, VBYTES>::NT, 1)
single_tile_kernel(const SingleTileParams<KBYTES, F> P) {
  using KeyU = typename UIntOf<KBYTES>::type;
  using W = typename WideOf<KBYTES>::type;
  using ValU = typename UIntOf<VBYTES ? VBYTES : 1>::type;
  using S = SingleTileShape<KBYTES, VBYTES>;
  constexpr int NT = S::NT, IPT = S::IPT, TILE = S::TILE, NW = S::NW;
  constexpr bool HAS_VALUES = VBYTES != 0;
  static_assert(NT >= RADIX, "one thread per digit needed");

  extern __shared__ __align__(128) unsigned char smem[];
  KeyU* sk = reinterpret_cast<KeyU*>(smem + S::OFF_KEYS);
  ValU* sv = reinterpret_cast<ValU*>(smem + S::OFF_VALS);
  unsigned int* whist = reinterpret_cast<unsigned int*>(smem + S::OFF_WHIST);
  unsigned int* s_wtot = reinterpret_cast<unsigned int*>(smem + S::OFF_MISC);  // [8]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = (int)P.n;

  {  // tile -> shared memory, padded with a key that orders last in every pass (never written out)
    const KeyU* gk = reinterpret_cast<const KeyU*>(P.keys_in);
    const ValU* gv = reinterpret_cast<const ValU*>(P.vals_in);
#pragma unroll 4
    for (int i = tid; i < TILE; i += NT) sk[i] = i < n ? gk[i] : (KeyU)P.pad_key;
    if (HAS_VALUES) {
#pragma unroll 4
      for (int i = tid; i < n; i += NT) sv[i] = gv[i];
    }
  }
  __syncthreads();

  const int warp_base = warp * 32 * IPT;
  unsigned int* myhist = whist + warp * RADIX;
  const unsigned int myhist_s = smem_u32(myhist);
  const unsigned int lt = lanemask_lt();
  auto op = P.op;

  for (int bit = P.begin_bit; bit < P.end_bit; bit += RADIX_BITS) {
    const int nbits = P.end_bit - bit < RADIX_BITS ? P.end_bit - bit : RADIX_BITS;
    op.bit = (uint32_t)bit;
    op.mask = (1u << nbits) - 1u;
    op.xor_digit = (uint32_t)(op.xor_mask >> bit) & op.mask;

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
    unsigned int m = match_ballot<RADIX_BITS, true>(d, P.ones);
    unsigned int bcast_prev = 0, below_prev = 0;
#pragma unroll
    for (int u = 0; u < IPT; ++u) {
      const unsigned int leader = bfind(m);
      const unsigned int below = __popc(m & lt);
      const unsigned int raw = atoms_add_if(lane == leader, myhist_s + d * 4, (unsigned int)__popc(m));
      unsigned int d_next = 0, m_next = 0;
      if (u + 1 < IPT) {
        d_next = op(key[u + 1]);
        m_next = match_ballot<RADIX_BITS, true>(d_next, P.ones);
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

  {  // sorted tile -> global memory (the padding sorted to the end)
    KeyU* gk = reinterpret_cast<KeyU*>(P.keys_out);
    ValU* gv = reinterpret_cast<ValU*>(P.vals_out);
    for (int i = tid; i < n; i += NT) gk[i] = sk[i];
    if (HAS_VALUES)
      for (int i = tid; i < n; i += NT) gv[i] = sv[i];
  }
}

}  // namespace b2s
