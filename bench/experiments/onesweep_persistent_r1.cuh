// b2s_onesweep2.cuh -- persistent, software-pipelined digit pass (the production kernel).
//
// Same contract as b2s_onesweep.cuh (stable partition of all n items by one <=8-bit digit into
// their global positions, chained-scan style), restructured after the round-1 ncu profiles
// (profiles/r1_onesweep_v1.md): with one TMA-staged tile per CTA, 10 % of the warp samples sat in
// the TMA wait for the tile's keys, ~20 % behind the serial look-back walk, ~4 % in CTA set-up,
// and the in-place value reorder needed an extra block barrier.
//
// Replaces (reference, for parity of RESULT only):
//   DeviceRadixSortOnesweepKernel  cub/device/dispatch/dispatch_radix_sort.cuh:580
//   AgentRadixSortOnesweep         cub/agent/agent_radix_sort_onesweep.cuh:98-688
//   BlockRadixRankMatchEarlyCounts cub/block/block_radix_rank.cuh:898-1192
//
// Structure:
//   * persistent CTAs (MINB per SM) claim tiles from a global counter in input order; a CTA
//     claims its NEXT tile while it still works on the current one and prefetches that tile's
//     keys straight into registers (coalesced warp-striped loads, in flight during the
//     write-out), so ranking of the next tile starts without a load bubble;
//   * values are loaded into registers right after the ranking sweep (in flight during the
//     digit scan and the look-back) and are scattered together with the keys: shared memory
//     holds only the reorder buffers, 4 block barriers per tile;
//   * the look-back starts IMMEDIATELY after the partial counts are published and reads a window
//     of LBW predecessor tiles per round trip (independent loads).  The time between a tile's
//     partial and inclusive publication is what makes successors walk further back; with the
//     window wider than (L2 round trip / time between consecutive tiles) the walk is one round trip;
//   * single ranking sweep (warp-private digit counters produced BY the match ranking), keys
//     and values reordered through shared memory, digit runs written out coalesced.
//
// Forward progress: tile ids are claimed in increasing order by running CTAs only, and a CTA
// holds at most one claimed-but-not-started tile, whose id is larger than the id of the tile
// it is working on; hence the smallest unfinished tile is always being worked on and never
// waits on anything.
//
// Stability: items are ranked in tile order (warp-striped rows, lane order inside a row,
// rows in program order); tiles are ordered by tile id == position in the input.
#pragma once
#include "b2s_common.cuh"
#include "b2s_onesweep.cuh"

namespace b2s {

template <int KBYTES, int VBYTES, int NT, int IPT>
struct Onesweep2Smem {
  static constexpr int TILE = NT * IPT;
  static constexpr int NW = NT / 32;
  static constexpr int KEY_BYTES = TILE * KBYTES;
  static constexpr int VAL_BYTES = TILE * VBYTES;
  static constexpr int OFF_KEYS = 0;
  static constexpr int OFF_VALS = (KEY_BYTES + 127) / 128 * 128;
  static constexpr int OFF_WHIST = OFF_VALS + (VAL_BYTES + 127) / 128 * 128;
  static constexpr int OFF_GOFF = OFF_WHIST + NW * RADIX * 4;
  static constexpr int OFF_MISC = OFF_GOFF + RADIX * 8;
  static constexpr int TOTAL = OFF_MISC + 128;
};

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// streaming global load (read once, do not keep in L1)
template <typename T>
__device__ __forceinline__ T ld_stream(const T* p) {
  return __ldcs(p);
}

template <int KBYTES, int VBYTES, bool IS_FLOAT, typename OffT, int NT, int IPT, int MINB, int LBW>
__global__ void __launch_bounds__(NT, MINB) onesweep2_kernel(const OnesweepParams<KBYTES, IS_FLOAT> P) {
  using KeyU = typename UIntOf<KBYTES>::type;
  using W = typename WideOf<KBYTES>::type;
  using ValU = typename UIntOf<VBYTES ? VBYTES : 1>::type;
  using L = Onesweep2Smem<KBYTES, VBYTES, NT, IPT>;
  constexpr int TILE = L::TILE;
  constexpr int NW = L::NW;
  constexpr bool HAS_VALUES = VBYTES != 0;
  constexpr int OBITS = sizeof(OffT) * 8;
  constexpr OffT FLAG_INCLUSIVE = OffT(1) << (OBITS - 1);
  constexpr OffT FLAG_PARTIAL = OffT(1) << (OBITS - 2);
  constexpr OffT FLAG_ANY = FLAG_INCLUSIVE | FLAG_PARTIAL;
  constexpr OffT VALUE_MASK = FLAG_PARTIAL - 1;
  static_assert(NT >= RADIX && NT % 32 == 0, "one thread per digit needed");
  static_assert(NT * IPT <= 65536, "tile positions are packed into 16 bits");

  extern __shared__ __align__(128) unsigned char smem[];
  KeyU* sort_k = reinterpret_cast<KeyU*>(smem + L::OFF_KEYS);
  ValU* sort_v = reinterpret_cast<ValU*>(smem + L::OFF_VALS);
  unsigned int* whist = reinterpret_cast<unsigned int*>(smem + L::OFF_WHIST);
  OffT* s_goff = reinterpret_cast<OffT*>(smem + L::OFF_GOFF);
  unsigned int* s_wtot = reinterpret_cast<unsigned int*>(smem + L::OFF_MISC + 16);  // [8]
  unsigned int* s_tile = reinterpret_cast<unsigned int*>(smem + L::OFF_MISC + 64);  // [2]: first, next

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const unsigned int warp_base = warp * 32 * IPT + lane;
  const unsigned long long n = P.n;
  const unsigned int num_tiles = (unsigned int)((n + TILE - 1) / TILE);
  const auto op = P.op;
  unsigned int* myhist = whist + warp * RADIX;
  const unsigned int lt = lanemask_lt();

  if (tid == 0) s_tile[0] = atomicAdd(P.tile_counter, 1u);
#pragma unroll
  for (int i = lane; i < RADIX; i += 32) myhist[i] = 0;
  __syncthreads();
  unsigned int tile = s_tile[0];
  if (tile >= num_tiles) return;

  W key[IPT];
  auto load_keys = [&](unsigned int t) {
    const unsigned long long base = (unsigned long long)t * TILE + warp_base;
    const KeyU* g = reinterpret_cast<const KeyU*>(P.keys_in) + base;
    if (base + (unsigned long long)(IPT - 1) * 32 < n) {
#pragma unroll
      for (int u = 0; u < IPT; ++u) key[u] = (W)ld_stream(g + u * 32);
    } else {
      // a partial tile is padded with a key whose digit is the largest one in every pass, so the
      // padding ranks after all real items and is never written out
#pragma unroll
      for (int u = 0; u < IPT; ++u) key[u] = (base + u * 32 < n) ? (W)ld_stream(g + u * 32) : (W)P.pad_key;
    }
  };
  load_keys(tile);

  while (true) {
    const unsigned long long tile_base = (unsigned long long)tile * TILE;
    const unsigned long long remain = n - tile_base;
    const bool full = remain >= (unsigned long long)TILE;
    const int valid = full ? TILE : (int)remain;

    // ---- P1: match-rank inside the warp (keys are in registers).  Ranks inside this warp's digit
    // bucket, two 16-bit ranks per register (the digit is recomputed from the key when needed:
    // registers, not ALU slots, are what limits the tile size)
    unsigned int rk2[(IPT + 1) / 2];
    {
      const unsigned int myhist_s = smem_u32(myhist);
      unsigned int d_next = op(key[0]);
      unsigned int m_next = match_ballot<RADIX_BITS>(d_next);
#pragma unroll
      for (int u = 0; u < IPT; ++u) {
        const unsigned int d = d_next;
        const unsigned int m = m_next;
        if (u + 1 < IPT) {
          d_next = op(key[u + 1]);
          m_next = match_ballot<RADIX_BITS>(d_next);
        }
        const unsigned int leader = bfind(m);  // highest peer lane adds the whole group
        unsigned int prev = atoms_add_if(lane == leader, myhist_s + d * 4, (unsigned int)__popc(m));
        prev = __shfl_sync(0xffffffffu, prev, leader);
        const unsigned int r = prev + __popc(m & lt);
        if (u & 1)
          rk2[u / 2] |= r << 16;
        else
          rk2[u / 2] = r;
      }
    }
    __syncthreads();  // S2: all warp histograms complete; everybody is done with the previous tile

    // values -> registers; in flight during the digit scan and the look-back
    ValU val[HAS_VALUES ? IPT : 1];
    if (HAS_VALUES) {
      const ValU* g = reinterpret_cast<const ValU*>(P.vals_in) + tile_base + warp_base;
      if (full) {
#pragma unroll
        for (int u = 0; u < IPT; ++u) val[u] = ld_stream(g + u * 32);
      } else {
#pragma unroll
        for (int u = 0; u < IPT; ++u)
          if (warp_base + u * 32 < (unsigned int)valid) val[u] = ld_stream(g + u * 32);
      }
    }

    // ---- P2: per-digit tile counts -> partial status; digit prefix; per-warp bases
    OffT* status = reinterpret_cast<OffT*>(P.status) + (size_t)tile * RADIX;
    unsigned int total = 0;
    if (tid < RADIX) {
#pragma unroll
      for (int w = 0; w < NW; ++w) total += whist[w * RADIX + tid];
      st_status(status + tid, (tile == 0 ? FLAG_INCLUSIVE : FLAG_PARTIAL) | (OffT)total);
      if (P.status_next) reinterpret_cast<OffT*>(P.status_next)[(size_t)tile * RADIX + tid] = 0;
    }
    unsigned int incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (tid < RADIX && lane == 31) s_wtot[warp] = incl;
    __syncthreads();  // S2b
    unsigned int tile_excl = 0;
    if (tid < RADIX) {
      unsigned int base = 0;
#pragma unroll
      for (int w = 0; w < RADIX / 32; ++w)
        if (w < warp) base += s_wtot[w];
      tile_excl = base + incl - total;
      // counts -> running bases (re-read instead of holding NW counts in registers across the scan)
      unsigned int run = tile_excl;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        const unsigned int c = whist[w * RADIX + tid];
        whist[w * RADIX + tid] = run;
        run += c;
      }
    }
    __syncthreads();  // S3: per-warp bases ready

    // ---- P3: reorder keys and values in shared memory
#pragma unroll
    for (int u = 0; u < IPT; ++u) {
      const unsigned int r = ((u & 1) ? (rk2[u / 2] >> 16) : (rk2[u / 2] & 0xffffu)) + myhist[op(key[u])];
      sort_k[r] = (KeyU)key[u];
      if (HAS_VALUES) sort_v[r] = val[u];
    }
    // this warp's counters are free again: clear them for the next tile
    __syncwarp();
#pragma unroll
    for (int i = lane; i < RADIX; i += 32) myhist[i] = 0;

    // Claim the next tile as late as possible: tiles must START in (nearly) the order of their ids, or a tile
    // spins on predecessors that were claimed earlier but are still waiting for their CTA to get to them.
    if (tid == NT - 1) s_tile[1] = atomicAdd(P.tile_counter, 1u);

    // ---- look-back: exclusive prefix of this tile for digit `tid`.  Each round trip reads the next LBW
    // predecessors with independent loads (issued only now, so that they see fresh state) and sums partial
    // counts up to the nearest inclusive prefix.
    if (tid < RADIX) {
      OffT excl = 0;
      if (tile > 0) {
        const OffT* p = status - RADIX + tid;  // first entry of the current window
        unsigned int left = tile;              // predecessors not yet examined
        bool done = false;
        while (true) {
          OffT win[LBW];
#pragma unroll
          for (int j = 0; j < LBW; ++j) win[j] = (left > (unsigned int)j) ? ld_status(p - j * RADIX) : FLAG_INCLUSIVE;
#pragma unroll
          for (int j = 0; j < LBW; ++j) {
            if (!done) {
              OffT v = win[j];
              while ((v & FLAG_ANY) == 0) v = ld_status(p - j * RADIX);
              excl += v & VALUE_MASK;
              if (v & FLAG_INCLUSIVE) done = true;
            }
          }
          if (done) break;
          p -= LBW * RADIX;
          left -= LBW;
        }
        st_status(status + tid, FLAG_INCLUSIVE | (excl + (OffT)total));
      }
      s_goff[tid] = reinterpret_cast<const OffT*>(P.bins)[tid] + excl - (OffT)tile_excl;
    }
    __syncthreads();  // S4

    // ---- prefetch the next tile's keys into registers: in flight during the write-out
    const unsigned int next = s_tile[1];
    const bool more = next < num_tiles;
    if (more) load_keys(next);

    // ---- P4: coalesced write-out of digit runs
    {
      KeyU* okeys = reinterpret_cast<KeyU*>(P.keys_out);
      ValU* ovals = reinterpret_cast<ValU*>(P.vals_out);
      if (full) {
#pragma unroll 4
        for (int u = 0; u < IPT; ++u) {
          const int pos = u * NT + tid;
          const KeyU k = sort_k[pos];
          const OffT dst = s_goff[op((W)k)] + (OffT)pos;
          okeys[dst] = k;
          if (HAS_VALUES) ovals[dst] = sort_v[pos];
        }
      } else {
#pragma unroll 1
        for (int pos = tid; pos < valid; pos += NT) {
          const KeyU k = sort_k[pos];
          const OffT dst = s_goff[op((W)k)] + (OffT)pos;
          okeys[dst] = k;
          if (HAS_VALUES) ovals[dst] = sort_v[pos];
        }
      }
    }
    if (!more) break;
    tile = next;
  }
}

}  // namespace b2s
