// EXPERIMENT (round 1, not built into libb2s.so) -- measured 44.8 GKeys/s (512x20) / 43.5 (512x16) on u32/u32 2^28
// against 49.1 for two independent CTAs per SM (gpurun_out/tune_r1k.jsonl): forcing the ranking phases to alternate
// does NOT help.  The ranking phase itself depends on the shared-memory pipe (one leader ATOMS + SHFL per row on
// its critical path), so running it against a partner that saturates that pipe with the reorder scatter slows it
// down more than the ALU/LSU overlap gains; and only 16 warps rank at a time.
//
// onesweep_pingpong_r1.cuh (was b2s_onesweep_pp.cuh) -- "ping-pong" digit pass: the one-tile digit pass of b2s_onesweep.cuh run by TWO
// persistent 512-thread groups inside one 1024-thread CTA per SM, with the ranking phase of the two groups
// forced to alternate.
//
// Why: the ranking phase (P1) is ALU/issue-bound (8 ballot rounds per row), the reorder / look-back / write-out
// phases (P2-P4) are bound by the shared-memory / LSU pipe.  Two independent CTAs per SM only overlap those by
// chance (ncu, profiles/r1_onesweep_production.ncu.txt: ALU pipe 50 % + LSU wavefront pipe 59 % busy, i.e. the two
// resources are used almost one after the other).  Here a token (a shared-memory turn word polled by one thread
// per group) lets only one group rank at a time; the other group is then in its LSU-bound phases, so both pipes
// stay busy by construction.
//
// Same contract, same per-tile algorithm, same look-back protocol and status layout as onesweep_kernel; tiles are
// claimed by the groups from the same global counter immediately before their loads (start order == tile order).
#pragma once
#include "b2s_common.cuh"
#include "b2s_onesweep.cuh"

namespace b2s {

__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <int KBYTES, int VBYTES, typename OpT, typename OffT, int NT, int IPT, int LBW, bool PEER>
__global__ void __launch_bounds__(2 * NT, 1) onesweep_pp_kernel(const OnesweepParams<KBYTES, OpT> P) {
  using KeyU = typename UIntOf<KBYTES>::type;
  using W = typename WideOf<KBYTES>::type;
  using ValU = typename UIntOf<VBYTES ? VBYTES : 1>::type;
  using L = OnesweepSmem<KBYTES, VBYTES, NT, IPT>;
  constexpr int TILE = L::TILE;
  constexpr int NW = L::NW;
  constexpr bool HAS_VALUES = VBYTES != 0;
  constexpr int OBITS = sizeof(OffT) * 8;
  constexpr OffT FLAG_INCLUSIVE = OffT(1) << (OBITS - 1);
  constexpr OffT FLAG_PARTIAL = OffT(1) << (OBITS - 2);
  constexpr OffT VALUE_MASK = FLAG_PARTIAL - 1;
  constexpr int GROUP_SMEM = (L::TOTAL + 127) / 128 * 128;
  static_assert(NT >= RADIX && NT % 32 == 0, "one thread per digit needed");
  static_assert(32 * IPT < 65536, "warp-bucket rank is packed into 16 bits");

  extern __shared__ __align__(128) unsigned char smem_all[];
  __shared__ volatile unsigned int s_done[2];  // group has no more tiles
  __shared__ volatile unsigned int s_turn;     // group that may rank next

  const int g = threadIdx.x >= NT ? 1 : 0;  // group
  const int tid = threadIdx.x - g * NT;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  unsigned char* smem = smem_all + g * GROUP_SMEM;
  unsigned char* stage_k = smem + L::OFF_KEYS;
  unsigned char* stage_v = smem + L::OFF_VALS;
  unsigned int* whist = reinterpret_cast<unsigned int*>(smem + L::OFF_WHIST);
  OffT* s_goff = reinterpret_cast<OffT*>(smem + L::OFF_GOFF);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::OFF_MISC);                  // [2]
  unsigned int* s_wtot = reinterpret_cast<unsigned int*>(smem + L::OFF_MISC + 16);  // [8]
  unsigned int* s_tile = reinterpret_cast<unsigned int*>(smem + L::OFF_MISC + 64);

  const int BAR_GROUP = 1 + g;  // this group's block barrier
  auto group_sync = [&]() { named_bar_sync(BAR_GROUP, NT); };

  const unsigned long long n = P.n;
  const unsigned int num_tiles = (unsigned int)((n + TILE - 1) / TILE);
  const auto op = P.op;
  unsigned int* myhist = whist + warp * RADIX;
  const unsigned int lt = lanemask_lt();
  const int warp_base = warp * 32 * IPT;

  if (tid == 0) {
    s_done[g] = 0;
    if (g == 0) s_turn = 0;
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();  // the only CTA-wide barrier: both s_done flags are initialised

  unsigned int phase = 0;  // parity of the two mbarriers (they complete once per bulk-staged tile)

  while (true) {
    // ---- P0: claim a tile (claim order == start order), stage it, clear counters
    if (tid == 0) *s_tile = atomicAdd(P.tile_counter, 1u);
#pragma unroll
    for (int i = tid; i < NW * RADIX; i += NT) whist[i] = 0;
    group_sync();  // also: every thread of the group is done with the previous tile's shared memory
    const unsigned int tile = *s_tile;
    if (tile >= num_tiles) break;
    const unsigned long long tile_base = (unsigned long long)tile * TILE;
    const unsigned long long remain = n - tile_base;
    const bool full = remain >= (unsigned long long)TILE;
    const int valid = full ? TILE : (int)remain;
    const KeyU* gkeys = reinterpret_cast<const KeyU*>(P.keys_in) + tile_base;
    const ValU* gvals = reinterpret_cast<const ValU*>(P.vals_in) + tile_base;
    const uintptr_t kaddr = reinterpret_cast<uintptr_t>(gkeys);
    const uintptr_t vaddr = reinterpret_cast<uintptr_t>(gvals);
    unsigned int kshift = (unsigned int)(kaddr & 15);
    unsigned int vshift = HAS_VALUES ? (unsigned int)(vaddr & 15) : 0;
    const unsigned int kbytes = (kshift + TILE * KBYTES + 15u) & ~15u;
    const unsigned int vbytes = (vshift + TILE * VBYTES + 15u) & ~15u;
    bool bulk = full && (tile > 0 || (kshift == 0 && vshift == 0));
    bulk = bulk && (kshift == 0 || remain * KBYTES >= (unsigned long long)kbytes - kshift) &&
           (vshift == 0 || remain * VBYTES >= (unsigned long long)vbytes - vshift);
    if (bulk) {
      if (tid == 0) {
        fence_proxy_async_smem();  // generic-proxy accesses of the previous tile before the async-proxy writes
        mbar_expect_tx(&bar[0], kbytes);
        bulk_g2s(stage_k, reinterpret_cast<const void*>(kaddr - kshift), kbytes, &bar[0]);
        if (HAS_VALUES) {
          mbar_expect_tx(&bar[1], vbytes);
          bulk_g2s(stage_v, reinterpret_cast<const void*>(vaddr - vshift), vbytes, &bar[1]);
        }
      }
    } else {
      kshift = 0;
      vshift = 0;
      KeyU* sk = reinterpret_cast<KeyU*>(stage_k);
      for (int i = tid; i < TILE; i += NT) sk[i] = i < valid ? gkeys[i] : (KeyU)P.pad_key;
      if (HAS_VALUES) {
        ValU* sv = reinterpret_cast<ValU*>(stage_v);
        for (int i = tid; i < valid; i += NT) sv[i] = gvals[i];
      }
      group_sync();
    }

    // ---- ranking token: only one group ranks at a time.  One thread polls the turn word (the others sleep in
    // the group barrier); a partner that ran out of tiles releases the token for good.
    if (tid == 0) {
      while (s_turn != (unsigned int)g && s_done[1 - g] == 0) __nanosleep(40);
    }
    group_sync();

    // ---- P1: keys -> registers (warp-striped rows), match-rank inside the warp
    W key[IPT];
    unsigned int rk[IPT];  // (digit << 16) | rank inside this warp's digit bucket
    {
      if (bulk) mbar_wait(&bar[0], phase);
      const KeyU* sk = reinterpret_cast<const KeyU*>(stage_k + kshift);
#pragma unroll
      for (int u = 0; u < IPT; ++u) key[u] = (W)sk[warp_base + u * 32 + lane];
    }
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
        const unsigned int leader = bfind(m);
        unsigned int prev = atoms_add_if(lane == leader, myhist_s + d * 4, (unsigned int)__popc(m));
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rk[u] = (prev + __popc(m & lt)) | (d << 16);
      }
    }
    group_sync();  // S2: all warp histograms complete, all staged keys consumed
    if (tid == 0) s_turn = 1 - g;  // pass the token

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
    group_sync();  // S2b
    unsigned int tile_excl = 0;
    if (tid < RADIX) {
      unsigned int base = 0;
#pragma unroll
      for (int w = 0; w < RADIX / 32; ++w)
        if (w < warp) base += s_wtot[w];
      tile_excl = base + incl - total;
      unsigned int run = tile_excl;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        const unsigned int c = whist[w * RADIX + tid];
        whist[w * RADIX + tid] = run;
        run += c;
      }
    }
    group_sync();  // S3: per-warp bases ready

    // ---- P3: reorder keys in shared memory (staged keys were all consumed before S2)
    {
      KeyU* sk = reinterpret_cast<KeyU*>(stage_k);
#pragma unroll
      for (int u = 0; u < IPT; ++u) {
        const unsigned int r = (rk[u] & 0xffffu) + myhist[rk[u] >> 16];
        rk[u] = r;
        sk[r] = (KeyU)key[u];
      }
    }
    ValU val[HAS_VALUES ? IPT : 1];
    if (HAS_VALUES) {
      if (bulk) mbar_wait(&bar[1], phase);
      const ValU* sv = reinterpret_cast<const ValU*>(stage_v + vshift);
#pragma unroll
      for (int u = 0; u < IPT; ++u) val[u] = sv[warp_base + u * 32 + lane];
    }
    if (bulk) phase ^= 1u;

    // ---- look-back (see b2s_onesweep.cuh)
    if (tid < RADIX) {
      OffT excl = 0;
      if (tile > 0) {
        const OffT* p = status - RADIX + tid;
        unsigned int left = tile;
        bool done = false;
        while (true) {
          OffT win[LBW];
#pragma unroll
          for (int j = 0; j < LBW; ++j) win[j] = (left > (unsigned int)j) ? ld_status(p - j * RADIX) : FLAG_INCLUSIVE;
#pragma unroll
          for (int j = 0; j < LBW; ++j) {
            if (!done) {
              OffT v = win[j];
              while ((v & (FLAG_INCLUSIVE | FLAG_PARTIAL)) == 0) v = ld_status(p - j * RADIX);
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
    if (HAS_VALUES) {
      group_sync();  // S3b: every staged value is in a register
      ValU* sv = reinterpret_cast<ValU*>(stage_v);
#pragma unroll
      for (int u = 0; u < IPT; ++u) sv[rk[u]] = val[u];
    }
    group_sync();  // S4

    // ---- P4: coalesced write-out of digit runs
    {
      const KeyU* sk = reinterpret_cast<const KeyU*>(stage_k);
      const ValU* sv = reinterpret_cast<const ValU*>(stage_v);
      KeyU* okeys = reinterpret_cast<KeyU*>(P.keys_out);
      ValU* ovals = reinterpret_cast<ValU*>(P.vals_out);
      auto emit = [&](int pos) {
        const KeyU k = sk[pos];
        const unsigned int d = op((W)k);
        const OffT dst = s_goff[d] + (OffT)pos;
        if (PEER) {
          okeys = reinterpret_cast<KeyU*>(P.peer_keys[d & (MAX_PEERS - 1)]);
          ovals = reinterpret_cast<ValU*>(P.peer_vals[d & (MAX_PEERS - 1)]);
        }
        okeys[dst] = k;
        if (HAS_VALUES) ovals[dst] = sv[pos];
      };
      if (full) {
#pragma unroll
        for (int u = 0; u < IPT; ++u) emit(u * NT + tid);
      } else {
#pragma unroll 1
        for (int pos = tid; pos < valid; pos += NT) emit(pos);
      }
    }
  }

  // ---- no more tiles for this group: release a partner that may be waiting for the token, for good
  if (tid == 0) s_done[g] = 1;
}

}  // namespace b2s
