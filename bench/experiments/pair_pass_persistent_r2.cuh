// bench/experiments/pair_pass_persistent_r2.cuh -- REJECTED round-2 experiment, not compiled into anything (profiles/
// r2_tune_persistent_pair_kernel.jsonl): 63.5 GKeys/s (value copy issued after the write-out, waited for before S2),
// 56 (values straight from global memory into registers: the LDGs clog the LSU queue), 44 (value registers loaded during
// the digit scan: 368 B of spills) against 69.2 for the one-tile-per-CTA kernel -- the key copy is hidden, the value copy
// is exposed instead, two more block barriers per tile, and no register is left for anything smarter.
//
// PERSISTENT digit pass for 4-byte keys with 4-byte values (the pair flow of b2s_pass.cuh: a
// (key, value) pair is ranked once and scattered as one 64-bit shared-memory store).
//
// Same result, same phases and same look-back protocol as digit_pass_kernel<4, 4, ..., PF_PAIR | PF_NOBR>; what differs
// is WHEN a tile's TMA copies are issued.  In the one-tile-per-CTA kernel a CTA starts by waiting for its own copies
// (~1-2 us of the ~12 us a tile takes: the copy cannot be faster than the SM's share of the HBM stream), with a third of
// the SM's registers and half of its shared memory idle.  Here a CTA loops over tiles (atomic tickets, so a tile's
// predecessors are always running or finished -- no reliance on dispatch order) and the NEXT tile's copies are issued
// while the current tile is still being written out:
//   * the sorted pairs of the current tile are read front to back by the write-out; once the first IPT/2 + 1 steps have
//     been read by every thread, the front half of the staging area (== the key staging buffer) is dead: the key copy of
//     the next tile goes there (ticket taken at the start of the write-out, so its latency is hidden too);
//   * after the last step the value copy follows.
// The next tile then starts with its keys already in shared memory; its value copy overlaps its counting sweep.
//
// Replaces (reference, for parity of RESULT only): AgentRadixSortOnesweep's tile loop,
// cub/agent/agent_radix_sort_onesweep.cuh:616-688 (that kernel is not persistent either; it claims one tile per CTA).
#pragma once
#include "b2s_pass.cuh"

namespace b2s {

template <typename OpT, typename OffT, int NT, int IPT, int MINB, int LBW, int STAG_NS = 0, int PFD = 222>
__global__ void __launch_bounds__(NT, MINB) pair_pass_persistent_kernel(const OnesweepParams<4, OpT> P) {
  constexpr int KBYTES = 4, VBYTES = 4;
  constexpr bool CONV = OpConverts<OpT>::value;
  using KeyU = uint32_t;
  using ValU = uint32_t;
  using W = uint32_t;
  using L = PassSmem<KBYTES, VBYTES, NT, IPT, false>;
  constexpr int TILE = L::TILE;
  constexpr int NW = L::NW;
  constexpr int OBITS = sizeof(OffT) * 8;
  constexpr OffT FLAG_INCLUSIVE = OffT(1) << (OBITS - 1);
  constexpr OffT FLAG_PARTIAL = OffT(1) << (OBITS - 2);
  // write-out steps after which the key staging buffer (plus the <= 32 bytes of an unaligned window) has been read
  constexpr int UH = IPT / 2 + 1;
  static_assert(NT == RADIX, "one thread per digit, every warp scans and looks back");
  static_assert(UH < IPT && (long)UH * NT * 8 >= (long)L::OFF_VALS, "front half of the sorted pairs covers the key staging buffer");

  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* stage_k = smem + L::OFF_KEYS;
  unsigned char* stage_v = smem + L::OFF_VALS;
  unsigned int* whist = reinterpret_cast<unsigned int*>(smem + L::OFF_WHIST);
  OffT* s_goff = reinterpret_cast<OffT*>(smem + L::OFF_GOFF);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::OFF_MISC);                   // [2] keys, values
  unsigned int* s_wtot = reinterpret_cast<unsigned int*>(smem + L::OFF_MISC + 16);   // [8]
  volatile unsigned int* s_tile = reinterpret_cast<unsigned int*>(smem + L::OFF_MISC + 64);  // [2] tile id per iteration parity
  volatile unsigned int* s_geom = reinterpret_cast<unsigned int*>(smem + L::OFF_MISC + 72);  // [2] tile summary per parity

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const unsigned long long num_tiles = (P.n + TILE - 1) / TILE;

  struct TileGeom {
    unsigned long long base, remain;
    uintptr_t kaddr, vaddr;
    unsigned int kshift, vshift, kbytes, vbytes;
    bool full, bulk;
  };
  auto geom = [&](unsigned long long tile) {
    TileGeom g;
    g.base = tile * TILE;
    g.remain = P.n - g.base;
    g.full = g.remain >= (unsigned long long)TILE;
    g.kaddr = reinterpret_cast<uintptr_t>(reinterpret_cast<const KeyU*>(P.keys_in) + g.base);
    g.vaddr = reinterpret_cast<uintptr_t>(reinterpret_cast<const ValU*>(P.vals_in) + g.base);
    g.kshift = (unsigned int)(g.kaddr & 15);
    g.vshift = (unsigned int)(g.vaddr & 15);
    g.kbytes = (g.kshift + TILE * KBYTES + 15u) & ~15u;
    g.vbytes = (g.vshift + TILE * VBYTES + 15u) & ~15u;
    g.bulk = g.full && (tile > 0 || (g.kshift == 0 && g.vshift == 0));
    g.bulk = g.bulk && (g.kshift == 0 || g.remain * KBYTES >= (unsigned long long)g.kbytes - g.kshift) &&
             (g.vshift == 0 || g.remain * VBYTES >= (unsigned long long)g.vbytes - g.vshift);
    return g;
  };
  // summary word of a tile: bit 0 bulk copies, bit 1 full, bit 2 mbarrier phase parity, bits 8-11 / 16-19 staging shifts
  auto summary = [](const TileGeom& g, unsigned int parity) {
    return (g.bulk ? 1u : 0u) | (g.full ? 2u : 0u) | (parity << 2) | (g.bulk ? (g.kshift << 8) | (g.vshift << 16) : 0u);
  };

  // thread 0 only: bulk tiles issued so far (an mbarrier completes one phase per bulk tile), pending next tile
  unsigned int nbulk = 0;
  unsigned int next_ticket = 0;

  // ---- prologue: first ticket, barriers, first tile's copies
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
    const unsigned int t0 = atomicAdd(P.tile_counter, 1u);
    s_tile[0] = t0;
    if ((unsigned long long)t0 < num_tiles) {
      const TileGeom g = geom(t0);
      s_geom[0] = summary(g, 0u);
      if (g.bulk) {
        mbar_expect_tx(&bar[0], g.kbytes);
        bulk_g2s(stage_k, reinterpret_cast<const void*>(g.kaddr - g.kshift), g.kbytes, &bar[0]);
        mbar_expect_tx(&bar[1], g.vbytes);
        bulk_g2s(stage_v, reinterpret_cast<const void*>(g.vaddr - g.vshift), g.vbytes, &bar[1]);
        nbulk = 1;
      }
    }
  }
#pragma unroll
  for (int i = tid; i < NW * RADIX; i += NT) whist[i] = 0;
  __syncthreads();

  // All CTAs of a persistent launch start in the same instant and would stay phase-locked (every SM counting, ranking and
  // writing out at the same time, both CTAs of an SM in the same phase).  Spread the starts over one tile period.
  if (STAG_NS > 0) {
    const unsigned int slot = (blockIdx.x * 2654435761u) >> 20;  // pseudo-random 12 bits
    unsigned int wait_ns = (unsigned int)(((unsigned long long)slot * STAG_NS) >> 12);
    while (wait_ns > 0) {
      const unsigned int step = wait_ns > 1000u ? 1000u : wait_ns;
      __nanosleep(step);
      wait_ns -= step;
    }
  }
  auto op = P.op;
  op.prepare();
  KeyU* okeys = reinterpret_cast<KeyU*>(P.keys_out);
  ValU* ovals = reinterpret_cast<ValU*>(P.vals_out);

  // ---- constant-digit pass: the partition is the identity.  Wait for the copies already in flight, then copy tile after
  // tile straight from global to global memory (converting when the encodings of input and output differ).
  if (P.skip_flag != nullptr && __ldg(P.skip_flag) != 0u) {
    unsigned int cur = 0;
    {
      const unsigned int gs = s_geom[0];
      if ((unsigned long long)s_tile[0] < num_tiles && (gs & 1u)) {
        mbar_wait(&bar[0], 0);
        mbar_wait(&bar[1], 0);
      }
    }
    while (true) {
      const unsigned long long tile = s_tile[cur];
      if (tile >= num_tiles) break;
      const TileGeom g = geom(tile);
      const int valid = g.full ? TILE : (int)g.remain;
      if (P.status_next) reinterpret_cast<OffT*>(P.status_next)[tile * RADIX + tid] = 0;
      const KeyU* ik = reinterpret_cast<const KeyU*>(g.kaddr);
      const ValU* iv = reinterpret_cast<const ValU*>(g.vaddr);
      int conv = 0;
      if constexpr (CONV) conv = P.op.raw_in == P.op.raw_out ? 0 : (P.op.raw_in ? 1 : 2);
#pragma unroll 4
      for (int i = tid; i < valid; i += NT) {
        KeyU k = ik[i];
        if constexpr (CONV) {
          if (conv == 1) k = (KeyU)P.op.to_image((W)k);
          else if (conv == 2) k = (KeyU)P.op.to_raw((W)k);
        }
        okeys[g.base + i] = k;
        ovals[g.base + i] = iv[i];
      }
      if (tid == 0) s_tile[cur ^ 1] = atomicAdd(P.tile_counter, 1u);
      __syncthreads();
      cur ^= 1;
    }
    return;
  }

  const int warp_base = warp * 32 * IPT;
  unsigned int* myhist = whist + warp * RADIX;
  const unsigned int myhist_s = smem_u32(myhist);
  const unsigned int lt = lanemask_lt();
  const unsigned int scratch = smem_u32(smem + L::OFF_DUMMY) + (unsigned int)tid * 4u;

  unsigned int cur = 0;  // iteration parity: which s_tile / s_geom slot describes the current tile
  while (true) {
    const unsigned long long tile = s_tile[cur];
    if (tile >= num_tiles) break;
    const unsigned int gs = s_geom[cur];

    if (PFD && tid == 32 && tile + PFD + 1 < num_tiles) {
      const TileGeom g = geom(tile);
      bulk_prefetch_l2(reinterpret_cast<const void*>((g.kaddr + (unsigned long long)PFD * TILE * KBYTES) & ~(uintptr_t)15),
                       (unsigned int)(TILE * KBYTES) & ~15u);
      bulk_prefetch_l2(reinterpret_cast<const void*>((g.vaddr + (unsigned long long)PFD * TILE * VBYTES) & ~(uintptr_t)15),
                       (unsigned int)(TILE * VBYTES) & ~15u);
    }
    if (!(gs & 1u)) {
      // edge tiles (first tile behind an unaligned pointer, partial last tile): element loads, padded with a key that
      // orders last in every pass
      const TileGeom g = geom(tile);
      const int valid = g.full ? TILE : (int)g.remain;
      const KeyU* gkeys = reinterpret_cast<const KeyU*>(g.kaddr);
      const ValU* gvals = reinterpret_cast<const ValU*>(g.vaddr);
      KeyU* sk = reinterpret_cast<KeyU*>(stage_k);
      ValU* sv = reinterpret_cast<ValU*>(stage_v);
      KeyU pad = (KeyU)P.pad_key;
      if constexpr (CONV) {
        if (!P.op.raw_in) pad = (KeyU)OpT::ONES;
      }
      for (int i = tid; i < TILE; i += NT) sk[i] = i < valid ? gkeys[i] : pad;
      for (int i = tid; i < valid; i += NT) sv[i] = gvals[i];
      __syncthreads();
    }

    // ---- P1: keys -> registers, counting sweep, values -> registers
    W key[IPT];
    ValU val[IPT];
    {
      if (gs & 1u) mbar_wait(&bar[0], (gs >> 2) & 1u);
      const KeyU* sk = reinterpret_cast<const KeyU*>(stage_k + ((gs >> 8) & 15u));
#pragma unroll
      for (int u = 0; u < IPT; ++u) key[u] = (W)sk[warp_base + u * 32 + lane];
    }
    if constexpr (CONV) {
      if (op.raw_in) {
#pragma unroll
        for (int u = 0; u < IPT; ++u) key[u] = op.to_image(key[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < IPT; ++u) red_shared_add(myhist_s + op(key[u]) * 4, 1u);
    __syncthreads();  // S2: warp histograms complete, staged keys consumed

    // ---- P2: tile digit counts -> PARTIAL status, digit prefix, per-warp bases (absolute slots of the sorted tile)
    unsigned int total = 0;
    OffT* status = reinterpret_cast<OffT*>(P.status) + tile * RADIX;
#pragma unroll
    for (int w = 0; w < NW; ++w) total += whist[w * RADIX + tid];
    st_status(status + tid, (tile == 0 ? (FLAG_INCLUSIVE | FLAG_PARTIAL) : FLAG_PARTIAL) | (OffT)total);
    if (P.status_next) reinterpret_cast<OffT*>(P.status_next)[tile * RADIX + tid] = 0;
    unsigned int incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();  // S2b
    {
      unsigned int base = 0;
#pragma unroll
      for (int w = 0; w < RADIX / 32; ++w)
        if (w < warp) base += s_wtot[w];
      unsigned int run = base + incl - total;
      s_goff[tid] = (OffT)run;  // parked until the look-back needs it
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        const unsigned int c = whist[w * RADIX + tid];
        whist[w * RADIX + tid] = run;
        run += c;
      }
    }
    {
      // staged values -> registers as late as possible (the value copy was issued at the very end of the previous tile),
      // but before S3: the pair scatter of the ranking sweep overwrites the value staging buffer
      if (gs & 1u) mbar_wait(&bar[1], (gs >> 2) & 1u);
      const ValU* sv = reinterpret_cast<const ValU*>(stage_v + ((gs >> 16) & 15u));
#pragma unroll
      for (int u = 0; u < IPT; ++u) val[u] = sv[warp_base + u * 32 + lane];
    }
    __syncthreads();  // S3: per-warp bases ready, staged values consumed

    // ---- P3: ranking fused with the pair scatter (b2s_pass.cuh, PF_PAIR | PF_NOBR)
    {
      auto place = [&](int u, unsigned int r) {
        reinterpret_cast<uint2*>(stage_k)[r] = make_uint2((unsigned int)key[u], (unsigned int)val[u]);
      };
      unsigned int d = op(key[0]);
      unsigned int m = match_ballot<RADIX_BITS, true>(d, P.ones);
      unsigned int bcast_prev = 0, below_prev = 0;
#pragma unroll
      for (int u = 0; u < IPT; ++u) {
        const unsigned int leader = bfind(m);
        const unsigned int below = __popc(m & lt);
        const unsigned int raw = atoms_add(lane == leader ? myhist_s + d * 4 : scratch, (unsigned int)__popc(m));
        unsigned int d_next = 0, m_next = 0;
        if (u + 1 < IPT) {
          d_next = op(key[u + 1]);
          m_next = match_ballot<RADIX_BITS, true>(d_next, P.ones);
        }
        if (u > 0) place(u - 1, bcast_prev + below_prev);
        bcast_prev = __shfl_sync(0xffffffffu, raw, leader);
        below_prev = below;
        d = d_next;
        m = m_next;
      }
      place(IPT - 1, bcast_prev + below_prev);
    }

    // ---- look-back: exclusive prefix of this tile for digit `tid`; thread 0 takes the next ticket first, so that its
    // round trip is over by the time the write-out is half done
    if (tid == 0) next_ticket = atomicAdd(P.tile_counter, 1u);
    {
      OffT excl = 0;
      if (tile > 0) {
        excl = lookback_exclusive<OffT, LBW>(status + tid, tile);
        st_status(status + tid, FLAG_INCLUSIVE | FLAG_PARTIAL | (excl + (OffT)total));
      }
      s_goff[tid] = reinterpret_cast<const OffT*>(P.bins)[tid] + excl - s_goff[tid];
    }
    __syncthreads();  // S4: sorted tile and global offsets complete

    // ---- P4: write-out of digit runs, front to back; the next tile's copies follow the reads
    auto emit = [&](int pos, auto to_raw) {
      const uint2 kv = reinterpret_cast<const uint2*>(stage_k)[pos];
      KeyU k = (KeyU)kv.x;
      const unsigned int d = op((W)k);
      const OffT dst = s_goff[d] + (OffT)pos;
      if constexpr (CONV && decltype(to_raw)::value) k = (KeyU)image_to_raw(op, (W)k);
      okeys[dst] = k;
      ovals[dst] = (ValU)kv.y;
    };
    auto emit_range = [&](auto lo, auto hi, auto to_raw) {
#pragma unroll
      for (int u = decltype(lo)::value; u < decltype(hi)::value; ++u) emit(u * NT + tid, to_raw);
    };
    auto emit_steps = [&](auto lo, auto hi) {
      if constexpr (CONV) {
        if (op.raw_out) emit_range(lo, hi, std::true_type{});
        else emit_range(lo, hi, std::false_type{});
      } else {
        emit_range(lo, hi, std::false_type{});
      }
    };
    auto emit_partial = [&](int lo, int hi) {  // partial tile: positions [lo, hi) of the valid items
      bool raw = false;
      if constexpr (CONV) raw = op.raw_out != 0;
#pragma unroll 1
      for (int pos = lo + tid; pos < hi; pos += NT) {
        if (raw) emit(pos, std::true_type{});
        else emit(pos, std::false_type{});
      }
    };
    const int valid = (gs & 2u) ? TILE : (int)(P.n - tile * TILE);
    if (gs & 2u) emit_steps(std::integral_constant<int, 0>{}, std::integral_constant<int, UH>{});
    else emit_partial(0, valid < UH * NT ? valid : UH * NT);
    __syncthreads();  // every thread has read the front half: the key staging buffer is dead
    if (tid == 0) {
      s_tile[cur ^ 1] = next_ticket;
      if ((unsigned long long)next_ticket < num_tiles) {
        const TileGeom g = geom(next_ticket);
        s_geom[cur ^ 1] = summary(g, nbulk & 1u);
        if (g.bulk) {
          fence_proxy_async();  // the buffer was read through the generic proxy and is now written by the TMA engine
          mbar_expect_tx(&bar[0], g.kbytes);
          bulk_g2s(stage_k, reinterpret_cast<const void*>(g.kaddr - g.kshift), g.kbytes, &bar[0]);
        }
      }
    }
    if (gs & 2u) emit_steps(std::integral_constant<int, UH>{}, std::integral_constant<int, IPT>{});
    else emit_partial(UH * NT, valid);
#pragma unroll
    for (int i = tid; i < NW * RADIX; i += NT) whist[i] = 0;  // for the next tile's counting sweep
    __syncthreads();  // every thread has read the whole sorted tile; next tile id and summary visible
    if (tid == 0 && (unsigned long long)next_ticket < num_tiles) {
      const TileGeom g = geom(next_ticket);
      if (g.bulk) {
        fence_proxy_async();
        mbar_expect_tx(&bar[1], g.vbytes);
        bulk_g2s(stage_v, reinterpret_cast<const void*>(g.vaddr - g.vshift), g.vbytes, &bar[1]);
        nbulk++;
      }
    }
    cur ^= 1;
  }
}

}  // namespace b2s
