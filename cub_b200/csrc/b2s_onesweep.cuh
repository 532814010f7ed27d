// b2s_onesweep.cuh -- LABORATORY version of the digit pass (tuning library only, -DB2S_TUNING; the product kernel is
// digit_pass_kernel in b2s_pass.cuh): the round-1 kernel with its ablation / trace / persistent / experiment branches,
// kept for A/B runs and the phase traces of bench/trace.py.
// One digit pass of the LSD sort: stable partition of all n items by one
// 8-bit digit into their global positions, chained-scan ("onesweep") style.
//
// Replaces (reference, for parity of RESULT only):
//   DeviceRadixSortOnesweepKernel  cub/device/dispatch/dispatch_radix_sort.cuh:580
//   AgentRadixSortOnesweep         cub/agent/agent_radix_sort_onesweep.cuh:98-688
//   BlockRadixRankMatchEarlyCounts cub/block/block_radix_rank.cuh:898-1192
//
// B200-first structure of the PRODUCTION flow (MODE = IMAD | BLOCKID | VAL_LATE | EARLY | FASTLB | prefetch, see the
// kernel's template comment; MODE = 0 is the classic flow kept for the multi-GPU partition pass and for A/B runs):
//   * the tile's keys AND values are staged global->shared by the TMA engine (cp.async.bulk + mbarrier; SASS UBLKCP),
//     issued by one thread as the very first thing the CTA does -- the tile id is the block index, so no claim round
//     trip precedes them -- plus an L2 prefetch (cp.async.bulk.prefetch.L2) of the tile ~one CTA lifetime ahead: no
//     per-item LDG instructions, no registers held by loads in flight; unaligned / partial tiles use element loads;
//   * three 12-warp CTAs per SM so that the load / count / rank / write-out phases of different tiles overlap;
//   * early counts: a counting sweep (one shared-memory reduction per item on warp-private counters) lets the tile
//     publish its digit counts BEFORE the ranking sweep; the digit scan then turns the counters into absolute positions
//     in the sorted tile, so the ranking sweep's leader atomic returns the final slot and every key is stored to it at
//     once -- no packed (digit, rank) word, no base gather, no separate reorder sweep;
//   * ranking by 8 ballot rounds whose complements run on the FMA pipe (IMAD), the ALU pipe being the half-rate one;
//   * look-back status words are gpu-scope relaxed, one array per pass parity (a pass clears the NEXT pass' array: one
//     memset per sort), 64-bit words when n >= 2^30 instead of <=2^28-item portions; the walk reads a window of LBW
//     predecessors with back-to-back loads at immediate offsets and sums it branch-free (AND-reduce "all published",
//     select chain up to the nearest inclusive word);
//   * values: staged values -> registers after the look-back, then in place to the slots remembered from the ranking;
//   * write-out: every thread walks the sorted tile with stride NT: coalesced stores of digit runs.
//
// Stability: items are ranked in tile order (warp-striped rows, lane order inside a row, rows in program order);
// tiles are ordered by their id == position in the input.
#pragma once
#include "b2s_common.cuh"
#include "b2s_pass.cuh"  // OnesweepParams, MAX_PEERS, look-back helpers shared with the production kernel

namespace b2s {

template <int KBYTES, int VBYTES, int NT, int IPT>
struct OnesweepSmem {
  static constexpr int TILE = NT * IPT;
  static constexpr int NW = NT / 32;
  static constexpr int KEY_BYTES = TILE * KBYTES + 16;
  static constexpr int VAL_BYTES = VBYTES ? TILE * VBYTES + 16 : 0;
  static constexpr int OFF_KEYS = 0;
  static constexpr int OFF_VALS = (KEY_BYTES + 127) / 128 * 128;
  static constexpr int OFF_WHIST = OFF_VALS + (VAL_BYTES + 127) / 128 * 128;
  static constexpr int OFF_GOFF = OFF_WHIST + NW * RADIX * 4;
  static constexpr int OFF_MISC = OFF_GOFF + RADIX * 8;
  static constexpr int TOTAL = OFF_MISC + 256;  // barriers, scan partials, tile id; +128: phase stamps of the trace variants
};

// One run of the sorted tile -> its destination, cooperatively by the CTA: single items up to the destination's first
// 16-byte boundary and after its last one, 16-byte stores in between (the shared-memory source has no such alignment: it
// is read item by item).
template <typename T>
__device__ __forceinline__ void copy_run_wide(T* dst, const T* src, int len, int tid, int nt) {
  constexpr int A = 16 / (int)sizeof(T);
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "4- or 8-byte items");
  int head = (int)(((16u - (unsigned int)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u) / sizeof(T));
  if (head > len) head = len;
  const int groups = (len - head) / A;
  const int tail_at = head + groups * A;
  if (tid < head) dst[tid] = src[tid];
  if (tid < len - tail_at) dst[tail_at + tid] = src[tail_at + tid];
  for (int g = tid; g < groups; g += nt) {
    const T* s = src + head + g * A;
    union {
      uint4 v;
      T t[A];
    } u;
#pragma unroll
    for (int i = 0; i < A; ++i) u.t[i] = s[i];
    *reinterpret_cast<uint4*>(dst + head + g * A) = u.v;
  }
}

// ABL: timing-only ablation switches for bench/tune.py (results are WRONG when non-zero; never used by the product):
//   1 = no global stores in P4, 2 = no P4 at all, 4 = no ranking sweep, 8 = no look-back walk, 16 = no P3/P4 value path
// MODE: bits 0-1 = 0 one tile per CTA (grid = tiles) | 1 persistent CTA (grid = resident CTAs), next tile claimed after the
//       write-out | 2 persistent, next tile claimed before the write-out (claim latency hidden behind it);
//       bit 2 = (experiment) cross-proxy fence before a persistent CTA re-fills its staging buffers;
//       bit 3 = complement of the ballots on the FMA pipe (IMAD) instead of the ALU pipe (LOP3);
//       bit 4 = (tuning) thread 0 records the SM clock at every phase boundary into P.trace;
//       bit 8 = branch-free look-back window (all LBW words summed with a select chain when every one is published);
//       bit 9 = (tuning, with bit 4) look-back statistics in trace slots 11-14;
//       bit 10 = (experiment, not measured yet) the key copy is issued as four chunks (one per quarter of the warps) with
//                their own mbarriers, so that a warp's counting sweep starts as soon as ITS keys have landed;
//       bit 11 = (experiment, not measured yet) the first look-back window is requested four rows before the end of the
//                ranking sweep -- the key registers that have died by then hold it -- so that its L2 round trip is hidden;
//       bit 12 = (experiment, not measured yet; splitter passes only) write-out per destination run with 16-byte stores;
//       bits 16+ = L2 prefetch distance in tiles (the CTA of tile t asks L2 for the keys/values of tile t + distance).
template <int KBYTES, int VBYTES, typename OpT, typename OffT, int NT, int IPT, int MINB, int LBW, bool PEER, int ABL = 0,
          int MODE = 0>
__global__ void __launch_bounds__(NT, MINB) onesweep_kernel(const OnesweepParams<KBYTES, OpT> P) {
  constexpr int PERSIST = MODE & 3;
  constexpr bool TRACE = (MODE & 16) != 0;
  constexpr bool TRACE_LB = TRACE && (MODE & 512) != 0;  // also count the look-back's round trips / words (costs registers)
  constexpr bool BLOCKID = (MODE & 32) != 0;   // tile id = blockIdx.x (CTAs are dispatched in index order) instead of a claim
  constexpr bool VAL_LATE = (MODE & 64) != 0;  // staged values -> registers after the look-back (frees registers for its window)
  // EARLY: tile digit counts come from a cheap counting sweep (one shared-memory reduction per item) and are published BEFORE
  // the ranking sweep, so successors never wait for a slow ranking phase of this tile; the ranking atomics then run on
  // counters that already hold absolute positions, which removes the per-item base gather and fuses the key reorder into
  // the ranking loop.
  constexpr bool EARLY = (MODE & 128) != 0;
  constexpr bool FASTLB = (MODE & 256) != 0;
  constexpr bool EARLYWIN = (MODE & 2048) != 0 && EARLY && FASTLB && IPT > 6;
  constexpr bool WIDE = (MODE & 4096) != 0;  // few long runs (<= MAX_PEERS destinations): 16-byte stores per run
  static_assert(!WIDE || (OpT::kMaxDigit < MAX_PEERS && !EARLY && (KBYTES == 4 || KBYTES == 8) && (VBYTES == 0 || VBYTES == 4 || VBYTES == 8)),
                "wide write-out is for the splitter pass of the classic flow");
  constexpr bool CHUNKED = (MODE & 1024) != 0 && (NT / 32) % 4 == 0 && (NT / 128 * 32 * IPT * KBYTES) % 16 == 0;
  static_assert(!((MODE & 1024) && PERSIST), "chunked key copies are for one-tile CTAs");
  static_assert(!(BLOCKID && PERSIST), "persistent CTAs claim their tiles");
  static_assert(!(EARLY && ABL), "ablations apply to the classic flow");
  // stamps go to shared memory (fixed address, no registers held) and are copied out once per tile
#define B2S_TRACE(slot)                                                                                      \
  do {                                                                                                       \
    if (TRACE && threadIdx.x == 0) reinterpret_cast<long long*>(smem + L::OFF_MISC + 128)[slot] = clock64(); \
  } while (0)
  constexpr int PFD = MODE >> 16;
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
  static_assert(NT >= RADIX && NT % 32 == 0, "one thread per digit needed");
  static_assert(32 * IPT < 65536, "warp-bucket rank is packed into 16 bits");

  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* stage_k = smem + L::OFF_KEYS;
  unsigned char* stage_v = smem + L::OFF_VALS;
  unsigned int* whist = reinterpret_cast<unsigned int*>(smem + L::OFF_WHIST);
  OffT* s_goff = reinterpret_cast<OffT*>(smem + L::OFF_GOFF);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::OFF_MISC);            // [2]
  unsigned int* s_wtot = reinterpret_cast<unsigned int*>(smem + L::OFF_MISC + 16);  // [8]
  unsigned int* s_tile = reinterpret_cast<unsigned int*>(smem + L::OFF_MISC + 64);
  uint64_t* kbar = reinterpret_cast<uint64_t*>(smem + L::OFF_MISC + 80);       // [4], CHUNKED only

  int tid = threadIdx.x;
  B2S_TRACE(11);  // CTA entry

  // ---- P0: claim a tile (launch order == input order), arm the barriers, clear counters
  if (tid == 0) {
    if (!BLOCKID) *s_tile = atomicAdd(P.tile_counter, 1u);
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    if (CHUNKED) {
#pragma unroll
      for (int g = 0; g < 4; ++g) mbar_init(&kbar[g], 1);
    }
    mbar_fence_init();
  }
  if (!BLOCKID) {
#pragma unroll
    for (int i = tid; i < NW * RADIX; i += NT) whist[i] = 0;
    __syncthreads();
  }

  const unsigned long long num_tiles = (P.n + TILE - 1) / TILE;
  unsigned int phase = 0;  // parity of the mbarrier phase the next bulk copies complete
  for (;;) {               // one iteration per tile; a single one unless PERSIST
  // persistent CTAs: everything derived from the thread index is re-derived per tile (opaque) instead of being hoisted
  // out of the loop, where it would cost registers for the whole tile
  if (PERSIST) tid = (int)opaque((unsigned int)tid);
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const unsigned long long tile = BLOCKID ? (unsigned long long)blockIdx.x : (unsigned long long)*s_tile;
  if (PERSIST && tile >= num_tiles) break;
  if (TRACE && threadIdx.x == 0) {
    unsigned long long gt;
    unsigned int smid;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    reinterpret_cast<unsigned long long*>(smem + L::OFF_MISC + 128)[0] = gt;
    reinterpret_cast<unsigned long long*>(smem + L::OFF_MISC + 128)[15] = smid;
  }
  B2S_TRACE(1);  // tile claimed
  const unsigned long long tile_base = tile * TILE;
  const unsigned long long remain = P.n - tile_base;
  const bool full = remain >= (unsigned long long)TILE;
  const int valid = full ? TILE : (int)remain;

  const KeyU* gkeys = reinterpret_cast<const KeyU*>(P.keys_in) + tile_base;
  const ValU* gvals = reinterpret_cast<const ValU*>(P.vals_in) + tile_base;

  // TMA path needs 16-byte aligned source windows that stay inside the arrays.
  const uintptr_t kaddr = reinterpret_cast<uintptr_t>(gkeys);
  const uintptr_t vaddr = reinterpret_cast<uintptr_t>(gvals);
  unsigned int kshift = (unsigned int)(kaddr & 15);
  unsigned int vshift = HAS_VALUES ? (unsigned int)(vaddr & 15) : 0;
  const unsigned int kbytes = (kshift + TILE * KBYTES + 15u) & ~15u;
  const unsigned int vbytes = (vshift + TILE * VBYTES + 15u) & ~15u;
  bool bulk = full && (tile > 0 || (kshift == 0 && vshift == 0));
  bulk = bulk && (kshift == 0 || remain * KBYTES >= (unsigned long long)kbytes - kshift) &&
         (vshift == 0 || remain * VBYTES >= (unsigned long long)vbytes - vshift);

  // kept as an integer: a predicate that lives across the ranking loop would collide with the seven the ballots need
  const unsigned int bulk_flag = PERSIST ? opaque(bulk ? 1u : 0u) : 0u;
  if (PFD && tid == 32) {
    // ask L2 for a tile that will be claimed about one CTA lifetime from now, so that its TMA copies hit L2
    if (tile + PFD + 1 < num_tiles) {
      bulk_prefetch_l2(reinterpret_cast<const void*>((kaddr + (unsigned long long)PFD * TILE * KBYTES) & ~(uintptr_t)15),
                       (unsigned int)(TILE * KBYTES) & ~15u);
      if (HAS_VALUES)
        bulk_prefetch_l2(reinterpret_cast<const void*>((vaddr + (unsigned long long)PFD * TILE * VBYTES) & ~(uintptr_t)15),
                         (unsigned int)(TILE * VBYTES) & ~15u);
    }
  }
  if (bulk) {
    if (tid == 0) {
      if (PERSIST && (MODE & 4)) fence_proxy_async();  // (experiment) cross-proxy fence before re-filling the buffers
      if (CHUNKED && kshift == 0) {
        constexpr unsigned int CH = NT / 128 * 32 * IPT * KBYTES;  // the keys of a quarter of the warps
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          mbar_expect_tx(&kbar[g], CH);
          bulk_g2s(stage_k + g * CH, reinterpret_cast<const void*>(kaddr + g * CH), CH, &kbar[g]);
        }
      } else {
        mbar_expect_tx(&bar[0], kbytes);
        bulk_g2s(stage_k, reinterpret_cast<const void*>(kaddr - kshift), kbytes, &bar[0]);
      }
      if (HAS_VALUES) {
        mbar_expect_tx(&bar[1], vbytes);
        bulk_g2s(stage_v, reinterpret_cast<const void*>(vaddr - vshift), vbytes, &bar[1]);
      }
    }
  } else {
    kshift = 0;
    vshift = 0;
    KeyU* sk = reinterpret_cast<KeyU*>(stage_k);
    // a partial tile is padded with a key whose digit is the largest one in every pass, so
    // the padding ranks after all real items and is never written out
    for (int i = tid; i < TILE; i += NT) sk[i] = i < valid ? gkeys[i] : (KeyU)P.pad_key;
    if (HAS_VALUES) {
      ValU* sv = reinterpret_cast<ValU*>(stage_v);
      for (int i = tid; i < valid; i += NT) sv[i] = gvals[i];
    }
    __syncthreads();
  }

  if (BLOCKID) {  // the tile id needed no round trip: the copies above are already in flight while the counters are cleared
#pragma unroll
    for (int i = tid; i < NW * RADIX; i += NT) whist[i] = 0;
    __syncthreads();
  }

  // ---- P1: keys -> registers (warp-striped rows), match-rank inside the warp
  const int warp_base = warp * 32 * IPT;
  W key[IPT];
  unsigned int rk[IPT];  // (digit << 16) | rank inside this warp's digit bucket
  {
    if (bulk) {
      if (CHUNKED && kshift == 0) mbar_wait(&kbar[warp / (NW / 4)], phase);
      else mbar_wait(&bar[0], phase);
    }
    B2S_TRACE(2);  // keys staged
    const KeyU* sk = reinterpret_cast<const KeyU*>(stage_k + kshift);
#pragma unroll
    for (int u = 0; u < IPT; ++u) key[u] = (W)sk[warp_base + u * 32 + lane];
  }
  auto op = P.op;
  op.prepare();  // no-op for DigitOp; loads the device-resident splitters for SplitterOp
  unsigned int* myhist = whist + warp * RADIX;
  const unsigned int myhist_s = smem_u32(myhist);
  const unsigned int lt = lanemask_lt();
  // Software pipeline over rows (a warp issues in order, so the ORDER below is what hides latency):
  //   leader atomic of row u  ->  8 ballot rounds of row u+1 (ALU work while the atomic is in flight)
  //   ->  rank of row u-1 from its broadcast (issued one iteration ago)  ->  SHFL broadcast of row u's atomic.
  // Neither the ATOMS->SHFL nor the SHFL->use latency is exposed; only the shared-memory pipe's throughput is.
  if (EARLY) {
    // ---- P1a: counting sweep -- per-warp digit counts, no ranks yet
#pragma unroll
    for (int u = 0; u < IPT; ++u) red_shared_add(myhist_s + op(key[u]) * 4, 1u);
  } else {
  unsigned int d = op(key[0]);
  unsigned int m = match_ballot<RADIX_BITS, (MODE & 8) != 0>(d, P.ones);
  unsigned int bcast_prev = 0, below_prev = 0, d_prev = 0;
  if (ABL & 4) {
#pragma unroll
    for (int u = 0; u < IPT; ++u) rk[u] = (unsigned int)(u * 32 + lane) | (op(key[u]) << 16);
    if (lane == 0) myhist[0] = 32 * IPT;  // keep the tile total consistent: everything counted as digit 0
  }
#pragma unroll
  for (int u = 0; u < ((ABL & 4) ? 0 : IPT); ++u) {
    const unsigned int leader = bfind(m);  // highest peer lane adds the whole group
    const unsigned int below = __popc(m & lt);
    const unsigned int raw = atoms_add_if(lane == leader, myhist_s + d * 4, (unsigned int)__popc(m));
    unsigned int d_next = 0, m_next = 0;
    if (u + 1 < IPT) {
      d_next = op(key[u + 1]);
      m_next = match_ballot<RADIX_BITS, (MODE & 8) != 0>(d_next, P.ones);
    }
    // opaque(): the packed word must be formed HERE; otherwise the compiler keeps rank parts and masks alive across the
    // block barriers and re-derives the digit from the key afterwards (11 instructions per item instead of 6)
    if (u > 0) rk[u - 1] = opaque((bcast_prev + below_prev) | (d_prev << 16));
    bcast_prev = __shfl_sync(0xffffffffu, raw, leader);
    below_prev = below;
    d_prev = d;
    d = d_next;
    m = m_next;
  }
  if (!(ABL & 4)) rk[IPT - 1] = opaque((bcast_prev + below_prev) | (d_prev << 16));
  }
  __syncthreads();  // S2: all warp histograms complete, all staged keys consumed
  B2S_TRACE(3);  // ranked

  // ---- P2: per-digit tile counts -> partial status; digit prefix; per-warp bases
  OffT* status = reinterpret_cast<OffT*>(P.status) + tile * RADIX;
  unsigned int total = 0;
  if (tid < RADIX) {
#pragma unroll
    for (int w = 0; w < NW; ++w) total += whist[w * RADIX + tid];
    st_status(status + tid, (tile == 0 ? (FLAG_INCLUSIVE | FLAG_PARTIAL) : FLAG_PARTIAL) | (OffT)total);
    if (P.status_next) reinterpret_cast<OffT*>(P.status_next)[tile * RADIX + tid] = 0;
  }
  unsigned int incl = total;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (tid < RADIX && lane == 31) s_wtot[warp] = incl;
  __syncthreads();  // S2b
  if (tid < RADIX) {
    unsigned int base = 0;
#pragma unroll
    for (int w = 0; w < RADIX / 32; ++w)
      if (w < warp) base += s_wtot[w];
    // the per-warp counts are read a second time rather than kept in NW registers across the barrier, and the digit's
    // first position in the tile is parked in its s_goff slot until the look-back needs it: registers are what limits
    // the items per thread
    unsigned int run = base + incl - total;
    s_goff[tid] = (OffT)run;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const unsigned int c = whist[w * RADIX + tid];
      whist[w * RADIX + tid] = run;
      run += c;
    }
  }
  __syncthreads();  // S3: per-warp bases ready
  B2S_TRACE(4);  // digit scan done

  // EARLYWIN: first look-back window, requested from inside the ranking sweep
  OffT win0[EARLYWIN ? LBW : 1];
  const bool pre_window = EARLYWIN && tid < RADIX && tile >= (unsigned long long)LBW;

  // ---- P3: reorder keys in shared memory (staged keys were all consumed before S2)
  if (EARLY) {
    // ranking sweep on counters that hold absolute tile positions: the leader's atomic returns the position of the first
    // peer, so a key goes to its sorted slot as soon as its row is ranked (same software pipeline as the classic loop)
    KeyU* sk = reinterpret_cast<KeyU*>(stage_k);
    unsigned int d = op(key[0]);
    unsigned int m = match_ballot<RADIX_BITS, (MODE & 8) != 0>(d, P.ones);
    unsigned int bcast_prev = 0, below_prev = 0;
#pragma unroll
    for (int u = 0; u < IPT; ++u) {
      const unsigned int leader = bfind(m);
      const unsigned int below = __popc(m & lt);
      const unsigned int raw = atoms_add_if(lane == leader, myhist_s + d * 4, (unsigned int)__popc(m));
      unsigned int d_next = 0, m_next = 0;
      if (u + 1 < IPT) {
        d_next = op(key[u + 1]);
        m_next = match_ballot<RADIX_BITS, (MODE & 8) != 0>(d_next, P.ones);
      }
      if (u > 0) {
        const unsigned int r = bcast_prev + below_prev;
        rk[u - 1] = r;
        sk[r] = (KeyU)key[u - 1];
      }
      if (EARLYWIN && u == IPT - 4) {
        if (pre_window)
          load_status_window<RADIX * (int)sizeof(OffT)>(status - RADIX + tid, win0, std::make_integer_sequence<int, LBW>{});
      }
      bcast_prev = __shfl_sync(0xffffffffu, raw, leader);
      below_prev = below;
      d = d_next;
      m = m_next;
    }
    const unsigned int r = bcast_prev + below_prev;
    rk[IPT - 1] = r;
    sk[r] = (KeyU)key[IPT - 1];
  } else {
    KeyU* sk = reinterpret_cast<KeyU*>(stage_k);
#pragma unroll
    for (int u = 0; u < IPT; ++u) {
      const unsigned int r = (rk[u] & 0xffffu) + myhist[rk[u] >> 16];
      rk[u] = r;
      sk[r] = (KeyU)key[u];
    }
  }
  // values: staged values -> registers (re-using the key registers), barrier, then in-place reorder
  ValU val[HAS_VALUES ? IPT : 1];
  auto load_values = [&]() {
    if (bulk) mbar_wait(&bar[1], phase);
    const ValU* sv = reinterpret_cast<const ValU*>(stage_v + vshift);
#pragma unroll
    for (int u = 0; u < IPT; ++u) val[u] = sv[warp_base + u * 32 + lane];
  };
  if (HAS_VALUES && !VAL_LATE) load_values();

  B2S_TRACE(5);  // keys reordered, values in registers
  // ---- look-back: exclusive prefix of this tile for digit `tid`.  Each round trip reads the next LBW
  // predecessors with independent loads -- issued only now, so that they see fresh state: a predecessor
  // publishes its inclusive prefix one round trip after ITS look-back starts, so a load issued before the
  // scatter above would only ever see partial counts -- and sums partial counts up to the nearest
  // inclusive prefix.  The walk takes ~ (L2 round trip)^2 / (LBW * time between consecutive tiles).
  if (tid < RADIX) {
    OffT excl = 0;
    if (tile > 0 && !(ABL & 8)) {
      const OffT* p = status - RADIX + tid;  // first entry of the current window
      unsigned long long left = tile;        // predecessors not yet examined
      bool done = false;
      bool have_window = pre_window;  // EARLYWIN: the first window is already on its way (or here)
      unsigned int n_trips = 0, n_spins = 0, n_walked = 0;  // TRACE only
      long long t_first = 0;
      while (true) {
        OffT win[LBW];
        if (TRACE_LB) ++n_trips;
        bool handled = false;
        if (FASTLB && left >= (unsigned long long)LBW) {
          // common case: the whole window exists and every word in it is published.  An inclusive word carries BOTH flag
          // bits, so "all published" is one AND-reduction; the sum up to the nearest inclusive word is a select chain
          // from the far end (no branches, no loads with computed addresses)
          if (EARLYWIN && have_window) {
#pragma unroll
            for (int j = 0; j < LBW; ++j) win[j] = win0[j];
          } else {
            load_status_window<RADIX * (int)sizeof(OffT)>(p, win, std::make_integer_sequence<int, LBW>{});
          }
          have_window = false;
          OffT all = win[0], any = win[0];
#pragma unroll
          for (int j = 1; j < LBW; ++j) {
            all &= win[j];
            any |= win[j];
          }
          if (all & FLAG_PARTIAL) {
            OffT acc = 0;
#pragma unroll
            for (int j = LBW - 1; j >= 0; --j) {
              const OffT stop = (win[j] & FLAG_INCLUSIVE) ? ~OffT(0) : OffT(0);
              acc = (win[j] & VALUE_MASK) + (acc & ~stop);
            }
            excl += acc;
            if (TRACE_LB) n_walked += LBW;
            if (TRACE_LB && n_trips == 1) t_first = clock64();
            done = (any & FLAG_INCLUSIVE) != 0;
            handled = true;
          }
        }
        if (!handled) {
#pragma unroll
          for (int j = 0; j < LBW; ++j) win[j] = (left > (unsigned long long)j) ? ld_status(p - j * RADIX) : FLAG_INCLUSIVE;
#pragma unroll
          for (int j = 0; j < LBW; ++j) {
            if (!done) {
              OffT v = win[j];
              while ((v & (FLAG_INCLUSIVE | FLAG_PARTIAL)) == 0) {
                v = ld_status(p - j * RADIX);
                if (TRACE_LB) ++n_spins;
              }
              if (TRACE_LB) ++n_walked;
              if (TRACE_LB && n_trips == 1 && j == 0) t_first = clock64();
              excl += v & VALUE_MASK;
              if (v & FLAG_INCLUSIVE) done = true;
            }
          }
        }
        if (done) break;
        p -= LBW * RADIX;
        left -= LBW;
      }
      st_status(status + tid, FLAG_INCLUSIVE | FLAG_PARTIAL | (excl + (OffT)total));
      if (TRACE_LB && tid == 0) {
        unsigned long long* tr = reinterpret_cast<unsigned long long*>(smem + L::OFF_MISC + 128);
        tr[12] = n_trips;
        tr[13] = n_spins;
        tr[14] = n_walked;
        tr[11] = (unsigned long long)t_first;  // overwrites the entry stamp: first status word in hand
      }
    }
    s_goff[tid] = reinterpret_cast<const OffT*>(P.bins)[tid] + excl - s_goff[tid];
  }
  B2S_TRACE(6);  // look-back of digit 0 done
  if (HAS_VALUES && VAL_LATE) load_values();
  if (HAS_VALUES) {
    __syncthreads();  // S3b: every staged value is in a register
    B2S_TRACE(7);  // everybody's look-back done
    ValU* sv = reinterpret_cast<ValU*>(stage_v);
#pragma unroll
    for (int u = 0; u < IPT; ++u) sv[rk[u]] = val[u];
  }
  __syncthreads();  // S4
  B2S_TRACE(8);  // values reordered

  unsigned int next_tile = 0;
  if (PERSIST == 2 && tid == 0) next_tile = atomicAdd(P.tile_counter, 1u);

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
        if ((unsigned long long)dst >= P.peer_capacity) return;
        okeys = reinterpret_cast<KeyU*>(P.peer_keys[d & (MAX_PEERS - 1)]);
        ovals = reinterpret_cast<ValU*>(P.peer_vals[d & (MAX_PEERS - 1)]);
      }
      if (ABL & 1) {
        if (dst == (OffT)0xfffffff1u) okeys[0] = k + (KeyU)(HAS_VALUES ? *reinterpret_cast<const unsigned char*>(&sv[pos]) : 0);
        return;
      }
      okeys[(ABL & 4) ? dst % (OffT)P.n : dst] = k;
      if (HAS_VALUES) ovals[(ABL & 4) ? dst % (OffT)P.n : dst] = sv[pos];
    };
    if (ABL & 2) return;
    if constexpr (WIDE) {
      // Splitter pass: at most MAX_PEERS runs of ~TILE / (ranks) items each.  Per run: the items up to the first
      // 16-byte boundary of the destination and after the last one go out singly, the middle as 16-byte stores (one warp
      // instruction covers 512 contiguous bytes of the destination -- full-size NVLink packets for the remote runs).
      // In the classic flow warp 0's counter row still holds every digit's first position in the sorted tile.
#pragma unroll 1
      for (int d = 0; d < MAX_PEERS; ++d) {
        const int s0 = (int)whist[d];
        int s1 = d + 1 < MAX_PEERS ? (int)whist[d + 1] : TILE;
        if (s1 > valid) s1 = valid;
        if (s1 <= s0) continue;
        const OffT g0 = s_goff[d] + (OffT)s0;
        long long len = s1 - s0;
        if (PEER) {
          if ((unsigned long long)g0 >= P.peer_capacity) continue;
          const unsigned long long room = P.peer_capacity - (unsigned long long)g0;
          if ((unsigned long long)len > room) len = (long long)room;
          okeys = reinterpret_cast<KeyU*>(P.peer_keys[d]);
          ovals = reinterpret_cast<ValU*>(P.peer_vals[d]);
        }
        copy_run_wide(okeys + g0, sk + s0, (int)len, tid, NT);
        if constexpr (HAS_VALUES) copy_run_wide(ovals + g0, sv + s0, (int)len, tid, NT);
      }
    } else if (full) {
#pragma unroll
      for (int u = 0; u < IPT; ++u) emit(u * NT + tid);
    } else {
#pragma unroll 1
      for (int pos = tid; pos < valid; pos += NT) emit(pos);
    }
  }
  B2S_TRACE(9);  // thread 0's stores issued
  if (TRACE && P.trace && threadIdx.x == 0 && !PERSIST) {
#pragma unroll
    for (int i = 0; i < 16; ++i) P.trace[tile * 16 + i] = reinterpret_cast<unsigned long long*>(smem + L::OFF_MISC + 128)[i];
  }
  if (!PERSIST) break;
  // ---- next tile of a persistent CTA: claim, clear the warp counters (unused since P3); the barrier also orders the
  // write-out's reads of the staging buffers before the next TMA copies into them
  if (tid == 0) *s_tile = PERSIST == 2 ? next_tile : atomicAdd(P.tile_counter, 1u);
#pragma unroll
  for (int i = tid; i < NW * RADIX; i += NT) whist[i] = 0;
  phase ^= bulk_flag;
  __syncthreads();
  B2S_TRACE(10);  // everybody's stores issued, next tile claimed
  if (TRACE && P.trace && threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) P.trace[tile * 16 + i] = reinterpret_cast<unsigned long long*>(smem + L::OFF_MISC + 128)[i];
  }
  B2S_TRACE(11);
  }  // tile loop
#undef B2S_TRACE
}

}  // namespace b2s
