// b2s_pass.cuh -- PRODUCTION digit pass of the LSD sort: stable partition of all n items by one 8-bit digit into
// their global positions, chained-scan ("onesweep") style.  This file holds the product kernel only; the round-1
// laboratory kernel with its ablation / trace / persistent branches lives in b2s_onesweep.cuh and is compiled into
// the tuning library alone.
//
// Replaces (reference, for parity of RESULT only):
//   DeviceRadixSortOnesweepKernel  cub/device/dispatch/dispatch_radix_sort.cuh:580
//   AgentRadixSortOnesweep         cub/agent/agent_radix_sort_onesweep.cuh:98-688
//   BlockRadixRankMatchEarlyCounts cub/block/block_radix_rank.cuh:898-1192
//
// Flow of one CTA == one tile (NT threads x IPT items, warp-striped rows of 32 consecutive items):
//   P0  one thread takes the tile id (block index, or a ticket when PF_CLAIM), arms two mbarriers and issues the TMA
//       copies (cp.async.bulk, SASS UBLKCP) keys / values -> shared memory, plus an L2 prefetch ~one CTA lifetime ahead.
//   P1  keys -> registers; counting sweep: one shared-memory reduction per item on the warp's private 256 counters.
//   P2  256 threads: tile digit counts -> PARTIAL status word (published BEFORE the ranking, "early counts"), clear the
//       next pass' status row, 256-wide scan; the warp counters are overwritten with absolute positions in the sorted tile.
//   P3  ranking fused with the reorder: 8 ballot rounds (complements on the FMA pipe), ONE leader atomic per (row, digit)
//       whose return value is the slot of the first peer; the item is stored to its slot at once.
//         split flow (default): key now, value after the look-back through the remembered slot (any K, V);
//         PF_PAIR (4-byte key + 4-byte value): (key, value) as one 64-bit store, no slot array, two barriers fewer.
//   LB  decoupled look-back over a window of LBW predecessor status words (branch-free when all are published).
//   P4  write-out of the sorted tile:
//         default: every thread walks the sorted tile with stride NT -> coalesced stores of digit runs;
//         PF_TMAW: the look-back runs BEFORE the ranking, every digit run is laid out in shared memory with the
//         16-byte phase of its global destination, and its 16-byte-aligned interior leaves the SM as ONE bulk
//         shared->global copy (TMA, SASS UBLKCP.G.S); only the <16-byte ends are stored by threads.  The LSU pipe --
//         the measured bound of this kernel -- no longer carries the write-out.
//   PF_PEER: "digit" d is a destination rank and its run is written to peer_keys[d] / peer_vals[d] (multi-GPU exchange
//         fused into the partition pass).
//
// Stability: items are ranked in tile order (rows in program order, lanes in order inside a row); tiles are ordered by
// their id == position in the input.
#pragma once
#include "b2s_common.cuh"

namespace b2s {

constexpr int MAX_PEERS = 8;

enum : int {
  PF_PAIR = 1,   // (key, value) scattered as one 64-bit shared-memory store (4-byte keys with 4-byte values only)
  PF_TMAW = 2,   // write-out by bulk shared->global copies (keys / values of 4 or 8 bytes)
  PF_CLAIM = 4,  // tile id = atomic ticket instead of the block index (no reliance on in-order CTA dispatch)
  PF_PEER = 8,   // per-destination output bases (multi-GPU partition pass)
  PF_NOBR = 16,  // ranking atomic issued by every lane (non-leaders add to a private scratch word): no branch around it
};

template <int KBYTES, typename OpT>
struct OnesweepParams {
  const void* keys_in;
  void* keys_out;
  const void* vals_in;
  void* vals_out;
  void* status;        // OffT[num_tiles][256], zero on entry
  void* status_next;   // OffT[num_tiles][256] cleared here for the next pass (may be null)
  const void* bins;    // OffT[256] exclusive digit offsets of this pass
  unsigned int* tile_counter;
  unsigned long long n;
  unsigned long long pad_key;  // raw key whose bit-ordered form is all ones
  unsigned long long* trace;   // tuning builds only (b2s_onesweep.cuh); unused here
  const unsigned int* skip_flag;  // DEVICE (may be null): non-zero = one digit holds every key of this pass -> the pass is a copy
  unsigned int ones;           // 0xffffffff, as a launch parameter so that the compiler cannot fold it (see agree_bit)
  OpT op;              // key -> digit of this pass (DigitOp), or key -> destination rank (SplitterOp)
  void* peer_keys[MAX_PEERS];
  void* peer_vals[MAX_PEERS];
  unsigned long long peer_capacity;  // items per receive buffer: stores at or beyond it are dropped
  // zero recording (ImageFloatOp, first pass of a full-range sort of floating keys; null otherwise): one word per row of 32
  // input keys -- bit l of zero_z[r] = key 32 r + l is +-0.0, bit l of zero_s[r] = its sign bit (b2s_fzero.cu restores them)
  unsigned int* zero_z;
  unsigned int* zero_s;
};

// Shared-memory plan.  TMAW pads every digit run to the 16-byte phase of its destination: A = items per 16 bytes of the
// narrower array; a run of c items occupies a slot of roundup(c + A - 1, A) items.
template <int KBYTES, int VBYTES, int NT, int IPT, bool TMAW>
struct PassSmem {
  static constexpr int TILE = NT * IPT;
  static constexpr int NW = NT / 32;
  static constexpr int MINB = VBYTES && VBYTES < KBYTES ? VBYTES : KBYTES;
  static constexpr int A = TMAW ? 16 / MINB : 1;
  static constexpr int SLOTS = TILE + (TMAW ? RADIX * (2 * A - 2) : 0);
  static constexpr int KEY_BYTES = SLOTS * KBYTES + 16;
  static constexpr int VAL_BYTES = VBYTES ? SLOTS * VBYTES + 16 : 0;
  static constexpr int OFF_KEYS = 0;
  static constexpr int OFF_VALS = (KEY_BYTES + 127) / 128 * 128;
  static constexpr int OFF_WHIST = OFF_VALS + (VAL_BYTES + 127) / 128 * 128;
  static constexpr int OFF_GOFF = OFF_WHIST + NW * RADIX * 4;   // OffT[256] (8 bytes reserved each)
  static constexpr int OFF_RUN = OFF_GOFF + RADIX * 8;          // TMAW: uint32[256] = slot start | length << 16
  static constexpr int OFF_MISC = OFF_RUN + (TMAW ? RADIX * 4 : 0);
  static constexpr int OFF_DUMMY = OFF_MISC + 128;              // PF_NOBR: one scratch word per lane and warp
  static constexpr int TOTAL = OFF_DUMMY + NW * 128;            // barriers, scan partials, tile id
  static_assert(SLOTS < 65536, "tile positions are 16-bit");
};

// Exclusive prefix of tile `tile` for the digit whose status row entry is `row` (== status + tile * RADIX + digit).
// Each round trip reads the next LBW predecessors with independent loads at immediate offsets; an inclusive word carries
// BOTH flag bits, so "all published" is one AND-reduction and the sum up to the nearest inclusive word is a select chain
// from the far end.  Words that are not published yet (rare with early counts) fall back to the per-word loop.
template <typename OffT, int LBW>
__device__ __forceinline__ OffT lookback_exclusive(const OffT* row, unsigned long long tile) {
  constexpr int OBITS = sizeof(OffT) * 8;
  constexpr OffT FLAG_INCLUSIVE = OffT(1) << (OBITS - 1);
  constexpr OffT FLAG_PARTIAL = OffT(1) << (OBITS - 2);
  constexpr OffT VALUE_MASK = FLAG_PARTIAL - 1;
  OffT excl = 0;
  const OffT* p = row - RADIX;  // first entry of the current window
  unsigned long long left = tile;
  bool done = false;
  while (true) {
    OffT win[LBW];
    bool handled = false;
    if (left >= (unsigned long long)LBW) {
      load_status_window<RADIX * (int)sizeof(OffT)>(p, win, std::make_integer_sequence<int, LBW>{});
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
          while ((v & (FLAG_INCLUSIVE | FLAG_PARTIAL)) == 0) v = ld_status(p - j * RADIX);
          excl += v & VALUE_MASK;
          if (v & FLAG_INCLUSIVE) done = true;
        }
      }
    }
    if (done) break;
    p -= LBW * RADIX;
    left -= LBW;
  }
  return excl;
}

// 1-D bulk async copy shared -> global (TMA engine, SASS UBLKCP.G.S); 16-byte aligned src/dst/size.
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// waits until the bulk copies of this thread have READ their shared-memory source (the CTA may then exit)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// One digit run (len items at shared-memory `src`, same 16-byte phase as `dst`) -> global memory by ONE thread:
// the 16-byte-aligned interior as a bulk copy, the ends item by item.
template <typename T>
__device__ __forceinline__ void store_run_bulk(T* dst, const T* src, unsigned int len) {
  constexpr unsigned int A = 16 / sizeof(T);
  unsigned int head = (unsigned int)((16u - (unsigned int)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u) / (unsigned int)sizeof(T);
  if (head > len) head = len;
  const unsigned int mid = (len - head) / A * A;
  if (mid) bulk_s2g(dst + head, src + head, mid * (unsigned int)sizeof(T));
#pragma unroll
  for (unsigned int j = 0; j < A - 1; ++j)
    if (j < head) dst[j] = src[j];
  const unsigned int t0 = head + mid;
#pragma unroll
  for (unsigned int j = 0; j < A - 1; ++j)
    if (t0 + j < len) dst[t0 + j] = src[t0 + j];
}

template <int KBYTES, int VBYTES, typename OpT, typename OffT, int NT, int IPT, int MINB, int LBW, int FLAGS, int PFD = 222>
__global__ void __launch_bounds__(NT, MINB) digit_pass_kernel(const OnesweepParams<KBYTES, OpT> P) {
  constexpr bool HAS_VALUES = VBYTES != 0;
  constexpr bool PAIR = (FLAGS & PF_PAIR) != 0 && KBYTES == 4 && VBYTES == 4;
  constexpr bool TMAW = (FLAGS & PF_TMAW) != 0 && !PAIR && KBYTES >= 4 && (VBYTES == 0 || VBYTES == 4 || VBYTES == 8);
  constexpr bool CLAIM = (FLAGS & PF_CLAIM) != 0;
  constexpr bool PEER = (FLAGS & PF_PEER) != 0;
  constexpr bool NOBR = (FLAGS & PF_NOBR) != 0;
  constexpr bool CONV = OpConverts<OpT>::value;  // floating keys travel as bit-ordered images between passes (OrderedFloatOp)
  static_assert(!(CONV && (TMAW || PEER)), "image-form keys: plain digit passes only");
  using KeyU = typename UIntOf<KBYTES>::type;
  using W = typename WideOf<KBYTES>::type;
  using ValU = typename UIntOf<VBYTES ? VBYTES : 1>::type;
  using L = PassSmem<KBYTES, VBYTES, NT, IPT, TMAW>;
  constexpr int TILE = L::TILE;
  constexpr int NW = L::NW;
  constexpr int OBITS = sizeof(OffT) * 8;
  constexpr OffT FLAG_INCLUSIVE = OffT(1) << (OBITS - 1);
  constexpr OffT FLAG_PARTIAL = OffT(1) << (OBITS - 2);
  static_assert(NT >= RADIX && NT % 32 == 0, "one thread per digit needed");
  static_assert(!PEER || OpT::kMaxDigit < MAX_PEERS, "peer launches map digits to destination ranks");

  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* stage_k = smem + L::OFF_KEYS;
  unsigned char* stage_v = smem + L::OFF_VALS;
  unsigned int* whist = reinterpret_cast<unsigned int*>(smem + L::OFF_WHIST);
  OffT* s_goff = reinterpret_cast<OffT*>(smem + L::OFF_GOFF);
  unsigned int* s_run = reinterpret_cast<unsigned int*>(smem + L::OFF_RUN);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::OFF_MISC);                  // [2]
  unsigned int* s_wtot = reinterpret_cast<unsigned int*>(smem + L::OFF_MISC + 16);  // [8]
  unsigned int* s_tile = reinterpret_cast<unsigned int*>(smem + L::OFF_MISC + 64);
  unsigned int* s_geom = reinterpret_cast<unsigned int*>(smem + L::OFF_MISC + 68);  // tile summary, written once by thread 0

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const unsigned long long num_tiles = (P.n + TILE - 1) / TILE;

  // Everything the copies need, as a function of the tile id (thread 0 evaluates it before the others know the id).
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
    // the TMA path needs 16-byte aligned source windows that stay inside the arrays
    g.kshift = (unsigned int)(g.kaddr & 15);
    g.vshift = HAS_VALUES ? (unsigned int)(g.vaddr & 15) : 0;
    g.kbytes = (g.kshift + TILE * KBYTES + 15u) & ~15u;
    g.vbytes = (g.vshift + TILE * VBYTES + 15u) & ~15u;
    g.bulk = g.full && (tile > 0 || (g.kshift == 0 && g.vshift == 0));
    g.bulk = g.bulk && (g.kshift == 0 || g.remain * KBYTES >= (unsigned long long)g.kbytes - g.kshift) &&
             (g.vshift == 0 || g.remain * VBYTES >= (unsigned long long)g.vbytes - g.vshift);
    return g;
  };

  // ---- P0: tile id, barriers, TMA copies in flight before anything else happens
  if (tid == 0) {
    const unsigned long long t0 = CLAIM ? (unsigned long long)atomicAdd(P.tile_counter, 1u) : (unsigned long long)blockIdx.x;
    if (CLAIM) *s_tile = (unsigned int)t0;
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
    const TileGeom g = geom(t0);
    // what the later phases need to know about the tile, so that nobody re-derives the 64-bit geometry
    *s_geom = (g.bulk ? 1u : 0u) | (g.full ? 2u : 0u) | (g.bulk ? (g.kshift << 8) | (g.vshift << 16) : 0u);
    if (g.bulk) {
      mbar_expect_tx(&bar[0], g.kbytes);
      bulk_g2s(stage_k, reinterpret_cast<const void*>(g.kaddr - g.kshift), g.kbytes, &bar[0]);
      if (HAS_VALUES) {
        mbar_expect_tx(&bar[1], g.vbytes);
        bulk_g2s(stage_v, reinterpret_cast<const void*>(g.vaddr - g.vshift), g.vbytes, &bar[1]);
      }
    }
  }
#pragma unroll
  for (int i = tid; i < NW * RADIX; i += NT) whist[i] = 0;
  __syncthreads();

  // The tile id is re-read in every phase instead of being carried in a register (special register with block-index ids,
  // one shared-memory load with tickets) -- the ranking sweep needs the registers.
  auto tile_now = [&]() -> unsigned long long {
    if (CLAIM) return (unsigned long long)*reinterpret_cast<volatile unsigned int*>(s_tile);
    return (unsigned long long)blockIdx.x;
  };
  // tile summary (bit 0 bulk copies in flight, bit 1 full tile, bits 8-11 / 16-19 byte shift of the staged keys / values):
  // one shared-memory load where it is needed instead of the 64-bit geometry or four live registers
  auto tile_summary = [&]() -> unsigned int { return *reinterpret_cast<volatile unsigned int*>(s_geom); };
  const bool t_bulk = (tile_summary() & 1u) != 0;
  if (PFD && tid == 32) {
    const unsigned long long tile = tile_now();
    const TileGeom g = geom(tile);
    if (tile + PFD + 1 < num_tiles) {
    // ask L2 for a tile that will start about one CTA lifetime from now, so that its TMA copies hit L2
    bulk_prefetch_l2(reinterpret_cast<const void*>((g.kaddr + (unsigned long long)PFD * TILE * KBYTES) & ~(uintptr_t)15),
                     (unsigned int)(TILE * KBYTES) & ~15u);
    if (HAS_VALUES)
      bulk_prefetch_l2(reinterpret_cast<const void*>((g.vaddr + (unsigned long long)PFD * TILE * VBYTES) & ~(uintptr_t)15),
                       (unsigned int)(TILE * VBYTES) & ~15u);
    }
  }
  if (!t_bulk) {
    const TileGeom g = geom(tile_now());
    const int valid = g.full ? TILE : (int)g.remain;
    const KeyU* gkeys = reinterpret_cast<const KeyU*>(g.kaddr);
    const ValU* gvals = reinterpret_cast<const ValU*>(g.vaddr);
    KeyU* sk = reinterpret_cast<KeyU*>(stage_k);
    // a partial tile is padded with a key whose digit is the largest one in every pass: the padding ranks after all
    // real items and is never written out
    KeyU pad = (KeyU)P.pad_key;
    if constexpr (CONV) {
      if (!P.op.raw_in) pad = (KeyU)OpT::ONES;  // keys arrive as images: all ones orders last
    }
    for (int i = tid; i < TILE; i += NT) sk[i] = i < valid ? gkeys[i] : pad;
    if (HAS_VALUES) {
      ValU* sv = reinterpret_cast<ValU*>(stage_v);
      for (int i = tid; i < valid; i += NT) sv[i] = gvals[i];
    }
    __syncthreads();
  }

  // ---- constant-digit pass (flag from the upfront histogram): the stable partition is the identity -- copy the staged
  // tile to the same positions of the output and leave.  (The reference short-circuits single-bin TILES,
  // cub/agent/agent_radix_sort_onesweep.cuh:344-420; here the upfront histogram already knows it for the whole pass.)
  if (P.skip_flag != nullptr && __ldg(P.skip_flag) != 0u) {
    const unsigned long long tile = tile_now();
    const TileGeom g = geom(tile);
    const int valid = g.full ? TILE : (int)g.remain;
    if (tid < RADIX && P.status_next) reinterpret_cast<OffT*>(P.status_next)[tile * RADIX + tid] = 0;
    if (g.bulk) {
      mbar_wait(&bar[0], 0);
      if (HAS_VALUES) mbar_wait(&bar[1], 0);
    }
    const KeyU* sk = reinterpret_cast<const KeyU*>(stage_k + (g.bulk ? g.kshift : 0u));
    KeyU* ok = reinterpret_cast<KeyU*>(P.keys_out) + g.base;
    int conv = 0;  // a copy still converts when the encodings of its input and output differ
    if constexpr (CONV) conv = P.op.raw_in == P.op.raw_out ? 0 : (P.op.raw_in ? 1 : 2);
    if (conv == 0) {
#pragma unroll 4
      for (int i = tid; i < valid; i += NT) ok[i] = sk[i];
    } else {
      if constexpr (CONV) {
#pragma unroll 4
        for (int i = tid; i < valid; i += NT) ok[i] = (KeyU)(conv == 1 ? P.op.to_image((W)sk[i]) : P.op.to_raw((W)sk[i]));
      }
    }
    if (HAS_VALUES) {
      const ValU* sv = reinterpret_cast<const ValU*>(stage_v + (g.bulk ? g.vshift : 0u));
      ValU* ov = reinterpret_cast<ValU*>(P.vals_out) + g.base;
#pragma unroll 4
      for (int i = tid; i < valid; i += NT) ov[i] = sv[i];
    }
    return;
  }

  // ---- P1: keys -> registers (warp-striped rows); counting sweep
  const int warp_base = warp * 32 * IPT;
  W key[IPT];
  {
    const unsigned int gs = tile_summary();
    if (gs & 1u) mbar_wait(&bar[0], 0);
    const KeyU* sk = reinterpret_cast<const KeyU*>(stage_k + ((gs >> 8) & 15u));  // the shift is zero without bulk copies
#pragma unroll
    for (int u = 0; u < IPT; ++u) key[u] = (W)sk[warp_base + u * 32 + lane];
  }
  auto op = P.op;
  op.prepare();  // no-op for DigitOp; loads the device-resident splitters for SplitterOp
  if constexpr (CONV) {
    if (op.raw_in) {
      bool converted = false;
      if constexpr (OpRecordsZeros<OpT>::value) {
        if (P.zero_z != nullptr) {
          // rows of this warp are consecutive rows of the input: lane j keeps the words of row (u & ~31) + j, one coalesced store
          // per 32 rows.  The sign ballot runs only for rows that hold a zero (warp-uniform branch).
          const unsigned long long row0 = tile_now() * (unsigned long long)(TILE / 32) + (unsigned long long)(warp * IPT);
          unsigned int zw = 0, sw = 0;
#pragma unroll
          for (int u = 0; u < IPT; ++u) {
            const W k = key[u];
            const bool z = OpT::is_zero(k);
            const unsigned int bz = __ballot_sync(0xffffffffu, z);
            unsigned int bs = 0;
            if (bz) bs = __ballot_sync(0xffffffffu, (k & OpT::HIGH) != 0);
            if (lane == (u & 31)) {
              zw = bz;
              sw = bs;
            }
            if ((u & 31) == 31 || u == IPT - 1) {
              if (lane <= (u & 31)) {
                P.zero_z[row0 + (u & ~31) + lane] = zw;
                P.zero_s[row0 + (u & ~31) + lane] = sw;
              }
            }
            key[u] = z ? OpT::HIGH : op.to_image_nz(k);
          }
          converted = true;
        }
      }
      if (!converted) {
#pragma unroll
        for (int u = 0; u < IPT; ++u) key[u] = op.to_image(key[u]);
      }
    }
  }
  unsigned int* myhist = whist + warp * RADIX;
  const unsigned int myhist_s = smem_u32(myhist);
  const unsigned int lt = lanemask_lt();
#pragma unroll
  for (int u = 0; u < IPT; ++u) red_shared_add(myhist_s + op(key[u]) * 4, 1u);

  ValU val[HAS_VALUES ? IPT : 1];
  auto load_values = [&]() {
    const unsigned int gs = tile_summary();
    if (gs & 1u) mbar_wait(&bar[1], 0);
    const ValU* sv = reinterpret_cast<const ValU*>(stage_v + ((gs >> 16) & 15u));
#pragma unroll
    for (int u = 0; u < IPT; ++u) val[u] = sv[warp_base + u * 32 + lane];
  };
  if (PAIR) load_values();  // the sorted pairs will overwrite both staging buffers
  __syncthreads();          // S2: all warp histograms complete, all staged keys (PAIR: and values) consumed

  // ---- P2: per-digit tile counts -> partial status; digit prefix; per-warp bases (absolute slots in the sorted tile)
  unsigned int total = 0;
  if (tid < RADIX) {
    const unsigned long long tile = tile_now();
    OffT* status = reinterpret_cast<OffT*>(P.status) + tile * RADIX;
#pragma unroll
    for (int w = 0; w < NW; ++w) total += whist[w * RADIX + tid];
    st_status(status + tid, (tile == 0 ? (FLAG_INCLUSIVE | FLAG_PARTIAL) : FLAG_PARTIAL) | (OffT)total);
    if (P.status_next) reinterpret_cast<OffT*>(P.status_next)[tile * RADIX + tid] = 0;
  }
  // TMAW: a run of c items takes a slot of roundup(c + A - 1, A) items (room for the phase shift), 0 when empty
  const unsigned int slot = TMAW ? (total ? (total + 2 * L::A - 2) / L::A * L::A : 0u) : total;
  unsigned int incl = slot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (tid < RADIX && lane == 31) s_wtot[warp] = incl;
  OffT gstart = 0;  // TMAW: global index of this tile's first item of digit `tid`
  if (TMAW && tid < RADIX) {
    const unsigned long long tile = tile_now();
    OffT* status = reinterpret_cast<OffT*>(P.status) + tile * RADIX;
    OffT excl = 0;
    if (tile > 0) {
      excl = lookback_exclusive<OffT, LBW>(status + tid, tile);
      st_status(status + tid, FLAG_INCLUSIVE | FLAG_PARTIAL | (excl + (OffT)total));
    }
    gstart = reinterpret_cast<const OffT*>(P.bins)[tid] + excl;
  }
  __syncthreads();  // S2b
  if (tid < RADIX) {
    unsigned int base = 0;
#pragma unroll
    for (int w = 0; w < RADIX / 32; ++w)
      if (w < warp) base += s_wtot[w];
    // the per-warp counts are read a second time rather than kept in NW registers across the barrier
    unsigned int run = base + incl - slot;
    if (TMAW) {
      // phase shift: the run starts at the same offset inside a 16-byte unit as its global destination.  Output
      // pointers of a TMAW launch are 16-byte aligned (the host checks), so the phase is a function of the index.
      run += (unsigned int)gstart & (unsigned int)(L::A - 1);
      unsigned int len = total;
      const TileGeom g = geom(tile_now());
      const int valid = g.full ? TILE : (int)g.remain;
      if (!g.full && tid == (int)op((W)(KeyU)P.pad_key)) len -= (unsigned int)(TILE - valid);  // padding is never written
      s_run[tid] = run | (len << 16);
      s_goff[tid] = gstart;
    } else {
      s_goff[tid] = (OffT)run;  // parked until the look-back needs it
    }
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const unsigned int c = whist[w * RADIX + tid];
      whist[w * RADIX + tid] = run;
      run += c;
    }
  }
  __syncthreads();  // S3: per-warp bases ready

  // ---- P3: ranking sweep on counters that hold absolute slots: the leader's atomic returns the slot of the first
  // peer, so an item goes to its sorted slot as soon as its row is ranked.  Software pipeline over rows (a warp issues
  // in order): leader atomic of row u -> ballots of row u+1 -> store of row u-1 -> SHFL broadcast of row u's atomic.
  unsigned int rk[(PAIR || !HAS_VALUES) ? 1 : IPT];
  {
    KeyU* sk = reinterpret_cast<KeyU*>(stage_k);
    auto place = [&](int u, unsigned int r) {
      if constexpr (PAIR) {
        reinterpret_cast<uint2*>(stage_k)[r] = make_uint2((unsigned int)key[u], (unsigned int)val[u]);
      } else {
        if constexpr (HAS_VALUES) rk[u] = r;
        sk[r] = (KeyU)key[u];
      }
    };
    unsigned int d = op(key[0]);
    unsigned int m = match_ballot<RADIX_BITS, true>(d, P.ones);
    unsigned int bcast_prev = 0, below_prev = 0;
#pragma unroll
    for (int u = 0; u < IPT; ++u) {
      const unsigned int leader = bfind(m);
      const unsigned int below = __popc(m & lt);
      unsigned int raw;
      if constexpr (NOBR) {
        const unsigned int scratch = smem_u32(smem + L::OFF_DUMMY) + (unsigned int)tid * 4u;
        raw = atoms_add(lane == leader ? myhist_s + d * 4 : scratch, (unsigned int)__popc(m));
      } else {
        raw = atoms_add_if(lane == leader, myhist_s + d * 4, (unsigned int)__popc(m));
      }
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

  // ---- look-back (unless it already ran before the ranking): exclusive prefix of this tile for digit `tid`
  if (!TMAW && tid < RADIX) {
    const unsigned long long tile = tile_now();
    OffT* status = reinterpret_cast<OffT*>(P.status) + tile * RADIX;
    OffT excl = 0;
    if (tile > 0) {
      excl = lookback_exclusive<OffT, LBW>(status + tid, tile);
      st_status(status + tid, FLAG_INCLUSIVE | FLAG_PARTIAL | (excl + (OffT)total));
    }
    s_goff[tid] = reinterpret_cast<const OffT*>(P.bins)[tid] + excl - s_goff[tid];
  }
  if (HAS_VALUES && !PAIR) {
    load_values();    // staged values -> registers (re-using the key registers)
    __syncthreads();  // S3b: every staged value is in a register
    ValU* sv = reinterpret_cast<ValU*>(stage_v);
#pragma unroll
    for (int u = 0; u < IPT; ++u) sv[rk[u]] = val[u];
  }
  if (TMAW) fence_proxy_async();  // the sorted tile was written by threads and is about to be read by the TMA engine
  __syncthreads();                // S4

  // ---- P4: write-out of digit runs
  KeyU* okeys = reinterpret_cast<KeyU*>(P.keys_out);
  ValU* ovals = reinterpret_cast<ValU*>(P.vals_out);
  if constexpr (TMAW) {
    if (tid < RADIX) {
      const unsigned int rw = s_run[tid];
      unsigned int len = rw >> 16;
      const unsigned int s0 = rw & 0xffffu;
      const OffT g0 = s_goff[tid];
      if (PEER) {
        okeys = reinterpret_cast<KeyU*>(P.peer_keys[tid & (MAX_PEERS - 1)]);
        ovals = reinterpret_cast<ValU*>(P.peer_vals[tid & (MAX_PEERS - 1)]);
        if ((unsigned long long)g0 >= P.peer_capacity) len = 0;
        else if ((unsigned long long)len > P.peer_capacity - (unsigned long long)g0) len = (unsigned int)(P.peer_capacity - (unsigned long long)g0);
      }
      if (len) {
        store_run_bulk(okeys + g0, reinterpret_cast<const KeyU*>(stage_k) + s0, len);
        if constexpr (HAS_VALUES) store_run_bulk(ovals + g0, reinterpret_cast<const ValU*>(stage_v) + s0, len);
      }
      bulk_commit();
      bulk_wait_read();  // shared memory must stay alive until the engine has read it
    }
  } else {
    const KeyU* sk = reinterpret_cast<const KeyU*>(stage_k);
    const ValU* sv = reinterpret_cast<const ValU*>(stage_v);
    auto emit = [&](int pos, auto to_raw) {
      KeyU k;
      ValU v{};
      if constexpr (PAIR) {
        const uint2 kv = reinterpret_cast<const uint2*>(stage_k)[pos];
        k = (KeyU)kv.x;
        v = (ValU)kv.y;
      } else {
        k = sk[pos];
        if constexpr (HAS_VALUES) v = sv[pos];
      }
      const unsigned int d = op((W)k);
      const OffT dst = s_goff[d] + (OffT)pos;
      if (PEER) {
        if ((unsigned long long)dst >= P.peer_capacity) return;
        okeys = reinterpret_cast<KeyU*>(P.peer_keys[d & (MAX_PEERS - 1)]);
        ovals = reinterpret_cast<ValU*>(P.peer_vals[d & (MAX_PEERS - 1)]);
      }
      if constexpr (CONV && decltype(to_raw)::value) k = (KeyU)image_to_raw(op, (W)k);
      okeys[dst] = k;
      if constexpr (HAS_VALUES) ovals[dst] = v;
    };
    auto emit_all = [&](auto to_raw) {
      if (tile_summary() & 2u) {
#pragma unroll
        for (int u = 0; u < IPT; ++u) emit(u * NT + tid, to_raw);
      } else {
        const int valid = (int)geom(tile_now()).remain;
#pragma unroll 1
        for (int pos = tid; pos < valid; pos += NT) emit(pos, to_raw);
      }
    };
    if constexpr (CONV) {
      if (op.raw_out) emit_all(std::true_type{});
      else emit_all(std::false_type{});
    } else {
      emit_all(std::false_type{});
    }
  }
}

}  // namespace b2s
