// b2s_kernels.cu -- kernel instantiations + launchers for ONE key width (-DB2S_K=1|2|4|8).
// Compiled four times so the instantiations build in parallel.
#include <cstdlib>
#include <mutex>
#include <set>
#include <type_traits>
#include <utility>

#include "b2s_histogram.cuh"
#include "b2s_internal.h"
#include "b2s_pass.cuh"
#ifdef B2S_TUNING
#include "b2s_onesweep.cuh"  // laboratory kernel: tuning library only
#endif
#include "b2s_segmented.cuh"
#include "b2s_single_tile.cuh"
#include "b2s_split.cuh"

#ifndef B2S_K
#error "compile with -DB2S_K=<key bytes>"
#endif

namespace b2s {
namespace {

constexpr int K = B2S_K;

// ---- tuning table -------------------------------------------------------------------------
// Variant 0 is the production tuning for (K, V).  A tuning build (-DB2S_TUNING) adds more points that bench/tune.py
// sweeps on the GPU: 1..29 = other shapes / flows of the production kernel (b2s_pass.cuh), 30.. = the round-1
// laboratory kernel (b2s_onesweep.cuh).  Table entries give items/thread for 4-byte keys with <=4-byte values; wider
// items scale it down by bytes (shared memory) and by registers.
#ifdef B2S_TUNING
constexpr int NUM_VARIANTS = 58;
#else
constexpr int NUM_VARIANTS = 1;
#endif

template <int V>
constexpr int scale_ipt(int ipt) {
  const int kw = K > 4 ? 2 : 1, vw = (V + 3) / 4;
  const int regs_per_item = (vw > kw ? vw : kw) + 1;  // key words (values re-use them) + slot
  const int by_bytes = (K + V <= 8) ? ipt : ipt * 8 / (K + V);
  const int by_regs = ipt * 2 / regs_per_item;
  const int r = by_bytes < by_regs ? by_bytes : by_regs;
  return r < 4 ? 4 : r;
}

// Fewer items per thread when the TMA write-out pads every digit run (b2s_pass.cuh PassSmem): SLOTS * (K + V) bytes.
template <int V>
constexpr int tmaw_ipt(int nt, int minb, int want) {
  int ipt = want;
  while (ipt > 4) {
    const int a = 16 / ((V && V < K) ? V : K);
    const long slots = (long)nt * ipt + 256 * (2 * a - 2);
    const long bytes = slots * (K + V) + 256 + (nt / 32) * 1024 + 2048 + 1024 + 128 + 1024;
    if (bytes * minb <= 232448 && slots < 65536) break;
    --ipt;
  }
  return ipt;
}

template <int V, bool F = false, bool OFF64 = false>
constexpr Variant variant_cfg(int vi) {
  // {threads, items/thread, min CTAs/SM, look-back window, -, lab mode, flow}.  flow >= 0: production kernel with these
  // PF_* flags; flow < 0: laboratory kernel with `mode` (tuning builds).  Shapes from the B200 sweeps in
  // profiles/r1_tune_sweep_*.jsonl and profiles/r2_*.jsonl.  A variant that spills loses 15-25 %.  Pairs want few threads
  // with many items each (2 CTAs x 256 threads: no warp idles while 256 threads scan and look back, larger tiles);
  // keys alone want more warps (3 CTAs x 320).  Floating keys use the integer shapes (OrderedFloatOp).
  const bool small_pairs = V > 0 && K + V <= 8;
  const bool pair44 = K == 4 && V == 4;
  // (round-2 sweeps: profiles/r2_tune_shapes_*.jsonl, r2_tune_nobr.jsonl)
  // keys alone: 256 threads x 60 items x 3 CTAs/SM with the branch-free ranking atomic (8-byte keys: 256 x 48 x 2): every
  // warp scans and looks back, 15360-key tiles; +11-12 % over 320 x 30 x 3 / 384 x 26 x 3 (profiles/r2_tune_shapes_keys_256thr.jsonl)
  const Variant d = V == 0                  ? (F ? Variant{384, scale_ipt<V>(K <= 4 ? 26 : 24), 3, 12, 0, 0, 0}  // floating keys: measured best (spills at 256 x 44+)
                                                 : K == 8 ? Variant{256, 48 - (OFF64 ? 4 : 0), 2, 12, 0, 0, PF_NOBR}
                                                          : Variant{256, 60 - (OFF64 ? 4 : 0), 3, 8, 0, 0, PF_NOBR})
                    : (K + V <= 6 && V >= 2) ? (F ? Variant{384, scale_ipt<V>(24), 3, 12, 0, 0, 0}
                                                  : Variant{320, scale_ipt<V>(30) - (OFF64 ? 4 : 0), 3, 12, 0, 0, PF_NOBR})
                    : pair44                ? Variant{256, (F ? 40 : 48) - (OFF64 ? 4 : 0), 2, 8, 0, 0, PF_PAIR | PF_NOBR}  // (key, value) as one 64-bit store
                    : small_pairs          ? Variant{512, scale_ipt<V>(22), 2, 12, 0, 0, 0}
                    : (K + V >= 12)         ? Variant{256, scale_ipt<V>(48) - (OFF64 ? 2 : 0), 2, 8, 0, 0, PF_NOBR}  // wide pairs: few threads, many items each
                                            : Variant{384, scale_ipt<V>(20), 3, 12, 0, 0, 0};
#ifdef B2S_TUNING
  constexpr int M = 8 | 32 | 64 | 128 | 256 | (222 << 16);  // lab kernel: the round-1 production flow
  switch (vi) {
    case 0: return d;
    // production kernel, split flow at the round-1 production shapes (A/B against the lab kernel, variant 30)
    case 1: return small_pairs && !F ? Variant{512, scale_ipt<V>(22), 2, 12, 0, 0, 0} : Variant{d.nt, d.ipt, d.minb, d.lbw, 0, 0, 0};
    case 2: return Variant{d.nt, d.ipt, d.minb, d.lbw, 0, 0, d.flow | PF_CLAIM};  // ticketed tile ids
    // pair flow shapes
    case 3: return Variant{512, scale_ipt<V>(20), 2, 12, 0, 0, PF_PAIR};
    case 4: return Variant{448, scale_ipt<V>(22), 2, 12, 0, 0, PF_PAIR};
    case 5: return Variant{384, scale_ipt<V>(26), 2, 12, 0, 0, PF_PAIR};
    case 6: return Variant{448, scale_ipt<V>(24), 2, 8, 0, 0, PF_PAIR};
    // TMA write-out (bulk shared->global copies per digit run)
    case 7: return Variant{512, tmaw_ipt<V>(512, 2, scale_ipt<V>(22)), 2, 12, 0, 0, PF_TMAW};
    case 8: return Variant{384, tmaw_ipt<V>(384, 3, scale_ipt<V>(20)), 3, 12, 0, 0, PF_TMAW};
    case 9: return Variant{448, tmaw_ipt<V>(448, 2, scale_ipt<V>(24)), 2, 12, 0, 0, PF_TMAW};
    case 10: return Variant{512, tmaw_ipt<V>(512, 2, scale_ipt<V>(22)), 2, 6, 0, 0, PF_TMAW};
    case 11: return Variant{1024, tmaw_ipt<V>(1024, 1, scale_ipt<V>(22)), 1, 12, 0, 0, PF_TMAW};
    case 12: return Variant{256, tmaw_ipt<V>(256, 4, scale_ipt<V>(22)), 4, 12, 0, 0, PF_TMAW};
    // keys alone / other shapes of the split flow
    case 13: return Variant{384, scale_ipt<V>(24), 3, 12, 0, 0, 0};
    case 14: return Variant{384, scale_ipt<V>(26), 3, 12, 0, 0, 0};
    case 15: return Variant{384, scale_ipt<V>(22), 3, 12, 0, 0, 0};
    case 16: return Variant{512, scale_ipt<V>(22), 2, 12, 0, 0, 0};
    case 17: return Variant{256, scale_ipt<V>(26), 4, 12, 0, 0, 0};
    // more CTAs per SM with the same per-thread shape (overlap of the latency phases: TMA wait, digit scan, look-back)
    case 18: return Variant{d.nt, d.ipt, d.minb, d.lbw, 0, 0, d.flow | PF_NOBR};
    case 19: return Variant{320, scale_ipt<V>(30), 3, 12, 0, 0, PF_NOBR};
    case 20: return Variant{256, scale_ipt<V>(44), 2, 8, 0, 0, PF_PAIR | PF_NOBR};
    case 21: return Variant{256, scale_ipt<V>(28), 3, 8, 0, 0, PF_PAIR};
    case 22: return Variant{256, scale_ipt<V>(48), 2, 8, 0, 0, PF_PAIR | PF_NOBR};
    case 23: return Variant{256, scale_ipt<V>(46), 2, 8, 0, 0, PF_PAIR | PF_NOBR};
    case 24: return Variant{256, scale_ipt<V>(44), 2, 8, 0, 0, PF_PAIR | PF_NOBR};
    case 25: return Variant{384, scale_ipt<V>(28), 2, 8, 0, 0, PF_PAIR | PF_NOBR};
    case 26: return Variant{256, scale_ipt<V>(28), 3, 8, 0, 0, PF_PAIR | PF_NOBR};
    case 27: return Variant{320, scale_ipt<V>(30), 3, 12, 0, 0, 0};
    case 28: return Variant{384, scale_ipt<V>(28), 3, 12, 0, 0, 0};
    case 29: return Variant{512, scale_ipt<V>(28), 2, 12, 0, 0, 0};
    // laboratory kernel (round 1): production flow, classic flow, the ladder, traces
    case 30: return small_pairs && !F ? Variant{512, scale_ipt<V>(22), 2, 12, 0, M, -1}
                                      : Variant{384, scale_ipt<V>(V == 0 ? (F ? 22 : 24) : (K + V <= 6 && V >= 2) ? (F ? 22 : 24) : (F ? 18 : 20)), 3, 12, 0, M, -1};
    case 31: return Variant{512, scale_ipt<V>(20), 2, 4, 0, 0, -1};                                   // classic flow
    case 32: return Variant{384, scale_ipt<V>(19), 3, 4, 0, 0, -1};
    case 33: return Variant{384, scale_ipt<V>(19), 3, 12, 0, 8 | 64 | 128 | 256 | (222 << 16), -1};   // early counts, claimed ids
    case 34: return Variant{512, scale_ipt<V>(22), 2, 12, 0, 16 | M, -1};                             // phase time stamps
    case 35: return Variant{512, scale_ipt<V>(22), 2, 12, 0, 16 | 512 | M, -1};                       // + look-back statistics
    // timing-only ablations of the classic flow (wrong results by design): see ABL in b2s_onesweep.cuh
    case 36: return Variant{512, scale_ipt<V>(20), 2, 4, 1, 0, -1};
    case 37: return Variant{512, scale_ipt<V>(20), 2, 4, 2, 0, -1};
    case 38: return Variant{512, scale_ipt<V>(20), 2, 4, 4, 0, -1};
    case 39: return Variant{512, scale_ipt<V>(20), 2, 4, 8, 0, -1};
    // round 2, second shape sweep around 256 x 28 x 3 (pairs) -- fewer threads, more items per thread
    case 40: return Variant{256, scale_ipt<V>(42), 2, 8, 0, 0, PF_NOBR};
    case 41: return Variant{256, scale_ipt<V>(45), 2, 8, 0, 0, PF_NOBR};
    case 42: return Variant{256, scale_ipt<V>(47), 2, 8, 0, 0, PF_NOBR};
    case 43: return Variant{256, scale_ipt<V>(48), 2, 8, 0, 0, PF_NOBR};
    case 44: return Variant{256, scale_ipt<V>(44), 2, 12, 0, 0, PF_NOBR};
    case 45: return Variant{256, scale_ipt<V>(46), 2, 8, 0, 0, PF_NOBR};
    case 46: return Variant{256, scale_ipt<V>(36), 3, 12, 0, 0, PF_NOBR};
    case 47: return Variant{256, scale_ipt<V>(44), 2, 8, 148, 0, PF_PAIR};
    case 50: return Variant{256, scale_ipt<V>(44), 2, 8, 296, 0, PF_PAIR};
    case 51: return Variant{256, scale_ipt<V>(56), 3, 12, 0, 0, PF_NOBR};
    case 48: return Variant{256, scale_ipt<V>(48), 3, 12, 0, 0, PF_NOBR};
    case 49: return Variant{256, scale_ipt<V>(60), 2, 12, 0, 0, PF_NOBR};
    // wave balance at mid sizes (keys alone, 3 CTAs/SM: 444 slots): tile sizes that fill whole waves at 2^24 keys
    case 52: return Variant{256, scale_ipt<V>(50), 3, 8, 0, 0, PF_NOBR};
    case 53: return Variant{256, scale_ipt<V>(52), 3, 8, 0, 0, PF_NOBR};
    case 54: return Variant{256, scale_ipt<V>(38), 3, 8, 0, 0, PF_NOBR};
    case 55: return Variant{256, scale_ipt<V>(40), 3, 8, 0, 0, PF_NOBR};
    case 56: return Variant{256, scale_ipt<V>(30), 3, 8, 0, 0, PF_NOBR};
    case 57: return Variant{256, scale_ipt<V>(44), 3, 8, 0, 0, PF_NOBR};
    default: return d;
  }
#else
  (void)vi;
  return d;
#endif
}

template <bool F>
DigitOp<K, F> make_op(const DigitConsts& dc, int bit, int nbits) {
  using W = typename WideOf<K>::type;
  DigitOp<K, F> op;
  op.xor_mask = (W)dc.xor_mask;
  op.zero_img = (W)dc.zero_img;
  op.bit = (uint32_t)bit;
  op.mask = nbits >= 32 ? 0xffffffffu : (1u << nbits) - 1u;
  op.xor_digit = (uint32_t)((W)dc.xor_mask >> bit) & op.mask;
  return op;
}

// Functor of the multi-pass sort's digit passes.  FM 0: integer keys (DigitOp); 1: floating keys travelling between passes as
// their bit-ordered image, zeros collapsed in every digit extraction (OrderedFloatOp); 2: first / last pass of a full-range sort
// of 4- / 8-byte floating keys with zero recording (ImageFloatOp) -- the passes in between run as FM 0.
template <int FM>
using PassOp = std::conditional_t<FM == 2, ImageFloatOp<K>, std::conditional_t<FM == 1, OrderedFloatOp<K>, DigitOp<K, false>>>;
template <int FM>
PassOp<FM> make_pass_op(const PassArgs& a) {
  if constexpr (FM != 0) {
    using W = typename WideOf<K>::type;
    PassOp<FM> op;
    op.xor_mask = (W)a.dc.xor_mask;
    op.bit = (uint32_t)a.bit;
    op.mask = a.nbits >= 32 ? 0xffffffffu : (1u << a.nbits) - 1u;
    op.raw_in = a.raw_in ? 1 : 0;
    op.raw_out = a.raw_out ? 1 : 0;
    return op;
  } else {
    return make_op<false>(a.dc, a.bit, a.nbits);
  }
}

// Opt a kernel in to its dynamic shared-memory size once per (kernel, device); thread-safe.
template <typename KernT>
cudaError_t ensure_smem(KernT kern, int bytes) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> done;
  int dev = 0;
  cudaGetDevice(&dev);
  const std::pair<const void*, int> key(reinterpret_cast<const void*>(kern), dev);
  std::lock_guard<std::mutex> lock(mu);
  if (done.count(key)) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done.insert(key);
  return e;
}

// CTAs of a persistent launch: SMs of the current device x CTAs per SM.
inline int resident_ctas(int per_sm) {
  static int sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (sms[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    sms[dev] = v > 0 ? v : 148;
  }
  return sms[dev] * per_sm;
}

template <typename OpT>
void fill_params(OnesweepParams<K, OpT>& p, const PassArgs& a, const OpT& op) {
  p.keys_in = a.keys_in;
  p.keys_out = a.keys_out;
  p.vals_in = a.vals_in;
  p.vals_out = a.vals_out;
  p.status = a.status;
  p.status_next = a.status_next;
  p.bins = a.bins;
  p.tile_counter = a.tile_counter;
  p.n = a.n;
  p.pad_key = a.dc.pad_key;
  p.ones = 0xffffffffu;
  p.trace = a.trace;
  p.skip_flag = a.skip_flag;
  p.op = op;
  for (int i = 0; i < MAX_PEERS; ++i) p.peer_keys[i] = p.peer_vals[i] = nullptr;
  p.peer_capacity = ~0ull;
  p.zero_z = a.zero_z;
  p.zero_s = a.zero_s;
}

template <int V, int F, typename OffT, int VI>
cudaError_t launch_one(const PassArgs& a, cudaStream_t s) {
  constexpr Variant c = variant_cfg<V, F == 1, sizeof(OffT) == 8>(VI);  // zero-recording float passes take the integer shapes
  constexpr int TILE = c.nt * c.ipt;
  OnesweepParams<K, PassOp<F>> p;
  fill_params(p, a, make_pass_op<F>(a));
  const unsigned long long tiles = (a.n + TILE - 1) / TILE;
  // 64-bit look-back words cost two registers each: half the window
  constexpr int LBW = (sizeof(OffT) == 8 && c.lbw > 4) ? c.lbw / 2 : c.lbw;
#ifdef B2S_TUNING
  if constexpr (c.flow < 0) {
    using L = OnesweepSmem<K, V, c.nt, c.ipt>;
    auto kern = onesweep_kernel<K, V, DigitOp<K, F == 1>, OffT, c.nt, c.ipt, c.minb, LBW, false, c.abl, c.mode>;
    cudaError_t e = ensure_smem(kern, L::TOTAL);
    if (e != cudaSuccess) return e;
    unsigned long long grid = tiles;
    if ((c.mode & 3) && grid > (unsigned long long)resident_ctas(c.minb)) grid = resident_ctas(c.minb);
    kern<<<(unsigned int)grid, c.nt, L::TOTAL, s>>>(p);
    return cudaGetLastError();
  } else
#endif
  {
    constexpr int FL = c.flow < 0 ? 0 : c.flow;
    constexpr bool TMAW = (FL & PF_TMAW) != 0 && !((FL & PF_PAIR) && K == 4 && V == 4) && K >= 4 && (V == 0 || V == 4 || V == 8);
    using L = PassSmem<K, V, c.nt, c.ipt, TMAW>;
    auto launch = [&](auto kern) {
      cudaError_t e = ensure_smem(kern, L::TOTAL);
      if (e != cudaSuccess) return e;
      kern<<<(unsigned int)tiles, c.nt, L::TOTAL, s>>>(p);
      return cudaGetLastError();
    };
#ifdef B2S_TUNING
    return launch(digit_pass_kernel<K, V, PassOp<F>, OffT, c.nt, c.ipt, c.minb, LBW, FL, (c.abl > 0 ? c.abl : 222)>);  // abl = L2 prefetch distance
#else
    // ticketed tile ids on request (b2s_set_tile_claim / B2S_TILE_CLAIM=1): no reliance on in-order CTA dispatch
    if (a.claim) return launch(digit_pass_kernel<K, V, PassOp<F>, OffT, c.nt, c.ipt, c.minb, LBW, FL | PF_CLAIM>);
    return launch(digit_pass_kernel<K, V, PassOp<F>, OffT, c.nt, c.ipt, c.minb, LBW, FL>);
#endif
  }
}

template <int V, int F, typename OffT, int... VI>
cudaError_t launch_vi(int variant, const PassArgs& a, cudaStream_t s, std::integer_sequence<int, VI...>) {
  cudaError_t r = cudaErrorInvalidValue;
  (void)((variant == VI ? (r = launch_one<V, F, OffT, VI>(a, s), true) : false) || ...);
  return r;
}

template <int V, int F>
cudaError_t launch_off(int variant, const PassArgs& a, cudaStream_t s) {
  using Seq = std::make_integer_sequence<int, NUM_VARIANTS>;
#ifdef B2S_TUNING
  if (a.off64) return cudaErrorInvalidValue;  // tuning builds carry 32-bit offsets only
  return launch_vi<V, F, uint32_t>(variant, a, s, Seq{});
#else
  return a.off64 ? launch_vi<V, F, unsigned long long>(variant, a, s, Seq{})
                 : launch_vi<V, F, uint32_t>(variant, a, s, Seq{});
#endif
}

template <int V>
cudaError_t launch_f(int variant, const PassArgs& a, cudaStream_t s) {
#ifndef B2S_TUNING
  if constexpr (K >= 4 && (V == 0 || V == 4)) {
    if (a.dc.is_float && a.zero_z != nullptr) return launch_off<V, 2>(variant, a, s);  // zero recording (b2s_fzero.cu)
  }
  if constexpr (K >= 2) {
    if (a.dc.is_float) return launch_off<V, 1>(variant, a, s);
  }
#endif
  return launch_off<V, 0>(variant, a, s);
}

template <bool F, typename OffT>
cudaError_t hist_one(const HistArgs& a, cudaStream_t s) {
  HistParams<K, F> p;
  p.keys = a.keys;
  p.n = a.n;
  p.op = make_op<F>(a.dc, 0, 8);
  p.begin_bit = a.begin_bit;
  p.end_bit = a.end_bit;
  p.num_passes = a.num_passes;
  p.ghist = a.ghist;
  p.done = a.done;
  p.flags = a.flags;
  const bool full = a.begin_bit == 0 && a.end_bit == K * 8;
  auto kern = full ? histogram_kernel<K, F, OffT, true> : histogram_kernel<K, F, OffT, false>;
  cudaError_t e = ensure_smem(kern, HistSmem<K>::BYTES);
  if (e != cudaSuccess) return e;
  kern<<<a.grid, HIST_THREADS, HistSmem<K>::BYTES, s>>>(p);
  return cudaGetLastError();
}

// ---- single-tile sort ---------------------------------------------------------------------------------------------
template <int V, bool F>
cudaError_t single_one(const SingleArgs& a, cudaStream_t s) {
  using S = SingleTileShape<K, V>;
  SingleTileParams<K, F> p;
  p.keys_in = a.keys_in;
  p.keys_out = a.keys_out;
  p.vals_in = a.vals_in;
  p.vals_out = a.vals_out;
  p.n = (unsigned int)a.n;
  p.pad_key = a.dc.pad_key;
  p.op = make_op<F>(a.dc, a.begin_bit, 8);
  p.begin_bit = a.begin_bit;
  p.end_bit = a.end_bit;
  p.ones = 0xffffffffu;
  auto kern = single_tile_kernel<K, V, F>;
  cudaError_t e = ensure_smem(kern, S::TOTAL);
  if (e != cudaSuccess) return e;
  kern<<<1, S::NT, S::TOTAL, s>>>(p);
  return cudaGetLastError();
}

template <int V>
cudaError_t single_v(const SingleArgs& a, cudaStream_t s) {
  if constexpr (K >= 2) {
    if (a.dc.is_float) return single_one<V, true>(a, s);
  }
  return single_one<V, false>(a, s);
}

// ---- segmented sort ------------------------------------------------------------------------------------------------
template <int V, bool F, typename SegOffT>
cudaError_t segmented_one(const SegmentedArgs& a, cudaStream_t s) {
  using S = TileSmem<K, V, SEG_NT, SEG_IPT>;
  SegmentedParams<K, F> p;
  p.keys_src = a.keys_src;
  p.keys_a = a.keys_a;
  p.keys_b = a.keys_b;
  p.vals_src = a.vals_src;
  p.vals_a = a.vals_a;
  p.vals_b = a.vals_b;
  p.begin_offsets = a.begin_offsets;
  p.end_offsets = a.end_offsets;
  p.pad_key = a.dc.pad_key;
  p.op = make_op<F>(a.dc, a.begin_bit, 8);
  p.begin_bit = a.begin_bit;
  p.end_bit = a.end_bit;
  p.passes = a.passes;
  p.ones = 0xffffffffu;
  auto kern = segmented_sort_kernel<K, V, F, SegOffT>;
  cudaError_t e = ensure_smem(kern, S::TOTAL);
  if (e != cudaSuccess) return e;
  // one CTA per segment; grids beyond 2^31 - 1 segments are split
  for (uint64_t first = 0; first < a.num_segments; first += 0x7fffffffull) {
    const uint64_t count = a.num_segments - first < 0x7fffffffull ? a.num_segments - first : 0x7fffffffull;
    p.begin_offsets = reinterpret_cast<const SegOffT*>(a.begin_offsets) + first;
    p.end_offsets = reinterpret_cast<const SegOffT*>(a.end_offsets) + first;
    kern<<<(unsigned int)count, SEG_NT, S::TOTAL, s>>>(p);
  }
  return cudaGetLastError();
}

template <int V>
cudaError_t segmented_v(const SegmentedArgs& a, cudaStream_t s) {
  if constexpr (K >= 2) {
    if (a.dc.is_float)
      return a.offset_bytes == 8 ? segmented_one<V, true, long long>(a, s) : segmented_one<V, true, int>(a, s);
  }
  return a.offset_bytes == 8 ? segmented_one<V, false, long long>(a, s) : segmented_one<V, false, int>(a, s);
}

// ---- multi-GPU partition pass (4- and 8-byte keys; values 0/4/8 bytes) -----------------------
#if (B2S_K == 4 || B2S_K == 8)
// Shapes of the partition pass.  0 (default): two 512-thread CTAs per SM; 1: one 1024-thread CTA per SM with a tile twice as
// large (runs twice as long: larger bulk copies over NVLink); 2: 512 threads with more items per thread.  B2S_SPLIT_SHAPE
// selects (read once); the temp-storage size depends on the tile, so the choice is per process.
struct SplitShape { int nt, ipt, minb; };
template <int V>
constexpr SplitShape split_shape(int which) {
  return which == 1   ? SplitShape{1024, tmaw_ipt<V>(1024, 1, scale_ipt<V>(16)), 1}
         : which == 2 ? SplitShape{512, tmaw_ipt<V>(512, 2, scale_ipt<V>(20)), 2}
                      : SplitShape{512, tmaw_ipt<V>(512, 2, scale_ipt<V>(16)), 2};  // the destination functor is register-hungry
}
inline int split_shape_choice() {
  static const int c = [] {
    const char* e = std::getenv("B2S_SPLIT_SHAPE");
    const int v = e ? std::atoi(e) : 0;
    return (v >= 0 && v <= 2) ? v : 0;
  }();
  return c;
}

template <bool F>
SplitterOp<K, F> make_splitter_op(const SplitArgs& a) {
  using W = typename WideOf<K>::type;
  SplitterOp<K, F> op;
  op.base = make_op<F>(a.pass.dc, a.pass.bit, 8);
  const int nb = a.end_bit - a.pass.bit;
  op.range_mask = nb >= K * 8 ? ~W(0) : (W)((W(1) << nb) - 1);
  op.count = a.num_splitters;
  op.d_keys = a.d_splitter_keys;
  op.d_ranks = a.d_splitter_ranks;
  op.my_rank = a.my_rank;
  op.tie = 0;
  for (int j = 0; j < SplitterOp<K, F>::MAX_SPLITTERS; ++j) op.s[j] = ~W(0);
  return op;
}

template <int V, bool F, bool PEER, int SHAPE>
cudaError_t split_shape_one(const SplitArgs& a, cudaStream_t s) {
  constexpr SplitShape sh = split_shape<V>(SHAPE);
  using Op = SplitterOp<K, F>;
  OnesweepParams<K, Op> p;
  fill_params(p, a.pass, make_splitter_op<F>(a));
  for (int i = 0; i < MAX_PEERS; ++i) {
    p.peer_keys[i] = a.peer_keys[i];
    p.peer_vals[i] = a.peer_vals[i];
  }
  p.peer_capacity = a.peer_capacity;
  const unsigned long long tiles = (a.pass.n + sh.nt * sh.ipt - 1) / (sh.nt * sh.ipt);
  constexpr int BASE = PEER ? PF_PEER : 0;
  // With <= 8 destinations the runs of a tile are ~1000 items long: every run leaves the SM as one bulk shared->global
  // (or shared->peer over NVLink) copy when the destinations are 16-byte aligned (`bulk`); item stores otherwise.
  if (a.bulk) {
    using L = PassSmem<K, V, sh.nt, sh.ipt, true>;
    auto kern = digit_pass_kernel<K, V, Op, unsigned long long, sh.nt, sh.ipt, sh.minb, 6, BASE | PF_TMAW>;
    cudaError_t e = ensure_smem(kern, L::TOTAL);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned int)tiles, sh.nt, L::TOTAL, s>>>(p);
  } else {
    using L = PassSmem<K, V, sh.nt, sh.ipt, false>;
    auto kern = digit_pass_kernel<K, V, Op, unsigned long long, sh.nt, sh.ipt, sh.minb, 6, BASE>;
    cudaError_t e = ensure_smem(kern, L::TOTAL);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned int)tiles, sh.nt, L::TOTAL, s>>>(p);
  }
  return cudaGetLastError();
}

template <int V, bool F, bool PEER>
cudaError_t split_one(const SplitArgs& a, cudaStream_t s) {
  switch (split_shape_choice()) {
    case 1: return split_shape_one<V, F, PEER, 1>(a, s);
    case 2: return split_shape_one<V, F, PEER, 2>(a, s);
    default: return split_shape_one<V, F, PEER, 0>(a, s);
  }
}

template <int V>
cudaError_t split_v(const SplitArgs& a, cudaStream_t s) {
  if (a.pass.dc.is_float)
    return a.peer ? split_one<V, true, true>(a, s) : split_one<V, true, false>(a, s);
  return a.peer ? split_one<V, false, true>(a, s) : split_one<V, false, false>(a, s);
}

template <bool F>
cudaError_t split_count_one(const SplitArgs& a, cudaStream_t s) {
  const unsigned long long per_cta = 512ull * 16 * (16 / K);
  unsigned long long grid = (a.pass.n + per_cta - 1) / per_cta;
  if (grid > 148ull * 4) grid = 148ull * 4;
  if (grid == 0) grid = 1;
  split_count_kernel<K, SplitterOp<K, F>><<<(unsigned int)grid, 512, 0, s>>>(a.pass.keys_in, a.pass.n,
                                                                           make_splitter_op<F>(a),
                                                                           reinterpret_cast<unsigned long long*>(a.counts));
  return cudaGetLastError();
}
#endif

}  // namespace

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

cudaError_t CAT(hist_launch_k, B2S_K)(const HistArgs& a, cudaStream_t s) {
  if constexpr (K >= 2) {
    if (a.dc.is_float) return a.off64 ? hist_one<true, unsigned long long>(a, s) : hist_one<true, uint32_t>(a, s);
  }
  return a.off64 ? hist_one<false, unsigned long long>(a, s) : hist_one<false, uint32_t>(a, s);
}

cudaError_t CAT(onesweep_launch_k, B2S_K)(int variant, const PassArgs& a, cudaStream_t s) {
  switch (a.vbytes) {
    case 0: return launch_f<0>(variant, a, s);
    case 4: return launch_f<4>(variant, a, s);
    case 8: return launch_f<8>(variant, a, s);
#ifndef B2S_TUNING
    case 1: return launch_f<1>(variant, a, s);
    case 2: return launch_f<2>(variant, a, s);
    case 16: return launch_f<16>(variant, a, s);
#endif
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t CAT(single_launch_k, B2S_K)(const SingleArgs& a, cudaStream_t s) {
  switch (a.vbytes) {
    case 0: return single_v<0>(a, s);
    case 4: return single_v<4>(a, s);
    case 8: return single_v<8>(a, s);
#ifndef B2S_TUNING
    case 1: return single_v<1>(a, s);
    case 2: return single_v<2>(a, s);
    case 16: return single_v<16>(a, s);
#endif
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t CAT(segmented_launch_k, B2S_K)(const SegmentedArgs& a, cudaStream_t s) {
  switch (a.vbytes) {
    case 0: return segmented_v<0>(a, s);
    case 4: return segmented_v<4>(a, s);
    case 8: return segmented_v<8>(a, s);
#ifndef B2S_TUNING
    case 1: return segmented_v<1>(a, s);
    case 2: return segmented_v<2>(a, s);
    case 16: return segmented_v<16>(a, s);
#endif
    default: return cudaErrorInvalidValue;
  }
}

int CAT(single_tile_items_k, B2S_K)(int vbytes) {
#ifdef B2S_TUNING
  if (!(vbytes == 0 || vbytes == 4 || vbytes == 8)) return 0;
#endif
  return (K + vbytes <= 16) ? SingleTileShape<K, 0>::NT * 16 : SingleTileShape<K, 0>::NT * 8;
}

Variant CAT(onesweep_variant_k, B2S_K)(int variant, int vbytes, bool is_float, bool off64) {
#define B2S_VCASE(V)                                                                                           \
  case V:                                                                                                      \
    return is_float ? (off64 ? variant_cfg<V, true, true>(variant) : variant_cfg<V, true, false>(variant))     \
                    : (off64 ? variant_cfg<V, false, true>(variant) : variant_cfg<V, false, false>(variant));
  switch (vbytes) {
    B2S_VCASE(0)
    B2S_VCASE(1)
    B2S_VCASE(2)
    B2S_VCASE(4)
    B2S_VCASE(8)
    B2S_VCASE(16)
    default: return Variant{0, 0, 0, 0, 0, 0, 0};
  }
#undef B2S_VCASE
}

int CAT(onesweep_tile_k, B2S_K)(int variant, int vbytes, bool is_float, bool off64) {
  const Variant v = CAT(onesweep_variant_k, B2S_K)(variant, vbytes, is_float, off64);
  return v.nt * v.ipt;
}

int CAT(onesweep_num_variants_k, B2S_K)() { return NUM_VARIANTS; }

#if (B2S_K == 4 || B2S_K == 8)
cudaError_t CAT(split_count_launch_k, B2S_K)(const SplitArgs& a, cudaStream_t s) {
  return a.pass.dc.is_float ? split_count_one<true>(a, s) : split_count_one<false>(a, s);
}
cudaError_t CAT(split_launch_k, B2S_K)(const SplitArgs& a, cudaStream_t s) {
  switch (a.pass.vbytes) {
    case 0: return split_v<0>(a, s);
    case 4: return split_v<4>(a, s);
    case 8: return split_v<8>(a, s);
    default: return cudaErrorInvalidValue;
  }
}
int CAT(split_tile_k, B2S_K)(int vbytes) {
  const int c = split_shape_choice();
  switch (vbytes) {
    case 0: return split_shape<0>(c).nt * split_shape<0>(c).ipt;
    case 4: return split_shape<4>(c).nt * split_shape<4>(c).ipt;
    case 8: return split_shape<8>(c).nt * split_shape<8>(c).ipt;
    default: return 0;
  }
}
#else
cudaError_t CAT(split_count_launch_k, B2S_K)(const SplitArgs&, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t CAT(split_launch_k, B2S_K)(const SplitArgs&, cudaStream_t) { return cudaErrorNotSupported; }
int CAT(split_tile_k, B2S_K)(int) { return 0; }
#endif

}  // namespace b2s
