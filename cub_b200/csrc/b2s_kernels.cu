// b2s_kernels.cu -- kernel instantiations + launchers for ONE key width (-DB2S_K=1|2|4|8).
// Compiled four times so the instantiations build in parallel.
#include <cstdlib>
#include <utility>

#include "b2s_histogram.cuh"
#include "b2s_internal.h"
#include "b2s_onesweep.cuh"
#include "b2s_onesweep2.cuh"

#ifndef B2S_K
#error "compile with -DB2S_K=<key bytes>"
#endif

namespace b2s {
namespace {

constexpr int K = B2S_K;

// ---- tuning table -------------------------------------------------------------------------
// Variant 0 is the production tuning for (K, V).  A tuning build (-DB2S_TUNING) adds more
// points that bench/tune.py sweeps on the GPU.  Table entries give items/thread for 4-byte keys
// with <=4-byte values; wider items scale it down by bytes (shared memory) and by registers.
#ifdef B2S_TUNING
constexpr int NUM_VARIANTS = 14;
#else
constexpr int NUM_VARIANTS = 1;
#endif

template <int V>
constexpr int scale_ipt(int ipt) {
  const int kw = K > 4 ? 2 : 1, vw = (V + 3) / 4;
  const int regs_per_item = (vw > kw ? vw : kw) + 1;  // key words (values re-use them) + packed rank
  const int by_bytes = (K + V <= 8) ? ipt : ipt * 8 / (K + V);
  const int by_regs = ipt * 2 / regs_per_item;
  const int r = by_bytes < by_regs ? by_bytes : by_regs;
  return r < 4 ? 4 : r;
}

template <int V>
constexpr Variant variant_cfg(int vi) {
  // {threads, items/thread, min CTAs/SM, match mode, kernel kind (0 one tile per CTA, 1 persistent), look-back window}
  const Variant d = Variant{512, scale_ipt<V>(20), 2, MATCH_BALLOT, 0, 4};
#ifdef B2S_TUNING
  switch (vi) {
    case 0: return d;
    case 1: return Variant{512, scale_ipt<V>(20), 2, MATCH_BALLOT, 0, 1};
    case 2: return Variant{512, scale_ipt<V>(20), 2, MATCH_BALLOT, 0, 2};
    case 3: return Variant{512, scale_ipt<V>(20), 2, MATCH_BALLOT, 0, 8};
    case 4: return Variant{512, scale_ipt<V>(16), 2, MATCH_BALLOT, 0, 4};
    case 5: return Variant{384, scale_ipt<V>(16), 3, MATCH_BALLOT, 0, 4};
    case 6: return Variant{384, scale_ipt<V>(19), 3, MATCH_BALLOT, 0, 4};
    case 7: return Variant{512, scale_ipt<V>(22), 2, MATCH_BALLOT, 0, 4};
    case 8: return Variant{256, scale_ipt<V>(16), 4, MATCH_BALLOT, 0, 4};
    case 9: return Variant{256, scale_ipt<V>(20), 4, MATCH_BALLOT, 0, 4};
    case 10: return Variant{1024, scale_ipt<V>(16), 1, MATCH_BALLOT, 0, 4};
    case 11: return Variant{256, scale_ipt<V>(32), 2, MATCH_BALLOT, 1, 4};
    case 12: return Variant{512, scale_ipt<V>(16), 2, MATCH_BALLOT, 1, 4};
    case 13: return Variant{640, scale_ipt<V>(16), 1, MATCH_BALLOT, 0, 4};
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
  op.zero_from = (W)dc.zero_from;
  op.zero_to = (W)dc.zero_to;
  op.bit = (uint32_t)bit;
  op.mask = (1u << nbits) - 1u;
  return op;
}

int sm_count_cached() {
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev];
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (n <= 0) n = 148;
  if (dev >= 0 && dev < 64) cached[dev] = n;
  return n;
}

template <int V, bool F, typename OffT, int VI>
cudaError_t launch_one(const PassArgs& a, cudaStream_t s) {
  constexpr Variant c = variant_cfg<V>(VI);
  constexpr int TILE = c.nt * c.ipt;
  constexpr int SMEM = c.kind == 1 ? Onesweep2Smem<K, V, c.nt, c.ipt>::TOTAL : OnesweepSmem<K, V, c.nt, c.ipt>::TOTAL;
  void (*kern)(const OnesweepParams<K, F>);
  if constexpr (c.kind == 1)
    kern = onesweep2_kernel<K, V, F, OffT, c.nt, c.ipt, c.minb, c.lbw>;
  else
    kern = onesweep_kernel<K, V, F, OffT, c.nt, c.ipt, c.minb, c.lbw>;
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  OnesweepParams<K, F> p;
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
  {
    static const int stagger_env = [] { const char* e = std::getenv("B2S_STAGGER_NS"); return e ? std::atoi(e) : -1; }();
    const unsigned int sms = (unsigned int)sm_count_cached();
    p.stagger_lo = sms;
    p.stagger_hi = c.minb > 1 ? sms * 2 : sms;
    p.stagger_ns = stagger_env >= 0 ? (unsigned int)stagger_env : 0u;
  }
  p.op = make_op<F>(a.dc, a.bit, a.nbits);
  unsigned long long grid = (a.n + TILE - 1) / TILE;
  if (c.kind == 1) {
    const unsigned long long resident = (unsigned long long)sm_count_cached() * c.minb;
    if (grid > resident) grid = resident;
  }
  kern<<<(unsigned int)grid, c.nt, SMEM, s>>>(p);
  return cudaGetLastError();
}

template <int V, bool F, typename OffT, int... VI>
cudaError_t launch_vi(int variant, const PassArgs& a, cudaStream_t s, std::integer_sequence<int, VI...>) {
  cudaError_t r = cudaErrorInvalidValue;
  (void)((variant == VI ? (r = launch_one<V, F, OffT, VI>(a, s), true) : false) || ...);
  return r;
}

template <int V, bool F>
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
  if constexpr (K >= 2) {
    if (a.dc.is_float) return launch_off<V, true>(variant, a, s);
  }
#endif
  return launch_off<V, false>(variant, a, s);
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
  auto kern = histogram_kernel<K, F, OffT>;
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, HistSmem<K>::BYTES);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  kern<<<a.grid, HIST_THREADS, HistSmem<K>::BYTES, s>>>(p);
  return cudaGetLastError();
}

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

Variant CAT(onesweep_variant_k, B2S_K)(int variant, int vbytes) {
  switch (vbytes) {
    case 0: return variant_cfg<0>(variant);
    case 1: return variant_cfg<1>(variant);
    case 2: return variant_cfg<2>(variant);
    case 4: return variant_cfg<4>(variant);
    case 8: return variant_cfg<8>(variant);
    case 16: return variant_cfg<16>(variant);
    default: return Variant{0, 0, 0, 0, 0, 0};
  }
}

int CAT(onesweep_tile_k, B2S_K)(int variant, int vbytes) {
  const Variant v = CAT(onesweep_variant_k, B2S_K)(variant, vbytes);
  return v.nt * v.ipt;
}

int CAT(onesweep_num_variants_k, B2S_K)() { return NUM_VARIANTS; }

}  // namespace b2s
