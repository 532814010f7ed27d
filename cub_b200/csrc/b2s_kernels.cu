// b2s_kernels.cu -- kernel instantiations + launchers for ONE key width (-DB2S_K=1|2|4|8).
// Compiled four times so the instantiations build in parallel.
#include <utility>

#include "b2s_histogram.cuh"
#include "b2s_internal.h"
#include "b2s_onesweep.cuh"

#ifndef B2S_K
#error "compile with -DB2S_K=<key bytes>"
#endif

namespace b2s {
namespace {

constexpr int K = B2S_K;

// ---- tuning table -------------------------------------------------------------------------
// Variant 0 is the production tuning for (K, V).  A tuning build (-DB2S_TUNING) adds more
// points that bench/tune.py sweeps on the GPU.
#ifdef B2S_TUNING
constexpr int NUM_VARIANTS = 12;
#else
constexpr int NUM_VARIANTS = 1;
#endif

template <int V>
constexpr Variant variant_cfg(int vi) {
  // registers held per item across the ranking phase: key words + packed rank (values re-use the key registers)
  int regs_per_item = (K > 4 ? 2 : 1) + 1;
  if ((V + 3) / 4 > (K > 4 ? 2 : 1)) regs_per_item = (V + 3) / 4 + 1;
  int cap = 32 / regs_per_item;  // items per thread that fit a 64-register budget
  if (cap > 16) cap = 16;
  if (cap < 4) cap = 4;
  // production tuning (B200 sweep, profiles/r1_tune_sweep_*.jsonl): big tiles win -- longer digit runs on
  // the scatter side matter more than occupancy
  Variant d = (K + V <= 8) ? Variant{512, cap, 2, MATCH_BALLOT} : Variant{384, cap > 12 ? 12 : cap, 3, MATCH_BALLOT};
#ifdef B2S_TUNING
  switch (vi) {
    case 0: return d;
    case 1: return Variant{512, cap, 2, MATCH_BALLOT};
    case 2: return Variant{384, cap, 3, MATCH_BALLOT};
    case 3: return Variant{1024, cap, 1, MATCH_BALLOT};
    case 4: return Variant{1024, cap > 12 ? 12 : cap, 1, MATCH_BALLOT};
    case 5: return Variant{768, cap, 1, MATCH_BALLOT};
    case 6: return Variant{512, cap + 4, 2, MATCH_BALLOT};
    case 7: return Variant{640, cap, 1, MATCH_BALLOT};
    case 8: return Variant{512, cap > 12 ? 12 : cap, 2, MATCH_BALLOT};
    case 9: return Variant{384, cap > 12 ? 12 : cap, 3, MATCH_BALLOT};
    case 10: return Variant{256, cap, 4, MATCH_BALLOT};
    case 11: return Variant{512, cap > 8 ? 8 : cap, 3, MATCH_BALLOT};
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

template <int V, bool F, typename OffT, int VI>
cudaError_t launch_one(const PassArgs& a, cudaStream_t s) {
  constexpr Variant c = variant_cfg<V>(VI);
  using L = OnesweepSmem<K, V, c.nt, c.ipt>;
  auto kern = onesweep_kernel<K, V, F, OffT, c.nt, c.ipt, c.minb, c.match>;
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
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
  p.op = make_op<F>(a.dc, a.bit, a.nbits);
  const unsigned long long tiles = (a.n + L::TILE - 1) / L::TILE;
  kern<<<(unsigned int)tiles, c.nt, L::TOTAL, s>>>(p);
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
  histogram_kernel<K, F, OffT><<<a.grid, HIST_THREADS, 0, s>>>(p);
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
    default: return Variant{0, 0, 0, 0};
  }
}

int CAT(onesweep_tile_k, B2S_K)(int variant, int vbytes) {
  const Variant v = CAT(onesweep_variant_k, B2S_K)(variant, vbytes);
  return v.nt * v.ipt;
}

int CAT(onesweep_num_variants_k, B2S_K)() { return NUM_VARIANTS; }

}  // namespace b2s
