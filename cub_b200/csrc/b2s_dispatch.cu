// b2s_dispatch.cu -- host side of the B200 radix sort + the C-ABI of include/b2s_radix_sort.h.
//
// Mirrors the BEHAVIOUR of the reference host path (not its code):
//   cub::DeviceRadixSort entry points                 cub/device/device_radix_sort.cuh:312,781,1214,1675,2106,2525,2921,3330
//   DispatchRadixSort::Invoke / InvokeCopy            cub/device/dispatch/dispatch_radix_sort.cuh:1939-1978, 1885-1934
//   DispatchRadixSort::InvokeOnesweep                 cub/device/dispatch/dispatch_radix_sort.cuh:1521-1727
//   AliasTemporaries (256-byte carving, size query)   cub/util_device.cuh:68-109
//
// What is different by design: no <=2^28-item "portions" (64-bit look-back words for
// n >= 2^30 instead), ONE memset per sort (counters + histogram + first status array; each
// pass clears the status array of the next one), histogram + exclusive scan in one launch:
// a sort of P digit passes is 1 memset + 1 + P kernel launches.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>
#include <cstring>

#include "../../include/b2s_radix_sort.h"
#include "b2s_internal.h"

namespace b2s {
namespace {

struct KeyInfo {
  int bytes;
  int category;  // 0 unsigned, 1 signed, 2 floating
};
const KeyInfo kKeyInfo[B2S_KEY_TYPE_COUNT] = {
    {1, 0}, {1, 1}, {2, 0}, {2, 1}, {2, 2}, {2, 2}, {4, 0}, {4, 1}, {4, 2}, {8, 0}, {8, 1}, {8, 2},
};

thread_local int g_last_launches = 0;

// Optional per-launch timing (bench.py's roofline leg): when enabled, an event is recorded on the
// sort's stream before the first and after every enqueued operation of a sort call.
constexpr int kMaxTimingEvents = 16;
bool g_timing = false;
cudaEvent_t g_events[kMaxTimingEvents] = {};
int g_events_used = 0;
void timing_mark(cudaStream_t s) {
  if (!g_timing || g_events_used >= kMaxTimingEvents) return;
  if (!g_events[g_events_used]) cudaEventCreate(&g_events[g_events_used]);
  cudaEventRecord(g_events[g_events_used++], s);
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

DigitConsts make_consts(const KeyInfo& ki, bool descending) {
  const int bits = ki.bytes * 8;
  const uint64_t ones = bits == 64 ? ~0ull : ((1ull << bits) - 1);
  const uint64_t high = 1ull << (bits - 1);
  DigitConsts dc{};
  dc.is_float = ki.category == 2;
  if (dc.is_float) {
    dc.xor_mask = descending ? ones : 0;
    // image of -0.0 (ascending) / +0.0 (descending) under DigitOp::ordered before the collapse; keys narrower than
    // the 32-bit register carry ones above the key when negative
    const uint64_t reg_ones = bits == 64 ? ~0ull : 0xffffffffull;
    dc.zero_img = descending ? (ones ^ high) : ((reg_ones ^ high) & (bits == 64 ? ~0ull : 0xffffffffull));
    dc.pad_key = descending ? ones : (ones ^ high);  // -NaN(all ones) / +NaN(0x7f..f) order last
  } else {
    dc.xor_mask = (ki.category == 1 ? high : 0) ^ (descending ? ones : 0);
    dc.pad_key = ones ^ dc.xor_mask;
  }
  return dc;
}

// Test / tuning switches below are std::atomic (relaxed): every sort call only READS them, the setters of the C-ABI
// (b2s_set_*) may run on another host thread.  The per-launch timing hook (g_timing) stays single-threaded by contract.
// Tuning hook: variant 0 is the production tuning; a -DB2S_TUNING build carries more points.
std::atomic<int> g_variant{[] {
  const char* e = std::getenv("B2S_VARIANT");
  return e ? std::atoi(e) : 0;
}()};
int tuning_variant() { return g_variant; }
// B2S_SINGLE_TILE=0 sends small sorts through the multi-kernel path too (A/B runs, tests of that path at small n).
std::atomic<bool> g_single_tile{[] {
  const char* e = std::getenv("B2S_SINGLE_TILE");
  return !(e && e[0] == '0');
}()};
// Tile ids of the digit pass: block index (default; CTAs are dispatched in index order, the assumption CUB's decoupled
// look-back scan makes too) or an atomic ticket taken by every CTA (B2S_TILE_CLAIM=1 / b2s_set_tile_claim(1)): with
// tickets a tile's predecessors are always running or finished whatever the dispatch order, at ~1.5 % of throughput.
// B2S_SPLIT_BULK=0 keeps the partition pass on item stores (A/B runs of the multi-GPU exchange)
std::atomic<bool> g_split_bulk{[] {
  const char* e = std::getenv("B2S_SPLIT_BULK");
  return !(e && e[0] == '0');
}()};
std::atomic<bool> g_claim{[] {
  const char* e = std::getenv("B2S_TILE_CLAIM");
  return e && e[0] == '1';
}()};
// B2S_SKIP_CONSTANT=0 disables the constant-digit short circuit (A/B runs): every pass then ranks and scatters
std::atomic<bool> g_skip_constant{[] {
  const char* e = std::getenv("B2S_SKIP_CONSTANT");
  return !(e && e[0] == '0');
}()};
// Keys-only sorts of 1- and 2-byte keys over all their bits run as a counting sort (b2s_narrow.cu) from this many items on;
// B2S_COUNTING_SORT=0 / b2s_set_counting_sort(0) sends them through the digit passes like every other sort (A/B runs),
// b2s_set_counting_min_items lowers the cut-over for tests.
std::atomic<bool> g_counting{[] {
  const char* e = std::getenv("B2S_COUNTING_SORT");
  return !(e && e[0] == '0');
}()};
std::atomic<unsigned long long> g_counting_min[3] = {{1ull << 16}, {1ull << 22}, {1ull << 23}};  // 1-byte keys, 2-byte integers, 2-byte floats (measured cut-overs)
// Full-range sorts of 4- / 8-byte floating keys (no values or 4-byte values) record their zeros in the first pass and restore
// them after the last one (b2s_fzero.cu); B2S_FLOAT_ZERO_RECORD=0 / b2s_set_float_zero_recording(0) keeps the zero collapse in
// every digit extraction instead (A/B runs).
std::atomic<bool> g_fzero{[] {
  const char* e = std::getenv("B2S_FLOAT_ZERO_RECORD");
  return !(e && e[0] == '0');
}()};
// Tuning hook: phase-timestamp buffer for the trace variants of the digit pass (MODE bit 4), one pass per sort.
unsigned long long* g_trace = nullptr;
int g_trace_pass = -1;

struct KernelSet {
  cudaError_t (*hist)(const HistArgs&, cudaStream_t);
  cudaError_t (*onesweep)(int, const PassArgs&, cudaStream_t);
  int (*tile)(int, int, bool, bool);
  int (*num_variants)();
  Variant (*variant)(int, int, bool, bool);
  cudaError_t (*split_count)(const SplitArgs&, cudaStream_t);
  cudaError_t (*split)(const SplitArgs&, cudaStream_t);
  int (*split_tile)(int);
  cudaError_t (*single)(const SingleArgs&, cudaStream_t);
  int (*single_items)(int);
  cudaError_t (*segmented)(const SegmentedArgs&, cudaStream_t);
};
const KernelSet* kernels_for(int kbytes) {
  static const KernelSet k1{hist_launch_k1, onesweep_launch_k1, onesweep_tile_k1, onesweep_num_variants_k1, onesweep_variant_k1,
                            split_count_launch_k1, split_launch_k1, split_tile_k1, single_launch_k1, single_tile_items_k1, segmented_launch_k1};
  static const KernelSet k2{hist_launch_k2, onesweep_launch_k2, onesweep_tile_k2, onesweep_num_variants_k2, onesweep_variant_k2,
                            split_count_launch_k2, split_launch_k2, split_tile_k2, single_launch_k2, single_tile_items_k2, segmented_launch_k2};
  static const KernelSet k4{hist_launch_k4, onesweep_launch_k4, onesweep_tile_k4, onesweep_num_variants_k4, onesweep_variant_k4,
                            split_count_launch_k4, split_launch_k4, split_tile_k4, single_launch_k4, single_tile_items_k4, segmented_launch_k4};
  static const KernelSet k8{hist_launch_k8, onesweep_launch_k8, onesweep_tile_k8, onesweep_num_variants_k8, onesweep_variant_k8,
                            split_count_launch_k8, split_launch_k8, split_tile_k8, single_launch_k8, single_tile_items_k8, segmented_launch_k8};
  switch (kbytes) {
    case 1: return &k1;
    case 2: return &k2;
    case 4: return &k4;
    case 8: return &k8;
    default: return nullptr;
  }
}

int sm_count() {
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

// Temp-storage carving.
struct Layout {
  size_t off_ctrs, off_hist, off_status0, off_status1, off_keys, off_vals, total;
  size_t zero_bytes;  // [off_ctrs, off_ctrs + zero_bytes) is cleared once per sort
  size_t off_fz_z, off_fz_s, off_fz_partial;  // zero recording (b2s_fzero.cu), when used
};

Layout carve(uint64_t n, int kbytes, int vbytes, int passes, int tile, bool off64, bool need_alt, bool fz = false) {
  Layout L{};
  const size_t osz = off64 ? 8 : 4;
  const uint64_t tiles = (n + tile - 1) / tile;
  size_t o = 0;
  L.off_ctrs = o;      o += align_up(sizeof(unsigned int) * (size_t)(1 + 2 * passes), 256);  // ticket, tile counters, pass flags
  L.off_hist = o;      o += align_up(osz * 256 * (size_t)passes, 256);
  L.off_status0 = o;   o += align_up(osz * 256 * tiles, 256);
  L.zero_bytes = o;
  L.off_status1 = o;   o += passes > 1 ? align_up(osz * 256 * tiles, 256) : 0;
  L.off_keys = o;      o += need_alt ? align_up((size_t)n * kbytes, 256) : 0;
  L.off_vals = o;      o += (need_alt && vbytes) ? align_up((size_t)n * vbytes, 256) : 0;
  const size_t fzw = fz ? align_up(fzero_plane_words(n) * 4, 256) : 0;
  L.off_fz_z = o;      o += fzw;
  L.off_fz_s = o;      o += fzw;
  L.off_fz_partial = o; o += fz ? 8 * 1024 : 0;
  L.total = o + 255;   // slack so any d_temp_storage alignment works
  return L;
}

// Counting sort of narrow keys (b2s_narrow.cu).  The result always lands in kbuf[1]: `out` in the pointer form, the
// alternate buffer in the DoubleBuffer form (selector flips; which buffer holds the result is "a function of the number of
// key bits and the targeted architecture" in the reference's contract too, cub/device/device_radix_sort.cuh DoubleBuffer notes).
bool narrow_eligible(uint64_t n, const KeyInfo& ki, int vbytes, int begin_bit, int end_bit) {
  if (!g_counting || vbytes != 0 || ki.bytes > 2) return false;
  if (begin_bit != 0 || end_bit != ki.bytes * 8) return false;
  return n >= g_counting_min[ki.bytes == 1 ? 0 : (ki.category == 2 ? 2 : 1)];
}

struct NarrowLayout {
  size_t off_ctrs, off_counts, off_zflag, off_zpartial, off_prefix, off_zmasks, total, zero_bytes;
};
NarrowLayout carve_narrow(uint64_t n, const KeyInfo& ki, bool off64) {
  NarrowLayout L{};
  size_t o = 0;
  L.off_ctrs = o;    o += 256;  // 1-byte keys: completion ticket of the histogram kernel
  // 1-byte keys: the histogram kernel turns its counts into offsets in place; 2-byte keys: counts (zeroed) and prefix
  L.off_counts = o;  o += ki.bytes == 1 ? align_up((off64 ? 8 : 4) * 256, 256) : 8 * 65536;
  L.off_zflag = o;   o += 256;
  L.off_zpartial = o; o += ki.category == 2 ? 8 * 1024 : 0;
  L.zero_bytes = o;
  L.off_prefix = ki.bytes == 1 ? L.off_counts : o;
  o += ki.bytes == 1 ? 0 : align_up(8 * 65537, 256);
  L.off_zmasks = o;  o += (ki.bytes == 2 && ki.category == 2) ? align_up(narrow_zero_mask_bytes(n), 256) : 0;  // n / 8 bytes
  L.total = o + 255;
  return L;
}

int narrow_sort(void* d_temp, size_t* temp_bytes, void* kbuf[2], int* selector_out, bool overwrite, uint64_t n, const KeyInfo& ki,
                bool descending, cudaStream_t stream) {
  const bool off64 = n >= (1ull << 30);
  const NarrowLayout L = carve_narrow(n, ki, off64);
  if (!d_temp) {
    *temp_bytes = L.total;
    return (int)cudaSuccess;
  }
  if (*temp_bytes < L.total) return (int)cudaErrorInvalidValue;
  unsigned char* base = reinterpret_cast<unsigned char*>(align_up(reinterpret_cast<uintptr_t>(d_temp), 256));
  if (g_timing) g_events_used = 0;
  timing_mark(stream);
  cudaError_t e = cudaMemsetAsync(base, 0, L.zero_bytes, stream);
  if (e != cudaSuccess) return (int)e;
  g_last_launches++;
  timing_mark(stream);

  NarrowArgs a{};
  a.keys_in = kbuf[0];
  a.keys_out = kbuf[1];
  a.n = n;
  a.dc = make_consts(ki, descending);
  a.kbytes = ki.bytes;
  a.counts = base + L.off_counts;
  a.prefix = base + L.off_prefix;
  a.prefix64 = ki.bytes == 2 || off64;
  a.zflag = reinterpret_cast<unsigned int*>(base + L.off_zflag);
  a.zpartial = reinterpret_cast<unsigned long long*>(base + L.off_zpartial);
  a.zmasks = base + L.off_zmasks;
  a.sms = sm_count();
  auto step = [&](NarrowStep st) -> cudaError_t {
    const cudaError_t r = narrow_step(st, a, stream);
    if (r == cudaSuccess) {
      g_last_launches++;
      timing_mark(stream);
    }
    return r;
  };
  if (ki.bytes == 1) {
    HistArgs h{};
    h.keys = kbuf[0];
    h.n = n;
    h.dc = a.dc;
    h.begin_bit = 0;
    h.end_bit = 8;
    h.num_passes = 1;
    h.ghist = a.prefix;
    h.done = reinterpret_cast<unsigned int*>(base + L.off_ctrs);
    h.flags = nullptr;
    h.off64 = off64;
    const uint64_t vecs = (n + 16 * 1024 - 1) / (16 * 1024);
    uint64_t g = (uint64_t)sm_count();
    if (g > vecs) g = vecs ? vecs : 1;
    h.grid = (int)g;
    e = kernels_for(1)->hist(h, stream);
    if (e != cudaSuccess) return (int)e;
    g_last_launches++;
    timing_mark(stream);
  } else {
    if ((e = step(NarrowStep::kHist16)) != cudaSuccess) return (int)e;
    if ((e = step(NarrowStep::kPrefix16)) != cudaSuccess) return (int)e;
    if (ki.category == 2 && (e = step(NarrowStep::kZeroCount)) != cudaSuccess) return (int)e;
  }
  if ((e = step(NarrowStep::kExpand)) != cudaSuccess) return (int)e;
  if (ki.bytes == 2 && ki.category == 2) {
    if ((e = step(NarrowStep::kZeroWrite)) != cudaSuccess) return (int)e;
  }
  if (selector_out) *selector_out = overwrite ? 1 : 0;
  return (int)cudaSuccess;
}

// The shared implementation of both API forms.
//   overwrite == false: pointer form  (kin -> kout, kin never written)
//   overwrite == true : DoubleBuffer form (k[0]/k[1] ping-pong, selector returned)
int sort_impl(void* d_temp, size_t* temp_bytes, void* kbuf[2], void* vbuf[2], int* selector_out, bool overwrite,
              uint64_t n, int key_type, int vbytes, bool descending, int begin_bit, int end_bit,
              cudaStream_t stream) {
  g_last_launches = 0;
  if (!temp_bytes) return (int)cudaErrorInvalidValue;
  if (key_type < 0 || key_type >= B2S_KEY_TYPE_COUNT) return (int)cudaErrorInvalidValue;
  if (!(vbytes == 0 || vbytes == 1 || vbytes == 2 || vbytes == 4 || vbytes == 8 || vbytes == 16))
    return (int)cudaErrorInvalidValue;
  const KeyInfo ki = kKeyInfo[key_type];
  const int kbytes = ki.bytes;
  const int num_bits = end_bit - begin_bit;
  if (selector_out) *selector_out = 0;

  // Trivial cases (dispatch_radix_sort.cuh:1945-1963)
  if (n == 0 || (num_bits <= 0 && overwrite)) {
    if (!d_temp) *temp_bytes = 1;
    return (int)cudaSuccess;
  }
  if (num_bits <= 0) {
    if (!d_temp) { *temp_bytes = 1; return (int)cudaSuccess; }
    cudaError_t e = cudaMemcpyAsync(kbuf[1], kbuf[0], (size_t)n * kbytes, cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return (int)e;
    g_last_launches++;
    if (vbytes) {
      e = cudaMemcpyAsync(vbuf[1], vbuf[0], (size_t)n * vbytes, cudaMemcpyDeviceToDevice, stream);
      if (e != cudaSuccess) return (int)e;
      g_last_launches++;
    }
    return (int)cudaSuccess;
  }

  if (narrow_eligible(n, ki, vbytes, begin_bit, end_bit))
    return narrow_sort(d_temp, temp_bytes, kbuf, selector_out, overwrite, n, ki, descending, stream);

  const KernelSet* ks = kernels_for(kbytes);
  int variant = tuning_variant();
  if (variant < 0 || variant >= ks->num_variants()) variant = 0;
  const int passes = (num_bits + 7) / 8;
  const bool off64 = n >= (1ull << 30);
  // zero recording: floating keys, every bit sorted -- the zeros' run is then found at digit 0x80 of the top pass
  const bool fz = g_fzero && ki.category == 2 && kbytes >= 4 && (vbytes == 0 || vbytes == 4) && begin_bit == 0 &&
                  end_bit == kbytes * 8 && passes >= 2;
  const int tile = ks->tile(variant, vbytes, ki.category == 2 && !fz, off64);  // recording sorts run on the integer shapes
  const bool need_alt = !overwrite && passes > 1;
  const Layout L = carve(n, kbytes, vbytes, passes, tile, off64, need_alt, fz);

  if (!d_temp) {
    *temp_bytes = L.total;
    return (int)cudaSuccess;
  }
  if (*temp_bytes < L.total) return (int)cudaErrorInvalidValue;

  // Small sorts are launch-bound: at most one tile goes through ONE kernel that runs every pass in shared memory
  // (the reference's InvokeSingleTile, dispatch_radix_sort.cuh:1272; its cut-over is 4864 items).  The DoubleBuffer form
  // leaves the result where the multi-pass path would (selector = parity of the pass count).
  if (n <= (uint64_t)ks->single_items(vbytes) && g_single_tile) {
    SingleArgs sa{};
    const int dst = overwrite ? (passes & 1) : 1;
    sa.keys_in = kbuf[0];
    sa.keys_out = kbuf[dst];
    sa.vals_in = vbytes ? vbuf[0] : nullptr;
    sa.vals_out = vbytes ? vbuf[dst] : nullptr;
    sa.n = n;
    sa.dc = make_consts(ki, descending);
    sa.begin_bit = begin_bit;
    sa.end_bit = end_bit;
    sa.vbytes = vbytes;
    if (g_timing) g_events_used = 0;
    timing_mark(stream);
    cudaError_t e1 = ks->single(sa, stream);
    if (e1 != cudaSuccess) return (int)e1;
    g_last_launches++;
    timing_mark(stream);
    if (selector_out) *selector_out = overwrite ? dst : 0;
    return (int)cudaSuccess;
  }

  unsigned char* base = reinterpret_cast<unsigned char*>(align_up(reinterpret_cast<uintptr_t>(d_temp), 256));
  unsigned int* ctrs = reinterpret_cast<unsigned int*>(base + L.off_ctrs);
  unsigned char* hist = base + L.off_hist;
  unsigned char* status[2] = {base + L.off_status0, base + L.off_status1};
  const size_t osz = off64 ? 8 : 4;

  if (g_timing) g_events_used = 0;  // the timing hook is single-threaded by contract; plain sorts never touch shared state
  timing_mark(stream);
  cudaError_t e = cudaMemsetAsync(base + L.off_ctrs, 0, L.zero_bytes, stream);
  if (e != cudaSuccess) return (int)e;
  g_last_launches++;
  timing_mark(stream);

  const DigitConsts dc = make_consts(ki, descending);

  HistArgs h{};
  h.keys = kbuf[0];
  h.n = n;
  h.dc = dc;
  h.begin_bit = begin_bit;
  h.end_bit = end_bit;
  h.num_passes = passes;
  h.ghist = hist;
  h.done = ctrs;
  h.flags = g_skip_constant ? ctrs + 1 + passes : nullptr;
  h.off64 = off64;
  {
    const uint64_t vecs = (n * kbytes + 16 * 1024 - 1) / (16 * 1024);  // CTAs worth of 128-bit loads
    uint64_t g = (uint64_t)sm_count();                                 // one 1024-thread CTA per SM
    if (g > vecs) g = vecs ? vecs : 1;
    h.grid = (int)g;
  }
  e = ks->hist(h, stream);
  if (e != cudaSuccess) return (int)e;
  g_last_launches++;
  timing_mark(stream);

  // Ping-pong plan.  Pointer form: in -> {tmp,out} alternating so that the last pass lands in
  // `out` and `in` is only ever read.  DoubleBuffer form: the two user buffers alternate.
  void* ktmp = need_alt ? base + L.off_keys : nullptr;
  void* vtmp = (need_alt && vbytes) ? base + L.off_vals : nullptr;
  const void* ksrc = kbuf[0];
  const void* vsrc = vbytes ? vbuf[0] : nullptr;
  int cur = 0;  // DoubleBuffer: index of the buffer holding the current data
  for (int p = 0; p < passes; ++p) {
    void* kdst;
    void* vdst;
    if (overwrite) {
      kdst = kbuf[cur ^ 1];
      vdst = vbytes ? vbuf[cur ^ 1] : nullptr;
    } else {
      const bool to_out = ((passes - 1 - p) & 1) == 0;
      kdst = to_out ? kbuf[1] : ktmp;
      vdst = vbytes ? (to_out ? vbuf[1] : vtmp) : nullptr;
    }
    PassArgs a{};
    const bool fpass = fz && (p == 0 || p == passes - 1);  // first / last pass convert (and record): ImageFloatOp kernels
    a.keys_in = ksrc;
    a.keys_out = kdst;
    a.vals_in = vsrc;
    a.vals_out = vdst;
    a.zero_z = fpass ? reinterpret_cast<unsigned int*>(base + L.off_fz_z) : nullptr;
    a.zero_s = fpass ? reinterpret_cast<unsigned int*>(base + L.off_fz_s) : nullptr;
    a.status = status[p & 1];
    a.status_next = (p + 1 < passes) ? status[(p + 1) & 1] : nullptr;
    a.bins = hist + osz * 256 * (size_t)p;
    a.tile_counter = ctrs + 1 + p;
    a.n = n;
    a.dc = dc;
    if (fz && !fpass) {  // between the first and the last pass the keys are plain bit-ordered integers
      a.dc = DigitConsts{};
      a.dc.pad_key = kbytes == 8 ? ~0ull : 0xffffffffull;
    }
    a.bit = begin_bit + 8 * p;
    a.nbits = (end_bit - a.bit) < 8 ? (end_bit - a.bit) : 8;
    a.off64 = off64;
    a.vbytes = vbytes;
    a.trace = (g_trace_pass == p) ? g_trace : nullptr;
    a.claim = g_claim;
    // the recording pass has to see every key: no constant-digit copy for it
    a.skip_flag = (g_skip_constant && !(fz && p == 0)) ? ctrs + 1 + passes + p : nullptr;
    a.raw_in = p == 0;
    a.raw_out = p == passes - 1;
    e = ks->onesweep(variant, a, stream);
    if (e != cudaSuccess) return (int)e;
    g_last_launches++;
    timing_mark(stream);
    ksrc = kdst;
    vsrc = vdst;
    cur ^= 1;
  }
  if (fz) {  // the zeros arrived as one run of equal keys in input order: give them their recorded signs back
    FzeroArgs f{};
    f.zero_z = reinterpret_cast<const unsigned int*>(base + L.off_fz_z);
    f.zero_s = reinterpret_cast<const unsigned int*>(base + L.off_fz_s);
    f.n = n;
    f.kbytes = kbytes;
    f.keys_out = const_cast<void*>(ksrc);
    f.top_bins = hist + osz * 256 * (size_t)(passes - 1);
    f.off64 = off64;
    f.partial = reinterpret_cast<unsigned long long*>(base + L.off_fz_partial);
    f.sms = sm_count();
    if ((e = fzero_count_launch(f, stream)) != cudaSuccess) return (int)e;
    g_last_launches++;
    timing_mark(stream);
    if ((e = fzero_write_launch(f, stream)) != cudaSuccess) return (int)e;
    g_last_launches++;
    timing_mark(stream);
  }
  if (selector_out) *selector_out = overwrite ? cur : 0;
  return (int)cudaSuccess;
}


// ---- segmented sort (cub::DeviceSegmentedRadixSort) -----------------------------------------------------------------
//   overwrite == false: pointer form (kin -> kout; kin never written; an alternate buffer in the temp storage when there is
//                       more than one pass);  overwrite == true: DoubleBuffer form (selector = parity of the pass count,
//                       dispatch_radix_sort.cuh:2343-2349)
int segmented_impl(void* d_temp, size_t* temp_bytes, void* kbuf[2], void* vbuf[2], int* selector_out, bool overwrite, uint64_t n,
                   uint64_t num_segments, const void* begin_offsets, const void* end_offsets, int offset_bytes, int key_type,
                   int vbytes, bool descending, int begin_bit, int end_bit, cudaStream_t stream) {
  if (!temp_bytes) return (int)cudaErrorInvalidValue;
  if (key_type < 0 || key_type >= B2S_KEY_TYPE_COUNT) return (int)cudaErrorInvalidValue;
  if (!(vbytes == 0 || vbytes == 1 || vbytes == 2 || vbytes == 4 || vbytes == 8 || vbytes == 16)) return (int)cudaErrorInvalidValue;
  if (!(offset_bytes == 4 || offset_bytes == 8)) return (int)cudaErrorInvalidValue;
  const KeyInfo ki = kKeyInfo[key_type];
  const int num_bits = end_bit - begin_bit;
  if (selector_out) *selector_out = 0;
  // dispatch_radix_sort.cuh:2369: empty problem, or no bits to sort with double buffering
  if (n == 0 || num_segments == 0 || (num_bits <= 0 && overwrite)) {
    if (!d_temp) *temp_bytes = 1;
    return (int)cudaSuccess;
  }
  const int passes = num_bits <= 0 ? 1 : (num_bits + 7) / 8;  // a zero-bit pass copies every segment (:2307)
  const bool need_alt = !overwrite && passes > 1;
  size_t o = 0;
  const size_t off_keys = o;  o += need_alt ? align_up((size_t)n * ki.bytes, 256) : 0;
  const size_t off_vals = o;  o += (need_alt && vbytes) ? align_up((size_t)n * vbytes, 256) : 0;
  const size_t total = o + 256;
  if (!d_temp) {
    *temp_bytes = total;
    return (int)cudaSuccess;
  }
  if (*temp_bytes < total) return (int)cudaErrorInvalidValue;
  unsigned char* base = reinterpret_cast<unsigned char*>(align_up(reinterpret_cast<uintptr_t>(d_temp), 256));
  SegmentedArgs a{};
  a.keys_src = kbuf[0];
  a.vals_src = vbytes ? vbuf[0] : nullptr;
  if (overwrite) {
    const int fin = passes & 1;  // pass p writes buffer (p + 1) & 1: the last one lands in buffer passes & 1
    a.keys_a = kbuf[fin];
    a.keys_b = kbuf[fin ^ 1];
    a.vals_a = vbytes ? vbuf[fin] : nullptr;
    a.vals_b = vbytes ? vbuf[fin ^ 1] : nullptr;
    if (selector_out) *selector_out = fin;
  } else {
    a.keys_a = kbuf[1];
    a.keys_b = need_alt ? base + off_keys : nullptr;
    a.vals_a = vbytes ? vbuf[1] : nullptr;
    a.vals_b = (need_alt && vbytes) ? base + off_vals : nullptr;
  }
  a.begin_offsets = begin_offsets;
  a.end_offsets = end_offsets;
  a.offset_bytes = offset_bytes;
  a.num_segments = num_segments;
  a.dc = make_consts(ki, descending);
  a.begin_bit = begin_bit;
  a.end_bit = end_bit;
  a.passes = passes;
  a.vbytes = vbytes;
  g_last_launches = 1;
  return (int)kernels_for(ki.bytes)->segmented(a, stream);
}

// ---- multi-GPU partition pass ---------------------------------------------------------------------------------
bool fill_split_args(SplitArgs& a, uint64_t n, int key_type, int vbytes, bool descending, int begin_bit, int end_bit,
                     const void* d_splitter_keys, const int* d_splitter_ranks, int num_splitters, int my_rank) {
  if (key_type < 0 || key_type >= B2S_KEY_TYPE_COUNT) return false;
  if (num_splitters < 0 || num_splitters > kMaxSplitters) return false;
  if (num_splitters && (!d_splitter_keys || !d_splitter_ranks)) return false;
  const KeyInfo ki = kKeyInfo[key_type];
  if (end_bit <= begin_bit || begin_bit < 0 || end_bit > ki.bytes * 8) return false;
  a = SplitArgs{};
  a.pass.n = n;
  a.pass.dc = make_consts(ki, descending);
  a.pass.bit = begin_bit;
  a.pass.nbits = 8;
  a.pass.off64 = true;
  a.pass.vbytes = vbytes;
  a.end_bit = end_bit;
  a.num_splitters = num_splitters;
  a.d_splitter_keys = d_splitter_keys;
  a.d_splitter_ranks = d_splitter_ranks;
  a.my_rank = my_rank;
  return true;
}

struct SplitLayout {
  size_t off_ctr, off_bins, off_status, total, zero_bytes;
};
SplitLayout carve_split(uint64_t n, int tile) {
  SplitLayout L{};
  const uint64_t tiles = (n + tile - 1) / tile;
  size_t o = 0;
  L.off_ctr = o;    o += 256;
  L.off_bins = o;   o += 8 * 256;
  L.off_status = o; o += align_up(8 * 256 * tiles, 256);
  L.zero_bytes = o;
  L.total = o + 255;
  return L;
}

}  // namespace
}  // namespace b2s

extern "C" {

int b2s_enable_peer_access(int peer_device) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (peer_device == dev) return (int)cudaSuccess;
  int can = 0;
  e = cudaDeviceCanAccessPeer(&can, dev, peer_device);
  if (e != cudaSuccess) return (int)e;
  if (!can) return (int)cudaErrorPeerAccessUnsupported;
  e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();  // clear the sticky-free error state
    e = cudaSuccess;
  }
  return (int)e;
}

int b2s_ipc_open(const void* handle64, void** d_ptr) {
  if (!handle64 || !d_ptr) return (int)cudaErrorInvalidValue;
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::memcpy(&h, handle64, sizeof(h));
  // opened with the CURRENT device being the one whose kernels will store into the buffer, so that the driver
  // sets up the peer mapping for it
  return (int)cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess);
}

int b2s_ipc_close(void* d_ptr) { return (int)cudaIpcCloseMemHandle(d_ptr); }

int b2s_split_count(const void* d_keys_in, uint64_t num_items, int key_type, int descending, int begin_bit, int end_bit,
                    const void* d_splitter_keys, const int* d_splitter_ranks, int num_splitters, int my_rank,
                    uint64_t* d_counts, b2s_stream_t stream) {
  b2s::SplitArgs a;
  if (!d_counts || !b2s::fill_split_args(a, num_items, key_type, 0, descending != 0, begin_bit, end_bit, d_splitter_keys,
                                         d_splitter_ranks, num_splitters, my_rank))
    return (int)cudaErrorInvalidValue;
  const b2s::KernelSet* ks = b2s::kernels_for(b2s::kKeyInfo[key_type].bytes);
  cudaError_t e = cudaMemsetAsync(d_counts, 0, sizeof(uint64_t) * (size_t)(num_splitters + 1), (cudaStream_t)stream);
  if (e != cudaSuccess || num_items == 0) return (int)e;
  a.pass.keys_in = d_keys_in;
  a.counts = reinterpret_cast<uint64_t*>(d_counts);
  return (int)ks->split_count(a, (cudaStream_t)stream);
}

int b2s_split_scatter(void* d_temp_storage, size_t* temp_storage_bytes, const void* d_keys_in, void* d_keys_out,
                      const void* d_values_in, void* d_values_out, uint64_t num_items, int key_type, int value_bytes,
                      int descending, int begin_bit, int end_bit, const void* d_splitter_keys,
                      const int* d_splitter_ranks, int num_splitters, int my_rank, const uint64_t* d_dest_offsets,
                      void* const* peer_keys, void* const* peer_vals, uint64_t peer_capacity, b2s_stream_t stream) {
  if (!temp_storage_bytes) return (int)cudaErrorInvalidValue;
  b2s::SplitArgs a;
  if (!b2s::fill_split_args(a, num_items, key_type, value_bytes, descending != 0, begin_bit, end_bit, d_splitter_keys,
                            d_splitter_ranks, num_splitters, my_rank))
    return (int)cudaErrorInvalidValue;
  const b2s::KernelSet* ks = b2s::kernels_for(b2s::kKeyInfo[key_type].bytes);
  const int tile = ks->split_tile(value_bytes);
  if (tile == 0) return (int)cudaErrorNotSupported;  // 4-/8-byte keys with 0-/4-/8-byte values only
  if (num_items == 0) {
    if (!d_temp_storage) *temp_storage_bytes = 1;
    return (int)cudaSuccess;
  }
  const b2s::SplitLayout L = b2s::carve_split(num_items, tile);
  if (!d_temp_storage) {
    *temp_storage_bytes = L.total;
    return (int)cudaSuccess;
  }
  if (*temp_storage_bytes < L.total || !d_dest_offsets) return (int)cudaErrorInvalidValue;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* base = reinterpret_cast<unsigned char*>(b2s::align_up(reinterpret_cast<uintptr_t>(d_temp_storage), 256));
  cudaError_t e = cudaMemsetAsync(base, 0, L.zero_bytes, s);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemcpyAsync(base + L.off_bins, d_dest_offsets, sizeof(uint64_t) * (size_t)(num_splitters + 1),
                      cudaMemcpyDeviceToDevice, s);
  if (e != cudaSuccess) return (int)e;
  a.pass.keys_in = d_keys_in;
  a.pass.keys_out = d_keys_out;
  a.pass.vals_in = d_values_in;
  a.pass.vals_out = d_values_out;
  a.pass.status = base + L.off_status;
  a.pass.status_next = nullptr;
  a.pass.bins = base + L.off_bins;
  a.pass.tile_counter = reinterpret_cast<unsigned int*>(base + L.off_ctr);
  a.peer = peer_keys != nullptr;
  a.peer_capacity = peer_capacity;
  uintptr_t align_or = reinterpret_cast<uintptr_t>(d_keys_out) | reinterpret_cast<uintptr_t>(d_values_out);
  if (a.peer) {
    align_or = 0;
    for (int d = 0; d <= num_splitters; ++d) {
      a.peer_keys[d] = peer_keys[d];
      a.peer_vals[d] = (value_bytes && peer_vals) ? peer_vals[d] : nullptr;
      align_or |= reinterpret_cast<uintptr_t>(a.peer_keys[d]) | reinterpret_cast<uintptr_t>(a.peer_vals[d]);
    }
  }
  a.bulk = (align_or & 15) == 0 && b2s::g_split_bulk;
  return (int)ks->split(a, s);
}

// 128-bit integer keys are a two-member composite on this little-endian machine: high word (signed for __int128), low word
static const b2s_field_t kFieldsU128[2] = {{8, B2S_U64}, {0, B2S_U64}};
static const b2s_field_t kFieldsI128[2] = {{8, B2S_I64}, {0, B2S_U64}};

int b2s_radix_sort(void* d_temp_storage, size_t* temp_storage_bytes, const void* d_keys_in, void* d_keys_out,
                   const void* d_values_in, void* d_values_out, uint64_t num_items, int key_type, int value_bytes,
                   int offset_bytes, int descending, int begin_bit, int end_bit, b2s_stream_t stream) {
  (void)offset_bytes;
  if (key_type == B2S_U128 || key_type == B2S_I128)
    return b2s_radix_sort_struct(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out, num_items, 16,
                                 key_type == B2S_U128 ? kFieldsU128 : kFieldsI128, 2, value_bytes, descending, begin_bit, end_bit,
                                 stream);
  void* k[2] = {const_cast<void*>(d_keys_in), d_keys_out};
  void* v[2] = {const_cast<void*>(d_values_in), d_values_out};
  return b2s::sort_impl(d_temp_storage, temp_storage_bytes, k, v, nullptr, false, num_items, key_type, value_bytes,
                        descending != 0, begin_bit, end_bit, (cudaStream_t)stream);
}

int b2s_radix_sort_db(void* d_temp_storage, size_t* temp_storage_bytes, void* key_bufs[2], int* key_selector,
                      void* val_bufs[2], int* val_selector, uint64_t num_items, int key_type, int value_bytes,
                      int offset_bytes, int descending, int begin_bit, int end_bit, b2s_stream_t stream) {
  (void)offset_bytes;
  if (key_type == B2S_U128 || key_type == B2S_I128)
    return b2s_radix_sort_struct_db(d_temp_storage, temp_storage_bytes, key_bufs, key_selector, val_bufs, val_selector, num_items,
                                    16, key_type == B2S_U128 ? kFieldsU128 : kFieldsI128, 2, value_bytes, descending, begin_bit,
                                    end_bit, stream);
  if (!key_bufs || !key_selector) return (int)cudaErrorInvalidValue;
  if (value_bytes && (!val_bufs || !val_selector)) return (int)cudaErrorInvalidValue;
  const int ks = *key_selector & 1;
  const int vs = value_bytes ? (*val_selector & 1) : 0;
  void* k[2] = {key_bufs[ks], key_bufs[ks ^ 1]};
  void* v[2] = {value_bytes ? val_bufs[vs] : nullptr, value_bytes ? val_bufs[vs ^ 1] : nullptr};
  int flipped = 0;
  const int r = b2s::sort_impl(d_temp_storage, temp_storage_bytes, k, v, &flipped, true, num_items, key_type,
                               value_bytes, descending != 0, begin_bit, end_bit, (cudaStream_t)stream);
  if (r == 0 && d_temp_storage) {
    *key_selector = ks ^ flipped;
    if (value_bytes) *val_selector = vs ^ flipped;
  }
  return r;
}

int b2s_segmented_radix_sort(void* d_temp_storage, size_t* temp_storage_bytes, const void* d_keys_in, void* d_keys_out,
                             const void* d_values_in, void* d_values_out, uint64_t num_items, uint64_t num_segments,
                             const void* d_begin_offsets, const void* d_end_offsets, int offset_bytes, int key_type,
                             int value_bytes, int descending, int begin_bit, int end_bit, b2s_stream_t stream) {
  void* k[2] = {const_cast<void*>(d_keys_in), d_keys_out};
  void* v[2] = {const_cast<void*>(d_values_in), d_values_out};
  return b2s::segmented_impl(d_temp_storage, temp_storage_bytes, k, v, nullptr, false, num_items, num_segments, d_begin_offsets,
                             d_end_offsets, offset_bytes, key_type, value_bytes, descending != 0, begin_bit, end_bit,
                             (cudaStream_t)stream);
}

int b2s_segmented_radix_sort_db(void* d_temp_storage, size_t* temp_storage_bytes, void* key_bufs[2], int* key_selector,
                                void* val_bufs[2], int* val_selector, uint64_t num_items, uint64_t num_segments,
                                const void* d_begin_offsets, const void* d_end_offsets, int offset_bytes, int key_type,
                                int value_bytes, int descending, int begin_bit, int end_bit, b2s_stream_t stream) {
  if (!key_bufs || !key_selector) return (int)cudaErrorInvalidValue;
  if (value_bytes && (!val_bufs || !val_selector)) return (int)cudaErrorInvalidValue;
  const int ks = *key_selector & 1;
  const int vs = value_bytes ? (*val_selector & 1) : 0;
  void* k[2] = {key_bufs[ks], key_bufs[ks ^ 1]};
  void* v[2] = {value_bytes ? val_bufs[vs] : nullptr, value_bytes ? val_bufs[vs ^ 1] : nullptr};
  int flipped = 0;
  const int r = b2s::segmented_impl(d_temp_storage, temp_storage_bytes, k, v, &flipped, true, num_items, num_segments,
                                    d_begin_offsets, d_end_offsets, offset_bytes, key_type, value_bytes, descending != 0, begin_bit,
                                    end_bit, (cudaStream_t)stream);
  if (r == 0 && d_temp_storage) {
    *key_selector = ks ^ flipped;
    if (value_bytes) *val_selector = vs ^ flipped;
  }
  return r;
}

int b2s_digit_histogram(const void* d_keys, uint64_t num_items, int key_type, int descending, int begin_bit, int end_bit,
                        uint64_t* d_offsets, b2s_stream_t stream) {
  if (!d_offsets || key_type < 0 || key_type >= B2S_KEY_TYPE_COUNT || end_bit <= begin_bit) return (int)cudaErrorInvalidValue;
  const b2s::KeyInfo ki = b2s::kKeyInfo[key_type];
  const int passes = (end_bit - begin_bit + 7) / 8;
  const b2s::KernelSet* ks = b2s::kernels_for(ki.bytes);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(d_offsets, 0, sizeof(uint64_t) * ((size_t)passes * 256 + 1), s);
  if (e != cudaSuccess || num_items == 0) return (int)e;
  b2s::HistArgs h{};
  h.keys = d_keys;
  h.n = num_items;
  h.dc = b2s::make_consts(ki, descending != 0);
  h.begin_bit = begin_bit;
  h.end_bit = end_bit;
  h.num_passes = passes;
  h.ghist = d_offsets;
  h.done = reinterpret_cast<unsigned int*>(d_offsets + (size_t)passes * 256);
  h.flags = nullptr;
  h.off64 = true;
  const uint64_t vecs = (num_items * ki.bytes + 16 * 1024 - 1) / (16 * 1024);
  uint64_t g = (uint64_t)b2s::sm_count();
  if (g > vecs) g = vecs ? vecs : 1;
  h.grid = (int)g;
  return (int)ks->hist(h, s);
}

int b2s_key_bytes(int key_type) {
  if (key_type == B2S_U128 || key_type == B2S_I128) return 16;
  return (key_type >= 0 && key_type < B2S_KEY_TYPE_COUNT) ? b2s::kKeyInfo[key_type].bytes : 0;
}

const char* b2s_version(void) { return "b2s 0.1 sm_100a"; }

int b2s_last_launch_count(void) { return b2s::g_last_launches; }

int b2s_timing_enable(int on) {
  const int old = b2s::g_timing;
  b2s::g_timing = on != 0;
  return old;
}

int b2s_timing_read(float* ms, int capacity) {
  // segment i = time between mark i and mark i+1 of the last timed sort: [memset, histogram, pass 0, pass 1, ...]
  const int segs = b2s::g_events_used - 1;
  if (segs <= 0) return 0;
  cudaError_t e = cudaEventSynchronize(b2s::g_events[b2s::g_events_used - 1]);
  if (e != cudaSuccess) return -(int)e;
  int i = 0;
  for (; i < segs && i < capacity; ++i) {
    e = cudaEventElapsedTime(&ms[i], b2s::g_events[i], b2s::g_events[i + 1]);
    if (e != cudaSuccess) return -(int)e;
  }
  return i;
}

int b2s_set_variant(int variant) {
  const int old = b2s::g_variant;
  b2s::g_variant = variant;
  return old;
}

int b2s_describe_variant(int key_bytes, int value_bytes, int variant, int* nt, int* ipt, int* minb, int* match) {
  const b2s::KernelSet* ks = b2s::kernels_for(key_bytes);
  if (!ks || variant < 0 || variant >= ks->num_variants()) return -1;
  const b2s::Variant v = ks->variant(variant, value_bytes, false, false);
  if (nt) *nt = v.nt;
  if (ipt) *ipt = v.ipt;
  if (minb) *minb = v.minb;
  if (match) *match = (v.lbw << 8) | (v.abl << 16);  // legacy slot: look-back window in bits 8+, ablation in 16+
  return ks->num_variants();
}

int b2s_set_single_tile(int enable) {
  const int old = b2s::g_single_tile ? 1 : 0;
  b2s::g_single_tile = enable != 0;
  return old;
}

int b2s_set_float_zero_recording(int enable) {
  const int old = b2s::g_fzero ? 1 : 0;
  b2s::g_fzero = enable != 0;
  return old;
}

int b2s_set_counting_sort(int enable) {
  const int old = b2s::g_counting ? 1 : 0;
  b2s::g_counting = enable != 0;
  return old;
}

uint64_t b2s_set_counting_min_items(int key_bytes, uint64_t min_items) {
  if (key_bytes < 1 || key_bytes > 2) return 0;
  const uint64_t old = b2s::g_counting_min[key_bytes - 1];
  b2s::g_counting_min[key_bytes - 1] = min_items;
  if (key_bytes == 2) b2s::g_counting_min[2] = min_items;  // floating 2-byte keys follow (their own default is higher)
  return old;
}

int b2s_set_trace(void* d_trace, int pass) {
  b2s::g_trace = reinterpret_cast<unsigned long long*>(d_trace);
  b2s::g_trace_pass = d_trace ? pass : -1;
  return 0;
}

int b2s_set_tile_claim(int enable) {
  const int old = b2s::g_claim ? 1 : 0;
  b2s::g_claim = enable != 0;
  return old;
}

int b2s_variant_flow(int key_bytes, int value_bytes, int variant) {
  const b2s::KernelSet* ks = b2s::kernels_for(key_bytes);
  if (!ks || variant < 0 || variant >= ks->num_variants()) return -2;
  return ks->variant(variant, value_bytes, false, false).flow;
}

int b2s_variant_mode(int key_bytes, int value_bytes, int variant) {
  const b2s::KernelSet* ks = b2s::kernels_for(key_bytes);
  if (!ks || variant < 0 || variant >= ks->num_variants()) return -1;
  return ks->variant(variant, value_bytes, false, false).mode;
}

}  // extern "C"
