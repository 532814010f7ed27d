// b2s_histogram.cuh -- upfront all-digits histogram + fused exclusive scan.
//
// Replaces (reference, for parity of RESULT only):
//   DeviceRadixSortHistogramKernel     cub/device/dispatch/dispatch_radix_sort.cuh:556
//     / AgentRadixSortHistogram        cub/agent/agent_radix_sort_histogram.cuh:87-280
//   DeviceRadixSortExclusiveSumKernel  cub/device/dispatch/dispatch_radix_sort.cuh:603-635
//
// B200 design: persistent grid (one 1024-thread CTA per SM), every key read exactly once with
// 128-bit coalesced loads (4 vectors in flight per thread), all passes' digits counted in
// shared-memory bins privatised per CTA and per LANE: bin (pass, digit) has one counter per lane
// (bank == lane), so a warp-wide shared-memory atomic never has a bank conflict nor two lanes on
// one address and costs one wavefront instead of ~3 (bench/micro/prim.cu: 3.0 -> ~1.2 cycles per
// warp instruction), which takes the kernel from atomics-bound to HBM-bound.  That needs
// passes x 256 x 32 x 4 B = 128 KB of shared memory for 4-byte keys; 8-byte keys use 16 lanes per
// bin (lanes l and l+16 share a column).  One global atomic per non-empty (pass, digit) per CTA;
// the last CTA to finish turns the counts into exclusive offsets in place (no second launch).
// Roofline: HBM read of n*K bytes; algorithmic bytes/key = K.
#pragma once
#include "b2s_common.cuh"

namespace b2s {

constexpr int HIST_THREADS = 1024;
constexpr int HIST_UNROLL = 4;  // 128-bit loads in flight per thread
template <int KBYTES>
struct HistSmem {
  static constexpr int PARTS = KBYTES == 8 ? 16 : 32;          // per-lane counters per (pass, digit)
  static constexpr int BYTES = KBYTES * RADIX * PARTS * 4;     // KBYTES passes at 8 bits per digit
};

template <int KBYTES, bool IS_FLOAT>
struct HistParams {
  const void* keys;
  unsigned long long n;
  DigitOp<KBYTES, IS_FLOAT> op;  // .bit/.mask unused here
  int begin_bit;
  int end_bit;
  int num_passes;
  void* ghist;          // OffT[num_passes][256]: counts, then exclusive offsets
  unsigned int* done;   // CTA completion ticket
  unsigned int* flags;  // [num_passes], zero on entry (may be null): flags[p] = 1 when ONE digit of pass p holds all n keys
};

// FULL: begin_bit == 0 and end_bit == key bits (the default arguments): every digit is a whole byte at a constant
// shift, which takes the kernel from 26 to ~16 instructions per 32 keys (it is issue-bound once the atomics are
// conflict-free).
template <int KBYTES, bool IS_FLOAT, typename OffT, bool FULL>
__global__ void __launch_bounds__(HIST_THREADS) histogram_kernel(const HistParams<KBYTES, IS_FLOAT> P) {
  using KeyU = typename UIntOf<KBYTES>::type;
  using W = typename WideOf<KBYTES>::type;
  constexpr int MAXP = KBYTES;                 // passes at 8 bits per digit
  constexpr int KPV = 16 / KBYTES;             // keys per 128-bit vector
  constexpr int HIST_PARTS = HistSmem<KBYTES>::PARTS;
  extern __shared__ __align__(128) unsigned char hist_smem[];
  unsigned int (*bins)[RADIX][HIST_PARTS] = reinterpret_cast<unsigned int (*)[RADIX][HIST_PARTS]>(hist_smem);
  __shared__ bool s_last;

  const int tid = threadIdx.x;
  for (int i = tid; i < MAXP * RADIX * HIST_PARTS; i += HIST_THREADS) (&bins[0][0][0])[i] = 0;
  __syncthreads();

  const int part = tid & (HIST_PARTS - 1);
  const int np = P.num_passes;
  const int last_bits = P.end_bit - (P.begin_bit + 8 * (np - 1));
  const unsigned int last_mask = (1u << last_bits) - 1u;
  const auto op = P.op;

  auto count_key = [&](W k) {
    const W o = op.ordered(k);
    if (FULL) {
#pragma unroll
      for (int p = 0; p < MAXP; ++p) atomicAdd(&bins[p][(unsigned int)(o >> (8 * p)) & 255u][part], 1u);
      return;
    }
#pragma unroll
    for (int p = 0; p < MAXP; ++p) {
      if (p < np) {
        const unsigned int d = (unsigned int)(o >> (P.begin_bit + 8 * p)) & (p == np - 1 ? last_mask : 255u);
        atomicAdd(&bins[p][d][part], 1u);
      }
    }
  };

  const KeyU* keys = reinterpret_cast<const KeyU*>(P.keys);
  const unsigned long long n = P.n;
  // Peel the unaligned head (pointers are only element-aligned in the reference API).
  const uintptr_t addr = reinterpret_cast<uintptr_t>(keys);
  unsigned long long head = ((16 - (addr & 15)) & 15) / KBYTES;
  if (head > n) head = n;
  const unsigned long long nvec = (n - head) / KPV;
  const unsigned long long tail_start = head + nvec * KPV;

  if (blockIdx.x == 0) {
    for (unsigned long long i = tid; i < head; i += HIST_THREADS) count_key((W)keys[i]);
    for (unsigned long long i = tail_start + tid; i < n; i += HIST_THREADS) count_key((W)keys[i]);
  }

  const uint4* vec = reinterpret_cast<const uint4*>(keys + head);
  const unsigned long long stride = (unsigned long long)gridDim.x * HIST_THREADS;
  unsigned long long v = (unsigned long long)blockIdx.x * HIST_THREADS + tid;

  auto count_vec = [&](const uint4& q) {
    const unsigned int w[4] = {q.x, q.y, q.z, q.w};
    if (KBYTES == 8) {
      count_key((W)(((unsigned long long)w[1] << 32) | w[0]));
      count_key((W)(((unsigned long long)w[3] << 32) | w[2]));
    } else if (KBYTES == 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) count_key((W)w[j]);
    } else if (KBYTES == 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { count_key((W)(w[j] & 0xffffu)); count_key((W)(w[j] >> 16)); }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int b = 0; b < 4; ++b) count_key((W)((w[j] >> (8 * b)) & 0xffu));
      }
    }
  };

  // main loop: HIST_UNROLL independent 128-bit loads in flight
  for (; v + (HIST_UNROLL - 1) * stride < nvec; v += HIST_UNROLL * stride) {
    uint4 q[HIST_UNROLL];
#pragma unroll
    for (int j = 0; j < HIST_UNROLL; ++j) q[j] = __ldcs(vec + v + j * stride);
#pragma unroll
    for (int j = 0; j < HIST_UNROLL; ++j) count_vec(q[j]);
  }
  for (; v < nvec; v += stride) count_vec(__ldcs(vec + v));

  __syncthreads();
  OffT* ghist = reinterpret_cast<OffT*>(P.ghist);
  for (int i = tid; i < np * RADIX; i += HIST_THREADS) {
    const int p = i >> RADIX_BITS, d = i & (RADIX - 1);
    unsigned int c = 0;
#pragma unroll
    for (int q = 0; q < HIST_PARTS; ++q) c += bins[p][d][(q + tid) & (HIST_PARTS - 1)];  // rotated: conflict-free
    if (c) atomicAdd(reinterpret_cast<OffT*>(&ghist[i]), (OffT)c);
  }

  // last CTA done: exclusive scan of every pass' 256 counts, in place
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(P.done, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  __shared__ OffT warp_tot[HIST_THREADS / 32];
  const int lane = tid & 31;
  const int total = np * RADIX;
  for (int base_i = 0; base_i < total; base_i += HIST_THREADS) {  // HIST_THREADS covers whole passes
    const int i = base_i + tid;
    const bool active = i < total;
    const OffT c = active ? __ldcg(&ghist[i]) : OffT(0);
    // a pass whose digit is the same for every key moves nothing: the digit pass turns into a plain copy (the reference's
    // single-bin short circuit, cub/agent/agent_radix_sort_onesweep.cuh:344-420, decided once per pass instead of per tile)
    if (active && P.flags && (unsigned long long)c == P.n) P.flags[i >> RADIX_BITS] = 1u;
    OffT incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const OffT t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[tid >> 5] = incl;
    __syncthreads();
    // warps [first_warp, my_warp) belong to the same pass and precede this one
    const int my_warp = tid >> 5;
    const int first_warp = my_warp & ~(RADIX / 32 - 1);
    OffT base = 0;
    for (int w = first_warp; w < my_warp; ++w) base += warp_tot[w];
    if (active) ghist[i] = base + incl - c;
    __syncthreads();
  }
}

}  // namespace b2s
