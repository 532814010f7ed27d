// b2s_narrow.cu -- keys-only sorts of 1- and 2-byte keys over ALL their bits: a counting sort.
//
// The reference sorts such keys like any others: one histogram read plus one 8-bit digit pass per key byte
// (cub/device/dispatch/dispatch_radix_sort.cuh:1521-1727), i.e. K + 2*K*K bytes of HBM traffic per key (10 for 16-bit keys).
// When there are no values and every key bit takes part in the sort, two keys with the same digits in all passes are the SAME
// bit pattern, so that the stable order of equal keys cannot be observed in the result: the output is a function of the joint
// histogram of the keys alone and is produced here by
//   1. joint_hist16_kernel / the ordinary histogram kernel (8-bit keys): one read of the keys -> counts of all 2^bits images,
//   2. prefix16_kernel: exclusive prefix over the images in sort order,
//   3. expand_kernel: every 16-byte piece of the output looks up its image (binary search in the prefix array, L1/L2
//      resident) and is written with one 128-bit store,
// i.e. 2*K bytes of traffic per key.  The ONE exception to "same digits => same bits" are the floating zeros: -0.0 and +0.0
// share their digits (cub/block/radix_rank_sort_operations.cuh:55-66, 79-89) but not their bits, so the reference leaves them
// interleaved in input order inside one run.  Their images are neighbours (ZERO_IMG, HIGH), the expansion writes the run, and
// when BOTH occur (device-side flag written by the prefix kernel) two more kernels re-create the input order: zero_count16
// reads the keys once more and leaves one "is a zero" bit per key plus one count per CTA range, zero_write16 walks the bit
// masks and copies the zeros input -> run in order (both exit at once when one of the two zeros is absent).  Results are
// bit-identical to the digit passes.
//
// Roofline: HBM, algorithmic bytes per key = 2*K (+ K + 1/4 for the two zero kernels when both zeros occur).
#include <cuda_runtime.h>

#include <cstdlib>
#include <mutex>
#include <set>
#include <utility>

#include "b2s_common.cuh"
#include "b2s_internal.h"

namespace b2s {
namespace {

constexpr int NH_THREADS = 1024;
constexpr int NH_HALF = 32768;  // images counted per CTA (32-bit counters: 128 KB of shared memory)
constexpr int NH_UNROLL = 4;

// Joint histogram of 16-bit images.  65536 32-bit counters do not fit one SM, so CTAs work in pairs on the same chunk
// of keys: CTA 2c counts the images with a clear top bit, CTA 2c+1 the others (the second reader of a line hits L2).
// The image transform runs on two packed keys at a time; the pair's half is folded into the XOR constant so that
// "mine" is bit 15 / bit 31 of the transformed word.
template <bool IS_FLOAT, int MODE>
__global__ void __launch_bounds__(NH_THREADS, 1) joint_hist16_kernel(const uint16_t* __restrict__ keys, unsigned long long n,
                                                                   unsigned int xor16, unsigned long long* __restrict__ gbins) {
  extern __shared__ __align__(16) unsigned int nh_bins[];
  __shared__ unsigned int nh_dummy[NH_THREADS];
  constexpr bool NOBR = (MODE & 1) != 0;  // every lane issues the atomic: keys of the other half add to a private word (bank == lane)
  constexpr bool PREF = (MODE & 2) != 0;  // the next batch of loads is issued before the current one is counted
  const int tid = threadIdx.x;
  const unsigned int half = blockIdx.x & 1u;
  const unsigned int chunk = blockIdx.x >> 1, chunks = gridDim.x >> 1;
  for (int i = tid; i < NH_HALF; i += NH_THREADS) nh_bins[i] = 0;
  __syncthreads();
  const unsigned int bins_s = smem_u32(nh_bins);
  const unsigned int x1 = (xor16 & 0xffffu) ^ (half ? 0u : 0x8000u);
  const unsigned int x2 = x1 | (x1 << 16);
  // MODE bit 2: one dummy word per warp (all lanes of the other half on ONE address) instead of one per lane
  const unsigned int dummy_s = smem_u32(nh_dummy + ((MODE & 4) ? (tid & ~31) : tid));
  auto count_word = [&](unsigned int w) {
    unsigned int t;
    if (IS_FLOAT) {
      const unsigned int s = (w >> 15) & 0x00010001u;  // sign of either key
      t = w ^ ((s * 0xffffu) | 0x80008000u) ^ x2;      // negative: all bits flip, else the sign bit (Traits<fp>::TwiddleIn)
    } else {
      t = w ^ x2;
    }
    if (NOBR) {
      // ptxas wraps a predicated shared-memory atomic into BSSY / BRA / ATOMS / BSYNC (6 of 23 instructions per 32 keys and
      // the top stall reason, branch resolving); an unconditional one on a select of two addresses has no branch
      red_shared_add((t & 0x8000u) ? bins_s + ((t & 0x7fffu) << 2) : dummy_s, 1u);
      red_shared_add((t & 0x80000000u) ? bins_s + ((t >> 14) & 0x1fffcu) : dummy_s, 1u);
    } else {
      if (t & 0x8000u) red_shared_add(bins_s + ((t & 0x7fffu) << 2), 1u);
      if (t & 0x80000000u) red_shared_add(bins_s + ((t >> 14) & 0x1fffcu), 1u);
    }
  };
  auto count_vec = [&](const uint4& q) {
    count_word(q.x);
    count_word(q.y);
    count_word(q.z);
    count_word(q.w);
  };

  const uintptr_t addr = reinterpret_cast<uintptr_t>(keys);
  unsigned long long head = ((16 - (addr & 15)) & 15) / 2;
  if (head > n) head = n;
  const unsigned long long nvec = (n - head) / 8;
  const unsigned long long tail_start = head + nvec * 8;
  if (chunk == 0) {
    // element-wise head and tail: one key in the low half, the high half is neutralised by testing only bit 15
    auto count_one = [&](unsigned int k) {
      unsigned int t;
      if (IS_FLOAT) {
        const unsigned int s = (k >> 15) & 1u;
        t = (k ^ ((s * 0xffffu) | 0x8000u) ^ x1) & 0xffffu;
      } else {
        t = (k ^ x1) & 0xffffu;
      }
      if (t & 0x8000u) red_shared_add(bins_s + ((t & 0x7fffu) << 2), 1u);
    };
    for (unsigned long long i = tid; i < head; i += NH_THREADS) count_one(keys[i]);
    for (unsigned long long i = tail_start + tid; i < n; i += NH_THREADS) count_one(keys[i]);
  }

  const uint4* vec = reinterpret_cast<const uint4*>(keys + head);
  const unsigned long long stride = (unsigned long long)chunks * NH_THREADS;
  unsigned long long v = (unsigned long long)chunk * NH_THREADS + tid;
  if (PREF) {
    uint4 q[NH_UNROLL];
    bool have = v + (NH_UNROLL - 1) * stride < nvec;
    if (have) {
#pragma unroll
      for (int j = 0; j < NH_UNROLL; ++j) q[j] = __ldg(vec + v + j * stride);
    }
    while (have) {
      const unsigned long long vn = v + NH_UNROLL * stride;
      const bool have_next = vn + (NH_UNROLL - 1) * stride < nvec;
      uint4 qn[NH_UNROLL];
      if (have_next) {
#pragma unroll
        for (int j = 0; j < NH_UNROLL; ++j) qn[j] = __ldg(vec + vn + j * stride);
      }
#pragma unroll
      for (int j = 0; j < NH_UNROLL; ++j) count_vec(q[j]);
#pragma unroll
      for (int j = 0; j < NH_UNROLL; ++j) q[j] = qn[j];
      v = vn;
      have = have_next;
    }
  } else {
    for (; v + (NH_UNROLL - 1) * stride < nvec; v += NH_UNROLL * stride) {
      uint4 q[NH_UNROLL];
#pragma unroll
      for (int j = 0; j < NH_UNROLL; ++j) q[j] = __ldg(vec + v + j * stride);
#pragma unroll
      for (int j = 0; j < NH_UNROLL; ++j) count_vec(q[j]);
    }
  }
  for (; v < nvec; v += stride) count_vec(__ldg(vec + v));
  __syncthreads();
  unsigned long long* mine = gbins + (size_t)half * NH_HALF;
  for (int i = tid; i < NH_HALF; i += NH_THREADS) {
    const unsigned int c = nh_bins[i];
    if (c) atomicAdd(mine + i, (unsigned long long)c);
  }
}

// counts of the 65536 images -> exclusive prefix (+ total at [65536]); floating keys: flag "both zeros occur".
// 64 CTAs of 1024 images; CTA c first sums the counts of ALL images before its own (c * 8 KB of L2-resident reads, all
// independent) and then scans its own 1024 -- no ordering between CTAs, which is why the prefix is a second array.
constexpr int PX_CTAS = 64;
__global__ void __launch_bounds__(1024) prefix16_kernel(const unsigned long long* __restrict__ counts,
                                                        unsigned long long* __restrict__ prefix, unsigned int* __restrict__ zflag) {
  __shared__ unsigned long long s_warp[2][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (zflag != nullptr && blockIdx.x == 0 && tid == 0) *zflag = (counts[0x7fff] != 0 && counts[0x8000] != 0) ? 1u : 0u;
  unsigned long long before = 0;
#pragma unroll 8
  for (unsigned int c = 0; c < blockIdx.x; ++c) before += __ldcg(counts + c * 1024 + tid);
  const unsigned long long mine = __ldcg(counts + blockIdx.x * 1024 + tid);
  unsigned long long incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
  if (lane == 31) s_warp[0][warp] = incl;
  if (lane == 0) s_warp[1][warp] = before;
  __syncthreads();
  unsigned long long base = 0;
#pragma unroll
  for (int w = 0; w < 32; ++w) {
    base += s_warp[1][w];
    if (w < warp) base += s_warp[0][w];
  }
  prefix[blockIdx.x * 1024 + tid] = base + incl - mine;
  if (blockIdx.x == PX_CTAS - 1 && tid == 1023) prefix[65536] = base + incl;
}

// ---- floating zeros: stable compaction of the zero-like input keys into their run of the output --------------------
// Input order = [head keys up to the first 16-byte boundary][128-bit vectors][tail keys].  The vectors are cut into tiles of
// Z_TILE_VECS, every CTA owns a contiguous range of tiles.  zero_count16 reads the keys once and leaves (a) one bit per key
// ("is +-0.0"; one byte per vector) and (b) one count per CTA; zero_write16 then works on the bit masks alone (1/16 of the
// keys' bytes): counts of the CTAs before it (at most 1024), a block-wide exclusive scan per tile, and one 2-byte copy
// input -> output per set bit.
constexpr int Z_THREADS = 512;
constexpr int Z_MPT = 16;                         // mask bytes (= vectors) per thread and tile
constexpr int Z_TILE_VECS = Z_THREADS * Z_MPT;    // 8192 vectors = 65536 keys
constexpr int Z_MAX_CTAS = 1024;

struct ZeroGeom {
  unsigned long long head, nvec, tail_start, tiles, tiles_per_cta;
};
__device__ __forceinline__ ZeroGeom zero_geom(const uint16_t* keys, unsigned long long n) {
  ZeroGeom g;
  const uintptr_t addr = reinterpret_cast<uintptr_t>(keys);
  g.head = ((16 - (addr & 15)) & 15) / 2;
  if (g.head > n) g.head = n;
  g.nvec = (n - g.head) / 8;
  g.tail_start = g.head + g.nvec * 8;
  g.tiles = (g.nvec + Z_TILE_VECS - 1) / Z_TILE_VECS;
  g.tiles_per_cta = (g.tiles + gridDim.x - 1) / gridDim.x;
  return g;
}
// two flag bits of a word of two packed keys: bit 0 = low key is +-0.0, bit 1 = high key is +-0.0
__device__ __forceinline__ unsigned int zero_bits(unsigned int w) {
  const unsigned int nz = ((w & 0x7fff7fffu) + 0x7fff7fffu) & 0x80008000u;  // bit 15 / 31: the half is NOT zero (no carry between halves)
  return (((nz >> 15) | (nz >> 30)) & 3u) ^ 3u;
}

__global__ void __launch_bounds__(Z_THREADS) zero_count16_kernel(const uint16_t* __restrict__ keys, unsigned long long n,
                                                                  const unsigned int* __restrict__ zflag,
                                                                  unsigned long long* __restrict__ partial,
                                                                  unsigned char* __restrict__ masks) {
  if (*zflag == 0) return;
  __shared__ unsigned int s_warp[Z_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const ZeroGeom g = zero_geom(keys, n);
  const uint4* vec = reinterpret_cast<const uint4*>(keys + g.head);
  unsigned long long t = (unsigned long long)blockIdx.x * g.tiles_per_cta, t_end = t + g.tiles_per_cta;
  if (t_end > g.tiles) t_end = g.tiles;
  unsigned int cnt = 0;  // a CTA's share stays far below 2^32
  for (; t < t_end; ++t) {
#pragma unroll 1
    for (int r = 0; r < Z_MPT; r += 4) {
      const unsigned long long v0 = t * Z_TILE_VECS + (unsigned long long)r * Z_THREADS + tid;
      uint4 q[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned long long v = v0 + (unsigned long long)j * Z_THREADS;
        q[j] = v < g.nvec ? __ldcs(vec + v) : make_uint4(0x00010001u, 0x00010001u, 0x00010001u, 0x00010001u);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned int m = zero_bits(q[j].x) | (zero_bits(q[j].y) << 2) | (zero_bits(q[j].z) << 4) | (zero_bits(q[j].w) << 6);
        cnt += __popc(m);
        masks[v0 + (unsigned long long)j * Z_THREADS] = (unsigned char)m;  // every byte of every tile is written (padding: 0)
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) s_warp[warp] = cnt;
  __syncthreads();
  if (tid == 0) {
    unsigned long long tot = 0;
    for (int w = 0; w < Z_THREADS / 32; ++w) tot += s_warp[w];
    partial[blockIdx.x] = tot;
  }
}

__global__ void __launch_bounds__(Z_THREADS) zero_write16_kernel(const uint16_t* __restrict__ keys, uint16_t* __restrict__ out,
                                                                  unsigned long long n, const unsigned int* __restrict__ zflag,
                                                                  const unsigned long long* __restrict__ partial,
                                                                  const unsigned char* __restrict__ masks,
                                                                  const unsigned long long* __restrict__ prefix) {
  if (*zflag == 0) return;
  __shared__ unsigned long long s_red[Z_THREADS / 32];
  __shared__ unsigned int s_warp[2][Z_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool last_cta = blockIdx.x == gridDim.x - 1;
  const ZeroGeom g = zero_geom(keys, n);
  unsigned long long base = prefix[0x7fff];  // the zeros' run starts at the smaller of their two images
  // head keys come first in input order
  unsigned int hz = 0;
  for (unsigned long long i = 0; i < g.head; ++i) hz += (keys[i] & 0x7fffu) == 0 ? 1u : 0u;
  if (blockIdx.x == 0 && tid == 0) {
    unsigned long long p = base;
    for (unsigned long long i = 0; i < g.head; ++i)
      if ((keys[i] & 0x7fffu) == 0) out[p++] = keys[i];
  }
  if (partial[blockIdx.x] == 0 && !last_cta) return;  // no zero in this CTA's tiles
  // zeros of the CTAs before this one
  unsigned long long before = 0;
  for (unsigned int c = tid; c < blockIdx.x; c += Z_THREADS) before += partial[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
  if (lane == 0) s_red[warp] = before;
  __syncthreads();
  before = 0;
#pragma unroll
  for (int w = 0; w < Z_THREADS / 32; ++w) before += s_red[w];
  base += hz + before;

  unsigned long long t = (unsigned long long)blockIdx.x * g.tiles_per_cta, t_end = t + g.tiles_per_cta;
  if (t_end > g.tiles) t_end = g.tiles;
  int flip = 0;
  for (; t < t_end; ++t, flip ^= 1) {
    // 16 consecutive mask bytes = 128 consecutive keys per thread: input order = thread order, then bit order
    const unsigned long long v0 = t * Z_TILE_VECS + (unsigned long long)tid * Z_MPT;
    const uint4 m4 = __ldg(reinterpret_cast<const uint4*>(masks + v0));
    const unsigned int c = __popc(m4.x) + __popc(m4.y) + __popc(m4.z) + __popc(m4.w);
    unsigned int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) s_warp[flip][warp] = incl;
    __syncthreads();  // one barrier per tile: the warp totals alternate between two arrays
    unsigned int wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < Z_THREADS / 32; ++w) {
      const unsigned int u = s_warp[flip][w];
      if (w < warp) wbase += u;
      total += u;
    }
    if (c) {
      unsigned long long p = base + wbase + incl - c;
      const uint16_t* src = keys + g.head + v0 * 8;
      const unsigned int mw[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        unsigned int m = mw[k];
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          out[p++] = src[k * 32 + b];
        }
      }
    }
    base += total;
  }
  if (last_cta && tid == 0) {  // tail keys come last
    for (unsigned long long i = g.tail_start; i < n; ++i)
      if ((keys[i] & 0x7fffu) == 0) out[base++] = keys[i];
  }
}

// ---- expansion -------------------------------------------------------------------------------------------------
constexpr int EX_THREADS = 512;

template <int KBYTES, bool IS_FLOAT, typename OffT>
__global__ void __launch_bounds__(EX_THREADS) expand_kernel(void* __restrict__ out_, unsigned long long n,
                                                            const OffT* __restrict__ prefix, unsigned int xor_mask) {
  using KeyU = typename UIntOf<KBYTES>::type;
  constexpr unsigned int NB = 1u << (8 * KBYTES);
  constexpr int KPC = 16 / KBYTES;  // keys per 16-byte piece
  KeyU* out = reinterpret_cast<KeyU*>(out_);

  auto end_of = [&](unsigned int b) -> unsigned long long { return b + 1 < NB ? (unsigned long long)__ldg(prefix + b + 1) : n; };
  // largest image b >= lo with prefix[b] <= p (prefix[lo] <= p holds on entry): the non-empty bin that contains position p
  auto find = [&](unsigned long long p, unsigned int lo) -> unsigned int {
    unsigned int hi = NB;
    while (hi - lo > 1) {
      const unsigned int mid = (lo + hi) >> 1;
      if ((unsigned long long)__ldg(prefix + mid) <= p) lo = mid;
      else hi = mid;
    }
    return lo;
  };
  auto raw_of = [&](unsigned int b) -> unsigned int {
    if constexpr (IS_FLOAT) {
      OrderedFloatOp<KBYTES> op{};
      op.xor_mask = xor_mask;
      return (unsigned int)op.to_raw(b);
    } else {
      return (b ^ xor_mask) & (NB - 1u);
    }
  };
  auto word_of = [&](unsigned int b) -> unsigned int { return raw_of(b) * (KBYTES == 2 ? 0x00010001u : 0x01010101u); };
  // position p lies at or after the end e of bin b: step to the bin that holds it -- the next few bins one by one (the
  // common case: neighbouring non-empty bins), then by bisection (long stretches of empty bins)
  auto advance = [&](unsigned long long p, unsigned int& b, unsigned long long& e) {
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
      ++b;
      e = end_of(b);
      if (p < e) return;
    }
    b = find(p, b + 1);
    e = end_of(b);
  };

  const uintptr_t addr = reinterpret_cast<uintptr_t>(out);
  unsigned long long head = ((16 - (addr & 15)) & 15) / KBYTES;
  if (head > n) head = n;
  const unsigned long long pieces = (n - head) / KPC;
  const unsigned long long tail_start = head + pieces * KPC;
  if (blockIdx.x == 0) {  // element-wise head and tail (at most 2 * (KPC - 1) keys)
    const unsigned long long extra = head + (n - tail_start);
    for (unsigned long long j = threadIdx.x; j < extra; j += EX_THREADS) {
      const unsigned long long p = j < head ? j : tail_start + (j - head);
      out[p] = (KeyU)raw_of(find(p, 0));
    }
  }

  // every warp owns a contiguous range of rows (a row = 32 pieces = 512 bytes): positions advance slowly, so that the
  // current bin is found by one compare most of the time
  const unsigned long long rows = (pieces + 31) / 32;
  const unsigned long long warps = (unsigned long long)gridDim.x * (EX_THREADS / 32);
  const unsigned long long rpw = (rows + warps - 1) / warps;
  const unsigned long long gw = (unsigned long long)blockIdx.x * (EX_THREADS / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  unsigned long long r = gw * rpw;
  unsigned long long r_end = r + rpw;
  if (r_end > rows) r_end = rows;
  if (r >= r_end) return;

  unsigned int b = 0;
  unsigned long long e = 0;  // end of bin b; 0 forces the first search
  unsigned int word = 0;
  for (; r < r_end; ++r) {
    const unsigned long long c = r * 32 + lane;
    if (c >= pieces) break;
    const unsigned long long p = head + c * KPC;
    if (p >= e) {
      if (e == 0) {
        b = find(p, 0);
        e = end_of(b);
      } else {
        advance(p, b, e);
      }
      word = word_of(b);
    }
    uint4 q;
    if (p + KPC <= e) {
      q = make_uint4(word, word, word, word);
    } else {
      unsigned int w[4] = {0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < KPC; ++j) {
        const unsigned long long pj = p + j;
        if (pj >= e) advance(pj, b, e);
        w[j / (KPC / 4)] |= raw_of(b) << (8 * KBYTES * (j % (KPC / 4)));
      }
      word = word_of(b);
      q = make_uint4(w[0], w[1], w[2], w[3]);
    }
    *reinterpret_cast<uint4*>(out + p) = q;
  }
}

template <int KBYTES, bool IS_FLOAT>
cudaError_t expand_launch(const NarrowArgs& a, cudaStream_t s) {
  const int grid = a.sms * 4;
  if (a.prefix64)
    expand_kernel<KBYTES, IS_FLOAT, unsigned long long><<<grid, EX_THREADS, 0, s>>>(
        a.keys_out, a.n, reinterpret_cast<const unsigned long long*>(a.prefix), (unsigned int)a.dc.xor_mask);
  else
    expand_kernel<KBYTES, IS_FLOAT, unsigned int><<<grid, EX_THREADS, 0, s>>>(
        a.keys_out, a.n, reinterpret_cast<const unsigned int*>(a.prefix), (unsigned int)a.dc.xor_mask);
  return cudaGetLastError();
}

}  // namespace

size_t narrow_zero_mask_bytes(uint64_t n) { return (size_t)((n / 8 + Z_TILE_VECS - 1) / Z_TILE_VECS) * Z_TILE_VECS; }

cudaError_t narrow_step(NarrowStep step, const NarrowArgs& a, cudaStream_t s) {
  const uint16_t* k16 = reinterpret_cast<const uint16_t*>(a.keys_in);
  unsigned long long* bins = reinterpret_cast<unsigned long long*>(a.prefix);      // exclusive prefix (what the later steps read)
  unsigned long long* counts = reinterpret_cast<unsigned long long*>(a.counts);  // 2-byte keys: the joint histogram
  const bool fl = a.dc.is_float;
  // CTAs of the two zero kernels (they must agree): four per SM, at most one per tile
  auto zero_grid = [&]() -> unsigned int {
    const uint64_t tiles = (a.n / 8 + Z_TILE_VECS - 1) / Z_TILE_VECS;
    uint64_t g = (uint64_t)a.sms * 4;
    if (g > (uint64_t)Z_MAX_CTAS) g = Z_MAX_CTAS;
    if (g > tiles) g = tiles;
    return (unsigned int)(g ? g : 1);
  };
  switch (step) {
    case NarrowStep::kHist16: {
      const size_t smem = (size_t)NH_HALF * 4;
      using Kern = void (*)(const uint16_t*, unsigned long long, unsigned int, unsigned long long*);
#ifdef B2S_TUNING
      // tuning library only: A/B switch of the histogram's inner loop (MODE bits at joint_hist16_kernel; bench/runs/counting_modes.sh)
      static const int mode = [] {
        const char* e = std::getenv("B2S_NH_MODE");
        return e ? (std::atoi(e) & 7) : 5;
      }();
      const Kern table[12] = {joint_hist16_kernel<false, 0>, joint_hist16_kernel<false, 1>, joint_hist16_kernel<false, 2>,
                              joint_hist16_kernel<false, 3>, joint_hist16_kernel<false, 5>, joint_hist16_kernel<false, 7>,
                              joint_hist16_kernel<true, 0>,  joint_hist16_kernel<true, 1>,  joint_hist16_kernel<true, 2>,
                              joint_hist16_kernel<true, 3>,  joint_hist16_kernel<true, 5>,  joint_hist16_kernel<true, 7>};
      const int slot = mode < 4 ? mode : (mode == 5 ? 4 : (mode == 7 ? 5 : 1));
      const Kern kern = table[(fl ? 6 : 0) + slot];
#else
      const Kern kern = fl ? joint_hist16_kernel<true, 5> : joint_hist16_kernel<false, 5>;  // branch-free atomics, one dummy word per warp
#endif
      {  // opt in to 128 KB of dynamic shared memory once per (kernel, device)
        static std::mutex mu;
        static std::set<std::pair<const void*, int>> done;
        int dev = 0;
        cudaGetDevice(&dev);
        const std::pair<const void*, int> key(reinterpret_cast<const void*>(kern), dev);
        std::lock_guard<std::mutex> lock(mu);
        if (!done.count(key)) {
          const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e != cudaSuccess) return e;
          done.insert(key);
        }
      }
      // pairs of CTAs; no more pairs than 16 KB chunks of keys
      uint64_t pairs = (uint64_t)(a.sms / 2 > 0 ? a.sms / 2 : 1);
      const uint64_t chunks = (a.n * 2 + 16 * 1024 - 1) / (16 * 1024);
      if (pairs > chunks) pairs = chunks ? chunks : 1;
      kern<<<(unsigned int)(2 * pairs), NH_THREADS, smem, s>>>(k16, a.n, (unsigned int)a.dc.xor_mask, counts);
      return cudaGetLastError();
    }
    case NarrowStep::kPrefix16:
      prefix16_kernel<<<PX_CTAS, 1024, 0, s>>>(counts, bins, fl ? a.zflag : nullptr);
      return cudaGetLastError();
    case NarrowStep::kZeroCount:
      zero_count16_kernel<<<zero_grid(), Z_THREADS, 0, s>>>(k16, a.n, a.zflag, a.zpartial, a.zmasks);
      return cudaGetLastError();
    case NarrowStep::kExpand:
      if (a.kbytes == 1) return expand_launch<1, false>(a, s);
      return fl ? expand_launch<2, true>(a, s) : expand_launch<2, false>(a, s);
    case NarrowStep::kZeroWrite:
      zero_write16_kernel<<<zero_grid(), Z_THREADS, 0, s>>>(k16, reinterpret_cast<uint16_t*>(a.keys_out), a.n, a.zflag,
                                                            a.zpartial, a.zmasks, bins);
      return cudaGetLastError();
  }
  return cudaErrorInvalidValue;
}

}  // namespace b2s
