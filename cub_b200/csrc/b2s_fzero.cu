// b2s_fzero.cu -- "zero recording" for full-range sorts of 4- and 8-byte floating keys: the restore step.
//
// The reference treats -0.0 and +0.0 as one key in every digit (cub/block/radix_rank_sort_operations.cuh:55-66, 79-89) but
// keeps their bits; it pays for that with a compare in every digit extraction of every pass.  Here the first digit pass
// (ImageFloatOp, b2s_common.cuh / b2s_pass.cuh) gives both zeros ONE image and writes two bit planes in input order -- Z: the
// key is a zero, S: its sign bit -- after which the zeros are ordinary equal keys: the passes are stable, so they arrive at the
// output as one run in input order, at the offset of digit 0x80 of the top pass (the zeros' image is the smallest one with that
// top digit).  The two kernels below give the run its signs back: per-CTA counts of the Z plane, then every CTA walks its
// contiguous range of the planes (one bit per key each: n/8 bytes) and stores +0.0 or -0.0 at run start + rank.
// Cost: two ballots per 32 keys in the first pass and ~n/4 bytes of extra traffic; gain: digits cost 2 instructions instead
// of 4 in every pass, the passes in between ARE the integer kernels with their shapes.
#include <cuda_runtime.h>

#include "b2s_common.cuh"
#include "b2s_internal.h"

namespace b2s {
namespace {

constexpr int FZ_THREADS = 512;
constexpr int FZ_WPT = 4;                        // plane words per thread and tile (one 128-bit load)
constexpr int FZ_TILE = FZ_THREADS * FZ_WPT;     // 2048 words = 65536 keys
constexpr int FZ_MAX_CTAS = 1024;

struct FzGeom {
  unsigned long long words, tiles, tiles_per_cta;
};
__device__ __forceinline__ FzGeom fz_geom(unsigned long long n) {
  FzGeom g;
  g.words = (n + 31) / 32;
  g.tiles = (g.words + FZ_TILE - 1) / FZ_TILE;
  g.tiles_per_cta = (g.tiles + gridDim.x - 1) / gridDim.x;
  return g;
}
__device__ __forceinline__ uint4 fz_load(const unsigned int* plane, unsigned long long w0, unsigned long long words) {
  // planes are 16-byte aligned and padded to whole tiles by the caller's layout, but only words < `words` were written
  uint4 q = make_uint4(0, 0, 0, 0);
  if (w0 + 3 < words) {
    q = __ldg(reinterpret_cast<const uint4*>(plane + w0));
  } else {
    if (w0 < words) q.x = __ldg(plane + w0);
    if (w0 + 1 < words) q.y = __ldg(plane + w0 + 1);
    if (w0 + 2 < words) q.z = __ldg(plane + w0 + 2);
  }
  return q;
}

__global__ void __launch_bounds__(FZ_THREADS) fzero_count_kernel(const unsigned int* __restrict__ zero_z, unsigned long long n,
                                                                 unsigned long long* __restrict__ partial) {
  __shared__ unsigned int s_warp[FZ_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const FzGeom g = fz_geom(n);
  unsigned long long t = (unsigned long long)blockIdx.x * g.tiles_per_cta, t_end = t + g.tiles_per_cta;
  if (t_end > g.tiles) t_end = g.tiles;
  unsigned int cnt = 0;  // a CTA's share of the keys stays far below 2^32
  for (; t < t_end; ++t) {
    const uint4 q = fz_load(zero_z, t * FZ_TILE + (unsigned long long)tid * FZ_WPT, g.words);
    cnt += __popc(q.x) + __popc(q.y) + __popc(q.z) + __popc(q.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) s_warp[warp] = cnt;
  __syncthreads();
  if (tid == 0) {
    unsigned long long tot = 0;
    for (int w = 0; w < FZ_THREADS / 32; ++w) tot += s_warp[w];
    partial[blockIdx.x] = tot;
  }
}

template <typename KeyU, typename OffT>
__global__ void __launch_bounds__(FZ_THREADS) fzero_write_kernel(const unsigned int* __restrict__ zero_z,
                                                                 const unsigned int* __restrict__ zero_s, unsigned long long n,
                                                                 const unsigned long long* __restrict__ partial,
                                                                 KeyU* __restrict__ out, const OffT* __restrict__ top_bins) {
  if (partial[blockIdx.x] == 0) return;  // no zero among this CTA's keys
  __shared__ unsigned long long s_red[FZ_THREADS / 32];
  __shared__ unsigned int s_warp[2][FZ_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const FzGeom g = fz_geom(n);
  unsigned long long before = 0;
  for (unsigned int c = tid; c < blockIdx.x; c += FZ_THREADS) before += partial[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
  if (lane == 0) s_red[warp] = before;
  __syncthreads();
  before = 0;
#pragma unroll
  for (int w = 0; w < FZ_THREADS / 32; ++w) before += s_red[w];
  // the zeros' image is the smallest one whose top digit is 0x80: their run starts where that digit starts
  unsigned long long base = (unsigned long long)top_bins[0x80] + before;
  constexpr KeyU NEG_ZERO = (KeyU)1 << (sizeof(KeyU) * 8 - 1);

  unsigned long long t = (unsigned long long)blockIdx.x * g.tiles_per_cta, t_end = t + g.tiles_per_cta;
  if (t_end > g.tiles) t_end = g.tiles;
  int flip = 0;
  for (; t < t_end; ++t, flip ^= 1) {
    // 4 consecutive plane words = 128 consecutive keys per thread: input order = thread order, then bit order
    const unsigned long long w0 = t * FZ_TILE + (unsigned long long)tid * FZ_WPT;
    const uint4 z4 = fz_load(zero_z, w0, g.words);
    const unsigned int c = __popc(z4.x) + __popc(z4.y) + __popc(z4.z) + __popc(z4.w);
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
    for (int w = 0; w < FZ_THREADS / 32; ++w) {
      const unsigned int u = s_warp[flip][w];
      if (w < warp) wbase += u;
      total += u;
    }
    if (c) {
      unsigned long long p = base + wbase + incl - c;
      const unsigned int zw[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        unsigned int m = zw[k];
        if (m) {
          const unsigned int sgn = __ldg(zero_s + w0 + k);
          while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            out[p++] = ((sgn >> b) & 1u) ? NEG_ZERO : (KeyU)0;
          }
        }
      }
    }
    base += total;
  }
}

unsigned int fz_grid(const FzeroArgs& a) {
  const uint64_t words = (a.n + 31) / 32;
  const uint64_t tiles = (words + FZ_TILE - 1) / FZ_TILE;
  uint64_t g = (uint64_t)a.sms * 4;
  if (g > (uint64_t)FZ_MAX_CTAS) g = FZ_MAX_CTAS;
  if (g > tiles) g = tiles;
  return (unsigned int)(g ? g : 1);
}

}  // namespace

// words per plane: one per row of 32 keys of every tile of the first pass (tiles are at most 2^16 items), rounded for the
// 128-bit loads of the restore kernels
size_t fzero_plane_words(uint64_t n) { return (size_t)(((n + 65536) / 32 + 64 + 3) & ~(uint64_t)3); }

cudaError_t fzero_count_launch(const FzeroArgs& a, cudaStream_t s) {
  fzero_count_kernel<<<fz_grid(a), FZ_THREADS, 0, s>>>(a.zero_z, a.n, a.partial);
  return cudaGetLastError();
}

cudaError_t fzero_write_launch(const FzeroArgs& a, cudaStream_t s) {
  const unsigned int grid = fz_grid(a);
  if (a.kbytes == 4) {
    if (a.off64)
      fzero_write_kernel<uint32_t, unsigned long long><<<grid, FZ_THREADS, 0, s>>>(
          a.zero_z, a.zero_s, a.n, a.partial, reinterpret_cast<uint32_t*>(a.keys_out), reinterpret_cast<const unsigned long long*>(a.top_bins));
    else
      fzero_write_kernel<uint32_t, unsigned int><<<grid, FZ_THREADS, 0, s>>>(
          a.zero_z, a.zero_s, a.n, a.partial, reinterpret_cast<uint32_t*>(a.keys_out), reinterpret_cast<const unsigned int*>(a.top_bins));
  } else {
    if (a.off64)
      fzero_write_kernel<unsigned long long, unsigned long long><<<grid, FZ_THREADS, 0, s>>>(
          a.zero_z, a.zero_s, a.n, a.partial, reinterpret_cast<unsigned long long*>(a.keys_out),
          reinterpret_cast<const unsigned long long*>(a.top_bins));
    else
      fzero_write_kernel<unsigned long long, unsigned int><<<grid, FZ_THREADS, 0, s>>>(
          a.zero_z, a.zero_s, a.n, a.partial, reinterpret_cast<unsigned long long*>(a.keys_out),
          reinterpret_cast<const unsigned int*>(a.top_bins));
  }
  return cudaGetLastError();
}

}  // namespace b2s
