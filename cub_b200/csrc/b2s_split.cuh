// b2s_split.cuh -- counting kernel of the multi-GPU partition pass (new functionality; the
// reference has no multi-GPU path).  counts[d] = number of local keys whose destination (SplitterOp) is d.
// The partition itself is the digit-pass kernel of b2s_onesweep.cuh instantiated with SplitterOp.
// Roofline: HBM read of n*K bytes.
#pragma once
#include "b2s_common.cuh"

namespace b2s {

// Keys are read once with 128-bit streaming loads (4 in flight per thread; element loads for the unaligned head and the
// tail).  Per key the thread adds the MAX_SPLITTERS "orders at or after splitter j" predicates to MAX_SPLITTERS counters --
// they are monotone in j, so counts[d] = ge[d-1] - ge[d] -- instead of deriving the destination and comparing it against
// every bucket: half the instructions of the obvious form, which is what bounds this kernel (not HBM) for 4-byte keys.
template <int KBYTES, typename OpT>
__global__ void __launch_bounds__(512) split_count_kernel(const void* keys_v, unsigned long long n, const OpT op_in,
                                                          unsigned long long* counts) {
  OpT op = op_in;
  op.prepare();
  using KeyU = typename UIntOf<KBYTES>::type;
  using W = typename WideOf<KBYTES>::type;
  constexpr int NS = OpT::MAX_SPLITTERS;
  constexpr int ND = NS + 1;
  constexpr int VEC = 16 / KBYTES;
  __shared__ unsigned long long s_ge[ND];  // [NS] = items seen
  if (threadIdx.x < ND) s_ge[threadIdx.x] = 0;
  __syncthreads();
  const KeyU* keys = reinterpret_cast<const KeyU*>(keys_v);
  unsigned int ge[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) ge[j] = 0;
  unsigned int seen = 0;
  const unsigned long long gtid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  unsigned long long head = ((16u - (unsigned int)(reinterpret_cast<uintptr_t>(keys) & 15u)) & 15u) / KBYTES;
  if (head > n) head = n;
  const unsigned long long nvec = (n - head) / VEC;
  const uint4* vk = reinterpret_cast<const uint4*>(keys + head);
  auto take = [&](W k) {
    op.add_ge(k, ge);
    ++seen;
  };
  auto take_vec = [&](const uint4& v) {
    if constexpr (KBYTES == 4) {
      take((W)v.x); take((W)v.y); take((W)v.z); take((W)v.w);
    } else {
      take((W)v.x | ((W)v.y << 32));
      take((W)v.z | ((W)v.w << 32));
    }
  };
  unsigned long long i = gtid;
  for (; i + 3 * stride < nvec; i += 4 * stride) {
    const uint4 a = __ldcs(vk + i), b = __ldcs(vk + i + stride), c = __ldcs(vk + i + 2 * stride), d = __ldcs(vk + i + 3 * stride);
    take_vec(a); take_vec(b); take_vec(c); take_vec(d);
  }
  for (; i < nvec; i += stride) take_vec(__ldcs(vk + i));
  if (gtid < head) take((W)keys[gtid]);
  const unsigned long long tail0 = head + nvec * VEC;
  if (tail0 + gtid < n) take((W)keys[tail0 + gtid]);
  // warp reduce, then one shared-memory atomic per warp and counter, one global atomic per CTA and counter
#pragma unroll
  for (int j = 0; j <= NS; ++j) {
    unsigned int v = j < NS ? ge[j < NS ? j : 0] : seen;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_ge[j], (unsigned long long)v);
  }
  __syncthreads();
  if (threadIdx.x < ND) {
    // counts[d] = ge[d-1] - ge[d] with ge[-1] = seen and ge[count..] = 0 (splitters beyond `count` never match)
    const int d = threadIdx.x;
    const unsigned long long hi = d == 0 ? s_ge[NS] : s_ge[d - 1];
    const unsigned long long lo = d < NS ? s_ge[d] : 0ull;
    if (hi != lo) atomicAdd(&counts[d], hi - lo);
  }
}

}  // namespace b2s
