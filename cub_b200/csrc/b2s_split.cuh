// b2s_split.cuh -- counting kernel of the multi-GPU partition pass (new functionality; the
// reference has no multi-GPU path).  counts[d] = number of local keys whose destination (SplitterOp) is d.
// The partition itself is the digit-pass kernel of b2s_onesweep.cuh instantiated with SplitterOp.
// Roofline: HBM read of n*K bytes.
#pragma once
#include "b2s_common.cuh"

namespace b2s {

template <int KBYTES, typename OpT>
__global__ void __launch_bounds__(1024) split_count_kernel(const void* keys_v, unsigned long long n, const OpT op_in,
                                                           unsigned long long* counts) {
  OpT op = op_in;
  op.prepare();
  using KeyU = typename UIntOf<KBYTES>::type;
  using W = typename WideOf<KBYTES>::type;
  constexpr int ND = OpT::MAX_SPLITTERS + 1;
  __shared__ unsigned long long s_counts[ND];
  if (threadIdx.x < ND) s_counts[threadIdx.x] = 0;
  __syncthreads();
  const KeyU* keys = reinterpret_cast<const KeyU*>(keys_v);
  unsigned int c[ND];
#pragma unroll
  for (int j = 0; j < ND; ++j) c[j] = 0;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned int d = op((W)__ldcs(keys + i));
#pragma unroll
    for (int j = 0; j < ND; ++j) c[j] += (d == (unsigned int)j) ? 1u : 0u;
  }
#pragma unroll
  for (int j = 0; j < ND; ++j) {
    unsigned int v = c[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_counts[j], (unsigned long long)v);
  }
  __syncthreads();
  if (threadIdx.x < ND && s_counts[threadIdx.x]) atomicAdd(&counts[threadIdx.x], s_counts[threadIdx.x]);
}

}  // namespace b2s
