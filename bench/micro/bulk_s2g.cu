// bench/micro/bulk_s2g.cu -- how fast can an SM issue SMALL shared->global bulk copies (cp.async.bulk.global.shared::cta)?
// Question behind it: could the digit pass write its ~256 digit runs per tile (avg ~120 B) with the TMA engine instead of
// LDS + STG through the LSU pipe?   nvcc -O3 -gencode arch=compute_100a,code=sm_100a bulk_s2g.cu -o bulk_s2g && ./bulk_s2g
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int MODE>  // 0: bulk copies, one per thread per array; 1: plain coalesced STG of the same bytes (baseline)
__global__ void __launch_bounds__(384, 3) k(char* out, int len, int copies_per_thread, int iters, long long* cycles) {
  extern __shared__ __align__(128) char smem[];
  const int tid = threadIdx.x;
  for (int i = tid; i < 60 * 1024 / 4; i += 384) reinterpret_cast<int*>(smem)[i] = i;
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  char* base = out + (size_t)blockIdx.x * 256 * copies_per_thread * len;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      if (tid < 256) {
        for (int c = 0; c < copies_per_thread; ++c) {
          const int slot = tid * copies_per_thread + c;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + (size_t)slot * len),
                       "r"(smem_u32(smem + (slot * len) % (60 * 1024 - 256))), "r"(len)
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    } else {
      const int total = 256 * copies_per_thread * len / 4;
      for (int i = tid; i < total; i += 384) reinterpret_cast<int*>(base)[i] = reinterpret_cast<int*>(smem)[i % (15 * 1024)];
    }
    __syncthreads();
  }
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
  const int grid = 148 * 3, iters = 200;
  char* out;
  long long* cyc;
  cudaMalloc(&out, (size_t)grid * 256 * 2 * 512 + (1 << 20));
  cudaMalloc(&cyc, grid * sizeof(long long));
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode)
    for (int len : {32, 64, 96, 128, 160, 256, 512}) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<grid, 384, 60 * 1024>>>(out, len, 2, iters, cyc);
        else k<1><<<grid, 384, 60 * 1024>>>(out, len, 2, iters, cyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double copies = (double)grid * 512 * iters;
      printf("%s len %3d B: %.3f ms  %.1f copies/us/SM  %.0f GB/s  (%s)\n", mode ? "STG " : "BULK", len, ms,
             copies / (ms * 1e3) / 148, copies * len / (ms * 1e6), cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
