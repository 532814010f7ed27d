// prim.cu -- microbenchmarks of the warp/shared-memory primitives the digit-pass kernel is built from.
// Reports cycles per warp-instruction per SM (all SMs busy, WARPS warps per SM) so that design choices
// (hardware MATCH.ANY vs ballot rounds, leader atomics, random shared-memory scatter width) rest on measurements.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o prim prim.cu && ./prim
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 512;
constexpr int NT = 512;

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}
template <int I>
__device__ __forceinline__ void match_bit(uint32_t& m, uint32_t d) {
  asm volatile("{\n .reg .pred p;\n .reg .b32 t, b;\n and.b32 t, %1, %2;\n setp.ne.u32 p, t, 0;\n"
               " vote.sync.ballot.b32 b, p, 0xffffffff;\n @!p not.b32 b, b;\n and.b32 %0, %0, b;\n}\n"
               : "+r"(m) : "r"(d), "n"(1u << I));
}
template <int BITS>
__device__ __forceinline__ uint32_t match_ballot(uint32_t d) {
  uint32_t m = 0xffffffffu;
  match_bit<0>(m, d);
  if (BITS > 1) match_bit<1>(m, d);
  if (BITS > 2) match_bit<2>(m, d);
  if (BITS > 3) match_bit<3>(m, d);
  if (BITS > 4) match_bit<4>(m, d);
  if (BITS > 5) match_bit<5>(m, d);
  if (BITS > 6) match_bit<6>(m, d);
  if (BITS > 7) match_bit<7>(m, d);
  if (BITS > 8) match_bit<8>(m, d);
  if (BITS > 9) match_bit<9>(m, d);
  if (BITS > 10) match_bit<10>(m, d);
  return m;
}

// MODE: 0 match.any (BITS-bit random digits), 1 ballot match, 2 leader ATOMS (+shfl), 3 STS scatter 32-bit,
//       4 STS scatter 64-bit, 5 LDS gather 32-bit, 6 full rank step (ballot), 7 full rank step (match.any),
//       8 all-lane ATOMS random, 9 LDS/STS conflict-free baseline, 10 match.any x4 independent
template <int MODE, int BITS>
__global__ void __launch_bounds__(NT) k(uint32_t* out, long long* cyc, uint32_t seed) {
  __shared__ uint32_t sm[8192 + 64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 8192; i += NT) sm[i] = 0;
  __syncthreads();
  uint32_t x = mix(seed + blockIdx.x * NT + tid);
  uint32_t acc = 0;
  const uint32_t mask = (1u << BITS) - 1;
  uint32_t* myh = sm + (warp & 3) * 2048;  // up to 2048 counters per warp-group
  long long t0 = clock64();
#pragma unroll 4
  for (int it = 0; it < ITERS; ++it) {
    x = x * 1664525u + 1013904223u;
    const uint32_t d = (x >> 13) & mask;
    if (MODE == 0) {
      acc += __match_any_sync(0xffffffffu, d);
    } else if (MODE == 10) {
      acc += __match_any_sync(0xffffffffu, d) ^ __match_any_sync(0xffffffffu, d ^ (x >> 3 & mask)) ^
             __match_any_sync(0xffffffffu, (x >> 5) & mask) ^ __match_any_sync(0xffffffffu, (x >> 20) & mask);
    } else if (MODE == 1) {
      acc += match_ballot<BITS>(d);
    } else if (MODE == 2) {
      // leader = lanes whose digit is unique-ish: emulate with a cheap predicate (~30 of 32 active)
      const bool leader = (x & 0x0f000000u) != 0;
      uint32_t old = 0;
      if (leader) old = atomicAdd(&myh[d], 1u);
      acc += __shfl_sync(0xffffffffu, old, (x >> 8) & 31);
    } else if (MODE == 8) {
      acc += atomicAdd(&myh[d], 1u);
    } else if (MODE == 3) {
      sm[(x >> 9) & 8191] = x;
    } else if (MODE == 4) {
      reinterpret_cast<uint2*>(sm)[(x >> 9) & 4095] = make_uint2(x, acc);
    } else if (MODE == 5) {
      acc += sm[(x >> 9) & 8191];
    } else if (MODE == 9) {
      sm[(it * 32 + tid) & 8191] = x;
      acc += sm[(it * 64 + tid) & 8191];
    } else if (MODE == 6 || MODE == 7) {
      const uint32_t m = MODE == 6 ? match_ballot<BITS>(d) : __match_any_sync(0xffffffffu, d);
      const uint32_t leader = 31 - __clz(m);
      uint32_t prev = 0;
      if (lane == leader) prev = atomicAdd(&myh[d], (uint32_t)__popc(m));
      prev = __shfl_sync(0xffffffffu, prev, leader);
      acc += prev + __popc(m & ((1u << lane) - 1));
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * NT + tid] = acc + sm[tid];
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE, int BITS>
int run(const char* name, int ctas_per_sm, int sms, uint32_t* out, long long* cyc) {
  const int grid = sms * ctas_per_sm;
  k<MODE, BITS><<<grid, NT>>>(out, cyc, 1);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE, BITS><<<grid, NT>>>(out, cyc, 2);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(grid);
  CK(cudaMemcpy(h.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0; for (auto v : h) avg += v; avg /= grid;
  const double warps_per_sm = ctas_per_sm * (NT / 32);
  // cycles per warp-iteration per SM = CTA cycles / (ITERS * warps on the SM)
  printf("%-34s bits=%2d ctas/SM=%d  cyc/iter/warp(latency-ish)=%7.2f  cyc/warp-iter/SM=%6.3f  (%.3f ms)\n", name, BITS,
         ctas_per_sm, avg / ITERS, avg / ITERS / warps_per_sm, ms);
  return 0;
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint32_t* out; long long* cyc;
  CK(cudaMalloc(&out, sizeof(uint32_t) * sms * 4 * NT));
  CK(cudaMalloc(&cyc, sizeof(long long) * sms * 4));
  for (int c = 1; c <= 2; ++c) {
    run<0, 8>("match.any", c, sms, out, cyc);
    run<0, 11>("match.any", c, sms, out, cyc);
    run<0, 2>("match.any", c, sms, out, cyc);
    run<10, 8>("match.any x4 indep", c, sms, out, cyc);
    run<1, 8>("ballot match", c, sms, out, cyc);
    run<1, 11>("ballot match", c, sms, out, cyc);
    run<2, 8>("leader ATOMS + SHFL", c, sms, out, cyc);
    run<8, 8>("all-lane ATOMS random", c, sms, out, cyc);
    run<8, 2>("all-lane ATOMS 4 addrs", c, sms, out, cyc);
    run<3, 8>("STS.32 random scatter", c, sms, out, cyc);
    run<4, 8>("STS.64 random scatter", c, sms, out, cyc);
    run<5, 8>("LDS.32 random gather", c, sms, out, cyc);
    run<9, 8>("STS+LDS conflict-free", c, sms, out, cyc);
    run<6, 8>("rank step ballot", c, sms, out, cyc);
    run<6, 11>("rank step ballot", c, sms, out, cyc);
    run<7, 8>("rank step match.any", c, sms, out, cyc);
    run<7, 11>("rank step match.any", c, sms, out, cyc);
  }
  return 0;
}
