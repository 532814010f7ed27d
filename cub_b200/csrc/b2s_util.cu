// b2s_util.cu -- small device helpers exported through the C-ABI: synthetic input generators
// (SURVEY.md §8d counter-based generator), an order/multiset checker for sizes beyond the CPU
// oracle, and a bit-ordered lower_bound used by the multi-GPU splitter step.
#include <cuda_runtime.h>

#include "../../include/b2s_mgpu.h"
#include "../../include/b2s_radix_sort.h"
#include "b2s_common.cuh"

namespace b2s {
namespace {

__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

template <typename T>
__global__ void fill_keys_kernel(T* out, unsigned long long n, unsigned long long seed, int and_rounds,
                                 unsigned long long first) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    unsigned long long v = ~0ull;
    for (int r = 0; r < and_rounds; ++r) v &= splitmix64((seed + r) * 0x100000001B3ull + first + i);
    out[i] = (T)v;
  }
}

template <typename T>
__global__ void fill_iota_kernel(T* out, unsigned long long n, unsigned long long first) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = (T)(first + i);
}

struct TypeConsts {
  unsigned long long ones, high, xor_mask, zero_from, zero_to;
  int is_float;
  int begin_bit, nbits;
};

__device__ __forceinline__ unsigned long long sort_key(unsigned long long k, const TypeConsts& c) {
  if (c.is_float) {
    k = (k == c.zero_from) ? c.zero_to : k;
    k ^= (k & c.high) ? c.ones : c.high;
  }
  k = (k ^ c.xor_mask) & c.ones;
  k >>= c.begin_bit;
  if (c.nbits < 64) k &= (1ull << c.nbits) - 1;
  return k;
}

template <typename T>
__device__ __forceinline__ unsigned long long load_bits(const void* p, unsigned long long i) {
  return (unsigned long long)reinterpret_cast<const T*>(p)[i];
}
__device__ __forceinline__ unsigned long long load_any(const void* p, unsigned long long i, int bytes) {
  switch (bytes) {
    case 1: return load_bits<unsigned char>(p, i);
    case 2: return load_bits<unsigned short>(p, i);
    case 4: return load_bits<unsigned int>(p, i);
    case 8: return load_bits<unsigned long long>(p, i);
    case 16: {
      const unsigned long long* q = reinterpret_cast<const unsigned long long*>(p) + 2 * i;
      return q[0] ^ splitmix64(q[1]);
    }
    default: return 0;
  }
}

__global__ void check_sorted_kernel(const void* keys, const void* vals, unsigned long long n, int kbytes, int vbytes,
                                    TypeConsts c, unsigned long long* result) {
  unsigned long long inv = 0, ksum = 0, psum = 0;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned long long k = load_any(keys, i, kbytes);
    if (i + 1 < n) {
      const unsigned long long k2 = load_any(keys, i + 1, kbytes);
      if (sort_key(k, c) > sort_key(k2, c)) inv++;
    }
    ksum += splitmix64(k);
    if (vals) {
      const unsigned long long v = load_any(vals, i, vbytes);
      psum += splitmix64(k ^ ((v << 32) | (v >> 32)));
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    inv += __shfl_down_sync(0xffffffffu, inv, o);
    ksum += __shfl_down_sync(0xffffffffu, ksum, o);
    psum += __shfl_down_sync(0xffffffffu, psum, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (inv) atomicAdd(&result[0], inv);
    atomicAdd(&result[1], ksum);
    if (vals) atomicAdd(&result[2], psum);
  }
}

// adjacent positions with equal sort keys whose values do not increase (values = input indices: 0 for a stable sort)
__global__ void check_stable_kernel(const void* keys, const void* vals, unsigned long long n, int kbytes, int vbytes,
                                    TypeConsts c, unsigned long long* result) {
  unsigned long long bad = 0;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i + 1 < n; i += stride) {
    if (sort_key(load_any(keys, i, kbytes), c) == sort_key(load_any(keys, i + 1, kbytes), c) &&
        load_any(vals, i, vbytes) >= load_any(vals, i + 1, vbytes))
      bad++;
  }
  for (int o = 16; o > 0; o >>= 1) bad += __shfl_down_sync(0xffffffffu, bad, o);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(result, bad);
}

__global__ void lower_bound_kernel(const void* keys, unsigned long long n, int kbytes, TypeConsts c,
                                   const void* splitters, int num, unsigned long long* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num) return;
  const unsigned long long s = sort_key(load_any(splitters, i, kbytes), c);
  unsigned long long lo = 0, hi = n;
  while (lo < hi) {
    const unsigned long long mid = lo + (hi - lo) / 2;
    if (sort_key(load_any(keys, mid, kbytes), c) < s) lo = mid + 1; else hi = mid;
  }
  out[i] = lo;
}

const int kBytes[B2S_KEY_TYPE_COUNT] = {1, 1, 2, 2, 2, 2, 4, 4, 4, 8, 8, 8};
const int kCat[B2S_KEY_TYPE_COUNT] = {0, 1, 0, 1, 2, 2, 0, 1, 2, 0, 1, 2};

bool make_type_consts(int key_type, int descending, int begin_bit, int end_bit, TypeConsts* c) {
  if (key_type < 0 || key_type >= B2S_KEY_TYPE_COUNT) return false;
  const int bits = kBytes[key_type] * 8;
  c->ones = bits == 64 ? ~0ull : ((1ull << bits) - 1);
  c->high = 1ull << (bits - 1);
  c->is_float = kCat[key_type] == 2;
  if (c->is_float) {
    c->xor_mask = descending ? c->ones : 0;
    c->zero_from = descending ? 0 : c->high;
    c->zero_to = descending ? c->high : 0;
  } else {
    c->xor_mask = (kCat[key_type] == 1 ? c->high : 0) ^ (descending ? c->ones : 0);
    c->zero_from = c->zero_to = 0;
  }
  c->begin_bit = begin_bit;
  c->nbits = end_bit - begin_bit;
  return true;
}

int grid_for(unsigned long long n, int threads) {
  unsigned long long g = (n + threads - 1) / threads;
  if (g > 148ull * 16) g = 148ull * 16;
  if (g == 0) g = 1;
  return (int)g;
}

}  // namespace
}  // namespace b2s

extern "C" {

int b2s_fill_keys(void* d_keys, uint64_t n, int key_bytes, uint64_t seed, int and_rounds, uint64_t first_index,
                  b2s_stream_t stream) {
  using namespace b2s;
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int g = grid_for(n, 256);
  switch (key_bytes) {
    case 1: fill_keys_kernel<<<g, 256, 0, s>>>((unsigned char*)d_keys, n, seed, and_rounds, first_index); break;
    case 2: fill_keys_kernel<<<g, 256, 0, s>>>((unsigned short*)d_keys, n, seed, and_rounds, first_index); break;
    case 4: fill_keys_kernel<<<g, 256, 0, s>>>((unsigned int*)d_keys, n, seed, and_rounds, first_index); break;
    case 8: fill_keys_kernel<<<g, 256, 0, s>>>((unsigned long long*)d_keys, n, seed, and_rounds, first_index); break;
    default: return (int)cudaErrorInvalidValue;
  }
  return (int)cudaGetLastError();
}

int b2s_fill_iota(void* d_values, uint64_t n, int value_bytes, uint64_t first_index, b2s_stream_t stream) {
  using namespace b2s;
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int g = grid_for(n, 256);
  switch (value_bytes) {
    case 1: fill_iota_kernel<<<g, 256, 0, s>>>((unsigned char*)d_values, n, first_index); break;
    case 2: fill_iota_kernel<<<g, 256, 0, s>>>((unsigned short*)d_values, n, first_index); break;
    case 4: fill_iota_kernel<<<g, 256, 0, s>>>((unsigned int*)d_values, n, first_index); break;
    case 8: fill_iota_kernel<<<g, 256, 0, s>>>((unsigned long long*)d_values, n, first_index); break;
    default: return (int)cudaErrorInvalidValue;
  }
  return (int)cudaGetLastError();
}

int b2s_check_stable(const void* d_keys, const void* d_values, uint64_t n, int key_type, int value_bytes, int descending,
                     int begin_bit, int end_bit, uint64_t* d_result, b2s_stream_t stream) {
  using namespace b2s;
  TypeConsts c;
  if (!make_type_consts(key_type, descending, begin_bit, end_bit, &c) || !d_values || !(value_bytes == 4 || value_bytes == 8))
    return (int)cudaErrorInvalidValue;
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(d_result, 0, sizeof(uint64_t), s);
  if (e != cudaSuccess) return (int)e;
  if (n < 2) return 0;
  check_stable_kernel<<<grid_for(n, 256), 256, 0, s>>>(d_keys, d_values, n, kBytes[key_type], value_bytes, c,
                                                       reinterpret_cast<unsigned long long*>(d_result));
  return (int)cudaGetLastError();
}

int b2s_check_sorted(const void* d_keys, const void* d_values, uint64_t n, int key_type, int value_bytes,
                     int descending, int begin_bit, int end_bit, uint64_t* d_result, b2s_stream_t stream) {
  using namespace b2s;
  TypeConsts c;
  if (!make_type_consts(key_type, descending, begin_bit, end_bit, &c)) return (int)cudaErrorInvalidValue;
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(d_result, 0, 3 * sizeof(uint64_t), s);
  if (e != cudaSuccess) return (int)e;
  if (n == 0) return 0;
  check_sorted_kernel<<<grid_for(n, 256), 256, 0, s>>>(d_keys, value_bytes ? d_values : nullptr, n,
                                                       kBytes[key_type], value_bytes, c,
                                                       reinterpret_cast<unsigned long long*>(d_result));
  return (int)cudaGetLastError();
}

int b2s_lower_bound(const void* d_sorted_keys, uint64_t n, int key_type, const void* d_splitters, int num_splitters,
                    uint64_t* d_out, b2s_stream_t stream) {
  using namespace b2s;
  TypeConsts c;
  if (!make_type_consts(key_type, 0, 0, kBytes[key_type < 0 || key_type >= B2S_KEY_TYPE_COUNT ? 0 : key_type] * 8, &c))
    return (int)cudaErrorInvalidValue;
  if (num_splitters <= 0) return 0;
  lower_bound_kernel<<<(num_splitters + 63) / 64, 64, 0, (cudaStream_t)stream>>>(
      d_sorted_keys, n, kBytes[key_type], c, d_splitters, num_splitters, reinterpret_cast<unsigned long long*>(d_out));
  return (int)cudaGetLastError();
}

}  // extern "C"
