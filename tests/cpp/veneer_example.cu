// tests/cpp/veneer_example.cu -- a reference-style call site compiled against include/b200/device_radix_sort.cuh.
// Shaped like examples/device/example_device_radix_sort.cu:190-198 and the test harness back-ends
// (test/test_device_radix_sort.cu:154-263): size query with d_temp_storage == nullptr, allocate, sort;
// pointer and DoubleBuffer forms; result checked against host std::stable_sort.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>

#include "b200/device_radix_sort.cuh"
#include "b200/device_segmented_radix_sort.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %d at %s:%d\n", (int)e_, __FILE__, __LINE__); return 2; } } while (0)

template <typename KeyT>
int run(int n, bool descending) {
  std::vector<KeyT> h_keys(n);
  std::vector<int> h_vals(n);
  unsigned s = 12345u;
  for (int i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; h_keys[i] = (KeyT)(int)(s >> 7) / (KeyT)3; h_vals[i] = i; }
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return descending ? h_keys[b] < h_keys[a] : h_keys[a] < h_keys[b]; });

  KeyT *d_keys_in, *d_keys_out; int *d_vals_in, *d_vals_out;
  CK(cudaMalloc(&d_keys_in, n * sizeof(KeyT))); CK(cudaMalloc(&d_keys_out, n * sizeof(KeyT)));
  CK(cudaMalloc(&d_vals_in, n * sizeof(int)));  CK(cudaMalloc(&d_vals_out, n * sizeof(int)));
  CK(cudaMemcpy(d_keys_in, h_keys.data(), n * sizeof(KeyT), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_vals_in, h_vals.data(), n * sizeof(int), cudaMemcpyHostToDevice));

  // pointer form
  void* d_temp_storage = nullptr; size_t temp_storage_bytes = 0;
  if (descending) { CK(b200::DeviceRadixSort::SortPairsDescending(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_vals_in, d_vals_out, n)); }
  else            { CK(b200::DeviceRadixSort::SortPairs(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_vals_in, d_vals_out, n)); }
  CK(cudaMalloc(&d_temp_storage, temp_storage_bytes));
  if (descending) { CK(b200::DeviceRadixSort::SortPairsDescending(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_vals_in, d_vals_out, n)); }
  else            { CK(b200::DeviceRadixSort::SortPairs(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_vals_in, d_vals_out, n)); }
  std::vector<int> got(n);
  CK(cudaMemcpy(got.data(), d_vals_out, n * sizeof(int), cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i) if (got[i] != order[i]) { printf("pointer form mismatch at %d\n", i); return 1; }
  CK(cudaFree(d_temp_storage));

  // DoubleBuffer form, keys only, explicit stream
  cudaStream_t stream; CK(cudaStreamCreate(&stream));
  b200::DoubleBuffer<KeyT> d_keys(d_keys_in, d_keys_out);
  d_temp_storage = nullptr; temp_storage_bytes = 0;
  CK(b200::DeviceRadixSort::SortKeys(d_temp_storage, temp_storage_bytes, d_keys, (size_t)n, 0, (int)sizeof(KeyT) * 8, stream));
  CK(cudaMalloc(&d_temp_storage, temp_storage_bytes));
  CK(b200::DeviceRadixSort::SortKeys(d_temp_storage, temp_storage_bytes, d_keys, (size_t)n, 0, (int)sizeof(KeyT) * 8, stream));
  CK(cudaStreamSynchronize(stream));
  std::vector<KeyT> gk(n);
  CK(cudaMemcpy(gk.data(), d_keys.Current(), n * sizeof(KeyT), cudaMemcpyDeviceToHost));
  std::vector<KeyT> ek(h_keys);
  std::stable_sort(ek.begin(), ek.end());
  for (int i = 0; i < n; ++i) if (!(gk[i] == ek[i])) { printf("DoubleBuffer form mismatch at %d\n", i); return 1; }
  cudaFree(d_temp_storage); cudaFree(d_keys_in); cudaFree(d_keys_out); cudaFree(d_vals_in); cudaFree(d_vals_out);
  cudaStreamDestroy(stream);
  return 0;
}

// User-defined key struct with a decomposer, shaped like the reference's documentation example
// (cub/device/device_radix_sort.cuh:128-195: custom_t { float f; long long lli; } + decomposer_t).
struct custom_t {
  float f;
  long long lli;
};
struct decomposer_t {
  __host__ __device__ ::cuda::std::tuple<float&, long long&> operator()(custom_t& key) const { return {key.f, key.lli}; }
};

int run_struct(int n) {
  std::vector<custom_t> h(n);
  std::vector<int> h_vals(n);
  unsigned s = 777u;
  for (int i = 0; i < n; ++i) {
    s = s * 1664525u + 1013904223u;
    h[i].f = (float)((int)(s >> 24) - 128) / 4.0f;  // few distinct values: the second member decides often
    s = s * 1664525u + 1013904223u;
    h[i].lli = (long long)(int)s * 1000003ll;
    h_vals[i] = i;
  }
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    return h[a].f < h[b].f || (h[a].f == h[b].f && h[a].lli < h[b].lli);
  });
  custom_t *d_in, *d_out; int *d_vin, *d_vout;
  CK(cudaMalloc(&d_in, n * sizeof(custom_t))); CK(cudaMalloc(&d_out, n * sizeof(custom_t)));
  CK(cudaMalloc(&d_vin, n * sizeof(int))); CK(cudaMalloc(&d_vout, n * sizeof(int)));
  CK(cudaMemcpy(d_in, h.data(), n * sizeof(custom_t), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_vin, h_vals.data(), n * sizeof(int), cudaMemcpyHostToDevice));
  void* d_temp_storage = nullptr; size_t temp_storage_bytes = 0;
  CK(b200::DeviceRadixSort::SortPairs(d_temp_storage, temp_storage_bytes, d_in, d_out, d_vin, d_vout, n, decomposer_t{}));
  CK(cudaMalloc(&d_temp_storage, temp_storage_bytes));
  CK(b200::DeviceRadixSort::SortPairs(d_temp_storage, temp_storage_bytes, d_in, d_out, d_vin, d_vout, n, decomposer_t{}));
  std::vector<int> got(n);
  std::vector<custom_t> gk(n);
  CK(cudaMemcpy(got.data(), d_vout, n * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(gk.data(), d_out, n * sizeof(custom_t), cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i)
    if (got[i] != order[i] || gk[i].f != h[order[i]].f || gk[i].lli != h[order[i]].lli) { printf("decomposer form mismatch at %d\n", i); return 1; }
  // descending, keys only, DoubleBuffer form, explicit bit range covering the whole 96-bit image
  b200::DoubleBuffer<custom_t> d_keys(d_in, d_out);
  CK(b200::DeviceRadixSort::SortKeysDescending(d_temp_storage, temp_storage_bytes, d_keys, n, decomposer_t{}, 0, 96));
  CK(cudaMemcpy(gk.data(), d_keys.Current(), n * sizeof(custom_t), cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i) {
    const custom_t& e = h[order[n - 1 - i]];  // all (f, lli) pairs are distinct with overwhelming probability: reverse order
    if (gk[i].f != e.f || gk[i].lli != e.lli) { printf("decomposer descending mismatch at %d\n", i); return 1; }
  }
  // the deprecated debug_synchronous overload still compiles and sorts
  unsigned *d_u, *d_u2;
  CK(cudaMalloc(&d_u, 1000 * sizeof(unsigned))); CK(cudaMalloc(&d_u2, 1000 * sizeof(unsigned)));
  CK(cudaMemset(d_u, 0, 1000 * sizeof(unsigned)));
  size_t tb = 0;
  CK(b200::DeviceRadixSort::SortKeys(nullptr, tb, d_u, d_u2, 1000, 0, 32, 0, true));
  cudaFree(d_u); cudaFree(d_u2);
  cudaFree(d_temp_storage); cudaFree(d_in); cudaFree(d_out); cudaFree(d_vin); cudaFree(d_vout);
  return 0;
}

// Segmented sort, shaped like the reference's documentation snippet (cub/device/device_segmented_radix_sort.cuh:60-100):
// d_offsets / d_offsets + 1 as begin / end offsets.
int run_segmented(int n) {
  std::vector<int> h_keys(n), h_vals(n), h_offsets;
  unsigned s = 4242u;
  for (int i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; h_keys[i] = (int)(s >> 4) - (1 << 27); h_vals[i] = i; }
  for (int o = 0; o < n; o += 1 + (int)((s = s * 1664525u + 1013904223u) >> 18) % 9000) h_offsets.push_back(o);
  h_offsets.push_back(n);
  const int num_segments = (int)h_offsets.size() - 1;
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  for (int g = 0; g < num_segments; ++g)
    std::stable_sort(order.begin() + h_offsets[g], order.begin() + h_offsets[g + 1], [&](int a, int b) { return h_keys[a] < h_keys[b]; });
  int *d_keys_in, *d_keys_out, *d_vals_in, *d_vals_out, *d_offsets;
  CK(cudaMalloc(&d_keys_in, n * sizeof(int))); CK(cudaMalloc(&d_keys_out, n * sizeof(int)));
  CK(cudaMalloc(&d_vals_in, n * sizeof(int))); CK(cudaMalloc(&d_vals_out, n * sizeof(int)));
  CK(cudaMalloc(&d_offsets, h_offsets.size() * sizeof(int)));
  CK(cudaMemcpy(d_keys_in, h_keys.data(), n * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_vals_in, h_vals.data(), n * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_offsets, h_offsets.data(), h_offsets.size() * sizeof(int), cudaMemcpyHostToDevice));
  void* d_temp_storage = nullptr; size_t temp_storage_bytes = 0;
  CK(b200::DeviceSegmentedRadixSort::SortPairs(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_vals_in, d_vals_out, n,
                                              num_segments, d_offsets, d_offsets + 1));
  CK(cudaMalloc(&d_temp_storage, temp_storage_bytes));
  CK(b200::DeviceSegmentedRadixSort::SortPairs(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_vals_in, d_vals_out, n,
                                              num_segments, d_offsets, d_offsets + 1));
  std::vector<int> got(n);
  CK(cudaMemcpy(got.data(), d_vals_out, n * sizeof(int), cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i) if (got[i] != order[i]) { printf("segmented form mismatch at %d\n", i); return 1; }
  cudaFree(d_temp_storage); cudaFree(d_keys_in); cudaFree(d_keys_out); cudaFree(d_vals_in); cudaFree(d_vals_out); cudaFree(d_offsets);
  return 0;
}

int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 1000003;
  int rc = 0;
  rc |= run<float>(n, false);
  rc |= run<int>(n, true);
  rc |= run<unsigned long long>(n / 3, false);
  rc |= run<double>(n / 5, true);
  rc |= run_struct(n / 2);
  rc |= run_segmented(n / 2);
  printf(rc == 0 ? "veneer example: OK\n" : "veneer example: FAILED\n");
  return rc;
}
