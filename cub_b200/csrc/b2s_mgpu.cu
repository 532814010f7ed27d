// b2s_mgpu.cu -- C++ host of the single-box multi-GPU SortPairs (include/b2s_mgpu.h): one process per GPU, NCCL for
// the metadata, CUDA IPC peer mappings for the payload, the partition kernel of b2s_pass.cuh (PF_PEER) as the all-to-all.
// New functionality: the reference (NVIDIA/cub) is single-GPU; the local building block is the DeviceRadixSort drop-in.
//
// NCCL is loaded lazily with dlopen("libnccl.so.2"): libb2s.so carries no link-time dependency on it, and inside a
// PyTorch process the already-loaded copy (same SONAME) is the one that is used.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <new>

#include "../../include/b2s_mgpu.h"

namespace b2s {
namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

const NcclApi* nccl_api() {
  static const NcclApi api = [] {
    NcclApi a;
    a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.handle) a.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!a.handle) return a;
    auto sym = [&](const char* n) { return dlsym(a.handle, n); };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.AllReduce) a.handle = nullptr;
    return a;
  }();
  return api.handle ? &api : nullptr;
}

constexpr int G_MAX = B2S_MGPU_MAX_RANKS;

template <typename T>
__global__ void sample_kernel(const T* keys, unsigned long long n, int s, T* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < s) out[i] = keys[(unsigned long long)(((unsigned __int128)i * n) / (unsigned)s)];
}
__global__ void source_rank_kernel(int* src, int total, int s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) src[i] = i / s;
}
template <typename T>
__global__ void pick_splitters_kernel(const T* sorted_keys, const int* sorted_ranks, int s, int world, T* sp_keys, int* sp_ranks) {
  const int j = threadIdx.x;
  if (j < world - 1) {
    sp_keys[j] = sorted_keys[(j + 1) * s];
    sp_ranks[j] = sorted_ranks[(j + 1) * s];
  }
}
// offsets[d] = position of MY segment in rank d's receive buffer = items the lower ranks send to d
__global__ void plan_kernel(const unsigned long long* matrix, int world, int me, unsigned long long* offsets) {
  const int d = threadIdx.x;
  if (d < world) {
    unsigned long long o = 0;
    for (int r = 0; r < me; ++r) o += matrix[r * world + d];
    offsets[d] = o;
  }
}

}  // namespace
}  // namespace b2s

struct b2s_mgpu_sorter {
  int rank = 0, world = 1, device = 0;
  ncclComm_t comm = nullptr;
  int key_type = 0, kbytes = 0, vbytes = 0, descending = 0, begin_bit = 0, end_bit = 0, samples = 0;
  uint64_t max_items = 0, capacity = 0;
  void* recv_k[2] = {nullptr, nullptr};
  void* recv_v[2] = {nullptr, nullptr};
  void* peer_k[b2s::G_MAX] = {};
  void* peer_v[b2s::G_MAX] = {};
  void* opened[2 * b2s::G_MAX] = {};
  int num_opened = 0;
  void* sample = nullptr;          // [s] keys
  void* gathered[2] = {};          // [world * s] keys, sort input / output
  int* src[2] = {};                // [world * s] source ranks, sort input / output
  void* sp_keys = nullptr;         // [world - 1]
  int* sp_ranks = nullptr;
  unsigned long long* counts = nullptr;   // [world]
  unsigned long long* matrix = nullptr;   // [world * world]
  unsigned long long* offsets = nullptr;  // [world]
  unsigned long long* h_matrix = nullptr; // pinned
  int* fence = nullptr;
  void* sort_temp = nullptr;
  size_t sort_temp_bytes = 0;
  void* sample_temp = nullptr;
  size_t sample_temp_bytes = 0;
  void* split_temp = nullptr;
  size_t split_temp_bytes = 0;
  cudaEvent_t ev[6] = {};
  cudaEvent_t ev_matrix = nullptr;
  bool timed = false;
  uint64_t items_sent = 0;
  char err[256] = {0};
};

namespace b2s {
namespace {

int fail(b2s_mgpu_sorter* s, int code, const char* what) {
  if (s) {
    const NcclApi* api = nccl_api();
    if (code >= 1000 && code < 2000 && api && api->GetErrorString)
      std::snprintf(s->err, sizeof(s->err), "%s: NCCL %s", what, api->GetErrorString((ncclResult_t)(code - 1000)));
    else if (code < 1000)
      std::snprintf(s->err, sizeof(s->err), "%s: %s", what, cudaGetErrorString((cudaError_t)code));
    else
      std::snprintf(s->err, sizeof(s->err), "%s (code %d)", what, code);
  }
  return code;
}

#define B2S_CUDA(call, what)                                  \
  do {                                                        \
    cudaError_t e_ = (call);                                  \
    if (e_ != cudaSuccess) return fail(S, (int)e_, what);     \
  } while (0)
#define B2S_NCCL(call, what)                                       \
  do {                                                             \
    ncclResult_t r_ = (call);                                      \
    if (r_ != ncclSuccess) return fail(S, 1000 + (int)r_, what);   \
  } while (0)
#define B2S_RC(call, what)                 \
  do {                                     \
    int rc_ = (call);                      \
    if (rc_ != 0) return fail(S, rc_, what); \
  } while (0)

void release(b2s_mgpu_sorter* S) {
  for (int i = 0; i < S->num_opened; ++i) cudaIpcCloseMemHandle(S->opened[i]);
  void* dev[] = {S->recv_k[0], S->recv_k[1], S->recv_v[0], S->recv_v[1], S->sample, S->gathered[0], S->gathered[1], S->src[0],
                 S->src[1], S->sp_keys, S->sp_ranks, S->counts, S->matrix, S->offsets, S->fence, S->sort_temp, S->sample_temp,
                 S->split_temp};
  for (void* p : dev)
    if (p) cudaFree(p);
  if (S->h_matrix) cudaFreeHost(S->h_matrix);
  for (cudaEvent_t e : S->ev)
    if (e) cudaEventDestroy(e);
  if (S->ev_matrix) cudaEventDestroy(S->ev_matrix);
  const NcclApi* api = nccl_api();
  if (S->comm && api) api->CommDestroy(S->comm);
  delete S;
}

}  // namespace
}  // namespace b2s

extern "C" {

int b2s_mgpu_unique_id(void* id128) {
  const b2s::NcclApi* api = b2s::nccl_api();
  if (!api) return B2S_MGPU_E_NCCL_MISSING;
  if (!id128) return B2S_MGPU_E_ARGUMENT;
  static_assert(sizeof(ncclUniqueId) == B2S_MGPU_ID_BYTES, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  ncclResult_t r = api->GetUniqueId(&id);
  if (r != ncclSuccess) return 1000 + (int)r;
  std::memcpy(id128, &id, sizeof(id));
  return 0;
}

int b2s_mgpu_create(b2s_mgpu_sorter_t** out, const void* id128, int rank, int world, uint64_t max_items, int key_type,
                    int value_bytes, int descending, int begin_bit, int end_bit, double slack, int samples_per_rank) {
  using namespace b2s;
  const NcclApi* api = nccl_api();
  if (!api) return B2S_MGPU_E_NCCL_MISSING;
  const int kbytes = b2s_key_bytes(key_type);
  if (!out || !id128 || world < 1 || world > G_MAX || rank < 0 || rank >= world || max_items < 1 ||
      !(kbytes == 4 || kbytes == 8) || !(value_bytes == 0 || value_bytes == 4 || value_bytes == 8) || begin_bit < 0 ||
      end_bit > kbytes * 8 || end_bit <= begin_bit || slack < 1.0 || samples_per_rank < 1 || samples_per_rank > 65536)
    return B2S_MGPU_E_ARGUMENT;
  b2s_mgpu_sorter* S = new (std::nothrow) b2s_mgpu_sorter();
  if (!S) return (int)cudaErrorMemoryAllocation;
  *out = S;  // returned even on failure so that the caller can read last_error and destroy
  S->rank = rank;
  S->world = world;
  S->key_type = key_type;
  S->kbytes = kbytes;
  S->vbytes = value_bytes;
  S->descending = descending != 0;
  S->begin_bit = begin_bit;
  S->end_bit = end_bit;
  S->samples = samples_per_rank;
  S->max_items = max_items;
  S->capacity = (uint64_t)((double)max_items * slack) + 1024;
  B2S_CUDA(cudaGetDevice(&S->device), "cudaGetDevice");
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  B2S_NCCL(api->CommInitRank(&S->comm, world, id, rank), "ncclCommInitRank");

  const size_t s = (size_t)samples_per_rank, gs = s * (size_t)world;
  for (int i = 0; i < 2; ++i) {
    B2S_CUDA(cudaMalloc(&S->recv_k[i], S->capacity * (size_t)kbytes), "cudaMalloc(receive keys)");
    if (value_bytes) B2S_CUDA(cudaMalloc(&S->recv_v[i], S->capacity * (size_t)value_bytes), "cudaMalloc(receive values)");
    B2S_CUDA(cudaMalloc(&S->gathered[i], gs * kbytes), "cudaMalloc(samples)");
    B2S_CUDA(cudaMalloc(reinterpret_cast<void**>(&S->src[i]), gs * sizeof(int)), "cudaMalloc(sample ranks)");
  }
  B2S_CUDA(cudaMalloc(&S->sample, s * kbytes), "cudaMalloc");
  B2S_CUDA(cudaMalloc(&S->sp_keys, (size_t)G_MAX * kbytes), "cudaMalloc");
  B2S_CUDA(cudaMalloc(reinterpret_cast<void**>(&S->sp_ranks), G_MAX * sizeof(int)), "cudaMalloc");
  B2S_CUDA(cudaMalloc(reinterpret_cast<void**>(&S->counts), G_MAX * sizeof(unsigned long long)), "cudaMalloc");
  B2S_CUDA(cudaMalloc(reinterpret_cast<void**>(&S->matrix), G_MAX * G_MAX * sizeof(unsigned long long)), "cudaMalloc");
  B2S_CUDA(cudaMalloc(reinterpret_cast<void**>(&S->offsets), G_MAX * sizeof(unsigned long long)), "cudaMalloc");
  B2S_CUDA(cudaMalloc(reinterpret_cast<void**>(&S->fence), 256), "cudaMalloc");
  B2S_CUDA(cudaMemset(S->fence, 0, 256), "cudaMemset");
  B2S_CUDA(cudaMallocHost(reinterpret_cast<void**>(&S->h_matrix), G_MAX * G_MAX * sizeof(unsigned long long)), "cudaMallocHost");
  for (auto& e : S->ev) B2S_CUDA(cudaEventCreate(&e), "cudaEventCreate");
  B2S_CUDA(cudaEventCreateWithFlags(&S->ev_matrix, cudaEventDisableTiming), "cudaEventCreate");
  source_rank_kernel<<<(unsigned)((gs + 255) / 256), 256>>>(S->src[0], (int)gs, (int)s);
  B2S_CUDA(cudaGetLastError(), "source_rank_kernel");

  // temp storage: local sort of `capacity` items (DoubleBuffer form), sample sort (pointer form), partition pass
  void* kb[2] = {S->recv_k[0], S->recv_k[1]};
  void* vb[2] = {S->recv_v[0], S->recv_v[1]};
  int ksel = 0, vsel = 0;
  B2S_RC(b2s_radix_sort_db(nullptr, &S->sort_temp_bytes, kb, &ksel, value_bytes ? vb : nullptr, value_bytes ? &vsel : nullptr,
                           S->capacity, key_type, value_bytes, 8, descending, begin_bit, end_bit, nullptr),
         "temp-storage query (local sort)");
  B2S_RC(b2s_radix_sort(nullptr, &S->sample_temp_bytes, S->gathered[0], S->gathered[1], S->src[0], S->src[1], gs, key_type, 4, 4,
                        descending, begin_bit, end_bit, nullptr),
         "temp-storage query (sample sort)");
  B2S_RC(b2s_split_scatter(nullptr, &S->split_temp_bytes, S->recv_k[0], nullptr, S->recv_v[0], nullptr, max_items, key_type,
                           value_bytes, descending, begin_bit, end_bit, S->sp_keys, S->sp_ranks, world - 1, rank, reinterpret_cast<const uint64_t*>(S->offsets),
                           S->peer_k, S->peer_v, S->capacity, nullptr),
         "temp-storage query (partition)");
  B2S_CUDA(cudaMalloc(&S->sort_temp, S->sort_temp_bytes), "cudaMalloc(sort temp)");
  B2S_CUDA(cudaMalloc(&S->sample_temp, S->sample_temp_bytes), "cudaMalloc(sample sort temp)");
  B2S_CUDA(cudaMalloc(&S->split_temp, S->split_temp_bytes), "cudaMalloc(partition temp)");

  // exchange the IPC handles of the receive buffers (buffer 0 of each array is the one peers store into)
  struct Handles {
    cudaIpcMemHandle_t k, v;
  };
  static_assert(sizeof(Handles) == 128, "two 64-byte handles");
  Handles mine;
  std::memset(&mine, 0, sizeof(mine));
  Handles all[G_MAX];
  if (world > 1) {
    B2S_CUDA(cudaIpcGetMemHandle(&mine.k, S->recv_k[0]), "cudaIpcGetMemHandle");
    if (value_bytes) B2S_CUDA(cudaIpcGetMemHandle(&mine.v, S->recv_v[0]), "cudaIpcGetMemHandle");
    Handles* d_h = nullptr;
    B2S_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_h), sizeof(Handles) * (G_MAX + 1)), "cudaMalloc");
    B2S_CUDA(cudaMemcpy(d_h + G_MAX, &mine, sizeof(mine), cudaMemcpyHostToDevice), "cudaMemcpy");
    ncclResult_t r = api->AllGather(d_h + G_MAX, d_h, sizeof(Handles), ncclUint8, S->comm, nullptr);
    cudaError_t e = r == ncclSuccess ? cudaStreamSynchronize(nullptr) : cudaSuccess;
    if (r == ncclSuccess && e == cudaSuccess) e = cudaMemcpy(all, d_h, sizeof(Handles) * world, cudaMemcpyDeviceToHost);
    cudaFree(d_h);
    if (r != ncclSuccess) return fail(S, 1000 + (int)r, "ncclAllGather(IPC handles)");
    if (e != cudaSuccess) return fail(S, (int)e, "IPC handle exchange");
  }
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      S->peer_k[r] = S->recv_k[0];
      S->peer_v[r] = S->recv_v[0];
      continue;
    }
    B2S_CUDA(cudaIpcOpenMemHandle(&S->peer_k[r], all[r].k, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle(keys)");
    S->opened[S->num_opened++] = S->peer_k[r];
    if (value_bytes) {
      B2S_CUDA(cudaIpcOpenMemHandle(&S->peer_v[r], all[r].v, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle(values)");
      S->opened[S->num_opened++] = S->peer_v[r];
    }
  }
  B2S_CUDA(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
  return 0;
}

int b2s_mgpu_sort(b2s_mgpu_sorter_t* S, const void* d_keys, const void* d_values, uint64_t n, void** d_keys_out,
                  void** d_values_out, uint64_t* out_count, uint64_t* counts_all, b2s_stream_t stream_) {
  using namespace b2s;
  const NcclApi* api = nccl_api();
  if (!S || !api) return B2S_MGPU_E_NCCL_MISSING;
  if (!d_keys || n < 1 || n > S->max_items || (S->vbytes != 0) != (d_values != nullptr) || !d_keys_out || !out_count)
    return fail(S, B2S_MGPU_E_ARGUMENT, "b2s_mgpu_sort arguments");
  cudaStream_t st = (cudaStream_t)stream_;
  const int G = S->world, me = S->rank, s = S->samples, kt = S->key_type, desc = S->descending, bb = S->begin_bit, eb = S->end_bit;
  S->timed = false;
  B2S_CUDA(cudaEventRecord(S->ev[0], st), "cudaEventRecord");

  // 1. regular samples -> all ranks -> the same world-1 splitters everywhere (device resident: nobody waits for them)
  if (G > 1) {
    if (S->kbytes == 4)
      sample_kernel<<<(s + 255) / 256, 256, 0, st>>>(static_cast<const unsigned int*>(d_keys), n, s, static_cast<unsigned int*>(S->sample));
    else
      sample_kernel<<<(s + 255) / 256, 256, 0, st>>>(static_cast<const unsigned long long*>(d_keys), n, s,
                                                      static_cast<unsigned long long*>(S->sample));
    B2S_CUDA(cudaGetLastError(), "sample_kernel");
    B2S_NCCL(api->AllGather(S->sample, S->gathered[0], (size_t)s * S->kbytes, ncclUint8, S->comm, st), "ncclAllGather(samples)");
    size_t tb = S->sample_temp_bytes;
    B2S_RC(b2s_radix_sort(S->sample_temp, &tb, S->gathered[0], S->gathered[1], S->src[0], S->src[1], (uint64_t)G * s, kt, 4, 4, desc,
                          bb, eb, st),
           "sample sort");
    if (S->kbytes == 4)
      pick_splitters_kernel<<<1, 32, 0, st>>>(static_cast<const unsigned int*>(S->gathered[1]), S->src[1], s, G,
                                              static_cast<unsigned int*>(S->sp_keys), S->sp_ranks);
    else
      pick_splitters_kernel<<<1, 32, 0, st>>>(static_cast<const unsigned long long*>(S->gathered[1]), S->src[1], s, G,
                                              static_cast<unsigned long long*>(S->sp_keys), S->sp_ranks);
    B2S_CUDA(cudaGetLastError(), "pick_splitters_kernel");
  }
  B2S_CUDA(cudaEventRecord(S->ev[1], st), "cudaEventRecord");

  // 2. counts per destination -> G x G matrix on every rank -> my offsets in every receive buffer (all on the device)
  B2S_RC(b2s_split_count(d_keys, n, kt, desc, bb, eb, S->sp_keys, S->sp_ranks, G - 1, me, reinterpret_cast<uint64_t*>(S->counts), st),
         "b2s_split_count");
  if (G > 1) {
    B2S_NCCL(api->AllGather(S->counts, S->matrix, (size_t)G, ncclUint64, S->comm, st), "ncclAllGather(counts)");
  } else {
    B2S_CUDA(cudaMemcpyAsync(S->matrix, S->counts, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st), "cudaMemcpyAsync");
  }
  plan_kernel<<<1, 32, 0, st>>>(S->matrix, G, me, S->offsets);
  B2S_CUDA(cudaGetLastError(), "plan_kernel");
  B2S_CUDA(cudaMemcpyAsync(S->h_matrix, S->matrix, sizeof(unsigned long long) * G * G, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync");
  B2S_CUDA(cudaEventRecord(S->ev_matrix, st), "cudaEventRecord");
  B2S_CUDA(cudaEventRecord(S->ev[2], st), "cudaEventRecord");

  // 3. partition == exchange: every run of the partitioned tile is copied straight into its destination's receive buffer.
  //    No barrier is needed BEFORE the stores: they are stream-ordered after the all-gathers above, which complete only
  //    once every peer's stream has reached them, i.e. has finished the local sort of its previous call (the last reader
  //    of its receive buffer on this stream).  AFTER them a one-word all-reduce is the fence: it completes on this rank
  //    only when every rank's partition kernel -- all stores into this rank's buffer -- has completed.
  size_t tb = S->split_temp_bytes;
  B2S_RC(b2s_split_scatter(S->split_temp, &tb, d_keys, nullptr, d_values, nullptr, n, kt, S->vbytes, desc, bb, eb, S->sp_keys,
                           S->sp_ranks, G - 1, me, reinterpret_cast<const uint64_t*>(S->offsets), S->peer_k, S->vbytes ? S->peer_v : nullptr,
                           S->capacity, st),
         "b2s_split_scatter");
  B2S_CUDA(cudaEventRecord(S->ev[3], st), "cudaEventRecord");
  if (G > 1) B2S_NCCL(api->AllReduce(S->fence, S->fence, 1, ncclInt32, ncclSum, S->comm, st), "ncclAllReduce(fence)");
  B2S_CUDA(cudaEventRecord(S->ev[4], st), "cudaEventRecord");

  // the host needs the size of the final sort: the matrix has been on its way since before the partition kernel started
  B2S_CUDA(cudaEventSynchronize(S->ev_matrix), "cudaEventSynchronize(count matrix)");
  uint64_t total = 0, worst = 0;
  for (int d = 0; d < G; ++d) {
    uint64_t c = 0;
    for (int r = 0; r < G; ++r) c += S->h_matrix[r * G + d];
    if (counts_all) counts_all[d] = c;
    if (d == me) total = c;
    if (c > worst) worst = c;
  }
  S->items_sent = n - S->h_matrix[me * G + me];
  if (worst > S->capacity) {  // the same matrix on every rank: every rank takes this exit
    std::snprintf(S->err, sizeof(S->err), "receive capacity %llu too small for %llu items: raise slack (nothing was overrun)",
                  (unsigned long long)S->capacity, (unsigned long long)worst);
    return B2S_MGPU_E_CAPACITY;
  }

  // 4. one local stable sort over the G received runs (they lie in source-rank order)
  void* kb[2] = {S->recv_k[0], S->recv_k[1]};
  void* vb[2] = {S->recv_v[0], S->recv_v[1]};
  int ksel = 0, vsel = 0;
  if (total > 0) {
    size_t sb = S->sort_temp_bytes;
    B2S_RC(b2s_radix_sort_db(S->sort_temp, &sb, kb, &ksel, S->vbytes ? vb : nullptr, S->vbytes ? &vsel : nullptr, total, kt, S->vbytes,
                             8, desc, bb, eb, st),
           "local sort");
  }
  B2S_CUDA(cudaEventRecord(S->ev[5], st), "cudaEventRecord");
  S->timed = true;
  *d_keys_out = kb[ksel];
  if (d_values_out) *d_values_out = S->vbytes ? vb[vsel] : nullptr;
  *out_count = total;
  return 0;
}

int b2s_mgpu_last_phases(b2s_mgpu_sorter_t* S, float* ms6, uint64_t* items_sent) {
  using namespace b2s;
  if (!S || !ms6 || !S->timed) return B2S_MGPU_E_ARGUMENT;
  B2S_CUDA(cudaEventSynchronize(S->ev[5]), "cudaEventSynchronize");
  for (int i = 0; i < 5; ++i) B2S_CUDA(cudaEventElapsedTime(&ms6[i], S->ev[i], S->ev[i + 1]), "cudaEventElapsedTime");
  B2S_CUDA(cudaEventElapsedTime(&ms6[5], S->ev[0], S->ev[5]), "cudaEventElapsedTime");
  if (items_sent) *items_sent = S->items_sent;
  return 0;
}

uint64_t b2s_mgpu_capacity(const b2s_mgpu_sorter_t* S) { return S ? S->capacity : 0; }
const char* b2s_mgpu_last_error(const b2s_mgpu_sorter_t* S) { return S ? S->err : "no sorter"; }

int b2s_mgpu_destroy(b2s_mgpu_sorter_t* S) {
  using namespace b2s;
  if (!S) return 0;
  const NcclApi* api = nccl_api();
  cudaDeviceSynchronize();
  if (api && S->comm && S->world > 1 && S->fence) {  // nobody may still be storing into a buffer that is about to go away
    api->AllReduce(S->fence, S->fence, 1, ncclInt32, ncclSum, S->comm, nullptr);
    cudaDeviceSynchronize();
  }
  release(S);
  return 0;
}

}  // extern "C"
