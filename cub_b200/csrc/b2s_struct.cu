// b2s_struct.cu -- DeviceRadixSort for user-defined key structs (the "decomposer" overloads), C-ABI b2s_radix_sort_struct[_db].
//
// Replaces (reference, for parity of RESULT only):
//   cub::DeviceRadixSort::{SortKeys,SortPairs}[Descending](..., decomposer[, begin_bit, end_bit])   16 overloads,
//       cub/device/device_radix_sort.cuh:486-530, 625-666, 922-962, 1055-1105, 1368-1426, 1515-1563, 1816-1856, 1949-...
//   the multi-field digit extractor machinery  cub/block/radix_rank_sort_operations.cuh:142-571
//   known answers  test/catch2_test_device_radix_sort_custom.cu:593-1690 (tests/golden/decomposer_kats.json)
//
// Semantics: a key is a struct; the decomposer names its arithmetic fields, most significant first.  The bit-ordered
// image of a key is the concatenation of the fields' bit-ordered images (bit 0 in the LAST field; per field: unsigned as
// is, signed with the sign bit flipped, floating x ^ (sign ? ~0 : HIGH)), complemented as a whole when descending, the two
// zeros of a floating field sharing one image as in the reference's onesweep path (see field_image); the sort is the stable sort on bits [begin_bit, end_bit) of that image.
//
// B200-first design (NOT the reference's): instead of dragging whole structs through every digit pass with a multi-field
// digit extractor, (1) a pack kernel writes bits [begin_bit, end_bit) of each key's image as one u64 word next to the
// item's index, (2) ONE stable SortPairs<u64, index> of this library over exactly the bits that exist (LSD over 64-bit
// words for wider ranges: pack word w through the current permutation, sort again), (3) a gather kernel moves structs
// and values through the sorted index.  Cost for the common <= 64-bit composite: pack + ceil(bits/8) passes over 12-byte
// pairs + gather, however large the struct is.
#include <cuda_runtime.h>

#include <cstring>

#include "../../include/b2s_radix_sort.h"

namespace b2s {
namespace {

constexpr int MAX_FIELDS = B2S_MAX_STRUCT_FIELDS;

struct Field {
  int offset;    // byte offset inside the struct
  int bytes;     // 1, 2, 4, 8
  int category;  // 0 unsigned, 1 signed, 2 floating
  int lo;        // bit position of the field's least significant bit inside the concatenated image
};
struct StructDesc {
  Field f[MAX_FIELDS];
  int num_fields;
  int struct_bytes;
  int descending;
};

const int kBytes[B2S_KEY_TYPE_COUNT] = {1, 1, 2, 2, 2, 2, 4, 4, 4, 8, 8, 8};
const int kCat[B2S_KEY_TYPE_COUNT] = {0, 1, 0, 1, 2, 2, 0, 1, 2, 0, 1, 2};

__device__ __forceinline__ unsigned long long load_field(const unsigned char* p, int bytes) {
  // fields are naturally aligned inside a naturally aligned struct in every sane layout; a packed struct takes the byte path
  if ((reinterpret_cast<uintptr_t>(p) & (uintptr_t)(bytes - 1)) == 0) {
    switch (bytes) {
      case 1: return *p;
      case 2: return *reinterpret_cast<const unsigned short*>(p);
      case 4: return *reinterpret_cast<const unsigned int*>(p);
      default: return *reinterpret_cast<const unsigned long long*>(p);
    }
  }
  unsigned long long v = 0;
  for (int b = 0; b < bytes; ++b) v |= (unsigned long long)p[b] << (8 * b);
  return v;
}

// bit-ordered image of one field (low 8*bytes bits)
__device__ __forceinline__ unsigned long long field_image(unsigned long long k, const Field& f, int descending) {
  const int bits = f.bytes * 8;
  const unsigned long long ones = bits == 64 ? ~0ull : ((1ull << bits) - 1);
  const unsigned long long high = 1ull << (bits - 1);
  if (f.category == 1) k ^= high;
  if (f.category == 2) k ^= (k & high) ? ones : high;
  if (descending) k ^= ones;
  k &= ones;
  // -0.0 and +0.0 share one image.  As in the reference's onesweep path (radix_rank_sort_operations.cuh:55-66, 79-89) it is
  // HIGH in BOTH directions: ascending -0.0 (image ~HIGH) is mapped onto +0.0, descending +0.0 (complemented image ~HIGH)
  // onto -0.0.  Only partial bit ranges can tell; the device-wide sort (DigitOp) follows the same rule.
  if (f.category == 2 && k == (ones ^ high)) k = high;
  return k;
}

// word[i] = bits [win_lo, win_lo + win_bits) of the image of key (perm ? perm[i] : i); win_bits <= 64
template <typename IdxT>
__global__ void pack_kernel(const unsigned char* keys, const IdxT* perm, unsigned long long n, StructDesc d, int win_lo,
                            int win_bits, unsigned long long* word, IdxT* iota_out) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned long long src = perm ? (unsigned long long)perm[i] : i;
    const unsigned char* rec = keys + src * (unsigned long long)d.struct_bytes;
    unsigned long long w = 0;
    for (int j = 0; j < d.num_fields; ++j) {
      const Field f = d.f[j];
      const int a = f.lo > win_lo ? f.lo : win_lo;
      const int fe = f.lo + f.bytes * 8, we = win_lo + win_bits;
      const int e = fe < we ? fe : we;
      if (e <= a) continue;
      const unsigned long long img = field_image(load_field(rec + f.offset, f.bytes), f, d.descending);
      const int take = e - a;
      const unsigned long long part = (img >> (a - f.lo)) & (take == 64 ? ~0ull : ((1ull << take) - 1));
      w |= part << (a - win_lo);
    }
    word[i] = w;
    if (iota_out) iota_out[i] = (IdxT)i;
  }
}

// out[i] = in[perm[i]] for items of `bytes` bytes
template <typename IdxT, typename T>
__global__ void gather_kernel(const T* in, const IdxT* perm, unsigned long long n, T* out) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[perm[i]];
}
template <typename IdxT>
__global__ void gather_bytes_kernel(const unsigned char* in, const IdxT* perm, unsigned long long n, int bytes, unsigned char* out) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned char* s = in + (unsigned long long)perm[i] * bytes;
    unsigned char* o = out + i * (unsigned long long)bytes;
    for (int b = 0; b < bytes; ++b) o[b] = s[b];
  }
}

struct A16 { unsigned long long a, b; };

template <typename IdxT>
cudaError_t gather(const void* in, const IdxT* perm, uint64_t n, int bytes, void* out, cudaStream_t s) {
  unsigned long long g = (n + 255) / 256;
  if (g > 148ull * 32) g = 148ull * 32;
  const uintptr_t al = reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out);
  if (bytes == 16 && (al & 7) == 0)
    gather_kernel<<<(unsigned)g, 256, 0, s>>>((const A16*)in, perm, n, (A16*)out);
  else if (bytes == 8 && (al & 7) == 0)
    gather_kernel<<<(unsigned)g, 256, 0, s>>>((const unsigned long long*)in, perm, n, (unsigned long long*)out);
  else if (bytes == 4 && (al & 3) == 0)
    gather_kernel<<<(unsigned)g, 256, 0, s>>>((const unsigned int*)in, perm, n, (unsigned int*)out);
  else if (bytes == 2 && (al & 1) == 0)
    gather_kernel<<<(unsigned)g, 256, 0, s>>>((const unsigned short*)in, perm, n, (unsigned short*)out);
  else if (bytes == 1)
    gather_kernel<<<(unsigned)g, 256, 0, s>>>((const unsigned char*)in, perm, n, (unsigned char*)out);
  else
    gather_bytes_kernel<<<(unsigned)g, 256, 0, s>>>((const unsigned char*)in, perm, n, bytes, (unsigned char*)out);
  return cudaGetLastError();
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

bool make_desc(StructDesc& d, int struct_bytes, const b2s_field_t* fields, int num_fields, int descending, int* total_bits) {
  if (!fields || num_fields < 1 || num_fields > MAX_FIELDS || struct_bytes < 1) return false;
  d.num_fields = num_fields;
  d.struct_bytes = struct_bytes;
  d.descending = descending != 0;
  int lo = 0;
  for (int j = num_fields - 1; j >= 0; --j) {  // the LAST field holds bit 0
    const int kt = fields[j].key_type;
    if (kt < 0 || kt >= B2S_KEY_TYPE_COUNT) return false;
    d.f[j].offset = fields[j].offset;
    d.f[j].bytes = kBytes[kt];
    d.f[j].category = kCat[kt];
    d.f[j].lo = lo;
    if (fields[j].offset < 0 || fields[j].offset + kBytes[kt] > struct_bytes) return false;
    lo += kBytes[kt] * 8;
  }
  *total_bits = lo;
  return true;
}

// shared implementation: kin -> kout (pointer form, kin never written); the DoubleBuffer wrapper flips the selector
int struct_sort_impl(void* d_temp, size_t* temp_bytes, const void* kin, void* kout, const void* vin, void* vout, uint64_t n,
                     int struct_bytes, const b2s_field_t* fields, int num_fields, int value_bytes, int descending,
                     int begin_bit, int end_bit, bool copy_when_empty_range, cudaStream_t stream) {
  if (!temp_bytes) return (int)cudaErrorInvalidValue;
  StructDesc d;
  int total_bits = 0;
  if (!make_desc(d, struct_bytes, fields, num_fields, descending, &total_bits)) return (int)cudaErrorInvalidValue;
  if (value_bytes < 0 || value_bytes > 64) return (int)cudaErrorInvalidValue;
  if (end_bit < 0) end_bit = total_bits;
  if (begin_bit < 0 || end_bit > total_bits) return (int)cudaErrorInvalidValue;
  const int nbits = end_bit - begin_bit;
  if (n == 0 || (nbits <= 0 && !copy_when_empty_range)) {
    if (!d_temp) *temp_bytes = 1;
    return (int)cudaSuccess;
  }
  if (nbits <= 0) {  // pointer form with an empty bit range: copy (dispatch_radix_sort.cuh:1955-1963)
    if (!d_temp) {
      *temp_bytes = 1;
      return (int)cudaSuccess;
    }
    cudaError_t e = cudaMemcpyAsync(kout, kin, (size_t)n * struct_bytes, cudaMemcpyDeviceToDevice, stream);
    if (e == cudaSuccess && value_bytes) e = cudaMemcpyAsync(vout, vin, (size_t)n * value_bytes, cudaMemcpyDeviceToDevice, stream);
    return (int)e;
  }
  const bool idx64 = n > 0xffffffffull;
  const int ib = idx64 ? 8 : 4;
  const int words = (nbits + 63) / 64;
  // temp carving: word[2][n] u64, idx[2][n], inner sort temp
  size_t inner = 0;
  {
    void* kb[2] = {nullptr, nullptr};
    void* vb[2] = {nullptr, nullptr};
    int ks = 0, vs = 0;
    const int rc = b2s_radix_sort_db(nullptr, &inner, kb, &ks, vb, &vs, n, B2S_U64, ib, 8, 0, 0, nbits < 64 ? nbits : 64, nullptr);
    if (rc != 0) return rc;
  }
  size_t o = 0;
  const size_t off_w0 = o;  o += align_up((size_t)n * 8, 256);
  const size_t off_w1 = o;  o += align_up((size_t)n * 8, 256);
  const size_t off_i0 = o;  o += align_up((size_t)n * ib, 256);
  const size_t off_i1 = o;  o += align_up((size_t)n * ib, 256);
  const size_t off_in = o;  o += inner;
  const size_t total = o + 255;
  if (!d_temp) {
    *temp_bytes = total;
    return (int)cudaSuccess;
  }
  if (*temp_bytes < total) return (int)cudaErrorInvalidValue;
  unsigned char* base = reinterpret_cast<unsigned char*>(align_up(reinterpret_cast<uintptr_t>(d_temp), 256));
  void* wbuf[2] = {base + off_w0, base + off_w1};
  void* ibuf[2] = {base + off_i0, base + off_i1};
  unsigned long long g = (n + 255) / 256;
  if (g > 148ull * 32) g = 148ull * 32;
  int isel = 0;
  for (int w = 0; w < words; ++w) {
    const int win_lo = begin_bit + 64 * w;
    const int win_bits = (end_bit - win_lo) < 64 ? (end_bit - win_lo) : 64;
    // the packed words of this round go next to the CURRENT permutation (first round: the identity, written by the kernel)
    if (idx64)
      pack_kernel<unsigned long long><<<(unsigned)g, 256, 0, stream>>>(
          (const unsigned char*)kin, w ? (const unsigned long long*)ibuf[isel] : nullptr, n, d, win_lo, win_bits,
          (unsigned long long*)wbuf[0], w ? nullptr : (unsigned long long*)ibuf[0]);
    else
      pack_kernel<unsigned int><<<(unsigned)g, 256, 0, stream>>>(
          (const unsigned char*)kin, w ? (const unsigned int*)ibuf[isel] : nullptr, n, d, win_lo, win_bits,
          (unsigned long long*)wbuf[0], w ? nullptr : (unsigned int*)ibuf[0]);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    void* kb[2] = {wbuf[0], wbuf[1]};
    void* vb[2] = {ibuf[isel], ibuf[isel ^ 1]};
    int ks = 0, vs = 0;
    size_t ibytes = inner;
    const int rc = b2s_radix_sort_db(base + off_in, &ibytes, kb, &ks, vb, &vs, n, B2S_U64, ib, 8, 0, 0, win_bits, stream);
    if (rc != 0) return rc;
    isel ^= vs;  // vb[] was passed in (current, alternate) order
  }
  cudaError_t e = idx64 ? gather<unsigned long long>(kin, (const unsigned long long*)ibuf[isel], n, struct_bytes, kout, stream)
                        : gather<unsigned int>(kin, (const unsigned int*)ibuf[isel], n, struct_bytes, kout, stream);
  if (e != cudaSuccess) return (int)e;
  if (value_bytes)
    e = idx64 ? gather<unsigned long long>(vin, (const unsigned long long*)ibuf[isel], n, value_bytes, vout, stream)
              : gather<unsigned int>(vin, (const unsigned int*)ibuf[isel], n, value_bytes, vout, stream);
  return (int)e;
}

}  // namespace
}  // namespace b2s

extern "C" {

int b2s_radix_sort_struct(void* d_temp_storage, size_t* temp_storage_bytes, const void* d_keys_in, void* d_keys_out,
                          const void* d_values_in, void* d_values_out, uint64_t num_items, int key_struct_bytes,
                          const b2s_field_t* fields, int num_fields, int value_bytes, int descending, int begin_bit, int end_bit,
                          b2s_stream_t stream) {
  return b2s::struct_sort_impl(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out, num_items,
                               key_struct_bytes, fields, num_fields, value_bytes, descending, begin_bit, end_bit, true,
                               (cudaStream_t)stream);
}

int b2s_radix_sort_struct_db(void* d_temp_storage, size_t* temp_storage_bytes, void* key_bufs[2], int* key_selector,
                             void* val_bufs[2], int* val_selector, uint64_t num_items, int key_struct_bytes,
                             const b2s_field_t* fields, int num_fields, int value_bytes, int descending, int begin_bit,
                             int end_bit, b2s_stream_t stream) {
  if (!key_bufs || !key_selector) return (int)cudaErrorInvalidValue;
  if (value_bytes && (!val_bufs || !val_selector)) return (int)cudaErrorInvalidValue;
  const int ks = *key_selector & 1;
  const int vs = value_bytes ? (*val_selector & 1) : 0;
  // total bits of the image, to recognise the empty range (no-op for the DoubleBuffer form, dispatch_radix_sort.cuh:1945)
  int total_bits = 0;
  for (int j = 0; j < num_fields && fields; ++j) total_bits += 8 * b2s_key_bytes(fields[j].key_type);
  const int eb = end_bit < 0 ? total_bits : end_bit;
  const int rc = b2s::struct_sort_impl(d_temp_storage, temp_storage_bytes, key_bufs[ks], key_bufs[ks ^ 1],
                                       value_bytes ? val_bufs[vs] : nullptr, value_bytes ? val_bufs[vs ^ 1] : nullptr, num_items,
                                       key_struct_bytes, fields, num_fields, value_bytes, descending, begin_bit, end_bit, false,
                                       (cudaStream_t)stream);
  if (rc == 0 && d_temp_storage && num_items > 0 && eb > begin_bit) {
    *key_selector = ks ^ 1;
    if (value_bytes) *val_selector = vs ^ 1;
  }
  return rc;
}

}  // extern "C"
