// oracle/ref_shim.cu -- TEST INFRASTRUCTURE, NOT PRODUCT.
//
// Thin extern "C" wrappers around the UNMODIFIED reference cub::DeviceRadixSort, compiled
// from the headers where they lie (-I/root/reference, CUB 2.2.0) into oracle/_ref/.
// Only tests/, __graft_entry__.smoke() and bench.py may load the resulting library; the
// product (cub_b200/) never links or loads it.  No reference source is copied here: this
// file only *calls* the reference's public API (cub/device/device_radix_sort.cuh:312, 781,
// 1214, 1675, 2106, 2525, 2921, 3330).
//
// The same file is compiled a second time against the CUDA toolkit's bundled CUB
// (2.8.2, has an SM100 tuning policy) as an informational speed comparator
// (-DREF_PREFIX=tk_cub, no -I/root/reference).
//
// The reference namespace is wrapped (CUB_WRAPPED_NAMESPACE) so the library can live in
// one process together with anything else that uses CUB (e.g. torch).
//
// One translation unit per key type (-DREF_KEY_GROUP=n) to parallelise the build.
#include <cub/device/device_radix_sort.cuh>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdint>

#ifndef REF_NS
#define REF_NS refcub
#endif

namespace rc = REF_NS::cub;

namespace {

struct V16 { unsigned long long a, b; };

template <typename KeyT, typename ValueT, typename NumItemsT>
cudaError_t sort_ptr(void* tmp, size_t& bytes, const void* kin, void* kout, const void* vin, void* vout,
                     uint64_t n, bool desc, int bb, int eb, cudaStream_t s) {
  if constexpr (std::is_same<ValueT, rc::NullType>::value) {
    return desc ? rc::DeviceRadixSort::SortKeysDescending(tmp, bytes, (const KeyT*)kin, (KeyT*)kout, (NumItemsT)n, bb, eb, s)
                : rc::DeviceRadixSort::SortKeys(tmp, bytes, (const KeyT*)kin, (KeyT*)kout, (NumItemsT)n, bb, eb, s);
  } else {
    return desc ? rc::DeviceRadixSort::SortPairsDescending(tmp, bytes, (const KeyT*)kin, (KeyT*)kout, (const ValueT*)vin,
                                                            (ValueT*)vout, (NumItemsT)n, bb, eb, s)
                : rc::DeviceRadixSort::SortPairs(tmp, bytes, (const KeyT*)kin, (KeyT*)kout, (const ValueT*)vin,
                                                  (ValueT*)vout, (NumItemsT)n, bb, eb, s);
  }
}

template <typename KeyT, typename ValueT, typename NumItemsT>
cudaError_t sort_db(void* tmp, size_t& bytes, void** kb, int* ksel, void** vb, int* vsel, uint64_t n, bool desc,
                    int bb, int eb, cudaStream_t s) {
  rc::DoubleBuffer<KeyT> dk((KeyT*)kb[0], (KeyT*)kb[1]);
  dk.selector = *ksel;
  cudaError_t e;
  if constexpr (std::is_same<ValueT, rc::NullType>::value) {
    e = desc ? rc::DeviceRadixSort::SortKeysDescending(tmp, bytes, dk, (NumItemsT)n, bb, eb, s)
             : rc::DeviceRadixSort::SortKeys(tmp, bytes, dk, (NumItemsT)n, bb, eb, s);
  } else {
    rc::DoubleBuffer<ValueT> dv((ValueT*)vb[0], (ValueT*)vb[1]);
    dv.selector = *vsel;
    e = desc ? rc::DeviceRadixSort::SortPairsDescending(tmp, bytes, dk, dv, (NumItemsT)n, bb, eb, s)
             : rc::DeviceRadixSort::SortPairs(tmp, bytes, dk, dv, (NumItemsT)n, bb, eb, s);
    if (tmp) *vsel = dv.selector;
  }
  if (tmp) *ksel = dk.selector;
  return e;
}

// value_bytes / offset_bytes fan-out for one key type. `Wide` enables the rarely used value widths.
template <typename KeyT, bool Wide>
int fan(bool db, void* tmp, size_t* bytes, const void* kin, void* kout, const void* vin, void* vout, void** kb,
        int* ksel, void** vb, int* vsel, uint64_t n, int vbytes, int obytes, bool desc, int bb, int eb,
        cudaStream_t s) {
#define GO(VT, NT)                                                                              \
  return (int)(db ? sort_db<KeyT, VT, NT>(tmp, *bytes, kb, ksel, vb, vsel, n, desc, bb, eb, s)  \
                  : sort_ptr<KeyT, VT, NT>(tmp, *bytes, kin, kout, vin, vout, n, desc, bb, eb, s))
  if (obytes == 4) {
    switch (vbytes) {
      case 0: GO(rc::NullType, uint32_t);
      case 4: GO(uint32_t, uint32_t);
      case 8: GO(unsigned long long, uint32_t);
      default: break;
    }
    if constexpr (Wide) {
      switch (vbytes) {
        case 1: GO(uint8_t, uint32_t);
        case 2: GO(uint16_t, uint32_t);
        case 16: GO(V16, uint32_t);
        default: break;
      }
    }
  } else if (obytes == 8) {
    if constexpr (Wide) {
      switch (vbytes) {
        case 0: GO(rc::NullType, unsigned long long);
        case 4: GO(uint32_t, unsigned long long);
        default: break;
      }
    }
  }
#undef GO
  return -1;  // combination not instantiated in the shim
}

}  // namespace

#ifndef REF_PREFIX
#define REF_PREFIX ref_cub
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

// Each TU defines ref_cub_group<N>(...) for the key types of its group.
#ifndef REF_KEY_GROUP
#error "REF_KEY_GROUP must be defined (0..5)"
#endif

extern "C" int CAT(CAT(REF_PREFIX, _group), REF_KEY_GROUP)(int db, void* tmp, size_t* bytes, const void* kin, void* kout,
                                                          const void* vin, void* vout, void** kb, int* ksel,
                                                          void** vb, int* vsel, uint64_t n, int key_type, int vbytes,
                                                          int obytes, int desc, int bb, int eb, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
#define ARGS db != 0, tmp, bytes, kin, kout, vin, vout, kb, ksel, vb, vsel, n, vbytes, obytes, desc != 0, bb, eb, s
  switch (key_type) {
#if REF_KEY_GROUP == 0
    case 0: return fan<uint8_t, false>(ARGS);
    case 1: return fan<int8_t, false>(ARGS);
#elif REF_KEY_GROUP == 1
    case 2: return fan<uint16_t, false>(ARGS);
    case 3: return fan<int16_t, false>(ARGS);
#elif REF_KEY_GROUP == 2
    case 4: return fan<__half, false>(ARGS);
    case 5: return fan<__nv_bfloat16, false>(ARGS);
#elif REF_KEY_GROUP == 3
    case 6: return fan<uint32_t, true>(ARGS);
#elif REF_KEY_GROUP == 4
    case 7: return fan<int32_t, false>(ARGS);
    case 8: return fan<float, false>(ARGS);
#elif REF_KEY_GROUP == 5
    case 9: return fan<unsigned long long, true>(ARGS);
#elif REF_KEY_GROUP == 6
    case 10: return fan<long long, false>(ARGS);
    case 11: return fan<double, false>(ARGS);
#endif
    default: return -1;
  }
#undef ARGS
}
