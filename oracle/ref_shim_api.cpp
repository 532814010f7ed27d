// oracle/ref_shim_api.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
// Front of oracle/_ref/lib<prefix>.so: same C signatures as include/b2s_radix_sort.h
// (b2s_radix_sort / b2s_radix_sort_db) so tests and bench can drive the reference and
// the product through one ctypes prototype.  Dispatches to the per-key-group TUs of
// ref_shim.cu.
#include <cstddef>
#include <cstdint>

#ifndef REF_PREFIX
#define REF_PREFIX ref_cub
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define GROUP(n) CAT(CAT(REF_PREFIX, _group), n)

#define DECL(n)                                                                                               \
  extern "C" int GROUP(n)(int, void*, size_t*, const void*, void*, const void*, void*, void**, int*, void**,  \
                          int*, uint64_t, int, int, int, int, int, int, void*);
DECL(0) DECL(1) DECL(2) DECL(3) DECL(4) DECL(5) DECL(6)

static int route(int db, void* tmp, size_t* bytes, const void* kin, void* kout, const void* vin, void* vout,
                 void** kb, int* ksel, void** vb, int* vsel, uint64_t n, int kt, int vbytes, int obytes, int desc,
                 int bb, int eb, void* s) {
#define CALL(g) return GROUP(g)(db, tmp, bytes, kin, kout, vin, vout, kb, ksel, vb, vsel, n, kt, vbytes, obytes, desc, bb, eb, s)
  switch (kt) {
    case 0: case 1: CALL(0);
    case 2: case 3: CALL(1);
    case 4: case 5: CALL(2);
    case 6: CALL(3);
    case 7: case 8: CALL(4);
    case 9: CALL(5);
    case 10: case 11: CALL(6);
    default: return -1;
  }
}

extern "C" int CAT(REF_PREFIX, _radix_sort)(void* tmp, size_t* bytes, const void* kin, void* kout, const void* vin,
                                            void* vout, uint64_t n, int kt, int vbytes, int obytes, int desc,
                                            int bb, int eb, void* s) {
  return route(0, tmp, bytes, kin, kout, vin, vout, nullptr, nullptr, nullptr, nullptr, n, kt, vbytes, obytes, desc,
               bb, eb, s);
}

extern "C" int CAT(REF_PREFIX, _radix_sort_db)(void* tmp, size_t* bytes, void** kb, int* ksel, void** vb, int* vsel,
                                               uint64_t n, int kt, int vbytes, int obytes, int desc, int bb, int eb,
                                               void* s) {
  return route(1, tmp, bytes, nullptr, nullptr, nullptr, nullptr, kb, ksel, vb, vsel, n, kt, vbytes, obytes, desc,
               bb, eb, s);
}
