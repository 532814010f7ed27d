// b2s_internal.h -- host-side interfaces between the dispatch layer and the per-key-width
// kernel translation units (split so the instantiations compile in parallel).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace b2s {

// Runtime constants of one sort call that parameterise the digit functor (see b2s_common.cuh).
struct DigitConsts {
  uint64_t xor_mask;
  uint64_t zero_img;  // floating keys: image of the zero that is collapsed onto the other one (see DigitOp)
  uint64_t pad_key;  // raw key that orders last (bit-ordered form all ones)
  bool is_float;
};

struct HistArgs {
  const void* keys;
  uint64_t n;
  DigitConsts dc;
  int begin_bit, end_bit, num_passes;
  void* ghist;
  unsigned int* done;
  unsigned int* flags;  // [num_passes] constant-digit flags (zeroed by the caller), may be null
  bool off64;
  int grid;
};

struct PassArgs {
  const void* keys_in;
  void* keys_out;
  const void* vals_in;
  void* vals_out;
  void* status;
  void* status_next;
  const void* bins;
  unsigned int* tile_counter;
  uint64_t n;
  DigitConsts dc;
  int bit;       // first bit of the digit
  int nbits;     // digit width of this pass (<= 8)
  bool off64;
  int vbytes;
  unsigned long long* trace;  // tuning builds: per-tile phase timestamps (u64[tiles][16]) of one selected pass, else null
  bool claim;                 // tile ids from an atomic ticket instead of the block index (see b2s_set_tile_claim)
  const unsigned int* skip_flag;  // DEVICE flag written by the histogram kernel: non-zero = every key has the same digit in this pass
  bool raw_in, raw_out;       // floating keys: this pass reads / writes the raw encoding (first / last pass); images in between
  // zero recording (full-range sorts of 4- / 8-byte floating keys, b2s_fzero.cu): non-null selects the ImageFloatOp kernels for
  // this pass; the first pass writes one word per 32 input keys to each plane
  unsigned int* zero_z;
  unsigned int* zero_s;
};

// Whole sort of one small tile in a single launch (b2s_single_tile.cuh).
struct SingleArgs {
  const void* keys_in;
  void* keys_out;   // may alias keys_in: the tile is held on chip between the load and the store
  const void* vals_in;
  void* vals_out;
  uint64_t n;
  DigitConsts dc;
  int begin_bit, end_bit;
  int vbytes;
};

// Segmented sort (b2s_segmented.cuh): one CTA per segment, all passes in one launch.
struct SegmentedArgs {
  const void* keys_src;
  void* keys_a;  // the last pass lands here
  void* keys_b;  // the other ping-pong buffer
  const void* vals_src;
  void* vals_a;
  void* vals_b;
  const void* begin_offsets;
  const void* end_offsets;
  int offset_bytes;  // 4 or 8: element type of the offset arrays
  uint64_t num_segments;
  DigitConsts dc;
  int begin_bit, end_bit, passes;
  int vbytes;
};

// Multi-GPU partition pass (b2s_split): destination = number of splitters ordering at or before the key.
constexpr int kMaxSplitters = 7;
struct SplitArgs {
  PassArgs pass;                     // .bit = begin_bit, .nbits unused; .bins = uint64 offsets per destination
  int end_bit;
  const void* d_splitter_keys;       // DEVICE: raw keys, ascending in sort order
  const int* d_splitter_ranks;       // DEVICE: source rank of every splitter
  int my_rank;
  int num_splitters;
  void* peer_keys[8];                // non-null => write destination d into peer_keys[d] / peer_vals[d]
  void* peer_vals[8];
  bool peer;
  bool bulk;                         // every destination base is 16-byte aligned: runs leave the SM as bulk copies
  uint64_t peer_capacity;            // items per peer receive buffer
  uint64_t* counts;                  // count launch: uint64[num_splitters + 1], zeroed by the caller
};

// Counting sort of 1- and 2-byte keys without values over all their bits (b2s_narrow.cu).
struct NarrowArgs {
  const void* keys_in;
  void* keys_out;     // never aliases keys_in
  uint64_t n;
  DigitConsts dc;
  int kbytes;
  void* counts;       // 2-byte keys: uint64[65536] joint histogram, zeroed by the caller
  void* prefix;       // 2-byte keys: uint64[65537] exclusive prefix + total; 1-byte keys: the histogram kernel's offsets
  bool prefix64;      // element type of `prefix`
  unsigned int* zflag;         // floating keys: device flag "both zeros occur"
  unsigned long long* zpartial; // floating keys: uint64[1024], zero-like keys per CTA of the zero kernels
  unsigned char* zmasks;        // floating keys: narrow_zero_mask_bytes(n) bytes, one "is +-0.0" bit per key
  int sms;
};
enum class NarrowStep { kHist16, kPrefix16, kZeroCount, kExpand, kZeroWrite };
cudaError_t narrow_step(NarrowStep step, const NarrowArgs& a, cudaStream_t s);
size_t narrow_zero_mask_bytes(uint64_t n);

// Zero recording for full-range sorts of 4- / 8-byte floating keys (b2s_fzero.cu): after the last digit pass the run of zeros
// [offset of digit 0x80 in the top pass, + number of zeros) of the output gets the recorded sign bits back, in input order.
struct FzeroArgs {
  const unsigned int* zero_z;   // planes written by the first digit pass
  const unsigned int* zero_s;
  uint64_t n;
  int kbytes;
  void* keys_out;               // final output of the sort (raw encoding)
  const void* top_bins;         // OffT[256]: exclusive digit offsets of the top pass
  bool off64;
  unsigned long long* partial;  // uint64[1024]
  int sms;
};
cudaError_t fzero_count_launch(const FzeroArgs& a, cudaStream_t s);
cudaError_t fzero_write_launch(const FzeroArgs& a, cudaStream_t s);
size_t fzero_plane_words(uint64_t n);

// One tuning point of the digit-pass kernel.
struct Variant {
  int nt, ipt, minb;
  int lbw;  // look-back window (predecessor tiles read per round trip)
  int abl;  // tuning builds only: timing ablation switches (0 in every product variant)
  int mode; // laboratory kernel only: flag bits 0-15 as documented at onesweep_kernel (b2s_onesweep.cuh); bits 16+: L2 prefetch distance (tiles)
  int flow; // >= 0: production kernel (b2s_pass.cuh) with these PF_* flags; < 0: laboratory kernel (tuning builds)
};

// Implemented once per key width in b2s_kernels.cu (-DB2S_K=1|2|4|8)
#define B2S_DECL_K(K)                                                                         \
  cudaError_t hist_launch_k##K(const HistArgs& a, cudaStream_t s);                            \
  cudaError_t onesweep_launch_k##K(int variant, const PassArgs& a, cudaStream_t s);           \
  int onesweep_tile_k##K(int variant, int vbytes, bool is_float, bool off64);                          \
  int onesweep_num_variants_k##K();                                                           \
  Variant onesweep_variant_k##K(int variant, int vbytes, bool is_float, bool off64);                   \
  cudaError_t split_count_launch_k##K(const SplitArgs& a, cudaStream_t s);                    \
  cudaError_t split_launch_k##K(const SplitArgs& a, cudaStream_t s);                          \
  int split_tile_k##K(int vbytes);                                                            \
  cudaError_t single_launch_k##K(const SingleArgs& a, cudaStream_t s);                        \
  cudaError_t segmented_launch_k##K(const SegmentedArgs& a, cudaStream_t s);                  \
  int single_tile_items_k##K(int vbytes);
B2S_DECL_K(1)
B2S_DECL_K(2)
B2S_DECL_K(4)
B2S_DECL_K(8)
#undef B2S_DECL_K

}  // namespace b2s
