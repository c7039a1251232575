// Host-side weight packing shared by the tcgen05 kernels: fp32 nn.Linear weights -> bf16 hi / lo = bf16(w - hi)
// planes cut into stages that are byte images of the shared-memory B-operand layout.
#pragma once
#include <stdint.h>
#include <string.h>

#include <cuda_fp16.h>

#include <vector>

namespace dfn {
namespace tc {

static inline uint16_t f2bf(float f) {  // round-to-nearest-even, as cvt.rn.bf16.f32
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

struct Packer {
  std::vector<uint8_t> hi32, lo32;   // [n rows x 32 K] stages, 64-byte swizzle (mlp_tc.cu)
  std::vector<uint8_t> h16;          // the hi32 images with fp16 instead of bf16 values (DFN_PREC_FP16), same offsets
  std::vector<uint8_t> l16;          // fp16 residuals fp16(w - fp16(w)), same offsets (DFN_PREC_FP16X3M)
  std::vector<uint8_t> hi2;          // [n/2 rows x 64 K] per CTA of a pair, 128-byte swizzle (mlp_pair.cu)
  std::vector<uint8_t> h16_2;        // the hi2 images with fp16 values (DFN_PREC_FP16), same offsets
  uint32_t last32 = 0;               // offset of the last layer added to hi32
  uint32_t last2 = 0;                // offset of the last layer added to hi2
  bool want2 = true;                 // build the cta_group::2 images
  // host-only introspection (dfn_*_program_host): dense fp32 weights, [layer][256 rows][6 input-block slots][64]
  static constexpr size_t kDenseLayer = (size_t)256 * 6 * 64;
  std::vector<float>* dense = nullptr;
  int n_dense = 0;
  // Appends the stages of one layer: for each K-block, for each chunk of <=128 output rows, a
  // [rows x 64] bf16 image in the swizzled K-major layout.  wfun(n, kbi, k) returns W[n][column
  // of K-block kbi, position k] or 0.
  template <class F>
  uint32_t add_layer(int n_out, int nkb, F wfun) {
    const uint32_t start = (uint32_t)hi32.size();
    if (dense) {
      dense->resize((size_t)(n_dense + 1) * kDenseLayer, 0.f);
      float* d = dense->data() + (size_t)n_dense * kDenseLayer;
      for (int r = 0; r < n_out; ++r)
        for (int kbi = 0; kbi < nkb; ++kbi)
          for (int k = 0; k < 64; ++k) d[((size_t)r * 6 + kbi) * 64 + k] = wfun(r, kbi, k);
      ++n_dense;
    }
    // K = 32 stages: for each K-block, for each half of it, all n_out rows x 32 K; 64-byte rows, 16-byte
    // chunk index XORed with (row >> 1) & 3 (cute Swizzle<2,4,3>), 8-row groups 512 bytes apart.
    last32 = (uint32_t)hi32.size();
    for (int kbi = 0; kbi < nkb; ++kbi) {
      for (int kh = 0; kh < 2; ++kh) {
        const size_t base = hi32.size();
        hi32.resize(base + (size_t)n_out * 64, 0);
        lo32.resize(base + (size_t)n_out * 64, 0);
        h16.resize(base + (size_t)n_out * 64, 0);
        l16.resize(base + (size_t)n_out * 64, 0);
        for (int r = 0; r < n_out; ++r) {
          for (int k = 0; k < 32; ++k) {
            const float w = wfun(r, kbi, kh * 32 + k);
            const uint16_t h = f2bf(w);
            const uint16_t l = f2bf(w - bf2f(h));
            const size_t o = base + (size_t)r * 64 + ((((size_t)k >> 3) ^ (((size_t)r >> 1) & 3)) << 4) + ((size_t)k & 7) * 2;
            memcpy(&hi32[o], &h, 2);
            memcpy(&lo32[o], &l, 2);
            const __half hh = __float2half_rn(w);
            memcpy(&h16[o], &hh, 2);
            const __half hl = __float2half_rn(w - __half2float(hh));
            memcpy(&l16[o], &hl, 2);
          }
        }
      }
    }
    // cta_group::2 stages (mlp_pair.cu): for each CTA rank of the pair, for each K-block, rows [rank*n/2, (rank+1)*n/2) x 64 K in the
    // 128-byte-swizzled K-major layout (the pair's MMA takes half of B's N rows from each CTA's shared memory).  Rank-major, so that a
    // CTA's halves of consecutive K-blocks are contiguous: one bulk copy fills a ring entry of two K-blocks.
    last2 = (uint32_t)hi2.size();
    if (want2 && n_out % 16 == 0) {
      const int half = n_out / 2;
      for (int rank = 0; rank < 2; ++rank) {
        for (int kbi = 0; kbi < nkb; ++kbi) {
          const size_t base = hi2.size();
          hi2.resize(base + (size_t)half * 128, 0);
          h16_2.resize(base + (size_t)half * 128, 0);
          for (int r = 0; r < half; ++r) {
            for (int k = 0; k < 64; ++k) {
              const float w = wfun(rank * half + r, kbi, k);
              const uint16_t h = f2bf(w);
              const size_t o = base + (size_t)r * 128 + ((((size_t)k >> 3) ^ ((size_t)r & 7)) << 4) + ((size_t)k & 7) * 2;
              memcpy(&hi2[o], &h, 2);
              const __half hh = __float2half_rn(w);
              memcpy(&h16_2[o], &hh, 2);
            }
          }
        }
      }
    }
    return start;
  }
};


}  // namespace tc
}  // namespace dfn
