// Epilogue building blocks shared by mlp_tc.cu and mlp_tc2.cu: accumulator (TMEM) -> + bias -> ReLU -> bf16 (hi[/lo])
// -> 128-byte-swizzled K-major activation blocks in shared memory.
#pragma once
#include "tc_ptx.cuh"

namespace dfn {
namespace tc {

static constexpr int KB_BYTES = TILE_M * 128;      // one activation K-block: 128 rows x 64 bf16

// Writes 8 consecutive columns (one 16-byte chunk) of this thread's row.
template <bool X3, bool F16 = false>
__device__ __forceinline__ void store_chunk(uint8_t* blk_hi, uint8_t* blk_lo, uint32_t row, uint32_t chunk,
                                            const float (&v)[8]) {
  uint4 h;
  if (F16) {
    h.x = pack_f16(v[0], v[1]);
    h.y = pack_f16(v[2], v[3]);
    h.z = pack_f16(v[4], v[5]);
    h.w = pack_f16(v[6], v[7]);
  } else {
    h.x = pack_bf16(v[0], v[1]);
    h.y = pack_bf16(v[2], v[3]);
    h.z = pack_bf16(v[4], v[5]);
    h.w = pack_bf16(v[6], v[7]);
  }
  *reinterpret_cast<uint4*>(blk_hi + swz(row, chunk)) = h;
  if (X3) {
    uint4 l;
    l.x = pack_bf16(v[0] - bf16_lo_f(h.x), v[1] - bf16_hi_f(h.x));
    l.y = pack_bf16(v[2] - bf16_lo_f(h.y), v[3] - bf16_hi_f(h.y));
    l.z = pack_bf16(v[4] - bf16_lo_f(h.z), v[5] - bf16_hi_f(h.z));
    l.w = pack_bf16(v[6] - bf16_lo_f(h.w), v[7] - bf16_hi_f(h.w));
    *reinterpret_cast<uint4*>(blk_lo + swz(row, chunk)) = l;
  }
}

// One 32-column chunk of this thread's row: + bias, ReLU, bf16 (hi[/lo]) and four 16-byte stores into the
// swizzled K-block.  `chunk32` = index of the 32-column chunk inside the layer output.
template <bool X3, bool GLOBAL_BIAS, bool F16 = false>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], int chunk32, const float* gbias, uint32_t sbias,
                                               uint8_t* arena_hi, uint8_t* arena_lo, uint32_t row) {
  uint8_t* dst_hi = arena_hi + (size_t)(chunk32 >> 1) * KB_BYTES;
  uint8_t* dst_lo = arena_lo + (size_t)(chunk32 >> 1) * KB_BYTES;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float b[8];
    if (GLOBAL_BIAS) ldg_f32x8(gbias + chunk32 * 32 + g * 8, b);
    else lds_f32x8(sbias + (uint32_t)(chunk32 * 32 + g * 8) * 4u, b);
    const uint32_t c16 = (uint32_t)((chunk32 & 1) * 4 + g);
    if (!X3) {
      uint4 h;
      if (F16) {
        h.x = add_relu_pack_f16(v[g * 8 + 0], v[g * 8 + 1], b[0], b[1]);
        h.y = add_relu_pack_f16(v[g * 8 + 2], v[g * 8 + 3], b[2], b[3]);
        h.z = add_relu_pack_f16(v[g * 8 + 4], v[g * 8 + 5], b[4], b[5]);
        h.w = add_relu_pack_f16(v[g * 8 + 6], v[g * 8 + 7], b[6], b[7]);
      } else {
        h.x = add_relu_pack(v[g * 8 + 0], v[g * 8 + 1], b[0], b[1]);
        h.y = add_relu_pack(v[g * 8 + 2], v[g * 8 + 3], b[2], b[3]);
        h.z = add_relu_pack(v[g * 8 + 4], v[g * 8 + 5], b[4], b[5]);
        h.w = add_relu_pack(v[g * 8 + 6], v[g * 8 + 7], b[6], b[7]);
      }
      *reinterpret_cast<uint4*>(dst_hi + swz(row, c16)) = h;
    } else {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaxf(__uint_as_float(v[g * 8 + e]) + b[e], 0.f);
      store_chunk<true>(dst_hi, dst_lo, row, c16, o);
    }
  }
}

// Decoder density head on the CUDA cores (TC_F_DOT_SIGMA, single-pass precisions): epilogue_chunk with one extra FMA
// per element, dot[] += relu(v + bias) * w[column] on the fp32 activations, four partial sums.  sdot: the head row in
// shared memory as packed bf16 / fp16 pairs (the precision the MMA path gives its weights; 512 bytes is what is left).
template <bool F16>
__device__ __forceinline__ void epilogue_chunk_dot1(const uint32_t (&v)[32], int chunk32, uint32_t sbias, uint32_t sdot,
                                                    uint8_t* arena_hi, uint32_t row, float (&dot)[4]) {
  uint8_t* dst_hi = arena_hi + (size_t)(chunk32 >> 1) * KB_BYTES;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int c = chunk32 * 32 + g * 8;
    float b[8], w[8], x[8];
    uint32_t wp[4];
    lds_f32x8(sbias + (uint32_t)c * 4u, b);
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(wp[0]), "=r"(wp[1]), "=r"(wp[2]), "=r"(wp[3]) : "r"(sdot + (uint32_t)c * 2u));
#pragma unroll
    for (int q = 0; q < 4; ++q) unpack_h2<F16>(wp[q], w[2 * q], w[2 * q + 1]);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      x[e] = fmaxf(__uint_as_float(v[g * 8 + e]) + b[e], 0.f);
      dot[e & 3] = fmaf(x[e], w[e], dot[e & 3]);
    }
    store_chunk<false, F16>(dst_hi, dst_hi, row, (uint32_t)((chunk32 & 1) * 4 + g), x);
  }
}

// Accumulator columns [0, ncols) of this thread's row -> next layer's activation blocks.  The TMEM load of
// chunk c+1 is in flight while chunk c is processed (tcgen05.wait::ld waits for all outstanding loads).
template <bool X3, bool GLOBAL_BIAS, bool F16 = false>
__device__ __forceinline__ void epilogue_relu(uint32_t acc, int ncols, const float* gbias, uint32_t sbias,
                                              uint8_t* arena_hi, uint8_t* arena_lo, uint32_t row) {
  uint32_t v0[32], v1[32];
  const int nch = ncols >> 5;  // even
  tmem_ld32(acc, v0);
  for (int c = 0; c < nch; c += 2) {
    tmem_ld_wait();
    tmem_ld32(acc + (c + 1) * 32, v1);
    epilogue_chunk<X3, GLOBAL_BIAS, F16>(v0, c, gbias, sbias, arena_hi, arena_lo, row);
    tmem_ld_wait();
    if (c + 2 < nch) tmem_ld32(acc + (c + 2) * 32, v0);
    epilogue_chunk<X3, GLOBAL_BIAS, F16>(v1, c + 1, gbias, sbias, arena_hi, arena_lo, row);
  }
}


}  // namespace tc
}  // namespace dfn
