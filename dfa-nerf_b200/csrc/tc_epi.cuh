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
    l.x = pack_lo<F16>(v[0], v[1], h.x);
    l.y = pack_lo<F16>(v[2], v[3], h.y);
    l.z = pack_lo<F16>(v[4], v[5], h.z);
    l.w = pack_lo<F16>(v[6], v[7], h.w);
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
      store_chunk<true, F16>(dst_hi, dst_lo, row, c16, o);
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


// ---- column-distributed epilogue (single-pass precisions) ----
// tcgen05.ld.16x256b hands thread t of a warp rows t/4 and t/4 + 8 of a 16-lane group and columns 8g + 2(t%4) + {0, 1} of every
// 8-column group g: a thread then needs only 16 of a K-block's 64 biases (eight 8-byte loads that are conflict-free across the warp: one
// wavefront each) instead of all of them through warp-broadcast LDS.128 (four wavefronts per 16 bytes -- 1,024 of the ~2,840 shared-memory
// wavefronts of a 256-wide tile-layer, and more than half of this epilogue's time when it runs alone: profiles/microbench/epilogue.cu).
// The packed pair goes out as one 4-byte store; the eight rows of a store instruction land in eight different 16-byte chunks of the
// 128-byte swizzle, i.e. 32 distinct banks.  Results are bit-identical to epilogue_relu (same fp32 add, same rounding).
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// one 16-row x 64-column piece: v as loaded by tmem_ld_16x256b_x8, b = this thread's 16 biases of the K-block
template <bool F16>
__device__ __forceinline__ void epilogue_piece_cd(const uint32_t (&v)[32], const float (&b)[16], uint8_t* blk, uint32_t row_lo, uint32_t lane) {
  const uint32_t sub = (lane & 3u) * 4u;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    uint32_t p0, p1;
    if (F16) {
      p0 = add_relu_pack_f16(v[4 * g + 0], v[4 * g + 1], b[2 * g], b[2 * g + 1]);
      p1 = add_relu_pack_f16(v[4 * g + 2], v[4 * g + 3], b[2 * g], b[2 * g + 1]);
    } else {
      p0 = add_relu_pack(v[4 * g + 0], v[4 * g + 1], b[2 * g], b[2 * g + 1]);
      p1 = add_relu_pack(v[4 * g + 2], v[4 * g + 3], b[2 * g], b[2 * g + 1]);
    }
    *reinterpret_cast<uint32_t*>(blk + swz(row_lo, (uint32_t)g) + sub) = p0;
    *reinterpret_cast<uint32_t*>(blk + swz(row_lo + 8u, (uint32_t)g) + sub) = p1;
  }
}

// The same piece for the split modes: relu(v + b) in fp32 -> hi = rn16(x) [, lo = rn16(x - hi)] as 4-byte stores into the hi / lo planes.
// need_lo = false when every consumer of the block is a single-pass layer (DFN_PREC_FP16X3M).
template <bool F16>
__device__ __forceinline__ void epilogue_piece_cd_split(const uint32_t (&v)[32], const float (&b)[16], uint8_t* blk_hi, uint8_t* blk_lo,
                                                        bool need_lo, uint32_t row_lo, uint32_t lane) {
  const uint32_t sub = (lane & 3u) * 4u;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float x0 = fmaxf(__uint_as_float(v[4 * g + 2 * h]) + b[2 * g], 0.f);
      const float x1 = fmaxf(__uint_as_float(v[4 * g + 2 * h + 1]) + b[2 * g + 1], 0.f);
      const uint32_t hi = F16 ? pack_f16(x0, x1) : pack_bf16(x0, x1);
      const uint32_t off = swz(row_lo + 8u * (uint32_t)h, (uint32_t)g) + sub;
      *reinterpret_cast<uint32_t*>(blk_hi + off) = hi;
      if (need_lo) *reinterpret_cast<uint32_t*>(blk_lo + off) = pack_lo<F16>(x0, x1, hi);
    }
  }
}

// K-blocks [kb_begin, kb_end) of the accumulator (64 columns each) of this warp's 32 lanes -> the next layer's activation blocks.
// acc: TMEM address of the warp's lane quarter (lane field = 32 * (warp % 4)) and the slot's first column; row0 = 32 * (warp % 4).
// DB: the next 16-lane piece is in flight while the current one is processed (64 data registers); !DB: one piece at a time, for
// kernels with two warps per lane quarter and a 96-register budget (the two warps overlap each other).
template <bool F16, bool DB = true>
__device__ __forceinline__ void epilogue_relu_cd(uint32_t acc, int kb_begin, int kb_end, uint32_t sbias, uint8_t* arena_hi, uint32_t row0,
                                                 uint32_t lane) {
  const uint32_t r = row0 + (lane >> 2);
  if (DB) {
    uint32_t v0[32], v1[32];
    tmem_ld_16x256b_x8(acc + (uint32_t)kb_begin * 64u, v0);
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      float b[16];
      const uint32_t ba = sbias + (uint32_t)(kb * 64 + 2 * (int)(lane & 3u)) * 4u;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(b[2 * g]), "=f"(b[2 * g + 1]) : "r"(ba + (uint32_t)g * 32u));
      uint8_t* blk = arena_hi + (size_t)kb * KB_BYTES;
      tmem_ld_wait();
      tmem_ld_16x256b_x8(acc + (16u << 16) + (uint32_t)kb * 64u, v1);
      epilogue_piece_cd<F16>(v0, b, blk, r, lane);
      tmem_ld_wait();
      if (kb + 1 < kb_end) tmem_ld_16x256b_x8(acc + (uint32_t)(kb + 1) * 64u, v0);
      epilogue_piece_cd<F16>(v1, b, blk, r + 16u, lane);
    }
  } else {
    uint32_t v[32];
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      tmem_ld_16x256b_x8(acc + (uint32_t)kb * 64u, v);
      float b[16];
      const uint32_t ba = sbias + (uint32_t)(kb * 64 + 2 * (int)(lane & 3u)) * 4u;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(b[2 * g]), "=f"(b[2 * g + 1]) : "r"(ba + (uint32_t)g * 32u));
      uint8_t* blk = arena_hi + (size_t)kb * KB_BYTES;
      tmem_ld_wait();
      epilogue_piece_cd<F16>(v, b, blk, r, lane);
      tmem_ld_16x256b_x8(acc + (16u << 16) + (uint32_t)kb * 64u, v);
      tmem_ld_wait();
      epilogue_piece_cd<F16>(v, b, blk, r + 16u, lane);
    }
  }
}

// Column-distributed epilogue of a ReLU layer that also evaluates a folded density head (TC_F_DOT_SIGMA, Decoder single-pass programs):
// K-blocks [kb_begin, kb_end) as epilogue_relu_cd, plus dot = sum_c relu(v + b)[c] * w[c] over THESE columns on the fp32 activations,
// w = the head row as packed 16-bit pairs in shared memory (sdot).  Returns the partial dot product of row row0 + lane (the row this
// thread owns in the row-per-thread epilogues); the caller adds the other column parts and the head's bias.
template <bool F16>
__device__ __forceinline__ float epilogue_relu_cd_dotpart(uint32_t acc, int kb_begin, int kb_end, uint32_t sbias, uint32_t sdot, uint8_t* arena_hi,
                                                          uint32_t row0, uint32_t lane) {
  const uint32_t r = row0 + (lane >> 2);
  const uint32_t sub = (lane & 3u) * 4u;
  float d[4] = {0.f, 0.f, 0.f, 0.f};   // rows r, r + 8, r + 16, r + 24
  uint32_t v[32];
  for (int kb = kb_begin; kb < kb_end; ++kb) {
    uint8_t* blk = arena_hi + (size_t)kb * KB_BYTES;
    const uint32_t c0 = (uint32_t)(kb * 64) + 2u * (lane & 3u);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      tmem_ld_16x256b_x8(acc + ((uint32_t)(16 * half) << 16) + (uint32_t)kb * 64u, v);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float b0, b1, w0, w1;
        uint32_t wp;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(b0), "=f"(b1) : "r"(sbias + (c0 + (uint32_t)g * 8u) * 4u));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wp) : "r"(sdot + (c0 + (uint32_t)g * 8u) * 2u));
        unpack_h2<F16>(wp, w0, w1);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float x0 = fmaxf(__uint_as_float(v[4 * g + 2 * h]) + b0, 0.f);
          const float x1 = fmaxf(__uint_as_float(v[4 * g + 2 * h + 1]) + b1, 0.f);
          d[2 * half + h] = fmaf(x1, w1, fmaf(x0, w0, d[2 * half + h]));
          *reinterpret_cast<uint32_t*>(blk + swz(r + (uint32_t)(16 * half + 8 * h), (uint32_t)g) + sub) = F16 ? pack_f16(x0, x1) : pack_bf16(x0, x1);
        }
      }
    }
  }
  float out = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    float t = d[a];
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t = __shfl_sync(0xffffffffu, t, (int)((lane & 7u) << 2));   // row (lane % 8) + 8 a lives in quad lane % 8
    if ((lane >> 3) == (uint32_t)a) out = t;
  }
  return out;
}

// TC_EPI_STAGE on the column-distributed layout: one 64-column block of the accumulator (acc: lane quarter + first column of the block)
// + bias, NO activation -> 16-bit pairs into the block `dst` (shared memory) and, when gdst != nullptr, into a byte image of the block
// in global memory.
template <bool F16>
__device__ __forceinline__ void epilogue_stage_cd(uint32_t acc, uint32_t sbias, uint8_t* dst, uint8_t* gdst, uint32_t row0, uint32_t lane) {
  const uint32_t r = row0 + (lane >> 2);
  const uint32_t sub = (lane & 3u) * 4u;
  float b[16];
#pragma unroll
  for (int g = 0; g < 8; ++g)
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(b[2 * g]), "=f"(b[2 * g + 1]) : "r"(sbias + (uint32_t)(g * 8 + 2 * (int)(lane & 3u)) * 4u));
  uint32_t v[32];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    tmem_ld_16x256b_x8(acc + ((uint32_t)(16 * half) << 16), v);
    tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 8; ++g) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float x0 = __uint_as_float(v[4 * g + 2 * h]) + b[2 * g], x1 = __uint_as_float(v[4 * g + 2 * h + 1]) + b[2 * g + 1];
        const uint32_t p = F16 ? pack_f16(x0, x1) : pack_bf16(x0, x1);
        const uint32_t off = swz(r + (uint32_t)(16 * half + 8 * h), (uint32_t)g) + sub;
        *reinterpret_cast<uint32_t*>(dst + off) = p;
        if (gdst != nullptr) *reinterpret_cast<uint32_t*>(gdst + off) = p;
      }
    }
  }
}

// Row-per-thread epilogue over columns [c_begin, c_end) (multiples of 32) with 16-column TMEM loads, the next one in flight (32 data
// registers): the per-ray-bias layer (GLOBAL_BIAS: every row has its own bias row in global memory) in the 96-register kernels.
template <bool GLOBAL_BIAS, bool F16>
__device__ __forceinline__ void epilogue_relu_rows16(uint32_t acc, int c_begin, int c_end, const float* gbias, uint32_t sbias,
                                                     uint8_t* arena_hi, uint32_t row) {
  uint32_t v0[16], v1[16];
  auto piece = [&](const uint32_t (&v)[16], int c) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      float b[8];
      if (GLOBAL_BIAS) ldg_f32x8(gbias + c + g * 8, b);
      else lds_f32x8(sbias + (uint32_t)(c + g * 8) * 4u, b);
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
      const int cc = c + g * 8;
      *reinterpret_cast<uint4*>(arena_hi + (size_t)(cc >> 6) * KB_BYTES + swz(row, (uint32_t)((cc & 63) >> 3))) = h;
    }
  };
  tmem_ld16(acc + (uint32_t)c_begin, v0);
  for (int c = c_begin; c < c_end; c += 32) {
    tmem_ld_wait();
    tmem_ld16(acc + (uint32_t)(c + 16), v1);
    piece(v0, c);
    tmem_ld_wait();
    if (c + 32 < c_end) tmem_ld16(acc + (uint32_t)(c + 32), v0);
    piece(v1, c + 16);
  }
}

}  // namespace tc
}  // namespace dfn
