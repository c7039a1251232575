// tcgen05 path of network_query_fn: pts = o + d*z -> positional encoding -> 8x256 skip-MLP ->
// raw (rgb, sigma), one persistent warp-specialised kernel (sm_100a).
//
// Reference arithmetic: HELP:21-52 (Embedder), HELP:275-299 (FaceNeRF.forward), HELP:372-396
// (NeRF.forward); the per-frame latent columns and the per-ray view-direction columns are folded
// into fp32 biases (exact), so the tensor cores see K = 64 (PE) and K = 256 (hidden) only.
//
// Per CTA (one per SM, 227 KB shared memory, all 512 TMEM columns):
//   warp 0     weight producer: cp.async.bulk (TMA) of pre-swizzled [<=128 x 64] bf16 weight
//              stages from L2 into a 4 x 16 KB ring, mbarrier complete_tx
//   warp 1     MMA issuer: one thread issues tcgen05.mma.kind::f16 (M=128, N<=128, K=16) with
//              A = activation K-block in shared memory, B = ring stage, D = TMEM accumulator
//   warp 2     TMEM allocator
//   warps 4-7  epilogue of tile slot 0, warps 8-11 of slot 1: positional encoding of the tile's
//              128 points into the PE K-block; per layer tcgen05.ld of the accumulator,
//              + bias, ReLU, bf16 (hi[/lo]) repack into the swizzled K-major activation blocks
//              that feed the next layer; last layer writes raw[128,4] to HBM.
// bf16 mode keeps two 128-point tiles in flight (ping-pong: the MMAs of one overlap the epilogue
// of the other); bf16x3 mode (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo) keeps one tile because the
// hi and lo activation planes fill the activation arena.
#include <math.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "model.h"

namespace dfn {
namespace tc {

static constexpr int TILE_M = 128;
static constexpr int KB_BYTES = TILE_M * 128;      // one activation K-block: 128 rows x 64 bf16
static constexpr int STAGE_BYTES = 128 * 128;      // one weight stage: <=128 rows x 64 bf16
static constexpr int N_STAGES = 4;
static constexpr int ARENA_BLOCKS = 2 * TC_KB_PER_TILE;
static constexpr int SMEM_RING = ARENA_BLOCKS * KB_BYTES;
static constexpr int SMEM_BIAS = SMEM_RING + N_STAGES * STAGE_BYTES;
static constexpr int SMEM_BAR = SMEM_BIAS + 2 * TC_BIAS_STRIDE * 4;
static constexpr int SMEM_TOTAL = SMEM_BAR + 128;
static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget");

struct Params {
  const uint8_t* w_hi;
  const uint8_t* w_lo;
  const float* bias;       // [n_layers][256], latent already folded
  const float* view_bias;  // [R][W/2]
  const float* rays_o;
  const float* rays_d;
  const float* z_vals;
  float* raw;
  int64_t n_points;
  int S;
  int n_tiles;
  int n_layers;
  int multires;
  int view_w;  // W/2
  TcLayer layers[TC_MAX_LAYERS];
};

// ------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded spin: a protocol bug must fault the launch (after ~2 s), never hang the GPU.
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FFu) == 0) {
      const uint64_t t = globaltimer_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 2000000000ull) __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tma_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// K-major, 128-byte swizzle, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor, version 1).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16: D=f32, A=B=bf16, both K-major, M=128 (cute::UMMA::InstrDescriptor).
__device__ __forceinline__ uint32_t make_idesc(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_lo_f(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16_hi_f(uint32_t p) { return __uint_as_float(p & 0xFFFF0000u); }

// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a swizzled K-block
__device__ __forceinline__ uint32_t swz(uint32_t row, uint32_t chunk) {
  return row * 128u + ((chunk ^ (row & 7u)) << 4);
}

// Writes 8 consecutive columns (one 16-byte chunk) of this thread's row.
template <bool X3>
__device__ __forceinline__ void store_chunk(uint8_t* blk_hi, uint8_t* blk_lo, uint32_t row, uint32_t chunk,
                                            const float (&v)[8]) {
  uint4 h;
  h.x = pack_bf16(v[0], v[1]);
  h.y = pack_bf16(v[2], v[3]);
  h.z = pack_bf16(v[4], v[5]);
  h.w = pack_bf16(v[6], v[7]);
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

// ------------------------------------------------------------------------------------ kernel
// X3 = false: bf16, two tile slots (384 threads).  X3 = true: split-bf16, one slot (256 threads).
template <bool X3>
__global__ void __launch_bounds__(X3 ? 256 : 384, 1) mlp_tc_kernel(const __grid_constant__ Params P) {
  constexpr int NSLOT = X3 ? 1 : 2;
  constexpr int NPART = X3 ? 2 : 1;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();  // swizzled descriptors need a 1024-byte aligned arena

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t bar_full = sbase + SMEM_BAR;          // [4]
  const uint32_t bar_empty = sbase + SMEM_BAR + 32;    // [4]
  const uint32_t bar_acc = sbase + SMEM_BAR + 64;      // [2] accumulator of slot s complete
  const uint32_t bar_aready = sbase + SMEM_BAR + 80;   // [2] activations of slot s written
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SMEM_BAR + 96);

  if (threadIdx.x == 0) {
    for (int i = 0; i < N_STAGES; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_acc + 8 * s, 1);
      mbar_init(bar_aready + 8 * s, TILE_M);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int G = gridDim.x;
  const int n_local = (int)blockIdx.x < P.n_tiles ? (P.n_tiles - (int)blockIdx.x + G - 1) / G : 0;
  const int n_iter = (n_local + NSLOT - 1) / NSLOT;

  if (warp == 0) {
    // ============================== weight producer (TMA) ===============================
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int j = 0; j < n_iter; ++j) {
        for (int l = 0; l < P.n_layers; ++l) {
          const TcLayer& L = P.layers[l];
          for (int s = 0; s < NSLOT; ++s) {
            if (j * NSLOT + s >= n_local) continue;
            uint32_t off = L.woff;
            for (int kbi = 0; kbi < L.nkb; ++kbi) {
              for (int c0 = 0; c0 < L.n; c0 += 128) {
                const uint32_t bytes = (uint32_t)min(128, (int)L.n - c0) * 128u;
                for (int part = 0; part < NPART; ++part) {
                  const uint32_t slot = cnt % N_STAGES, par = (cnt / N_STAGES) & 1u;
                  mbar_wait(bar_empty + 8 * slot, par ^ 1u);
                  mbar_expect_tx(bar_full + 8 * slot, bytes);
                  tma_bulk_load(sbase + SMEM_RING + slot * STAGE_BYTES, (part == 0 ? P.w_hi : P.w_lo) + off, bytes,
                                bar_full + 8 * slot);
                  ++cnt;
                }
                off += bytes;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================= MMA issuer =========================================
    if (lane == 0) {
      uint32_t cnt = 0;
      uint32_t apar[2] = {0u, 0u};
      for (int j = 0; j < n_iter; ++j) {
        for (int l = 0; l < P.n_layers; ++l) {
          const TcLayer& L = P.layers[l];
          for (int s = 0; s < NSLOT; ++s) {
            if (j * NSLOT + s >= n_local) continue;
            mbar_wait(bar_aready + 8 * s, apar[s]);
            apar[s] ^= 1u;
            tcgen05_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)s * 256u;
            for (int kbi = 0; kbi < L.nkb; ++kbi) {
              const uint32_t a_hi = sbase + (uint32_t)(s * TC_KB_PER_TILE + L.kb[kbi]) * KB_BYTES;
              const uint64_t adesc_hi = make_smem_desc(a_hi);
              const uint64_t adesc_lo = make_smem_desc(a_hi + TC_KB_PER_TILE * KB_BYTES);
              for (int c0 = 0; c0 < L.n; c0 += 128) {
                const uint32_t nn = (uint32_t)min(128, (int)L.n - c0);
                const uint32_t idesc = make_idesc(nn);
                for (int part = 0; part < NPART; ++part) {
                  const uint32_t slot = cnt % N_STAGES, par = (cnt / N_STAGES) & 1u;
                  mbar_wait(bar_full + 8 * slot, par);
                  tcgen05_fence_after();
                  const uint64_t bdesc = make_smem_desc(sbase + SMEM_RING + slot * STAGE_BYTES);
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) {
                    umma_bf16(acc + c0, adesc_hi + 2 * ks, bdesc + 2 * ks, idesc,
                              (kbi | part | ks) != 0 ? 1u : 0u);
                  }
                  if (X3 && part == 0) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_bf16(acc + c0, adesc_lo + 2 * ks, bdesc + 2 * ks, idesc, 1u);
                  }
                  umma_commit(bar_empty + 8 * slot);
                  ++cnt;
                }
              }
            }
            umma_commit(bar_acc + 8 * s);
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue warps =======================================
    const int s = (warp - 4) >> 2;                    // tile slot
    const uint32_t row = (uint32_t)((warp & 3) * 32 + lane);
    const int tid_s = (int)row;                       // 0..127 within the slot
    uint8_t* arena_hi = smem + (size_t)s * TC_KB_PER_TILE * KB_BYTES;
    uint8_t* arena_lo = arena_hi + (size_t)TC_KB_PER_TILE * KB_BYTES;
    float* bias_s = reinterpret_cast<float*>(smem + SMEM_BIAS) + s * TC_BIAS_STRIDE;
    const uint32_t acc = tmem_base + (uint32_t)s * 256u + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc_par = 0u;

    for (int i = s; i < n_local; i += NSLOT) {
      const int tile = (int)blockIdx.x + i * G;
      int64_t pt = (int64_t)tile * TILE_M + row;
      const bool valid = pt < P.n_points;
      if (!valid) pt = P.n_points - 1;
      const int64_t ray = pt / P.S;

      // stage the first layer's bias while the encoding is computed
      {
        const float2 b2 = reinterpret_cast<const float2*>(P.bias)[tid_s];
        reinterpret_cast<float2*>(bias_s)[tid_s] = b2;
      }
      // ---- positional encoding of x = o + d*z into the PE K-block (HELP:42-52) ----
      {
        const float z = P.z_vals[pt];
        float x[3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
          x[c] = __fadd_rn(P.rays_o[ray * 3 + c], __fmul_rn(P.rays_d[ray * 3 + c], z));
        uint8_t* pe_hi = arena_hi + TC_KB_PE * KB_BYTES;
        uint8_t* pe_lo = arena_lo + TC_KB_PE * KB_BYTES;
        auto put = [&](int e, float v) {
          const uint32_t o = swz(row, (uint32_t)e >> 3) + ((uint32_t)e & 7u) * 2u;
          const __nv_bfloat16 h = __float2bfloat16_rn(v);
          *reinterpret_cast<__nv_bfloat16*>(pe_hi + o) = h;
          if (X3) *reinterpret_cast<__nv_bfloat16*>(pe_lo + o) = __float2bfloat16_rn(v - __bfloat162float(h));
        };
        put(0, x[0]);
        put(1, x[1]);
        put(2, x[2]);
        float f = 1.0f;
        for (int k = 0; k < P.multires; ++k) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float sv, cv;
            sincosf(__fmul_rn(x[c], f), &sv, &cv);
            put(3 + 6 * k + c, sv);
            put(6 + 6 * k + c, cv);
          }
          f *= 2.0f;
        }
        for (int e = 3 + 6 * P.multires; e < 64; ++e) put(e, 0.0f);
      }
      fence_proxy_async();
      mbar_arrive(bar_aready + 8 * s);
      named_bar_sync(1 + s, TILE_M);  // bias_s visible to the slot's four warps

      float alpha = 0.f;
      for (int l = 0; l < P.n_layers; ++l) {
        const TcLayer& L = P.layers[l];
        // prefetch next layer's shared bias (latency hidden behind the accumulator wait)
        float2 nb = make_float2(0.f, 0.f);
        if (l + 1 < P.n_layers) nb = reinterpret_cast<const float2*>(P.bias + (l + 1) * TC_BIAS_STRIDE)[tid_s];

        mbar_wait(bar_acc + 8 * s, acc_par);
        acc_par ^= 1u;
        tcgen05_fence_after();

        if (L.epi == TC_EPI_RGB) {
          uint32_t v[16];
          tmem_ld16(acc, v);
          tmem_ld_wait();
          if (valid) {
            float4 o;
            o.x = __uint_as_float(v[0]) + bias_s[0];
            o.y = __uint_as_float(v[1]) + bias_s[1];
            o.z = __uint_as_float(v[2]) + bias_s[2];
            o.w = alpha;
            reinterpret_cast<float4*>(P.raw)[pt] = o;
          }
          tcgen05_fence_before();
        } else {
          const bool per_ray = L.epi == TC_EPI_VIEW0;
          const int n_relu = per_ray ? P.view_w : (int)L.n;
          const float* rb = P.view_bias + ray * P.view_w;  // per-ray bias (TC_EPI_VIEW0 only)
          for (int blk = 0; blk < n_relu / 64; ++blk) {
            uint32_t va[32], vb[32];
            tmem_ld32(acc + blk * 64, va);
            tmem_ld32(acc + blk * 64 + 32, vb);
            uint8_t* dst_hi = arena_hi + (size_t)blk * KB_BYTES;
            uint8_t* dst_lo = arena_lo + (size_t)blk * KB_BYTES;
            const float* bsrc = per_ray ? rb + blk * 64 : bias_s + blk * 64;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              float4 pb[8];
              if (per_ray) {
#pragma unroll
                for (int q = 0; q < 8; ++q) pb[q] = __ldg(reinterpret_cast<const float4*>(bsrc + half * 32) + q);
              }
              if (half == 0) tmem_ld_wait();
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float4 b0, b1;
                if (per_ray) {
                  b0 = pb[2 * g];
                  b1 = pb[2 * g + 1];
                } else {
                  b0 = reinterpret_cast<const float4*>(bsrc + half * 32)[2 * g];
                  b1 = reinterpret_cast<const float4*>(bsrc + half * 32)[2 * g + 1];
                }
                const uint32_t* vv = half == 0 ? va : vb;
                float o[8];
                o[0] = fmaxf(__uint_as_float(vv[g * 8 + 0]) + b0.x, 0.f);
                o[1] = fmaxf(__uint_as_float(vv[g * 8 + 1]) + b0.y, 0.f);
                o[2] = fmaxf(__uint_as_float(vv[g * 8 + 2]) + b0.z, 0.f);
                o[3] = fmaxf(__uint_as_float(vv[g * 8 + 3]) + b0.w, 0.f);
                o[4] = fmaxf(__uint_as_float(vv[g * 8 + 4]) + b1.x, 0.f);
                o[5] = fmaxf(__uint_as_float(vv[g * 8 + 5]) + b1.y, 0.f);
                o[6] = fmaxf(__uint_as_float(vv[g * 8 + 6]) + b1.z, 0.f);
                o[7] = fmaxf(__uint_as_float(vv[g * 8 + 7]) + b1.w, 0.f);
                store_chunk<X3>(dst_hi, dst_lo, row, (uint32_t)(half * 4 + g), o);
              }
            }
          }
          if (L.epi == TC_EPI_VIEW0) {
            uint32_t v[16];
            tmem_ld16(acc + P.view_w, v);
            tmem_ld_wait();
            alpha = __uint_as_float(v[0]) + bias_s[P.view_w];
          }
          tcgen05_fence_before();
          fence_proxy_async();
          mbar_arrive(bar_aready + 8 * s);
        }
        // swap in the next layer's bias
        if (l + 1 < P.n_layers) {
          named_bar_sync(1 + s, TILE_M);
          reinterpret_cast<float2*>(bias_s)[tid_s] = nb;
          named_bar_sync(1 + s, TILE_M);
        }
      }
      named_bar_sync(1 + s, TILE_M);  // everyone done with bias_s before the next tile restages it
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------- per-call prep kernels
// bias_out[l][n] = bias[l][n] (+ fold_w[which][n][:] . latent for the two latent-consuming layers):
// the per-frame audio/expression latent (HELP:276, columns input_ch..input_ch+dim_aud) is constant
// over the frame, so its contribution is a bias.  fp32, sequential over j.
__global__ void fold_latent_kernel(int n_layers, int W, int dim_aud, const float* __restrict__ bias,
                                   const float* __restrict__ fold_w, const float* __restrict__ latent,
                                   int fold0, int fold1, float* __restrict__ bias_out) {
  const int l = blockIdx.x, n = threadIdx.x;
  if (l >= n_layers || n >= TC_BIAS_STRIDE) return;
  float v = bias[l * TC_BIAS_STRIDE + n];
  const int which = l == fold0 ? 0 : (l == fold1 ? 1 : -1);
  if (which >= 0 && latent != nullptr && n < W) {
    const float* w = fold_w + ((size_t)which * W + n) * dim_aud;
    float acc = 0.f;
    for (int j = 0; j < dim_aud; ++j) acc = fmaf(w[j], latent[j], acc);
    v += acc;
  }
  bias_out[l * TC_BIAS_STRIDE + n] = v;
}

// view_bias[r][n] = b[n] + sum_j Wv[n][j] * PE(viewdir_r)[j]: the view-direction columns of
// views_linears.0 (HELP:288-292) are constant along a ray.  One block walks rays; thread n owns
// output n.  PE layout as HELP:42-52 with multires_views frequencies.
__global__ void view_bias_kernel(int64_t R, int Wh, int ncol, int L, const float* __restrict__ viewdirs,
                                 const float* __restrict__ vw, const float* __restrict__ vb,
                                 float* __restrict__ out) {
  extern __shared__ float sm[];
  float* w_s = sm;                 // [Wh][ncol]
  float* pe = sm + Wh * ncol;      // [ncol]
  for (int i = threadIdx.x; i < Wh * ncol; i += blockDim.x) w_s[i] = vw[i];
  const float bn = threadIdx.x < Wh ? vb[threadIdx.x] : 0.f;
  for (int64_t r = blockIdx.x; r < R; r += gridDim.x) {
    __syncthreads();
    if ((int)threadIdx.x < ncol) {
      const int j = threadIdx.x;
      float v;
      if (j < 3) {
        v = viewdirs[r * 3 + j];
      } else {
        const int k = (j - 3) / 6, q = (j - 3) % 6;
        const float a = __fmul_rn(viewdirs[r * 3 + (q % 3)], pow2i(k));
        v = q < 3 ? sinf(a) : cosf(a);
      }
      pe[j] = v;
    }
    __syncthreads();
    if ((int)threadIdx.x < Wh) {
      float acc = 0.f;
      const float* w = w_s + threadIdx.x * ncol;
      for (int j = 0; j < ncol; ++j) acc = fmaf(w[j], pe[j], acc);
      out[r * Wh + threadIdx.x] = bn + acc;
    }
  }
}

// ------------------------------------------------------------------------ host: weight packing
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
  std::vector<uint8_t> hi, lo;
  // Appends the stages of one layer: for each K-block, for each chunk of <=128 output rows, a
  // [rows x 64] bf16 image in the swizzled K-major layout.  wfun(n, kbi, k) returns W[n][column
  // of K-block kbi, position k] or 0.
  template <class F>
  uint32_t add_layer(int n_out, int nkb, F wfun) {
    const uint32_t start = (uint32_t)hi.size();
    for (int kbi = 0; kbi < nkb; ++kbi) {
      for (int c0 = 0; c0 < n_out; c0 += 128) {
        const int rows = n_out - c0 < 128 ? n_out - c0 : 128;
        const size_t base = hi.size();
        hi.resize(base + (size_t)rows * 128, 0);
        lo.resize(base + (size_t)rows * 128, 0);
        for (int r = 0; r < rows; ++r) {
          for (int k = 0; k < 64; ++k) {
            const float w = wfun(c0 + r, kbi, k);
            const uint16_t h = f2bf(w);
            const uint16_t l = f2bf(w - bf2f(h));
            const size_t o = base + (size_t)r * 128 + ((((size_t)k >> 3) ^ ((size_t)r & 7)) << 4) + ((size_t)k & 7) * 2;
            memcpy(&hi[o], &h, 2);
            memcpy(&lo[o], &l, 2);
          }
        }
      }
    }
    return start;
  }
};

}  // namespace tc

void tc_free_model(dfn_model* m) {
  cudaFree(m->tc_hi);
  cudaFree(m->tc_lo);
  cudaFree(m->tc_bias);
  cudaFree(m->tc_fold_w);
  cudaFree(m->tc_view_w);
  cudaFree(m->tc_view_b);
  m->tc_hi = m->tc_lo = nullptr;
  m->tc_bias = m->tc_fold_w = m->tc_view_w = m->tc_view_b = nullptr;
}

// Builds the layer program and the packed blobs from the reference tensors (order of dfn.h).
int tc_pack_model(dfn_model* m, const float* const* t, cudaStream_t st) {
  const dfn_model_desc& d = m->desc;
  if (d.W != 256 || d.input_ch > 63 || d.input_ch != 3 + 6 * d.multires || d.D < 2 || d.D + m->n_views + 1 > TC_MAX_LAYERS ||
      d.skip < 0 || d.skip >= d.D - 1) {
    set_error("tcgen05 path supports W=256, input_ch=3+6*multires<=63, one skip before the last trunk layer");
    return DFN_E_UNSUPPORTED;
  }
  const int W = d.W, Wh = W / 2, n_pts = d.input_ch + d.dim_aud;
  auto Wt = [&](int i) { return t[2 * i]; };
  auto Bt = [&](int i) { return t[2 * i + 1]; };
  const int i_views0 = d.D, i_feature = d.D + m->n_views, i_alpha = i_feature + 1, i_rgb = i_feature + 2;

  tc::Packer pk;
  TcProgram& pg = m->prog;
  pg = TcProgram();
  std::vector<float> bias((size_t)TC_MAX_LAYERS * TC_BIAS_STRIDE, 0.f);
  std::vector<float> fold_w((size_t)2 * W * (d.dim_aud > 0 ? d.dim_aud : 1), 0.f);
  int nl = 0, nfold = 0;

  // trunk (HELP:277-283): layer 0 reads the PE block; layer skip+1 reads [PE | h]; others read h.
  for (int i = 0; i < d.D; ++i) {
    TcLayer L;
    memset(&L, 0, sizeof(L));
    L.n = (uint16_t)W;
    L.epi = TC_EPI_RELU;
    const float* w = Wt(i);
    const bool has_in = (i == 0) || (i - 1 == d.skip);
    const int ld = (i == 0) ? n_pts : (has_in ? n_pts + W : W);
    int nkb = 0;
    if (has_in) L.kb[nkb++] = TC_KB_PE;
    if (i != 0)
      for (int q = 0; q < 4; ++q) L.kb[nkb++] = (uint8_t)(TC_KB_H0 + q);
    L.nkb = (uint8_t)nkb;
    const int hoff = has_in ? n_pts : 0;  // first weight column multiplying h
    L.woff = pk.add_layer(W, nkb, [&](int n, int kbi, int k) -> float {
      const int blk = L.kb[kbi];
      if (blk == TC_KB_PE) return k < d.input_ch ? w[(size_t)n * ld + k] : 0.f;
      return w[(size_t)n * ld + hoff + (blk - TC_KB_H0) * 64 + k];
    });
    for (int n = 0; n < W; ++n) bias[(size_t)nl * TC_BIAS_STRIDE + n] = Bt(i)[n];
    if (has_in && d.dim_aud > 0) {
      if (nfold >= 2) {
        set_error("more than two latent-consuming layers");
        return DFN_E_UNSUPPORTED;
      }
      for (int n = 0; n < W; ++n)
        for (int j = 0; j < d.dim_aud; ++j)
          fold_w[((size_t)nfold * W + n) * d.dim_aud + j] = w[(size_t)n * ld + d.input_ch + j];
      pg.fold_layer[nfold++] = nl;
    }
    pg.layers[nl++] = L;
  }
  // views_linears.0 (+ alpha_linear as output row Wh).  NeRF applies feature_linear first
  // (HELP:384) with no activation in between, so it is composed into views_linears.0 here
  // (fp64 on the host): Wv[:, :W] @ Wf and bv + Wv[:, :W] @ bf.
  {
    const float* wv = Wt(i_views0);
    const int ldv = W + d.input_ch_views;
    std::vector<float> wc((size_t)Wh * W);
    std::vector<float> bc(Wh);
    if (d.kind == DFN_MODEL_NERF) {
      const float* wf = Wt(i_feature);
      const float* bf = Bt(i_feature);
      for (int n = 0; n < Wh; ++n) {
        for (int k = 0; k < W; ++k) {
          double a = 0.0;
          for (int q = 0; q < W; ++q) a += (double)wv[(size_t)n * ldv + q] * (double)wf[(size_t)q * W + k];
          wc[(size_t)n * W + k] = (float)a;
        }
        double b = Bt(i_views0)[n];
        for (int q = 0; q < W; ++q) b += (double)wv[(size_t)n * ldv + q] * (double)bf[q];
        bc[n] = (float)b;
      }
    } else {
      for (int n = 0; n < Wh; ++n) {
        for (int k = 0; k < W; ++k) wc[(size_t)n * W + k] = wv[(size_t)n * ldv + k];
        bc[n] = Bt(i_views0)[n];
      }
    }
    TcLayer L;
    memset(&L, 0, sizeof(L));
    L.n = (uint16_t)(Wh + 16);
    L.epi = TC_EPI_VIEW0;
    L.nkb = 4;
    for (int q = 0; q < 4; ++q) L.kb[q] = (uint8_t)(TC_KB_H0 + q);
    const float* wa = Wt(i_alpha);
    L.woff = pk.add_layer(Wh + 16, 4, [&](int n, int kbi, int k) -> float {
      const int col = kbi * 64 + k;
      if (n < Wh) return wc[(size_t)n * W + col];
      if (n == Wh) return wa[col];
      return 0.f;
    });
    bias[(size_t)nl * TC_BIAS_STRIDE + Wh] = Bt(i_alpha)[0];
    pg.layers[nl++] = L;
    // view-direction columns + composed bias for view_bias_kernel
    std::vector<float> vw((size_t)Wh * d.input_ch_views);
    for (int n = 0; n < Wh; ++n)
      for (int j = 0; j < d.input_ch_views; ++j) vw[(size_t)n * d.input_ch_views + j] = wv[(size_t)n * ldv + W + j];
    DFN_CUDA(cudaMalloc(&m->tc_view_w, vw.size() * 4));
    DFN_CUDA(cudaMalloc(&m->tc_view_b, bc.size() * 4));
    DFN_CUDA(cudaMemcpyAsync(m->tc_view_w, vw.data(), vw.size() * 4, cudaMemcpyHostToDevice, st));
    DFN_CUDA(cudaMemcpyAsync(m->tc_view_b, bc.data(), bc.size() * 4, cudaMemcpyHostToDevice, st));
    DFN_CUDA(cudaStreamSynchronize(st));
  }
  for (int i = 1; i < m->n_views; ++i) {
    TcLayer L;
    memset(&L, 0, sizeof(L));
    L.n = (uint16_t)Wh;
    L.epi = TC_EPI_RELU;
    L.nkb = 2;
    L.kb[0] = TC_KB_H0;
    L.kb[1] = TC_KB_H0 + 1;
    const float* w = Wt(i_views0 + i);
    L.woff = pk.add_layer(Wh, 2, [&](int n, int kbi, int k) -> float { return w[(size_t)n * Wh + kbi * 64 + k]; });
    for (int n = 0; n < Wh; ++n) bias[(size_t)nl * TC_BIAS_STRIDE + n] = Bt(i_views0 + i)[n];
    pg.layers[nl++] = L;
  }
  {
    TcLayer L;
    memset(&L, 0, sizeof(L));
    L.n = 16;
    L.epi = TC_EPI_RGB;
    L.nkb = 2;
    L.kb[0] = TC_KB_H0;
    L.kb[1] = TC_KB_H0 + 1;
    const float* w = Wt(i_rgb);
    L.woff = pk.add_layer(16, 2, [&](int n, int kbi, int k) -> float { return n < 3 ? w[(size_t)n * Wh + kbi * 64 + k] : 0.f; });
    for (int n = 0; n < 3; ++n) bias[(size_t)nl * TC_BIAS_STRIDE + n] = Bt(i_rgb)[n];
    pg.layers[nl++] = L;
  }
  pg.n_layers = nl;

  m->tc_blob_bytes = (int64_t)pk.hi.size();
  DFN_CUDA(cudaMalloc(&m->tc_hi, pk.hi.size()));
  DFN_CUDA(cudaMalloc(&m->tc_lo, pk.lo.size()));
  DFN_CUDA(cudaMalloc(&m->tc_bias, bias.size() * 4));
  DFN_CUDA(cudaMalloc(&m->tc_fold_w, fold_w.size() * 4));
  DFN_CUDA(cudaMemcpyAsync(m->tc_hi, pk.hi.data(), pk.hi.size(), cudaMemcpyHostToDevice, st));
  DFN_CUDA(cudaMemcpyAsync(m->tc_lo, pk.lo.data(), pk.lo.size(), cudaMemcpyHostToDevice, st));
  DFN_CUDA(cudaMemcpyAsync(m->tc_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice, st));
  DFN_CUDA(cudaMemcpyAsync(m->tc_fold_w, fold_w.data(), fold_w.size() * 4, cudaMemcpyHostToDevice, st));
  DFN_CUDA(cudaStreamSynchronize(st));
  return 0;
}

static inline int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }

int64_t tc_query_workspace_bytes(const dfn_model* m, int64_t R, int S) {
  (void)S;
  return align256((int64_t)TC_MAX_LAYERS * TC_BIAS_STRIDE * 4) + align256(R * (m->desc.W / 2) * 4);
}

int tc_query_points(const dfn_model* m, int64_t R, int S, const float* rays_o, const float* rays_d,
                    const float* viewdirs, const float* z_vals, const float* latent, float* raw,
                    int precision, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  const dfn_model_desc& d = m->desc;
  if (workspace_bytes < tc_query_workspace_bytes(m, R, S)) {
    set_error("dfn_query_points: workspace %lld < %lld bytes", (long long)workspace_bytes,
              (long long)tc_query_workspace_bytes(m, R, S));
    return DFN_E_WORKSPACE;
  }
  if (d.dim_aud > 0 && latent == nullptr) {
    set_error("dfn_query_points: FaceNeRF needs a latent");
    return DFN_E_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(raw) & 15) != 0) {
    set_error("dfn_query_points: raw must be 16-byte aligned");
    return DFN_E_ARG;
  }
  float* bias_ws = reinterpret_cast<float*>(workspace);
  float* vbias_ws = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + align256((int64_t)TC_MAX_LAYERS * TC_BIAS_STRIDE * 4));
  const int Wh = d.W / 2;

  tc::fold_latent_kernel<<<m->prog.n_layers, TC_BIAS_STRIDE, 0, st>>>(
      m->prog.n_layers, d.W, d.dim_aud, m->tc_bias, m->tc_fold_w, latent, m->prog.fold_layer[0], m->prog.fold_layer[1], bias_ws);
  DFN_LAUNCH_CHECK();
  {
    int64_t blocks = R < (int64_t)num_sms() * 8 ? R : (int64_t)num_sms() * 8;
    size_t sm = ((size_t)Wh * d.input_ch_views + d.input_ch_views) * sizeof(float);
    tc::view_bias_kernel<<<(int)blocks, 128, sm, st>>>(R, Wh, d.input_ch_views, d.multires_views, viewdirs, m->tc_view_w,
                                                        m->tc_view_b, vbias_ws);
    DFN_LAUNCH_CHECK();
  }

  tc::Params P;
  memset(&P, 0, sizeof(P));
  P.w_hi = m->tc_hi;
  P.w_lo = m->tc_lo;
  P.bias = bias_ws;
  P.view_bias = vbias_ws;
  P.rays_o = rays_o;
  P.rays_d = rays_d;
  P.z_vals = z_vals;
  P.raw = raw;
  P.n_points = R * S;
  P.S = S;
  P.n_tiles = (int)((P.n_points + tc::TILE_M - 1) / tc::TILE_M);
  P.n_layers = m->prog.n_layers;
  P.multires = d.multires;
  P.view_w = Wh;
  for (int i = 0; i < m->prog.n_layers; ++i) P.layers[i] = m->prog.layers[i];

  const int grid = P.n_tiles < num_sms() ? P.n_tiles : num_sms();
  // algorithmic MACs per point with the latent / view-direction columns folded into biases
  double macs_pt = 0.0;
  for (int i = 0; i < d.D; ++i) {
    const bool has_in = (i == 0) || (i - 1 == d.skip);
    macs_pt += (double)d.W * ((i == 0 ? 0 : d.W) + (has_in ? d.input_ch : 0));
  }
  macs_pt += (double)d.W * Wh + d.W;                       // views_linears.0 (+ composed feature) and alpha
  macs_pt += (double)(m->n_views - 1) * Wh * Wh + 3.0 * Wh;  // remaining view layers and rgb
  const bool prof = profile_begin(st, macs_pt * (double)P.n_points);
  if (precision == DFN_PREC_BF16) {
    static bool attr_done = false;
    if (!attr_done) {
      DFN_CUDA(cudaFuncSetAttribute(tc::mlp_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_TOTAL));
      attr_done = true;
    }
    tc::mlp_tc_kernel<false><<<grid, 384, tc::SMEM_TOTAL, st>>>(P);
  } else if (precision == DFN_PREC_BF16X3) {
    static bool attr_done = false;
    if (!attr_done) {
      DFN_CUDA(cudaFuncSetAttribute(tc::mlp_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_TOTAL));
      attr_done = true;
    }
    tc::mlp_tc_kernel<true><<<grid, 256, tc::SMEM_TOTAL, st>>>(P);
  } else {
    set_error("tc_query_points: precision %d is not a tensor-core mode", precision);
    return DFN_E_ARG;
  }
  if (prof) profile_end(st);
  DFN_LAUNCH_CHECK();
  return 0;
}

}  // namespace dfn
