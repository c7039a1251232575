// Strided tcgen05 GEMM with in-kernel fp32 -> split-bf16 operand conversion: the workhorse of the training step
// (what `loss.backward()` at MAIN:923 and every nn.Linear forward of MAIN:855-880 spend their time in).
//
//   C[m, n] = act( sum_k A(m, k) * B(n, k) + bias[n] (+ addend[m, n]) ) (+ addend[m, n]) (+ C[m, n])
//   A(m, k) = A[m * a_ld_r + k * a_ld_k] (* mask'(A_mask[same index])),   B(n, k) = B[n * b_ld_r + k * b_ld_k]
//
// Both operands are addressed with a (row, k) stride pair, one of which is 1, so the three GEMMs of a layer need no
// transposed copies of anything:
//   forward      Y  = X W^T        A = X  (ld, 1)        B = W      (K, 1)
//   data grad    dX = dA W         A = dH (ld, 1) masked B = W^T    (1, K)      dA = dH * act'(Y), formed by the loader
//   weight grad  dW = dA^T X       A = dH^T (1, ld) masked, B = X^T (1, ld)     split over the batch, fp32 atomics
// Per CTA: one 128 x N (N <= 256) output tile; K walked in chunks of 64 through a two-stage shared-memory pipeline.
//   warps 0-3   A loaders (thread = tile row): 64 fp32 from global, optional activation-derivative mask, split into bf16
//               hi / lo = bf16(v - hi), 16-byte stores into the 128-byte-swizzled K-major operand block; then the epilogue
//   warps 4-11  B loaders (thread = output column n)
//   warp 12     TMEM allocation + MMA issue: tcgen05.mma.kind::f16, A_hi B_hi + A_lo B_hi + A_hi B_lo (bf16x3, fp32-level
//               products; DFN_PREC_BF16 issues the first term only), fp32 accumulator in tensor memory
// The kernel is bound by the loaders' conversion work, not by the tensor pipe (a training step is ~1 TFLOP: milliseconds).
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_epi.cuh"

namespace dfn {
namespace gemm {

using namespace dfn::tc;

static constexpr int BM = 128, BK = 64, BN_MAX = 256;
static constexpr int A_BYTES = BM * 128;        // one bf16 plane of the A tile
static constexpr int B_BYTES = BN_MAX * 128;
static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // hi + lo of both operands: 96 KB
static constexpr int N_STAGE = 2;
static constexpr int SMEM_BAR = N_STAGE * STAGE_BYTES;
static constexpr int SMEM_TOTAL = SMEM_BAR + 128;
static constexpr int A_THREADS = 128, B_THREADS = 256, LOADERS = A_THREADS + B_THREADS, THREADS = LOADERS + 32;
static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget");

struct Params {
  dfn_gemm_desc d;
  int n_pad;       // N rounded up to 16 (MMA granularity)
  int m_tiles;
  int chunks;      // ceil(K / 64)
  int per_split;   // chunks per K split
};

__device__ __forceinline__ float mask_factor(int mode, float y) {
  if (mode == DFN_MASK_RELU) return y > 0.f ? 1.f : 0.f;
  if (mode == DFN_MASK_LEAKY) return y > 0.f ? 1.f : 0.02f;
  return y * (1.f - y);   // DFN_MASK_SIGMOID: y is the sigmoid's output
}
// the same without control flow (selects on the uniform mode): the loaders issue all of a half-chunk's loads before the first use --
// with mask_factor's branches every mask load sat in its own basic block and paid its own memory latency (a 64-wide K chunk of the
// weight-gradient GEMM took 63,000 cycles)
__device__ __forceinline__ float mask_factor_sel(bool sig, float lo, float y) {
  const float step = y > 0.f ? 1.f : lo;
  return sig ? y * (1.f - y) : step;
}

// 32 consecutive k of one operand row -> registers (zero outside the matrix), optionally times act'(mask)
__device__ __forceinline__ void fetch32(const float* __restrict__ src, const float* __restrict__ msk, int mmode, int64_t row_off,
                                        int64_t ld_k, int k0, int K, bool row_ok, float (&v)[32]) {
  if (!row_ok) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
    return;
  }
  const float* p = src + row_off;
  const bool sig = mmode == DFN_MASK_SIGMOID;
  const float lo = mmode == DFN_MASK_LEAKY ? 0.02f : 0.f;
  if (ld_k == 1 && k0 + 32 <= K && ((reinterpret_cast<uintptr_t>(p + k0) & 15) == 0)) {
    const float4* q = reinterpret_cast<const float4*>(p + k0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = __ldg(q + j);
      v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
    if (msk) {
      const float4* qm = reinterpret_cast<const float4*>(msk + row_off + k0);
#pragma unroll
      for (int b = 0; b < 2; ++b) {       // sixteen mask values in flight at a time (register budget)
        float4 t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] = __ldg(qm + 4 * b + j);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int e = 16 * b + 4 * j;
          v[e] *= mask_factor_sel(sig, lo, t[j].x); v[e + 1] *= mask_factor_sel(sig, lo, t[j].y);
          v[e + 2] *= mask_factor_sel(sig, lo, t[j].z); v[e + 3] *= mask_factor_sel(sig, lo, t[j].w);
        }
      }
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int k = k0 + j;
    v[j] = k < K ? __ldg(p + (int64_t)k * ld_k) : 0.f;
  }
  if (msk) {
    const float* pm = msk + row_off;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      float y[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int k = k0 + 16 * b + j;
        y[j] = k < K ? __ldg(pm + (int64_t)k * ld_k) : 1.f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) v[16 * b + j] *= mask_factor_sel(sig, lo, y[j]);
    }
  }
}

// one tile row, one 64-wide K chunk: global fp32 -> bf16 hi [/ lo] planes in the swizzled K-major block
template <bool X3>
__device__ __forceinline__ void load_row(const float* __restrict__ src, const float* __restrict__ msk, int mmode, int64_t ld_r,
                                         int64_t ld_k, int64_t row_g, bool row_ok, int k0, int K, uint8_t* hi, uint8_t* lo,
                                         uint32_t row) {
  float v[32];
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    fetch32(src, msk, mmode, row_g * ld_r, ld_k, k0 + 32 * h, K, row_ok, v);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = v[ch * 8 + e];
      store_chunk<X3>(hi, lo, row, (uint32_t)(4 * h + ch), o);
    }
  }
}

// The same for an operand whose k stride is 1 (X, dH, W as [N, K]): 32 tile rows of this warp, one 64-wide K chunk.  Sixteen lanes
// cover a row's 256 bytes (a float4 each), so a load instruction reads two rows' full lines (4 wavefronts) where the row-per-thread
// mapping of load_row touches 32 rows x 16 bytes (32 wavefronts, half of every sector unused); the 8-byte bf16 stores of a row fill its
// swizzled 128-byte line.  Same values as load_row (same mask product, same hi / lo split).
template <bool X3>
__device__ __forceinline__ void load_rows_kcontig(const float* __restrict__ src, const float* __restrict__ msk, int mmode, int64_t ld_r,
                                                  int64_t row_g0, int64_t rows_valid, int k0, uint8_t* hi, uint8_t* lo, uint32_t row_l0,
                                                  uint32_t lane) {
  const bool sig = mmode == DFN_MASK_SIGMOID;
  const float mlo = mmode == DFN_MASK_LEAKY ? 0.02f : 0.f;
  const uint32_t piece = lane & 15u;           // 4 consecutive k
#pragma unroll 1
  for (int i0 = 0; i0 < 16; i0 += 8) {
    float4 v[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t r = row_g0 + 2 * (i0 + i) + (lane >> 4);
      const bool ok = r < rows_valid;
      const float4* p = reinterpret_cast<const float4*>(src + r * ld_r + k0) + piece;
      v[i] = ok ? __ldg(p) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (msk) y[i] = ok ? __ldg(reinterpret_cast<const float4*>(msk + r * ld_r + k0) + piece) : make_float4(1.f, 1.f, 1.f, 1.f);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 t = v[i];
      if (msk) {
        t.x *= mask_factor_sel(sig, mlo, y[i].x); t.y *= mask_factor_sel(sig, mlo, y[i].y);
        t.z *= mask_factor_sel(sig, mlo, y[i].z); t.w *= mask_factor_sel(sig, mlo, y[i].w);
      }
      const uint32_t row = row_l0 + 2u * (uint32_t)(i0 + i) + (lane >> 4);
      const uint32_t off = swz(row, piece >> 1) + (piece & 1u) * 8u;
      uint2 h;
      h.x = pack_bf16(t.x, t.y);
      h.y = pack_bf16(t.z, t.w);
      *reinterpret_cast<uint2*>(hi + off) = h;
      if (X3) {
        uint2 l;
        l.x = pack_lo<false>(t.x, t.y, h.x);
        l.y = pack_lo<false>(t.z, t.w, h.y);
        *reinterpret_cast<uint2*>(lo + off) = l;
      }
    }
  }
}

template <bool X3>
__global__ void __launch_bounds__(THREADS, 1) gemm_tc_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();
  const dfn_gemm_desc& d = P.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = sbase + SMEM_BAR;        // [2] stage filled by all loader threads
  const uint32_t bar_empty = sbase + SMEM_BAR + 16;  // [2] stage consumed by the MMAs
  const uint32_t bar_acc = sbase + SMEM_BAR + 32;    // accumulator complete
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SMEM_BAR + 48);

  if (threadIdx.x == 0) {
    for (int s = 0; s < N_STAGE; ++s) {
      mbar_init(bar_full + 8 * s, LOADERS);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 12) tmem_alloc(smem_u32(tmem_ptr_smem), 256);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int mt = (int)blockIdx.x % P.m_tiles, ks = (int)blockIdx.x / P.m_tiles;
  const int kc0 = ks * P.per_split;
  const int kc1 = min(P.chunks, kc0 + P.per_split);
  const int n_it = kc1 - kc0;     // >= 1 (host guarantees)
  const int64_t m0 = (int64_t)mt * BM;

  if (warp < 12) {
    // =========================================== operand loaders ===========================================
    const bool is_a = warp < 4;
    const uint32_t row = is_a ? (uint32_t)threadIdx.x : (uint32_t)(threadIdx.x - A_THREADS);
    const int64_t row_g = is_a ? m0 + row : (int64_t)row;
    const bool row_ok = is_a ? row_g < d.M : (int)row < d.N;
    const bool row_used = is_a || (int)row < P.n_pad;
    const float* src = is_a ? d.A : d.B;
    const float* msk = is_a ? d.A_mask : nullptr;
    const int64_t ld_r = is_a ? d.a_ld_r : d.b_ld_r, ld_k = is_a ? d.a_ld_k : d.b_ld_k;
    // k-contiguous operand with 16-byte aligned rows: the coalesced mapping (whole chunks only; the K tail takes the generic path)
    const bool kcontig = ld_k == 1 && (ld_r & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
                         (msk == nullptr || (reinterpret_cast<uintptr_t>(msk) & 15) == 0);
    const uint32_t row_l0 = (row & ~31u);                       // this warp's first tile row
    const int64_t rows_valid = is_a ? d.M : (int64_t)d.N;
    const bool warp_used = is_a || (int)row_l0 < P.n_pad;
    for (int it = 0; it < n_it; ++it) {
      const int s = it % N_STAGE;
      const uint32_t par = (uint32_t)(it / N_STAGE) & 1u;
      mbar_wait(bar_empty + 8 * s, par ^ 1u);
      uint8_t* st = smem + (size_t)s * STAGE_BYTES;
      uint8_t* hi = is_a ? st : st + 2 * A_BYTES;
      uint8_t* lo = is_a ? st + A_BYTES : st + 2 * A_BYTES + B_BYTES;
      const int k0 = (kc0 + it) * BK;
      if (kcontig && k0 + BK <= d.K) {
        if (warp_used) load_rows_kcontig<X3>(src, msk, d.a_mask_mode, ld_r, (is_a ? m0 : 0) + row_l0, rows_valid, k0, hi, lo, row_l0, (uint32_t)lane);
      } else if (row_used) load_row<X3>(src, msk, d.a_mask_mode, ld_r, ld_k, row_g, row_ok, k0, d.K, hi, lo, row);
      fence_proxy_async();
      mbar_arrive(bar_full + 8 * s);
    }
  } else {
    // ============================================= MMA issuer ==============================================
    const uint32_t idesc = make_idesc((uint32_t)P.n_pad);
    for (int it = 0; it < n_it; ++it) {
      const int s = it % N_STAGE;
      const uint32_t par = (uint32_t)(it / N_STAGE) & 1u;
      mbar_wait(bar_full + 8 * s, par);
      tcgen05_fence_after();
      const uint32_t st = sbase + (uint32_t)s * STAGE_BYTES;
      const uint64_t a_hi = make_smem_desc(st), a_lo = make_smem_desc(st + A_BYTES);
      const uint64_t b_hi = make_smem_desc(st + 2 * A_BYTES), b_lo = make_smem_desc(st + 2 * A_BYTES + B_BYTES);
#pragma unroll
      for (int q = 0; q < 4; ++q) umma_bf16(tmem_base, a_hi + 2 * q, b_hi + 2 * q, idesc, (it | q) != 0 ? 1u : 0u);
      if (X3) {
#pragma unroll
        for (int q = 0; q < 4; ++q) umma_bf16(tmem_base, a_lo + 2 * q, b_hi + 2 * q, idesc, 1u);
#pragma unroll
        for (int q = 0; q < 4; ++q) umma_bf16(tmem_base, a_hi + 2 * q, b_lo + 2 * q, idesc, 1u);
      }
      umma_commit(bar_empty + 8 * s);
    }
    umma_commit(bar_acc);
  }

  if (warp < 12) {
    // ================================================ epilogue ================================================
    // All twelve loader warps (they are idle by now): warp w reads TMEM lane quarter w % 4 and takes every third 16-column group --
    // with four warps this readout (bias, addend, activation and store per element) was two thirds of a CTA's 113k cycles.
    mbar_wait(bar_acc, 0u);
    tcgen05_fence_after();
    const int qw = warp & 3, cg = warp >> 2;     // lane quarter, column-group phase
    const uint32_t acc = tmem_base + ((uint32_t)(qw * 32) << 16);
    const bool atomic = d.k_splits > 1;
    const int act = d.act & 3;
    const bool pre_add = (d.act & 4) != 0;
    const float* const bias_p = ks == 0 ? d.bias : nullptr;       // (kernel parameters read once, not per element)
    const float* const add_p = ks == 0 ? d.addend : nullptr;
    const int64_t add_r = d.add_ld_r, add_c = d.add_ld_c;
    auto finish = [&](float x, int64_t m, int n) -> float {     // bias, addend, activation of one element (first K split only)
      if (bias_p != nullptr) x += __ldg(bias_p + n);
      const float add = add_p != nullptr ? add_p[m * add_r + (int64_t)n * add_c] : 0.f;
      if (pre_add) x += add;
      if (act == 1) x = fmaxf(x, 0.f);
      else if (act == 2) x = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
      else if (act == 3) x = x > 0.f ? x : __fmul_rn(0.02f, x);
      if (!pre_add) x += add;
      return x;
    };
    const bool pairs = d.c_ld_c == 1 && (d.c_ld_r & 1) == 0 && (reinterpret_cast<uintptr_t>(d.C) & 7) == 0;
    if (!atomic && !pairs) {
      // row per thread: 16 consecutive columns per TMEM load (outputs that are not row-major pairs: transposed or odd-pitched views)
      const int64_t m = m0 + (int64_t)(qw * 32 + lane);
      for (int c0 = cg * 16; c0 < P.n_pad; c0 += 48) {
        uint32_t v[16];
        tmem_ld16(acc + (uint32_t)c0, v);
        tmem_ld_wait();
        if (m < d.M) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = c0 + j;
            if (n >= d.N) continue;
            const float x = finish(__uint_as_float(v[j]), m, n);
            float* dst = d.C + m * d.c_ld_r + (int64_t)n * d.c_ld_c;
            *dst = d.beta ? *dst + x : x;
          }
        }
      }
    } else {
      // Column-distributed readout (tcgen05.ld.16x256b): a thread holds two adjacent columns of rows r and r + 8, so a row-major C takes
      // ONE 8-byte store (or, for a split-K partial tile, one 8-byte fp32 reduction) per pair and a warp's store instruction covers eight
      // rows x 32 contiguous bytes -- the row-per-thread readout above issues 4-byte stores at a row pitch per lane, 32 partial sectors per
      // instruction, and was a third of this kernel's time on the forward GEMM.
      for (int c0 = cg * 16; c0 < P.n_pad; c0 += 48) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[8];
          asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                       : "r"(acc + ((uint32_t)(16 * half) << 16) + (uint32_t)c0)
                       : "memory");
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 2; ++g) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int64_t m = m0 + qw * 32 + 16 * half + 8 * h + (lane >> 2);
              const int n = c0 + 8 * g + 2 * (lane & 3);
              if (m >= d.M || n >= d.N) continue;
              const bool two = n + 1 < d.N;
              const float x0 = finish(__uint_as_float(v[4 * g + 2 * h]), m, n);
              const float x1 = two ? finish(__uint_as_float(v[4 * g + 2 * h + 1]), m, n + 1) : 0.f;
              float* dst = d.C + m * d.c_ld_r + (int64_t)n * d.c_ld_c;
              if (atomic) {
                if (pairs && two) {
                  atomicAdd(reinterpret_cast<float2*>(dst), make_float2(x0, x1));
                } else {
                  atomicAdd(dst, x0);
                  if (two) atomicAdd(dst + d.c_ld_c, x1);
                }
              } else if (two) {
                float2* p2 = reinterpret_cast<float2*>(dst);
                float2 o = make_float2(x0, x1);
                if (d.beta) {
                  const float2 old = *p2;
                  o.x += old.x;
                  o.y += old.y;
                }
                *p2 = o;
              } else {
                *dst = d.beta ? *dst + x0 : x0;
              }
            }
          }
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 12) tmem_dealloc(tmem_base, 256);
}


// ---- DFN_PREC_FP32: the same contract on the CUDA cores (FFMA, round-to-nearest fp32 accumulation) ----
// The tensor core's fp32 accumulator truncates (measured: ~3e-8 of the running sum per accumulating MMA, biased -- splitting the
// operands into three bf16 pieces and issuing six products instead of three only moved the GEMM error from 4.7e-6 to 2.2e-6), and the
// backward pass of the deformation path is ill-conditioned: its gradient sums over the batch cancel to ~1e-3 of their terms, so a GEMM
// error of 4e-6 becomes 5e-3 on a dozen of the 74 gradient tensors (profiles/diag_train_precision.py; with exact GEMMs the same tape
// agrees with autograd to 3e-6 on all of them).  This kernel is the reference-exact mode of the training step, as the FFMA kernels of
// mlp_fp32.cu are for rendering: 64 x 64 output tile per CTA, 16-wide K steps through shared memory, 4 x 4 outputs per thread.
static constexpr int FM = 64, FN = 64, FK = 16;

// Acc = double: the per-frame vector products (M = 1 matrix-vector products, K = 1 outer products: at most 2^20 MACs) in EVERY precision.
// Their results (fc_z(z_shape), the signal encoders' outputs, ...) enter all points of the batch as biases, so their rounding error is
// coherent over the batch: 5e-7 there moved the first torso layer's bias-like gradients, sums that cancel to ~1e-4 of their terms, by 5e-3.
template <typename Acc>
__global__ void __launch_bounds__(256) gemm_fp32_kernel(const __grid_constant__ Params P) {
  __shared__ float As[FK][FM + 1];
  __shared__ float Bs[FK][FN + 1];
  const dfn_gemm_desc& d = P.d;
  const int n_tiles = (d.N + FN - 1) / FN;
  const int tile = (int)blockIdx.x % (P.m_tiles * n_tiles), ks = (int)blockIdx.x / (P.m_tiles * n_tiles);
  const int64_t m0 = (int64_t)(tile / n_tiles) * FM;
  const int n0 = (tile % n_tiles) * FN;
  const int k_begin = ks * P.per_split * FK, k_end = min(d.K, (ks + 1) * P.per_split * FK);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  Acc acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = (Acc)0;
  for (int k0 = k_begin; k0 < k_end; k0 += FK) {
    // A tile [64 x 16] and B tile [64 x 16]: 1,024 elements each, four per thread; the index split follows the unit stride
    for (int e = threadIdx.x; e < FM * FK; e += 256) {
      int r, k;
      if (d.a_ld_k == 1) { r = e / FK; k = e % FK; } else { r = e % FM; k = e / FM; }
      const int64_t m = m0 + r;
      float v = 0.f;
      if (m < d.M && k0 + k < k_end) {
        const int64_t off = m * d.a_ld_r + (int64_t)(k0 + k) * d.a_ld_k;
        v = d.A[off];
        if (d.A_mask != nullptr) v *= mask_factor(d.a_mask_mode, d.A_mask[off]);
      }
      As[k][r] = v;
    }
    for (int e = threadIdx.x; e < FN * FK; e += 256) {
      int r, k;
      if (d.b_ld_k == 1) { r = e / FK; k = e % FK; } else { r = e % FN; k = e / FN; }
      const int n = n0 + r;
      Bs[k][r] = (n < d.N && k0 + k < k_end) ? d.B[(int64_t)n * d.b_ld_r + (int64_t)(k0 + k) * d.b_ld_k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < FK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma((Acc)a[i], (Acc)b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool atomic = d.k_splits > 1;
  const int act = d.act & 3;
  const bool pre_add = (d.act & 4) != 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= d.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= d.N) continue;
      Acc xa = acc[i][j];
      if (d.bias != nullptr && ks == 0) xa += (Acc)d.bias[n];
      const float add = d.addend != nullptr && ks == 0 ? d.addend[m * d.add_ld_r + (int64_t)n * d.add_ld_c] : 0.f;
      if (pre_add) xa += (Acc)add;
      float x = (float)xa;
      if (act == 1) x = fmaxf(x, 0.f);
      else if (act == 2) x = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
      else if (act == 3) x = x > 0.f ? x : __fmul_rn(0.02f, x);
      if (!pre_add) x += add;
      float* dst = d.C + m * d.c_ld_r + (int64_t)n * d.c_ld_c;
      if (atomic) atomicAdd(dst, x);
      else *dst = d.beta ? *dst + x : x;
    }
  }
}

}  // namespace gemm
}  // namespace dfn

using namespace dfn;

extern "C" int dfn_gemm(const dfn_gemm_desc* d, void* stream) {
  DFN_CHECK_ARG(d && d->A && d->B && d->C && d->M > 0 && d->N > 0 && d->K > 0, "dfn_gemm: null operand or empty shape");
  DFN_CHECK_ARG(d->N <= gemm::BN_MAX, "dfn_gemm: N = %d > %d output columns per call", d->N, gemm::BN_MAX);
  DFN_CHECK_ARG(d->precision == DFN_PREC_BF16X3 || d->precision == DFN_PREC_BF16 || d->precision == DFN_PREC_FP32,
                "dfn_gemm: precision must be FP32, BF16X3 or BF16");
  DFN_CHECK_ARG(d->act >= 0 && d->act <= 7 && d->a_mask_mode >= 0 && d->a_mask_mode <= 3, "dfn_gemm: bad act / mask mode");
  DFN_CHECK_ARG(d->A_mask == nullptr || d->a_mask_mode != 0, "dfn_gemm: A_mask given without a mask mode");
  gemm::Params P;
  memset(&P, 0, sizeof(P));
  P.d = *d;
  if (P.d.A_mask == nullptr) P.d.a_mask_mode = 0;
  P.n_pad = (d->N + 15) / 16 * 16;
  const bool vec = (double)d->M * (double)d->N * (double)d->K <= 1048576.0;    // per-frame vector products: CUDA cores, fp64 accumulation
  const bool fp32 = d->precision == DFN_PREC_FP32 || vec;
  P.m_tiles = fp32 ? (d->M + gemm::FM - 1) / gemm::FM : (d->M + gemm::BM - 1) / gemm::BM;
  P.chunks = fp32 ? (d->K + gemm::FK - 1) / gemm::FK : (d->K + gemm::BK - 1) / gemm::BK;
  int splits = d->k_splits < 1 ? 1 : d->k_splits;
  if (splits > P.chunks) splits = P.chunks;
  P.per_split = (P.chunks + splits - 1) / splits;
  splits = (P.chunks + P.per_split - 1) / P.per_split;   // no empty split
  P.d.k_splits = splits;
  DFN_CHECK_ARG(splits == 1 || ((d->act & 3) == 0 && d->beta == 1),
                "dfn_gemm: a split-K call accumulates into C with atomics: it needs act = 0 and beta = 1 (C zeroed or running)");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t grid = (int64_t)P.m_tiles * splits * (fp32 ? (d->N + gemm::FN - 1) / gemm::FN : 1);
  DFN_CHECK_ARG(grid < (1ll << 31), "dfn_gemm: grid too large");
  static bool attr_done[2] = {false, false};
  const bool prof = profile_begin(st, (double)d->M * (double)d->N * (double)d->K);
  if (vec) {
    gemm::gemm_fp32_kernel<double><<<(int)grid, 256, 0, st>>>(P);
  } else if (fp32) {
    gemm::gemm_fp32_kernel<float><<<(int)grid, 256, 0, st>>>(P);
  } else if (d->precision == DFN_PREC_BF16X3) {
    if (!attr_done[0]) {
      DFN_CUDA(cudaFuncSetAttribute(gemm::gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm::SMEM_TOTAL));
      attr_done[0] = true;
    }
    gemm::gemm_tc_kernel<true><<<(int)grid, gemm::THREADS, gemm::SMEM_TOTAL, st>>>(P);
  } else {
    if (!attr_done[1]) {
      DFN_CUDA(cudaFuncSetAttribute(gemm::gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm::SMEM_TOTAL));
      attr_done[1] = true;
    }
    gemm::gemm_tc_kernel<false><<<(int)grid, gemm::THREADS, gemm::SMEM_TOTAL, st>>>(P);
  }
  if (prof) profile_end(st);
  DFN_LAUNCH_CHECK();
  return 0;
}
