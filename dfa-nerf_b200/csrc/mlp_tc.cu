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
#include "tc_ptx.cuh"
#include "tc_pack.h"
#include "tc_epi.cuh"
#include "pe.cuh"

namespace dfn {
namespace tc {

static constexpr int STAGE_BYTES = 128 * 128;      // one weight stage: <=128 rows x 64 bf16
static constexpr int N_STAGES = 4;
static constexpr int PE_HELPERS = 64;   // warps 2-3 encode the next tile while the current one is in its last layers
static constexpr int N_PAIRS = N_STAGES / 2;
static constexpr int ARENA_BLOCKS = 2 * TC_KB_PER_TILE;
static constexpr int SMEM_RING = ARENA_BLOCKS * KB_BYTES;
static constexpr int SMEM_BIAS = SMEM_RING + N_STAGES * STAGE_BYTES;
static constexpr int SMEM_BAR = SMEM_BIAS + 2 * TC_BIAS_STRIDE * 4;
static constexpr int SMEM_TOTAL = SMEM_BAR + 192;
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
  unsigned long long* trace;  // debug: per-role clock64 records of CTA 0 (null in production)
  int trace_tiles;            // local tiles per slot recorded
  TcLayer layers[TC_MAX_LAYERS];
};

// ------------------------------------------------------------------------------------ kernel
// X3 = false: bf16, two tile slots (384 threads).  X3 = true: split-bf16, one slot (256 threads).
// F16 (bf16-mode schedule only): fp16 operands -- P.w_hi then points at the fp16 weight stages.
template <bool X3, bool F16 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(X3 ? 256 : 384, 1)
    mlp_tc_kernel(const __grid_constant__ Params P) {
  static_assert(!(X3 && F16), "fp16 operands run in the single-pass schedule");
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
  const uint32_t bar_pefree = sbase + SMEM_BAR + 112;  // [2] the PE block of slot s is no longer read (last PE layer's MMAs done)
  const uint32_t bar_peready = sbase + SMEM_BAR + 128; // [2] the PE block of slot s holds the next tile's encoding

  if (threadIdx.x == 0) {
    for (int i = 0; i < N_STAGES; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 2);  // released by the MMA commits of both CTAs of the cluster
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_acc + 8 * s, 1);
      mbar_init(bar_aready + 8 * s, TILE_M);
      mbar_init(bar_pefree + 8 * s, TILE_M);
      mbar_init(bar_peready + 8 * s, PE_HELPERS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before any multicast copy or remote commit targets them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // Cluster of two CTAs: every weight stage is fetched from L2 once and multicast into both CTAs' rings
  // (the kernel is otherwise L2->SM bandwidth bound: each 128-point tile re-reads all 1.1 MB of weights).
  // The two CTAs therefore walk the same (iteration, layer, slot) sequence; tile group g = (j*C + c)*NSLOT + s
  // holds tiles 2g (rank 0) and 2g+1 (rank 1); a tile index past the end is computed on a clamped point
  // and not stored.
  const uint32_t crank = cluster_ctarank();
  const int C = gridDim.x >> 1, c = (int)blockIdx.x >> 1;
  const int n_groups = (P.n_tiles + 1) >> 1;
  const int n_local = c * NSLOT < n_groups ? (n_groups - c * NSLOT + C * NSLOT - 1) / (C * NSLOT) * NSLOT : 0;  // slots incl. tail
  const int n_iter = n_local / NSLOT;
  auto group_of = [&](int j, int s) { return (j * C + c) * NSLOT + s; };

  if (warp == 0) {
    // ============================== weight producer (TMA) ===============================
    {
      uint32_t cnt = 0;
      for (int j = 0; j < n_iter; ++j) {
        for (int l = 0; l < P.n_layers; ++l) {
          const TcLayer& L = P.layers[l];
          for (int s = 0; s < NSLOT; ++s) {
            if (group_of(j, s) >= n_groups) continue;
            // One ring entry = one K-block of weights (all L.n rows x 64 K) as two [rows x 32 K] 64-byte-swizzled
            // images in adjacent 16 KB slots, on ONE mbarrier: the issuing warp pays one barrier round trip per
            // four MMAs instead of per two (its loop, not the tensor pipe, was the limiter).
            uint32_t off = L.woff;
            const uint32_t bytes = (uint32_t)L.n * 64u;
            for (int kbi = 0; kbi < L.nkb; ++kbi) {
              if (!X3) {   // bf16: four single-stage entries ([n rows x 32 K], 16 KB), each refilled as soon as its two MMAs
                           // retire (+2.5 % over two 32 KB entries; the issuer is otherwise paced by the tensor pipe's
                           // operand fetch, not by its barrier waits -- a look-ahead barrier test changed nothing)
                for (int kh = 0; kh < 2; ++kh) {
                  const uint32_t e = cnt % N_STAGES, par = (cnt / N_STAGES) & 1u;
                  mbar_wait(bar_empty + 8 * e, par ^ 1u);
                  if (elect_one_sync()) {
                    mbar_expect_tx(bar_full + 8 * e, bytes);
                    if ((cnt & 1u) == crank)
                      tma_bulk_load_mc(sbase + SMEM_RING + e * STAGE_BYTES, P.w_hi + off + kh * bytes, bytes, bar_full + 8 * e, (uint16_t)3);
                  }
                  __syncwarp();
                  ++cnt;
                }
              } else
              for (int part = 0; part < NPART; ++part) {
                const uint32_t pair = cnt % N_PAIRS, par = (cnt / N_PAIRS) & 1u;
                mbar_wait(bar_empty + 8 * pair, par ^ 1u);
                if (elect_one_sync()) {
                  mbar_expect_tx(bar_full + 8 * pair, 2u * bytes);  // both CTAs arm their own barrier ...
                  const uint8_t* src = (part == 0 ? P.w_hi : P.w_lo) + off;
                  const uint32_t dst = sbase + SMEM_RING + pair * 2u * STAGE_BYTES;
                  if ((cnt & 1u) == crank) {                        // ... and take turns issuing the multicast copies
                    tma_bulk_load_mc(dst, src, bytes, bar_full + 8 * pair, (uint16_t)3);
                    tma_bulk_load_mc(dst + STAGE_BYTES, src + bytes, bytes, bar_full + 8 * pair, (uint16_t)3);
                  }
                }
                __syncwarp();
                ++cnt;
              }
              off += 2u * bytes;
            }
          }
        }
      }
      // tail: wait for the final release of every ring entry (it needs the peer CTA's remote arrival too)
      const uint32_t n_ent = X3 ? N_PAIRS : N_STAGES;
      for (uint32_t k = 0; k < n_ent && k < cnt; ++k) {
        const uint32_t u = cnt - 1u - k;
        mbar_wait(bar_empty + 8 * (u % n_ent), (u / n_ent) & 1u);
      }
    }
  } else if (warp == 1) {
    // ================================= MMA issuer =========================================
    {
      uint32_t cnt = 0;
      uint32_t apar[2] = {0u, 0u};
      for (int j = 0; j < n_iter; ++j) {
        for (int l = 0; l < P.n_layers; ++l) {
          const TcLayer& L = P.layers[l];
          for (int s = 0; s < NSLOT; ++s) {
            if (group_of(j, s) >= n_groups) continue;
            const bool tr = P.trace != nullptr && blockIdx.x == 0 && j < P.trace_tiles;
            long long t_w0 = 0, t_w1 = 0, t_full = 0;
            if (tr) t_w0 = clock64();
            mbar_wait(bar_aready + 8 * s, apar[s]);
            apar[s] ^= 1u;
            tcgen05_fence_after();
            if (tr) t_w1 = clock64();
            const uint32_t acc = tmem_base + (uint32_t)s * 256u;
            const uint32_t idesc = F16 ? make_idesc_f16(L.n) : make_idesc(L.n);
            for (int kbi = 0; kbi < L.nkb; ++kbi) {
              const uint32_t a_hi = sbase + (uint32_t)(s * TC_KB_PER_TILE + L.kb[kbi]) * KB_BYTES;
              const uint64_t adesc_hi = make_smem_desc(a_hi);
              const uint64_t adesc_lo = make_smem_desc(a_hi + TC_KB_PER_TILE * KB_BYTES);
              if (!X3) {
#pragma unroll
                for (int kh = 0; kh < 2; ++kh) {
                  const uint32_t e = cnt % N_STAGES, par = (cnt / N_STAGES) & 1u;
                  long long t_f0 = 0;
                  if (tr) t_f0 = clock64();
                  mbar_wait(bar_full + 8 * e, par);
                  tcgen05_fence_after();
                  if (tr) t_full += clock64() - t_f0;
                  const uint64_t bdesc = make_smem_desc_sw64(sbase + SMEM_RING + e * STAGE_BYTES);
#pragma unroll
                  for (int ks = 0; ks < 2; ++ks)
                    umma_bf16(acc, adesc_hi + 2 * (2 * kh + ks), bdesc + (uint64_t)(ks * 2), idesc, (kbi | kh | ks) != 0 ? 1u : 0u);
                  umma_commit_mc(bar_empty + 8 * e, (uint16_t)3);
                  ++cnt;
                }
              } else
              for (int part = 0; part < NPART; ++part) {
                const uint32_t pair = cnt % N_PAIRS, par = (cnt / N_PAIRS) & 1u;
                long long t_f0 = 0;
                if (tr) t_f0 = clock64();
                mbar_wait(bar_full + 8 * pair, par);
                tcgen05_fence_after();
                if (tr) t_full += clock64() - t_f0;
                const uint64_t bdesc = make_smem_desc_sw64(sbase + SMEM_RING + pair * 2u * STAGE_BYTES);
#pragma unroll
                for (int q = 0; q < 4; ++q) {  // q = 2*kh + ks: A advances 32 bytes per K-step, B 32 bytes inside a 16 KB image
                  const uint64_t bd = bdesc + (uint64_t)((q >> 1) * (STAGE_BYTES >> 4) + (q & 1) * 2);
                  umma_bf16(acc, adesc_hi + 2 * q, bd, idesc, (kbi | part | q) != 0 ? 1u : 0u);
                }
                if (X3 && part == 0) {
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    const uint64_t bd = bdesc + (uint64_t)((q >> 1) * (STAGE_BYTES >> 4) + (q & 1) * 2);
                    umma_bf16(acc, adesc_lo + 2 * q, bd, idesc, 1u);
                  }
                }
                umma_commit_mc(bar_empty + 8 * pair, (uint16_t)3);
                ++cnt;
              }
            }
            umma_commit(bar_acc + 8 * s);
            if (tr && lane == 0) {
              unsigned long long* r = P.trace + ((size_t)(j * P.n_layers + l) * 2 + s) * 4;
              r[0] = (unsigned long long)t_w0;
              r[1] = (unsigned long long)t_w1;
              r[2] = (unsigned long long)clock64();
              r[3] = (unsigned long long)t_full;
            }
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
    int last_pe_layer = 0;
    for (int l2 = 0; l2 < P.n_layers; ++l2)
      for (int k = 0; k < P.layers[l2].nkb; ++k)
        if (P.layers[l2].kb[k] == TC_KB_PE) last_pe_layer = l2;

    for (int j = 0; j < n_iter; ++j) {
      if (group_of(j, s) >= n_groups) break;
      const int i = j * NSLOT + s;
      const int tile = 2 * group_of(j, s) + (int)crank;
      const long long t_tile0 = clock64();
      int64_t pt = (int64_t)tile * TILE_M + row;
      const bool valid = pt < P.n_points;
      if (!valid) pt = P.n_points - 1;
      const int64_t ray = pt / P.S;

      // stage the first layer's bias while the encoding is computed
      {
        const float2 b2 = reinterpret_cast<const float2*>(P.bias)[tid_s];
        reinterpret_cast<float2*>(bias_s)[tid_s] = b2;
      }
      // the tile's positional encoding was written into the PE K-block by the helper warps (below), one tile ahead
      mbar_wait(bar_peready + 8 * s, (uint32_t)j & 1u);
      fence_proxy_async();
      mbar_arrive(bar_aready + 8 * s);
      named_bar_sync(1 + s, TILE_M);  // bias_s visible to the slot's four warps

      float alpha = 0.f;
      for (int l = 0; l < P.n_layers; ++l) {
        const TcLayer& L = P.layers[l];
        // prefetch next layer's shared bias (latency hidden behind the accumulator wait)
        float2 nb = make_float2(0.f, 0.f);
        if (l + 1 < P.n_layers) nb = reinterpret_cast<const float2*>(P.bias + (l + 1) * TC_BIAS_STRIDE)[tid_s];

        const bool tr = P.trace != nullptr && blockIdx.x == 0 && tid_s == 0 && (i / NSLOT) < P.trace_tiles;
        long long t_e0 = 0, t_e1 = 0;
        if (tr) t_e0 = clock64();
        if (L.epi == TC_EPI_VIEW0) prefetch_row_l1(P.view_bias + ray * P.view_w, P.view_w);   // hidden behind the wait
        mbar_wait(bar_acc + 8 * s, acc_par);
        acc_par ^= 1u;
        tcgen05_fence_after();
        if (tr) t_e1 = clock64();
        if (l == last_pe_layer) mbar_arrive(bar_pefree + 8 * s);   // its MMAs were the last readers of the PE block

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
          if (L.epi == TC_EPI_VIEW0)
            epilogue_relu<X3, true, F16>(acc, P.view_w, P.view_bias + ray * P.view_w, 0u, arena_hi, arena_lo, row);
          else
            epilogue_relu<X3, false, F16>(acc, (int)L.n, nullptr, smem_u32(bias_s), arena_hi, arena_lo, row);
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
        if (tr) {
          unsigned long long* r = P.trace + (size_t)P.trace_tiles * P.n_layers * 8 +
                                  ((size_t)((i / NSLOT) * P.n_layers + l) * 2 + s) * 4;
          r[0] = (unsigned long long)t_e0;
          r[1] = (unsigned long long)t_e1;
          r[2] = (unsigned long long)clock64();
          r[3] = (unsigned long long)t_tile0;
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

  if (warp == 2 || warp == 3) {
    // ============================ positional-encoding helper warps ==============================
    // x = o + d*z -> [x | sin(2^k x) | cos(2^k x)] (HELP:42-52, pe.cuh) for the NEXT tile of each slot, written straight
    // into the slot's PE K-block as soon as the last layer that reads it (the skip layer) has finished its MMAs: the
    // ~2,800 cycles of encoding leave the slot's dependency chain.  Two rows per thread and slot.
    const int t = (warp - 2) * 32 + lane;
    for (int j = 0; j < n_iter; ++j) {
      for (int s = 0; s < NSLOT; ++s) {
        if (group_of(j, s) >= n_groups) continue;
        if (j > 0) mbar_wait(bar_pefree + 8 * s, (uint32_t)(j - 1) & 1u);
        uint8_t* pe_hi = smem + (size_t)(s * TC_KB_PER_TILE + TC_KB_PE) * KB_BYTES;
        uint8_t* pe_lo = pe_hi + (size_t)TC_KB_PER_TILE * KB_BYTES;
        const int tile = 2 * group_of(j, s) + (int)crank;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const uint32_t row = (uint32_t)(t + 64 * h);
          int64_t pt = (int64_t)tile * TILE_M + row;
          if (pt >= P.n_points) pt = P.n_points - 1;
          const int64_t ray = pt / P.S;
          float pe[64], x[3];
          sample_point(P.rays_o, P.rays_d, ray, P.z_vals[pt], x);
          pe_embedder(x, P.multires, pe);
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = pe[ch * 8 + e];
            store_chunk<X3, F16>(pe_hi, pe_lo, row, (uint32_t)ch, o);
          }
        }
        fence_proxy_async();
        mbar_arrive(bar_peready + 8 * s);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still multicast into this CTA's ring / arrive on its barriers until it is done
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
// views_linears.0 (HELP:288-292) are constant along a ray.  A block walks rays four at a time; thread n owns
// output n of each.  PE layout as HELP:42-52 with multires_views frequencies.
static constexpr int VB_RAYS = 4;
__global__ void view_bias_kernel(int64_t R, int Wh, int ncol, int L, const float* __restrict__ viewdirs,
                                 const float* __restrict__ vw, const float* __restrict__ vb,
                                 float* __restrict__ out) {
  extern __shared__ float sm[];
  float* w_s = sm;                 // [Wh][ncol]
  float* pe = sm + Wh * ncol;      // [VB_RAYS][ncol]
  for (int i = threadIdx.x; i < Wh * ncol; i += blockDim.x) w_s[i] = vw[i];
  const float bn = (int)threadIdx.x < Wh ? vb[threadIdx.x] : 0.f;
  for (int64_t r0 = (int64_t)blockIdx.x * VB_RAYS; r0 < R; r0 += (int64_t)gridDim.x * VB_RAYS) {
    __syncthreads();
    for (int i = threadIdx.x; i < VB_RAYS * ncol; i += blockDim.x) {
      const int q = i / ncol, j = i % ncol;
      const int64_t r = r0 + q < R ? r0 + q : R - 1;
      float v;
      if (j < 3) {
        v = viewdirs[r * 3 + j];
      } else {
        const int k = (j - 3) / 6, c = (j - 3) % 6;
        const float a = __fmul_rn(viewdirs[r * 3 + (c % 3)], pow2i(k));
        v = c < 3 ? sinf(a) : cosf(a);
      }
      pe[i] = v;
    }
    __syncthreads();
    if ((int)threadIdx.x < Wh) {
      float acc[VB_RAYS];
#pragma unroll
      for (int q = 0; q < VB_RAYS; ++q) acc[q] = 0.f;
      const float* w = w_s + threadIdx.x * ncol;
      for (int j = 0; j < ncol; ++j) {
        const float wj = w[j];
#pragma unroll
        for (int q = 0; q < VB_RAYS; ++q) acc[q] = fmaf(wj, pe[q * ncol + j], acc[q]);
      }
#pragma unroll
      for (int q = 0; q < VB_RAYS; ++q)
        if (r0 + q < R) out[(r0 + q) * Wh + threadIdx.x] = bn + acc[q];
    }
  }
}


// One launch in front of a hierarchical render (dfn_render_rays): the per-call preparation of BOTH networks and the coarse
// depths -- fold_latent_kernel x 2, view_bias_kernel x 2 and z_vals_kernel in one grid.  Blocks [0, nl_a + nl_b) fold the
// per-frame latent into the layer biases of network a / b; the other blocks walk the rays VB_RAYS at a time: the view
// direction is encoded once, threads [0, Wh) form network a's per-ray bias row, threads [Wh, 2 Wh) network b's, and the
// block writes the rays' coarse depth rows (MAIN:617-619 + optional stratified jitter), with the arithmetic of the three
// kernels it replaces.
struct PrepNet {
  const float* bias;     // [n_layers][256]
  const float* fold_w;   // [2][W][dim_aud]
  const float* view_w;   // [Wh][ncol]
  const float* view_b;   // [Wh]
  float* bias_out;       // [n_layers][256]
  float* vbias_out;      // [R][Wh]
  int n_layers, fold0, fold1;
};

// NCOL = input_ch_views at compile time (27 for multires_views = 4: a thread keeps its output's view-direction weights in registers and
// reads the encodings of PREP_RAYS rays as two broadcast 16-byte loads per column -- the first version re-read the weight row and four
// scalar encodings from shared memory per column and was bound by those loads, 305 us per 202,500-ray frame) or 0 (any width: weights in
// shared memory).  Same fmaf chain over the columns in both, so the rows are bit-identical.
static constexpr int PREP_RAYS = 8;
template <int NCOL>
__global__ void __launch_bounds__(256) render_prep_kernel(PrepNet A, PrepNet B, int W, int dim_aud, const float* __restrict__ latent, int64_t R, int Wh,
                                   int ncol_rt, const float* __restrict__ viewdirs, int Nc, const float* __restrict__ t_vals,
                                   const float* __restrict__ near, const float* __restrict__ far,
                                   const float* __restrict__ rnd, float* __restrict__ z_out) {
  const int n_fold_blocks = A.n_layers + B.n_layers;
  if ((int)blockIdx.x < n_fold_blocks) {
    const bool second = (int)blockIdx.x >= A.n_layers;
    const PrepNet& N = second ? B : A;
    const int l = second ? (int)blockIdx.x - A.n_layers : (int)blockIdx.x, n = threadIdx.x;
    if (n >= TC_BIAS_STRIDE) return;
    float v = N.bias[l * TC_BIAS_STRIDE + n];
    const int which = l == N.fold0 ? 0 : (l == N.fold1 ? 1 : -1);
    if (which >= 0 && latent != nullptr && n < W) {
      const float* w = N.fold_w + ((size_t)which * W + n) * dim_aud;
      float acc = 0.f;
      for (int j = 0; j < dim_aud; ++j) acc = fmaf(w[j], latent[j], acc);
      v += acc;
    }
    N.bias_out[l * TC_BIAS_STRIDE + n] = v;
    return;
  }
  const int ncol = NCOL ? NCOL : ncol_rt;
  extern __shared__ __align__(16) float sm[];
  float* pe = sm;                          // [ncol][PREP_RAYS]
  float* w_s = sm + ncol * PREP_RAYS;      // NCOL == 0: [2][Wh][ncol]
  const int net = (int)threadIdx.x / Wh, nn = (int)threadIdx.x % Wh;      // blockDim = 2 * Wh
  const PrepNet& N = net == 0 ? A : B;
  float wreg[NCOL ? NCOL : 1];
  if (NCOL) {
#pragma unroll
    for (int j = 0; j < (NCOL ? NCOL : 1); ++j) wreg[j] = N.view_w[nn * ncol + j];
  } else {
    for (int i = threadIdx.x; i < Wh * ncol; i += blockDim.x) {
      w_s[i] = A.view_w[i];
      w_s[Wh * ncol + i] = B.view_w[i];
    }
  }
  const float bn = N.view_b[nn];
  const int ray_blocks = (int)gridDim.x - n_fold_blocks;
  for (int64_t r0 = (int64_t)((int)blockIdx.x - n_fold_blocks) * PREP_RAYS; r0 < R; r0 += (int64_t)ray_blocks * PREP_RAYS) {
    __syncthreads();
    for (int i = threadIdx.x; i < PREP_RAYS * ncol; i += blockDim.x) {
      const int q = i / ncol, j = i % ncol;
      const int64_t r = r0 + q < R ? r0 + q : R - 1;
      float v;
      if (j < 3) {
        v = viewdirs[r * 3 + j];
      } else {
        const int k = (j - 3) / 6, c = (j - 3) % 6;
        const float a = __fmul_rn(viewdirs[r * 3 + (c % 3)], pow2i(k));
        v = c < 3 ? sinf(a) : cosf(a);
      }
      pe[j * PREP_RAYS + q] = v;
    }
    __syncthreads();
    {
      float acc[PREP_RAYS];
#pragma unroll
      for (int q = 0; q < PREP_RAYS; ++q) acc[q] = 0.f;
      auto column = [&](float wj, int j) {
        const float4 p0 = *reinterpret_cast<const float4*>(pe + j * PREP_RAYS);
        const float4 p1 = *reinterpret_cast<const float4*>(pe + j * PREP_RAYS + 4);
        acc[0] = fmaf(wj, p0.x, acc[0]); acc[1] = fmaf(wj, p0.y, acc[1]); acc[2] = fmaf(wj, p0.z, acc[2]); acc[3] = fmaf(wj, p0.w, acc[3]);
        acc[4] = fmaf(wj, p1.x, acc[4]); acc[5] = fmaf(wj, p1.y, acc[5]); acc[6] = fmaf(wj, p1.z, acc[6]); acc[7] = fmaf(wj, p1.w, acc[7]);
      };
      if (NCOL) {
#pragma unroll
        for (int j = 0; j < (NCOL ? NCOL : 1); ++j) column(wreg[j], j);
      } else {
        const float* w = w_s + (size_t)net * Wh * ncol + nn * ncol;
        for (int j = 0; j < ncol; ++j) column(w[j], j);
      }
#pragma unroll
      for (int q = 0; q < PREP_RAYS; ++q)
        if (r0 + q < R) N.vbias_out[(r0 + q) * Wh + nn] = bn + acc[q];
    }
    // coarse depths of these rays (z_vals_kernel)
    for (int i = threadIdx.x; i < PREP_RAYS * Nc; i += blockDim.x) {
      const int q = i / Nc, sidx = i % Nc;
      const int64_t r = r0 + q;
      if (r >= R) break;
      const float nr = near[r], fr = far[r];
      auto zf = [&](int k) {
        const float t = t_vals[k];
        return __fadd_rn(__fmul_rn(nr, __fsub_rn(1.0f, t)), __fmul_rn(fr, t));
      };
      float z = zf(sidx);
      if (rnd) {
        const float lower = sidx == 0 ? z : __fmul_rn(0.5f, __fadd_rn(z, zf(sidx - 1)));
        const float upper = sidx == Nc - 1 ? z : __fmul_rn(0.5f, __fadd_rn(zf(sidx + 1), z));
        z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), rnd[r * Nc + sidx]));
      }
      z_out[r * Nc + sidx] = z;
    }
  }
}

}  // namespace tc

void tc_free_model(dfn_model* m) {
  cudaFree(m->tc2_hi);
  cudaFree(m->tc2_h16);
  m->tc2_hi = m->tc2_h16 = nullptr;
  cudaFree(m->tc_hi);
  cudaFree(m->tc_lo);
  cudaFree(m->tc_h16);
  cudaFree(m->tc_l16);
  m->tc_h16 = m->tc_l16 = nullptr;
  cudaFree(m->tc_bias);
  cudaFree(m->tc_fold_w);
  cudaFree(m->tc_view_w);
  cudaFree(m->tc_view_b);
  m->tc_hi = m->tc_lo = nullptr;
  m->tc_bias = m->tc_fold_w = m->tc_view_w = m->tc_view_b = nullptr;
}

// Builds the layer program and the packed blobs from the reference tensors (order of dfn.h).
// dump != nullptr: host-only dry run for dfn_model_program_host (no CUDA calls; the program, biases and fold / view
// matrices are returned in *dump).
int tc_pack_model(dfn_model* m, const float* const* t, cudaStream_t st, TcHostDump* dump) {
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
  if (dump) pk.dense = &dump->dense;
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
    m->tc32_woff[nl] = pk.last32;
    m->tc2_woff[nl] = pk.last2;
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
    m->tc32_woff[nl] = pk.last32;
    m->tc2_woff[nl] = pk.last2;
    pg.layers[nl++] = L;
    // view-direction columns + composed bias for view_bias_kernel
    std::vector<float> vw((size_t)Wh * d.input_ch_views);
    for (int n = 0; n < Wh; ++n)
      for (int j = 0; j < d.input_ch_views; ++j) vw[(size_t)n * d.input_ch_views + j] = wv[(size_t)n * ldv + W + j];
    if (dump) {
      dump->view_w = vw;
      dump->view_b = bc;
    } else {
      DFN_CUDA(cudaMalloc(&m->tc_view_w, vw.size() * 4));
      DFN_CUDA(cudaMalloc(&m->tc_view_b, bc.size() * 4));
      DFN_CUDA(cudaMemcpyAsync(m->tc_view_w, vw.data(), vw.size() * 4, cudaMemcpyHostToDevice, st));
      DFN_CUDA(cudaMemcpyAsync(m->tc_view_b, bc.data(), bc.size() * 4, cudaMemcpyHostToDevice, st));
      DFN_CUDA(cudaStreamSynchronize(st));
    }
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
    m->tc32_woff[nl] = pk.last32;
    m->tc2_woff[nl] = pk.last2;
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
    m->tc32_woff[nl] = pk.last32;
    m->tc2_woff[nl] = pk.last2;
    pg.layers[nl++] = L;
  }
  pg.n_layers = nl;
  m->tc_blob_bytes = (int64_t)pk.hi32.size();
  if (dump) {
    dump->bias = bias;
    dump->fold_w = fold_w;
    return 0;
  }
  DFN_CUDA(cudaMalloc(&m->tc2_hi, pk.hi2.size()));
  DFN_CUDA(cudaMalloc(&m->tc2_h16, pk.h16_2.size()));
  DFN_CUDA(cudaMemcpyAsync(m->tc2_hi, pk.hi2.data(), pk.hi2.size(), cudaMemcpyHostToDevice, st));
  DFN_CUDA(cudaMemcpyAsync(m->tc2_h16, pk.h16_2.data(), pk.h16_2.size(), cudaMemcpyHostToDevice, st));
  DFN_CUDA(cudaMalloc(&m->tc_hi, pk.hi32.size()));
  DFN_CUDA(cudaMalloc(&m->tc_lo, pk.lo32.size()));
  DFN_CUDA(cudaMalloc(&m->tc_h16, pk.h16.size()));
  DFN_CUDA(cudaMemcpyAsync(m->tc_h16, pk.h16.data(), pk.h16.size(), cudaMemcpyHostToDevice, st));
  DFN_CUDA(cudaMalloc(&m->tc_l16, pk.l16.size()));
  DFN_CUDA(cudaMemcpyAsync(m->tc_l16, pk.l16.data(), pk.l16.size(), cudaMemcpyHostToDevice, st));
  DFN_CUDA(cudaMalloc(&m->tc_bias, bias.size() * 4));
  DFN_CUDA(cudaMalloc(&m->tc_fold_w, fold_w.size() * 4));
  DFN_CUDA(cudaMemcpyAsync(m->tc_hi, pk.hi32.data(), pk.hi32.size(), cudaMemcpyHostToDevice, st));
  DFN_CUDA(cudaMemcpyAsync(m->tc_lo, pk.lo32.data(), pk.lo32.size(), cudaMemcpyHostToDevice, st));
  DFN_CUDA(cudaMemcpyAsync(m->tc_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice, st));
  DFN_CUDA(cudaMemcpyAsync(m->tc_fold_w, fold_w.data(), fold_w.size() * 4, cudaMemcpyHostToDevice, st));
  DFN_CUDA(cudaStreamSynchronize(st));
  return 0;
}

static inline int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }

// debug timeline hook (dfn_debug_trace): device buffer of 2 * trace_tiles * n_layers * 8 uint64
static int g_impl = -1;  // -1 auto; debug (dfn_debug_set_impl): low 4 bits 1 mlp_tc.cu, 2 mlp_pp.cu, 3 mlp_pair.cu (8: four epilogue warps per slot); high bits: pair flags + 1
void tc_set_impl(int impl) { g_impl = impl; }
static void* g_trace_ptr = nullptr;
static int g_trace_tiles = 0;
void tc_set_trace(void* dev_ptr, int tiles) {
  g_trace_ptr = dev_ptr;
  g_trace_tiles = dev_ptr ? tiles : 0;
}

void tc_get_trace(void** dev_ptr, int* tiles) {
  *dev_ptr = g_trace_ptr;
  *tiles = g_trace_tiles;
}

int64_t tc_query_workspace_bytes(const dfn_model* m, int64_t R, int S) {
  (void)S;
  return align256((int64_t)TC_MAX_LAYERS * TC_BIAS_STRIDE * 4) + align256(R * (m->desc.W / 2) * 4) + align256(pp_scratch_bytes());
}

// The preparation launch of a hierarchical render: biases (latent folded) and per-ray view-bias rows of both networks + the
// coarse depth rows.  pre_a / pre_b: tc_prep_bytes(R) each.  Returns DFN_E_UNSUPPORTED when the two networks cannot share it.
int64_t tc_prep_bytes(const dfn_model* m, int64_t R) {
  return align256((int64_t)TC_MAX_LAYERS * TC_BIAS_STRIDE * 4) + align256(R * (m->desc.W / 2) * 4);
}

int tc_prep_launch(const dfn_model* a, const dfn_model* b, int64_t R, int Nc, const float* viewdirs, const float* latent,
                   const float* t_vals, const float* near, const float* far, const float* rnd, float* z0, void* pre_a, void* pre_b,
                   cudaStream_t st) {
  const dfn_model_desc& da = a->desc;
  const dfn_model_desc& db = b->desc;
  if (a->tc_hi == nullptr || b->tc_hi == nullptr || da.W != db.W || da.input_ch_views != db.input_ch_views || da.dim_aud != db.dim_aud ||
      da.multires_views != db.multires_views || da.W / 2 > 128 || (da.dim_aud > 0 && latent == nullptr))
    return DFN_E_UNSUPPORTED;
  const int Wh = da.W / 2;
  auto net = [&](const dfn_model* m, void* pre) {
    tc::PrepNet n;
    n.bias = m->tc_bias;
    n.fold_w = m->tc_fold_w;
    n.view_w = m->tc_view_w;
    n.view_b = m->tc_view_b;
    n.bias_out = reinterpret_cast<float*>(pre);
    n.vbias_out = reinterpret_cast<float*>(reinterpret_cast<char*>(pre) + align256((int64_t)TC_MAX_LAYERS * TC_BIAS_STRIDE * 4));
    n.n_layers = m->prog.n_layers;
    n.fold0 = m->prog.fold_layer[0];
    n.fold1 = m->prog.fold_layer[1];
    return n;
  };
  int64_t ray_blocks = (R + tc::PREP_RAYS - 1) / tc::PREP_RAYS;
  if (ray_blocks > (int64_t)num_sms() * 8) ray_blocks = (int64_t)num_sms() * 8;
  const int grid = a->prog.n_layers + b->prog.n_layers + (int)ray_blocks;
  const int ncol = da.input_ch_views;
  if (ncol == 27) {
    tc::render_prep_kernel<27><<<grid, 2 * Wh, (size_t)tc::PREP_RAYS * ncol * sizeof(float), st>>>(
        net(a, pre_a), net(b, pre_b), da.W, da.dim_aud, latent, R, Wh, ncol, viewdirs, Nc, t_vals, near, far, rnd, z0);
  } else {
    const size_t sm = ((size_t)2 * Wh * ncol + tc::PREP_RAYS * ncol) * sizeof(float);
    tc::render_prep_kernel<0><<<grid, 2 * Wh, sm, st>>>(net(a, pre_a), net(b, pre_b), da.W, da.dim_aud, latent, R, Wh, ncol, viewdirs, Nc,
                                                        t_vals, near, far, rnd, z0);
  }
  DFN_LAUNCH_CHECK();
  return 0;
}

template <class K>
static int launch_tc(K kernel, int grid, int threads, const tc::Params& P, cudaStream_t st) {
  DFN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_TOTAL));
  kernel<<<grid, threads, tc::SMEM_TOTAL, st>>>(P);
  return 0;
}

int tc_query_points(const dfn_model* m, int64_t R, int S, const float* rays_o, const float* rays_d,
                    const float* viewdirs, const float* z_vals, const float* latent, float* raw,
                    int precision, void* workspace, int64_t workspace_bytes, cudaStream_t st, const void* pre) {
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
  void* const scratch_ws = reinterpret_cast<char*>(vbias_ws) + align256(R * (int64_t)Wh * 4);

  if (pre != nullptr) {       // biases and view-bias rows already formed by tc_prep_launch
    bias_ws = reinterpret_cast<float*>(const_cast<void*>(pre));
    vbias_ws = reinterpret_cast<float*>(reinterpret_cast<char*>(const_cast<void*>(pre)) + align256((int64_t)TC_MAX_LAYERS * TC_BIAS_STRIDE * 4));
  } else {
  tc::fold_latent_kernel<<<m->prog.n_layers, TC_BIAS_STRIDE, 0, st>>>(
      m->prog.n_layers, d.W, d.dim_aud, m->tc_bias, m->tc_fold_w, latent, m->prog.fold_layer[0], m->prog.fold_layer[1], bias_ws);
  DFN_LAUNCH_CHECK();
  {
    int64_t blocks = (R + tc::VB_RAYS - 1) / tc::VB_RAYS;
    if (blocks > (int64_t)num_sms() * 8) blocks = (int64_t)num_sms() * 8;
    size_t sm = ((size_t)Wh * d.input_ch_views + tc::VB_RAYS * d.input_ch_views) * sizeof(float);
    tc::view_bias_kernel<<<(int)blocks, 128, sm, st>>>(R, Wh, d.input_ch_views, d.multires_views, viewdirs, m->tc_view_w,
                                                        m->tc_view_b, vbias_ws);
    DFN_LAUNCH_CHECK();
  }
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
  P.trace = reinterpret_cast<unsigned long long*>(g_trace_ptr);
  P.trace_tiles = g_trace_tiles;
  for (int i = 0; i < m->prog.n_layers; ++i) {
    P.layers[i] = m->prog.layers[i];
    P.layers[i].woff = m->tc32_woff[i];
  }

  int grid = P.n_tiles < num_sms() ? P.n_tiles : num_sms();
  grid = (grid + 1) & ~1;  // clusters of two CTAs
  // algorithmic MACs per point with the latent / view-direction columns folded into biases
  double macs_pt = 0.0;
  for (int i = 0; i < d.D; ++i) {
    const bool has_in = (i == 0) || (i - 1 == d.skip);
    macs_pt += (double)d.W * ((i == 0 ? 0 : d.W) + (has_in ? d.input_ch : 0));
  }
  macs_pt += (double)d.W * Wh + d.W;                       // views_linears.0 (+ composed feature) and alpha
  macs_pt += (double)(m->n_views - 1) * Wh * Wh + 3.0 * Wh;  // remaining view layers and rgb
  const bool prof = profile_begin(st, macs_pt * (double)P.n_points);
  // default: bf16x3 -> mlp_pp.cu (2), bf16 / fp16 -> the CTA-pair kernel mlp_pair.cu (3)
  const int impl = g_impl >= 0 ? (g_impl & 15) : ((precision == DFN_PREC_BF16X3 || precision == DFN_PREC_FP16X3M) ? 2 : 3);
  pair_set_flags(g_impl >= 16 ? (g_impl >> 4) - 1 : 7);   // debug: flags + 1 in the high bits (bit 3: Decoder head stays on mlp_pp.cu)
  if (precision == DFN_PREC_FP16X3M || (impl == 2 && (precision == DFN_PREC_BF16 || precision == DFN_PREC_BF16X3))) {
    int rc = pp_launch(m, bias_ws, vbias_ws, scratch_ws, R, S, rays_o, rays_d, z_vals, raw, precision, st);
    if (rc) return rc;
  } else if ((impl == 3 || impl == 8) && (precision == DFN_PREC_BF16 || precision == DFN_PREC_FP16)) {
    pair_set_epilogue_warps(impl == 8 ? 4 : 8);
    int rc = pair_launch(m, bias_ws, vbias_ws, R, S, rays_o, rays_d, z_vals, raw, precision, st);
    if (rc) return rc;
  } else if (precision == DFN_PREC_BF16 || precision == DFN_PREC_FP16 || precision == DFN_PREC_BF16X3) {
    // the 1-CTA generation (impl 1): clusters of two CTAs that only share the weight stream by multicast
    int rc;
    if (precision == DFN_PREC_BF16X3) rc = launch_tc(tc::mlp_tc_kernel<true>, grid, 256, P, st);
    else if (precision == DFN_PREC_FP16) {
      P.w_hi = m->tc_h16;
      rc = launch_tc(tc::mlp_tc_kernel<false, true>, grid, 384, P, st);
    } else rc = launch_tc(tc::mlp_tc_kernel<false>, grid, 384, P, st);
    if (rc) return rc;
  } else {
    set_error("tc_query_points: precision %d is not a tensor-core mode", precision);
    return DFN_E_ARG;
  }
  if (prof) profile_end(st);
  DFN_LAUNCH_CHECK();
  return 0;
}

}  // namespace dfn
