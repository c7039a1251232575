// tcgen05 path, CTA-pair generation ("tc2"): the schedule of mlp_tc.cu (two 128-point tiles per CTA in flight,
// activations as 128-byte-swizzled K-major blocks in shared memory, accumulators in TMEM) with the MMAs issued as
// cta_group::2 by the leader CTA of a 2-CTA cluster: one instruction covers M = 256 points (128 per CTA) and takes
// HALF of the weight rows from each CTA's shared memory.
//
// Why: the 1-CTA kernels sit at their shared-memory roofline.  Per 256x256 layer-tile each SM
// moved A reads 64 KB + B reads 128 KB + TMA weight writes 128 KB + epilogue stores 64 KB = 384 KB (3,072 cycles at
// 128 B/clk) against 2,048 tensor-pipe cycles.  With the pair sharing B, each SM stores and reads only its half of
// every weight K-block: 64 + 64 + 64 + 64 = 256 KB = 2,048 cycles, and the same 64 KB ring holds four K-blocks
// instead of two.
//
// Roles per CTA (384 threads): warp 0 weight producer (cp.async.bulk of this CTA's half stage, local `full` barrier);
// warp 1: leader = MMA issuer, peer = relay (forwards its `full` completions to the leader's `pfull` barriers);
// warp 2 TMEM allocator; warps 4-11 cooperative epilogue (positional encoding, bias + ReLU + repack).
// Cross-CTA signalling: epilogue warps of both CTAs arrive on the LEADER's `aready` barrier (remote arrive with
// cluster-scope release); the leader's tcgen05.commit multicasts to both CTAs' `empty` and `acc` barriers.
//
// MEASURED (B200, 20,000 rays x 192 samples, bit-identical output to mlp_tc.cu): NOT faster, so this kernel is an
// experiment behind dfn_debug_set_impl(3), not the default.  What the timelines (dfn_debug_trace) showed:
//   * with a cluster-scope release on the relay's remote arrive the relay forwards one stage per ~800 cycles (the MMAs
//     need one per 512) -> 777 TFLOP/s; a CTA-scope release there (the relay only forwards a TMA completion it
//     observed) removes that: the issuer's weight waits drop to the polling minimum (2,200 cycles per layer-slot
//     against 2,700 in mlp_tc.cu, whose 64 KB ring holds only two K-blocks against the L2 latency);
//   * the step is then bound by the per-slot dependency chain MMA (2,200) -> accumulator readout + repack (3,400;
//     TMEM reads are 64 B/clk = 2,048 cycles per 128x256 fp32 tile) -> cross-CTA `aready` signal (~700: the peer's
//     epilogue needs a cluster-scope release): 985 TFLOP/s with per-slot epilogue warps, 807 with the cooperative
//     epilogue below (the readout does not get faster with eight warps -- it is TMEM-port bound -- and the two slots
//     serialise), against 1,078 for mlp_tc.cu.
// The pair halves shared-memory traffic as intended, but MMA, TMEM readout and the chain latency are co-limits at
// ~2,048 cycles each per tile-layer; the cross-CTA hop lengthens exactly the chain.  bf16 only.
// Reference arithmetic: HELP:21-52 (Embedder), HELP:275-299 (FaceNeRF.forward), HELP:372-396 (NeRF.forward).
#include <string.h>

#include "common.cuh"
#include "model.h"
#include "tc_ptx.cuh"
#include "tc_epi.cuh"
#include "pe.cuh"

namespace dfn {
namespace tc2 {

using namespace dfn::tc;

static constexpr int STAGE_BYTES = 128 * 128;      // one ring entry: this CTA's <=128 weight rows x 64 K (bf16)
static constexpr int N_STAGES = 4;
static constexpr int ARENA_BLOCKS = 2 * TC_KB_PER_TILE;
static constexpr int SMEM_RING = ARENA_BLOCKS * KB_BYTES;
static constexpr int SMEM_BIAS = SMEM_RING + N_STAGES * STAGE_BYTES;
static constexpr int SMEM_BAR = SMEM_BIAS + 2 * TC_BIAS_STRIDE * 4;
static constexpr int SMEM_TOTAL = SMEM_BAR + 256;
static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget");

struct Params {
  const uint8_t* w;        // cta-pair stage images (tc_pack.h: hi2)
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
  int view_w;
  unsigned long long* trace;  // debug: per-role clock64 records of CTA 0 (null in production), format of mlp_tc.cu
  int trace_tiles;
  TcLayer layers[TC_MAX_LAYERS];
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1) mlp_tc2_kernel(const __grid_constant__ Params P) {
  constexpr int NSLOT = 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t bar_full = sbase + SMEM_BAR;           // [4] this CTA's half of ring entry e has landed
  const uint32_t bar_empty = sbase + SMEM_BAR + 32;     // [4] ring entry e consumed by the pair's MMAs
  const uint32_t bar_pfull = sbase + SMEM_BAR + 64;     // [4] leader only: the PEER's half of entry e has landed
  const uint32_t bar_acc = sbase + SMEM_BAR + 96;       // [2] accumulator of slot s complete
  const uint32_t bar_aready = sbase + SMEM_BAR + 112;   // [2] leader only: both CTAs' activations of slot s written
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SMEM_BAR + 128);
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < N_STAGES; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
      mbar_init(bar_pfull + 8 * i, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_acc + 8 * s, 1);
      mbar_init(bar_aready + 8 * s, 16);  // eight epilogue warps of each CTA
    }
    fence_barrier_init();
  }
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers exist before the pair allocates TMEM / anything remote targets them
  if (warp == 2) tmem_alloc2(smem_u32(tmem_ptr_smem), 512);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // tile group g = (j*C + c)*NSLOT + s holds tiles 2g (leader) and 2g+1 (peer); a tile index past the end is computed
  // on a clamped point and not stored.
  const int C = gridDim.x >> 1, c = (int)blockIdx.x >> 1;
  const int n_groups = (P.n_tiles + 1) >> 1;
  const int n_local = c * NSLOT < n_groups ? (n_groups - c * NSLOT + C * NSLOT - 1) / (C * NSLOT) * NSLOT : 0;
  const int n_iter = n_local / NSLOT;
  auto group_of = [&](int j, int s) { return (j * C + c) * NSLOT + s; };

  if (warp == 0) {
    // ============================== weight producer: this CTA's half of every K-block ===================
    uint32_t cnt = 0;
    for (int j = 0; j < n_iter; ++j) {
      for (int l = 0; l < P.n_layers; ++l) {
        const TcLayer& L = P.layers[l];
        const uint32_t bytes = (uint32_t)L.n * 64u;   // n/2 rows x 128 bytes
        for (int s = 0; s < NSLOT; ++s) {
          if (group_of(j, s) >= n_groups) continue;
          const uint8_t* src = P.w + L.woff + crank * bytes;
          for (int kbi = 0; kbi < L.nkb; ++kbi) {
            const uint32_t e = cnt % N_STAGES, par = (cnt / N_STAGES) & 1u;
            mbar_wait(bar_empty + 8 * e, par ^ 1u);
            if (elect_one_sync()) {
              mbar_expect_tx(bar_full + 8 * e, bytes);
              tma_bulk_load(sbase + SMEM_RING + e * STAGE_BYTES, src, bytes, bar_full + 8 * e);
            }
            __syncwarp();
            src += 2u * bytes;
            ++cnt;
          }
        }
      }
    }
    // the leader's last commits still target this CTA's `empty` barriers: wait for the final release of every entry
    for (uint32_t k = 0; k < (uint32_t)N_STAGES && k < cnt; ++k) {
      const uint32_t u = cnt - 1u - k;
      mbar_wait(bar_empty + 8 * (u % N_STAGES), (u / N_STAGES) & 1u);
    }
  } else if (warp == 1 && !leader) {
    // ============================== relay (peer CTA): local `full` -> leader's `pfull` ==================
    uint32_t cnt = 0;
    for (int j = 0; j < n_iter; ++j) {
      for (int l = 0; l < P.n_layers; ++l) {
        const int nkb = P.layers[l].nkb;
        for (int s = 0; s < NSLOT; ++s) {
          if (group_of(j, s) >= n_groups) continue;
          for (int kbi = 0; kbi < nkb; ++kbi) {
            const uint32_t e = cnt % N_STAGES, par = (cnt / N_STAGES) & 1u;
            mbar_wait(bar_full + 8 * e, par);
            if (elect_one_sync()) mbar_arrive_remote_light(bar_pfull + 8 * e, 0u);
            __syncwarp();
            ++cnt;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================= MMA issuer (leader CTA) ==================================
    uint32_t cnt = 0;
    uint32_t apar[2] = {0u, 0u};
    for (int j = 0; j < n_iter; ++j) {
      for (int l = 0; l < P.n_layers; ++l) {
        const TcLayer& L = P.layers[l];
        const uint32_t idesc = make_idesc_m256(L.n);
        for (int s = 0; s < NSLOT; ++s) {
          if (group_of(j, s) >= n_groups) continue;
          const bool tr = P.trace != nullptr && blockIdx.x == 0 && j < P.trace_tiles;
          long long t_w0 = 0, t_w1 = 0, t_full = 0, t_pfull = 0;
          if (tr) t_w0 = clock64();
          mbar_wait_cluster(bar_aready + 8 * s, apar[s]);
          apar[s] ^= 1u;
          tcgen05_fence_after();
          if (tr) t_w1 = clock64();
          const uint32_t acc = tmem_base + (uint32_t)s * 256u;
          for (int kbi = 0; kbi < L.nkb; ++kbi) {
            const uint64_t adesc = make_smem_desc(sbase + (uint32_t)(s * TC_KB_PER_TILE + L.kb[kbi]) * KB_BYTES);
            const uint32_t e = cnt % N_STAGES, par = (cnt / N_STAGES) & 1u;
            long long t_f0 = 0;
            if (tr) t_f0 = clock64();
            mbar_wait(bar_full + 8 * e, par);
            long long t_f1 = 0;
            if (tr) t_f1 = clock64();
            mbar_wait_cluster(bar_pfull + 8 * e, par);
            tcgen05_fence_after();
            if (tr) {
              t_full += t_f1 - t_f0;
              t_pfull += clock64() - t_f1;
            }
            const uint64_t bdesc = make_smem_desc(sbase + SMEM_RING + e * STAGE_BYTES);
#pragma unroll
            for (int q = 0; q < 4; ++q)   // K = 16 per instruction: both operands advance 32 bytes inside the swizzle atom
              umma_bf16_2cta(acc, adesc + 2 * q, bdesc + 2 * q, idesc, (kbi | q) != 0 ? 1u : 0u);
            umma_commit2_mc(bar_empty + 8 * e, (uint16_t)3);
            ++cnt;
          }
          umma_commit2_mc(bar_acc + 8 * s, (uint16_t)3);
          if (tr && lane == 0) {
            unsigned long long* r = P.trace + ((size_t)(j * P.n_layers + l) * 2 + s) * 4;
            r[0] = (unsigned long long)t_w0;
            r[1] = (unsigned long long)t_w1;
            r[2] = (unsigned long long)clock64();
            r[3] = (unsigned long long)t_full | ((unsigned long long)t_pfull << 32);
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue warps (cooperative) ================================
    // All eight warps drain whichever slot's accumulator completes next (MMA order: slot 0, slot 1 of a layer): two warps
    // per TMEM lane quarter, each thread owns half of its row's columns.  The two slots' epilogues are then never
    // concurrent (they would only share the 64 B/clk TMEM read port), each takes half as long, and it runs under the
    // other slot's MMAs.
    const int q = warp & 3;                 // TMEM lane quarter of this warp
    const int hf = (warp - 4) >> 2;         // which half of the columns
    const int et = (warp - 4) * 32 + lane;  // 0..255
    const uint32_t row = (uint32_t)(q * 32 + lane);
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float* bias_s = reinterpret_cast<float*>(smem + SMEM_BIAS);   // [2][256], double-buffered per layer
    uint32_t acc_par[2] = {0u, 0u};
    // this warp's writes to slot s are done and its accumulator reads have completed: tell the leader's MMA issuer
    auto signal_ready = [&](int s) {
      if (lane == 0) {
        if (leader) mbar_arrive(bar_aready + 8 * s);
        else mbar_arrive_remote(bar_aready + 8 * s, 0u);
      }
    };
    const int NL = P.n_layers;

    for (int j = 0; j < n_iter; ++j) {
      int64_t pt[2], ray[2];
      bool valid[2], live[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        live[s] = group_of(j, s) < n_groups;
        const int tile = 2 * group_of(j, s) + (int)crank;
        pt[s] = (int64_t)tile * TILE_M + row;
        valid[s] = live[s] && pt[s] < P.n_points;
        if (pt[s] >= P.n_points) pt[s] = P.n_points - 1;
        ray[s] = pt[s] / P.S;
      }
      const bool tr = P.trace != nullptr && blockIdx.x == 0 && et == 0 && j < P.trace_tiles;
      const long long t_tile0 = clock64();
      // first layer's bias into buffer (j*NL) & 1
      bias_s[((j * NL) & 1) * TC_BIAS_STRIDE + et] = P.bias[et];
      // ---- positional encoding (HELP:42-52): warps 4-7 encode slot 0's rows, warps 8-11 slot 1's ----
      if (live[hf]) {
        float pe[64], x[3];
        sample_point(P.rays_o, P.rays_d, ray[hf], P.z_vals[pt[hf]], x);
        pe_embedder(x, P.multires, pe);
        uint8_t* pe_hi = smem + (size_t)(hf * TC_KB_PER_TILE + TC_KB_PE) * KB_BYTES;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = pe[ch * 8 + e];
          store_chunk<false>(pe_hi, pe_hi, row, (uint32_t)ch, o);
        }
      }
      tcgen05_fence_before();
      fence_proxy_async_all();
      named_bar_sync(1, 256);      // both slots' PE blocks and the first bias are in shared memory
      if (live[0]) signal_ready(0);
      if (live[1]) signal_ready(1);

      float alpha[2] = {0.f, 0.f};
      for (int l = 0; l < NL; ++l) {
        const TcLayer& L = P.layers[l];
        const int gl = j * NL + l;
        const float* bl = bias_s + (gl & 1) * TC_BIAS_STRIDE;
        const uint32_t sbias = smem_u32(bl);
        // everyone is done with the other bias buffer (layer gl-1): refill it for layer gl+1
        if (l > 0) named_bar_sync(1, 256);
        if (l + 1 < NL) bias_s[((gl + 1) & 1) * TC_BIAS_STRIDE + et] = P.bias[(l + 1) * TC_BIAS_STRIDE + et];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (!live[s]) continue;
          uint8_t* arena = smem + (size_t)s * TC_KB_PER_TILE * KB_BYTES;
          const uint32_t acc = lane_base + (uint32_t)s * 256u;
          long long t_e0 = 0, t_e1 = 0;
          if (tr) t_e0 = clock64();
          mbar_wait(bar_acc + 8 * s, acc_par[s]);
          acc_par[s] ^= 1u;
          tcgen05_fence_after();
          if (tr) t_e1 = clock64();
          if (L.epi == TC_EPI_RGB) {
            if (hf == 0) {
              uint32_t v[16];
              tmem_ld16(acc, v);
              tmem_ld_wait();
              if (valid[s]) {
                float4 o;
                o.x = __uint_as_float(v[0]) + bl[0];
                o.y = __uint_as_float(v[1]) + bl[1];
                o.z = __uint_as_float(v[2]) + bl[2];
                o.w = alpha[s];
                reinterpret_cast<float4*>(P.raw)[pt[s]] = o;
              }
            }
            tcgen05_fence_before();
          } else {
            const bool per_ray = L.epi == TC_EPI_VIEW0;
            const int nch = (per_ray ? P.view_w : (int)L.n) >> 6;   // 32-column chunks per thread: 4 or 2
            const int ch0 = hf * nch;
            const float* rb = P.view_bias + ray[s] * P.view_w;
            uint32_t v0[32], v1[32];
            tmem_ld32(acc + ch0 * 32, v0);
            for (int cc = 0; cc < nch; cc += 2) {
              tmem_ld_wait();
              tmem_ld32(acc + (ch0 + cc + 1) * 32, v1);
              if (per_ray) epilogue_chunk<false, true>(v0, ch0 + cc, rb, 0u, arena, arena, row);
              else epilogue_chunk<false, false>(v0, ch0 + cc, nullptr, sbias, arena, arena, row);
              tmem_ld_wait();
              if (cc + 2 < nch) tmem_ld32(acc + (ch0 + cc + 2) * 32, v0);
              if (per_ray) epilogue_chunk<false, true>(v1, ch0 + cc + 1, rb, 0u, arena, arena, row);
              else epilogue_chunk<false, false>(v1, ch0 + cc + 1, nullptr, sbias, arena, arena, row);
            }
            if (per_ray && hf == 0) {   // density head: accumulator column view_w, no activation
              uint32_t v[16];
              tmem_ld16(acc + P.view_w, v);
              tmem_ld_wait();
              alpha[s] = __uint_as_float(v[0]) + bl[P.view_w];
            }
            tcgen05_fence_before();
            fence_proxy_async_all();
            __syncwarp();
            signal_ready(s);
          }
          if (tr) {
            unsigned long long* r = P.trace + (size_t)P.trace_tiles * NL * 8 + ((size_t)(j * NL + l) * 2 + s) * 4;
            r[0] = (unsigned long long)t_e0;
            r[1] = (unsigned long long)t_e1;
            r[2] = (unsigned long long)clock64();
            r[3] = (unsigned long long)t_tile0;
          }
        }
      }
      named_bar_sync(1, 256);   // all reads of this tile's last bias buffer are done before the next tile restages
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // the pair's MMAs read both CTAs' shared memory and TMEM until the leader is done
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc2(tmem_base, 512);
}

}  // namespace tc2

int tc2_launch(const dfn_model* m, const float* bias_ws, const float* vbias_ws, int64_t R, int S, const float* rays_o,
               const float* rays_d, const float* z_vals, float* raw, int precision, cudaStream_t st) {
  if (precision != DFN_PREC_BF16 || m->tc2_hi == nullptr) {
    set_error("tc2_launch: the cta_group::2 kernel covers DFN_PREC_BF16 only");
    return DFN_E_UNSUPPORTED;
  }
  const dfn_model_desc& d = m->desc;
  tc2::Params P;
  memset(&P, 0, sizeof(P));
  P.w = m->tc2_hi;
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
  P.view_w = d.W / 2;
  tc_get_trace(reinterpret_cast<void**>(&P.trace), &P.trace_tiles);
  for (int i = 0; i < m->prog.n_layers; ++i) {
    P.layers[i] = m->prog.layers[i];
    P.layers[i].woff = m->tc2_woff[i];
  }
  int grid = P.n_tiles < num_sms() ? P.n_tiles : num_sms();
  grid = (grid + 1) & ~1;
  if (grid > num_sms()) grid = num_sms() & ~1;
  static bool attr_done = false;
  if (!attr_done) {
    DFN_CUDA(cudaFuncSetAttribute(tc2::mlp_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_TOTAL));
    attr_done = true;
  }
  tc2::mlp_tc2_kernel<<<grid, 384, tc2::SMEM_TOTAL, st>>>(P);
  return 0;
}

}  // namespace dfn
