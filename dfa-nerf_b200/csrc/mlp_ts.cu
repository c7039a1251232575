// tcgen05 path, second generation: activations live in TENSOR MEMORY.
//
// The first kernel (mlp_tc.cu) keeps activations in shared memory; ncu and the in-kernel timeline
// showed it bound by shared-memory bandwidth (A + B operand reads, TMA weight writes and epilogue
// stores all cross the same 128 B/clk port) and by a 4-stage weight ring.  Here
//   * the A operand of every 256-wide layer is read from TMEM (tcgen05.mma with [a_tmem]): the
//     epilogue writes relu(acc + bias) back as packed bf16x2 with tcgen05.st, so activations never
//     touch shared memory; only the positional-encoding block (K = 64) is a shared-memory operand;
//   * shared memory is therefore free for an 11-stage (bf16) / 9-stage (bf16x3) weight ring;
//   * a layer is issued as two N = 128 halves with their own commit barriers, and the activation
//     buffer is double-buffered in TMEM (ACC 256 + A0 128 + A1 128 = 512 columns), so the epilogue
//     of half 0 runs under the MMAs of half 1, and the next layer's first K-blocks start while the
//     epilogue of half 1 is still draining (dependencies tracked per K-block with mbarriers);
//   * the positional encoding of the next tile is produced by four dedicated warps into a second
//     PE buffer while the current tile is in the MLP.
// bf16x3 keeps A_hi and A_lo in TMEM (in place, no half overlap: the tensor pipe is 3x longer busy).
//
// Warp roles (512 threads): 0 weight producer (TMA), 1 MMA issuer, 2 TMEM allocator, 4-11 epilogue
// (lane quarter = warp % 4, 64-column share = (warp-4)/4), 12-15 positional encoding.
// Reference arithmetic: HELP:21-52, HELP:275-299, HELP:372-396 (see mlp_tc.cu header).
#include <string.h>

#include <vector>

#include "common.cuh"
#include "model.h"
#include "tc_ptx.cuh"
#include "pe.cuh"

namespace dfn {
namespace ts {

using namespace dfn::tc;

static constexpr int STAGE_BYTES = 128 * 128;
static constexpr int PE_BYTES = TILE_M * 128;
static constexpr int N_BAR_BYTES = 512;
static constexpr int BIAS_BYTES = TC_MAX_LAYERS * TC_BIAS_STRIDE * 4;

template <bool X3>
struct Cfg {
  static constexpr int PE_PLANES = X3 ? 2 : 1;
  static constexpr int SMEM_PE = 0;                                  // [2 buffers][planes][16 KB]
  static constexpr int SMEM_RING = 2 * PE_PLANES * PE_BYTES;
  static constexpr int N_STAGES = (227 * 1024 - SMEM_RING - BIAS_BYTES - N_BAR_BYTES) / STAGE_BYTES;
  static constexpr int SMEM_BIAS = SMEM_RING + N_STAGES * STAGE_BYTES;
  static constexpr int SMEM_BAR = SMEM_BIAS + BIAS_BYTES;
  static constexpr int SMEM_TOTAL = SMEM_BAR + N_BAR_BYTES;
  static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget");
  static_assert(N_STAGES >= 4 && N_STAGES <= 16, "ring depth");
};

static constexpr uint32_t TM_ACC = 0, TM_A0 = 256, TM_A1 = 384;  // TMEM column map

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
  int view_w;
  int last_pe_layer;  // last layer whose K-blocks include the PE block
  unsigned long long* trace;  // debug timeline of CTA 0 (see dfn_debug_trace)
  int trace_tiles;
  TcLayer layers[TC_MAX_LAYERS];  // woff = offsets into the half-major blobs
};

__device__ __forceinline__ int n_halves(const TcLayer& L) { return L.n > 128 ? 2 : 1; }
__device__ __forceinline__ int half_rows(const TcLayer& L, int h) { return h == 0 ? min((int)L.n, 128) : (int)L.n - 128; }

template <bool X3>
__global__ void __launch_bounds__(512, 1) mlp_ts_kernel(const __grid_constant__ Params P) {
  using C = Cfg<X3>;
  constexpr int NS = C::N_STAGES;
  constexpr int NPART = X3 ? 2 : 1;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t bar0 = sbase + C::SMEM_BAR;
  const uint32_t bar_full = bar0;                 // [NS]
  const uint32_t bar_empty = bar0 + 8 * 16;       // [NS]
  const uint32_t bar_acc = bar0 + 8 * 32;         // [2]  accumulator half h complete
  const uint32_t bar_akb = bar0 + 8 * 34;         // [4]  activation K-block kb written (and its acc columns drained)
  const uint32_t bar_pe_ready = bar0 + 8 * 38;    // [2]
  const uint32_t bar_pe_free = bar0 + 8 * 40;     // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + C::SMEM_BAR + 8 * 42);
  float* bias_s = reinterpret_cast<float*>(smem + C::SMEM_BIAS);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(bar_acc + 8 * h, 1);
      mbar_init(bar_pe_ready + 8 * h, TILE_M);
      mbar_init(bar_pe_free + 8 * h, 1);
    }
    for (int k = 0; k < 4; ++k) mbar_init(bar_akb + 8 * k, TILE_M);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
  // all layers' (frame-constant) biases stay resident in shared memory
  for (int i = threadIdx.x; i < P.n_layers * TC_BIAS_STRIDE; i += blockDim.x) bias_s[i] = P.bias[i];
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int G = gridDim.x;
  const int n_local = (int)blockIdx.x < P.n_tiles ? (P.n_tiles - (int)blockIdx.x + G - 1) / G : 0;

  if (warp == 0) {
    // ============================== weight producer (TMA) ===============================
    {
      uint32_t cnt = 0;
      for (int i = 0; i < n_local; ++i) {
        for (int l = 0; l < P.n_layers; ++l) {
          const TcLayer& L = P.layers[l];
          uint32_t off = L.woff;
          for (int h = 0; h < n_halves(L); ++h) {
            const uint32_t bytes = (uint32_t)half_rows(L, h) * 128u;
            for (int kbi = 0; kbi < L.nkb; ++kbi) {
              for (int part = 0; part < NPART; ++part) {
                const uint32_t slot = cnt % NS, par = (cnt / NS) & 1u;
                mbar_wait(bar_empty + 8 * slot, par ^ 1u);
                if (elect_one_sync()) {
                  mbar_expect_tx(bar_full + 8 * slot, bytes);
                  tma_bulk_load(sbase + C::SMEM_RING + slot * STAGE_BYTES, (part == 0 ? P.w_hi : P.w_lo) + off, bytes,
                                bar_full + 8 * slot);
                }
                __syncwarp();
                ++cnt;
              }
              off += bytes;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================= MMA issuer =========================================
    {
      uint32_t cnt = 0;
      uint32_t akb_par[4] = {0u, 0u, 0u, 0u};
      for (int i = 0; i < n_local; ++i) {
        const int buf = i & 1;
        mbar_wait(bar_pe_ready + 8 * buf, (uint32_t)(i >> 1) & 1u);
        tcgen05_fence_after();
        const uint32_t pe_hi = sbase + C::SMEM_PE + (uint32_t)(buf * C::PE_PLANES) * PE_BYTES;
        for (int l = 0; l < P.n_layers; ++l) {
          const TcLayer& L = P.layers[l];
          // what the previous layer's epilogue produces (the last layer of the previous tile for l == 0)
          const bool has_prev = !(i == 0 && l == 0);
          const int prev_nkb_out = !has_prev ? 0 : 2 * n_halves(P.layers[l == 0 ? P.n_layers - 1 : l - 1]);
          bool waited[4] = {false, false, false, false};
          const bool tr = P.trace != nullptr && blockIdx.x == 0 && i < P.trace_tiles;
          long long t_need = 0, t_full = 0;
          auto need = [&](int kb) {
            if (kb < prev_nkb_out && !waited[kb]) {
              long long t0 = 0;
              if (tr) t0 = clock64();
              mbar_wait(bar_akb + 8 * kb, akb_par[kb]);
              akb_par[kb] ^= 1u;
              waited[kb] = true;
              tcgen05_fence_after();
              if (tr) t_need += clock64() - t0;
            }
          };
          // activations of layer l: bf16 mode alternates TMEM buffers, bf16x3 keeps (hi, lo) in place
          const uint32_t a_in = tmem_base + (X3 ? TM_A0 : (((l - 1) & 1) ? TM_A1 : TM_A0));
          const uint32_t a_in_lo = tmem_base + TM_A1;
          for (int h = 0; h < n_halves(L); ++h) {
            const uint32_t nn = (uint32_t)half_rows(L, h);
            const uint32_t idesc = make_idesc(nn);
            const uint32_t d = tmem_base + TM_ACC + 128u * (uint32_t)h;
            long long t_h0 = 0;
            if (tr) {
              t_h0 = clock64();
              t_need = 0;
              t_full = 0;
            }
            if (h == 0) {  // accumulator columns [0,128) were drained by the writers of K-blocks 0 and 1
              need(0);
              need(1);
              if (X3) {    // in-place activations: everything of the previous layer must be consumed
                need(2);
                need(3);
              }
            } else {
              need(2);
              need(3);
            }
            for (int kbi = 0; kbi < L.nkb; ++kbi) {
              const int kb = L.kb[kbi];
              if (kb != TC_KB_PE) need(kb);
              for (int part = 0; part < NPART; ++part) {
                const uint32_t slot = cnt % NS, par = (cnt / NS) & 1u;
                long long t_f0 = 0;
                if (tr) t_f0 = clock64();
                mbar_wait(bar_full + 8 * slot, par);
                tcgen05_fence_after();
                if (tr) t_full += clock64() - t_f0;
                const uint64_t bdesc = make_smem_desc(sbase + C::SMEM_RING + slot * STAGE_BYTES);
                if (kb == TC_KB_PE) {
                  const uint64_t adesc = make_smem_desc(pe_hi);
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    umma_bf16(d, adesc + 2 * ks, bdesc + 2 * ks, idesc, (kbi | part | ks) != 0 ? 1u : 0u);
                  if (X3 && part == 0) {
                    const uint64_t adesc_lo = make_smem_desc(pe_hi + PE_BYTES);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_bf16(d, adesc_lo + 2 * ks, bdesc + 2 * ks, idesc, 1u);
                  }
                } else {
                  const uint32_t a0 = a_in + (uint32_t)kb * 32u;
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    umma_bf16_ts(d, a0 + 8 * ks, bdesc + 2 * ks, idesc, (kbi | part | ks) != 0 ? 1u : 0u);
                  if (X3 && part == 0) {
                    const uint32_t a1 = a_in_lo + (uint32_t)kb * 32u;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_bf16_ts(d, a1 + 8 * ks, bdesc + 2 * ks, idesc, 1u);
                  }
                }
                umma_commit(bar_empty + 8 * slot);
                ++cnt;
              }
            }
            umma_commit(bar_acc + 8 * h);
            if (tr && lane == 0) {
              unsigned long long* r = P.trace + ((size_t)(i * P.n_layers + l) * 2 + h) * 4;
              r[0] = (unsigned long long)t_h0;
              r[1] = (unsigned long long)t_need;
              r[2] = (unsigned long long)clock64();
              r[3] = (unsigned long long)t_full;
            }
          }
          // keep the K-block barrier phases in step even when this layer did not read them all
          for (int kb = 0; kb < 4; ++kb) need(kb);
          if (l == P.last_pe_layer) umma_commit(bar_pe_free + 8 * buf);
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ================================ epilogue warps =======================================
    const int q = warp & 3;                 // TMEM lane quarter
    const int cs = (warp - 4) >> 2;         // which 64 columns of a 128-column half
    const uint32_t row = (uint32_t)(q * 32 + lane);
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t acc_par[2] = {0u, 0u};

    for (int i = 0; i < n_local; ++i) {
      const int tile = (int)blockIdx.x + i * G;
      int64_t pt = (int64_t)tile * TILE_M + row;
      const bool valid = pt < P.n_points;
      if (!valid) pt = P.n_points - 1;
      const int64_t ray = pt / P.S;
      float alpha = 0.f;

      for (int l = 0; l < P.n_layers; ++l) {
        const TcLayer& L = P.layers[l];
        const float* bl = bias_s + l * TC_BIAS_STRIDE;
        const uint32_t a_out = lane_base + (X3 ? TM_A0 : ((l & 1) ? TM_A1 : TM_A0));
        const uint32_t a_out_lo = lane_base + TM_A1;
        const int nh = n_halves(L);
        if (X3) {  // in-place activations: wait until every MMA of the layer has read them
          for (int h = 0; h < nh; ++h) {
            mbar_wait(bar_acc + 8 * h, acc_par[h]);
            acc_par[h] ^= 1u;
          }
          tcgen05_fence_after();
        }
        for (int h = 0; h < nh; ++h) {
          const bool tr = P.trace != nullptr && blockIdx.x == 0 && row == 0 && cs == 0 && i < P.trace_tiles;
          long long t_e0 = 0, t_e1 = 0;
          if (tr) t_e0 = clock64();
          if (!X3) {
            mbar_wait(bar_acc + 8 * h, acc_par[h]);
            acc_par[h] ^= 1u;
            tcgen05_fence_after();
          }
          if (tr) t_e1 = clock64();
          const uint32_t acc = lane_base + TM_ACC + 128u * (uint32_t)h + 64u * (uint32_t)cs;
          const int col0 = 128 * h + 64 * cs;   // first output column handled by this thread
          if (L.epi == TC_EPI_RGB) {
            if (cs == 0) {
              uint32_t v[16];
              tmem_ld16(acc, v);
              tmem_ld_wait();
              if (valid) {
                float4 o;
                o.x = __uint_as_float(v[0]) + bl[0];
                o.y = __uint_as_float(v[1]) + bl[1];
                o.z = __uint_as_float(v[2]) + bl[2];
                o.w = alpha;
                reinterpret_cast<float4*>(P.raw)[pt] = o;
              }
            }
          } else if (L.epi == TC_EPI_VIEW0 && h == 1) {
            if (cs == 0) {  // density head: accumulator column view_w, no activation
              uint32_t v[16];
              tmem_ld16(acc, v);
              tmem_ld_wait();
              alpha = __uint_as_float(v[0]) + bl[P.view_w];
            }
          } else {
            // 64 columns -> one K-block (32 packed TMEM columns) of the next layer's activations
            uint32_t v0[32], v1[32];
            tmem_ld32(acc, v0);
            tmem_ld32(acc + 32, v1);
            const bool per_ray = L.epi == TC_EPI_VIEW0;
            const float* gb = P.view_bias + ray * P.view_w + col0;
            const uint32_t kb_out = (uint32_t)(2 * h + cs);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 2; ++j) {       // 32 columns -> 16 packed TMEM columns
              uint32_t o[16], ol[16];
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float b[8];
                if (per_ray) ldg_f32x8(gb + j * 32 + g * 8, b);
                else {
#pragma unroll
                  for (int e = 0; e < 8; ++e) b[e] = bl[col0 + j * 32 + g * 8 + e];
                }
#pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) {
                  const int c = g * 8 + e2 * 2;
                  const uint32_t x0 = j == 0 ? v0[c] : v1[c];
                  const uint32_t x1 = j == 0 ? v0[c + 1] : v1[c + 1];
                  if (!X3) {
                    o[c >> 1] = add_relu_pack(x0, x1, b[e2 * 2], b[e2 * 2 + 1]);
                  } else {
                    const float f0 = fmaxf(__uint_as_float(x0) + b[e2 * 2], 0.f);
                    const float f1 = fmaxf(__uint_as_float(x1) + b[e2 * 2 + 1], 0.f);
                    const uint32_t hp = pack_bf16(f0, f1);
                    o[c >> 1] = hp;
                    ol[c >> 1] = pack_bf16(f0 - bf16_lo_f(hp), f1 - bf16_hi_f(hp));
                  }
                }
              }
              tmem_st16(a_out + kb_out * 32u + (uint32_t)j * 16u, o);
              if (X3) tmem_st16(a_out_lo + kb_out * 32u + (uint32_t)j * 16u, ol);
            }
            tmem_st_wait();
          }
          tcgen05_fence_before();
          mbar_arrive(bar_akb + 8 * (2 * h + cs));
          if (tr) {
            unsigned long long* r = P.trace + (size_t)P.trace_tiles * P.n_layers * 8 + ((size_t)(i * P.n_layers + l) * 2 + h) * 4;
            r[0] = (unsigned long long)t_e0;
            r[1] = (unsigned long long)t_e1;
            r[2] = (unsigned long long)clock64();
            r[3] = 0;
          }
        }
      }
    }
  } else if (warp >= 12) {
    // ============================ positional-encoding warps ===================================
    const uint32_t row = (uint32_t)((warp - 12) * 32 + lane);
    for (int i = 0; i < n_local; ++i) {
      const int buf = i & 1;
      mbar_wait(bar_pe_free + 8 * buf, ((uint32_t)(i >> 1) & 1u) ^ 1u);
      const int tile = (int)blockIdx.x + i * G;
      int64_t pt = (int64_t)tile * TILE_M + row;
      if (pt >= P.n_points) pt = P.n_points - 1;
      const int64_t ray = pt / P.S;
      float pe[64], x[3];   // same evaluation as mlp_tc.cu (pe.cuh)
      sample_point(P.rays_o, P.rays_d, ray, P.z_vals[pt], x);
      pe_embedder(x, P.multires, pe);
      uint8_t* pe_hi = smem + C::SMEM_PE + (size_t)(buf * C::PE_PLANES) * PE_BYTES;
      uint8_t* pe_lo = pe_hi + PE_BYTES;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint4 h;
        h.x = pack_bf16(pe[ch * 8 + 0], pe[ch * 8 + 1]);
        h.y = pack_bf16(pe[ch * 8 + 2], pe[ch * 8 + 3]);
        h.z = pack_bf16(pe[ch * 8 + 4], pe[ch * 8 + 5]);
        h.w = pack_bf16(pe[ch * 8 + 6], pe[ch * 8 + 7]);
        *reinterpret_cast<uint4*>(pe_hi + swz(row, (uint32_t)ch)) = h;
        if (X3) {
          uint4 l;
          l.x = pack_bf16(pe[ch * 8 + 0] - bf16_lo_f(h.x), pe[ch * 8 + 1] - bf16_hi_f(h.x));
          l.y = pack_bf16(pe[ch * 8 + 2] - bf16_lo_f(h.y), pe[ch * 8 + 3] - bf16_hi_f(h.y));
          l.z = pack_bf16(pe[ch * 8 + 4] - bf16_lo_f(h.z), pe[ch * 8 + 5] - bf16_hi_f(h.z));
          l.w = pack_bf16(pe[ch * 8 + 6] - bf16_lo_f(h.w), pe[ch * 8 + 7] - bf16_hi_f(h.w));
          *reinterpret_cast<uint4*>(pe_lo + swz(row, (uint32_t)ch)) = l;
        }
      }
      fence_proxy_async();
      mbar_arrive(bar_pe_ready + 8 * buf);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace ts

// ---------------------------------------------------------------------------- host side
// Half-major blobs for the TMEM-activation kernel: for layer, for 128-row half, for K-block:
// a [rows x 64] bf16 image in the 128-byte-swizzled K-major layout (same image format as mlp_tc.cu).
int ts_pack_from_tc(dfn_model* m, const std::vector<uint8_t>& hi, const std::vector<uint8_t>& lo, cudaStream_t st) {
  const TcProgram& pg = m->prog;
  std::vector<uint8_t> thi, tlo;
  thi.reserve(hi.size());
  tlo.reserve(lo.size());
  for (int l = 0; l < pg.n_layers; ++l) {
    const TcLayer& L = pg.layers[l];
    m->ts_woff[l] = (uint32_t)thi.size();
    const int n = L.n;
    // source order (mlp_tc.cu): for kbi: for chunk c0 (128 rows max)
    auto src_off = [&](int kbi, int half) {
      size_t o = L.woff;
      const size_t per_kb = (size_t)n * 128;
      o += (size_t)kbi * per_kb + (half == 0 ? 0 : (size_t)128 * 128);
      return o;
    };
    for (int h = 0; h < (n > 128 ? 2 : 1); ++h) {
      const size_t bytes = (size_t)(h == 0 ? (n < 128 ? n : 128) : n - 128) * 128;
      for (int kbi = 0; kbi < L.nkb; ++kbi) {
        const size_t so = src_off(kbi, h);
        thi.insert(thi.end(), hi.begin() + so, hi.begin() + so + bytes);
        tlo.insert(tlo.end(), lo.begin() + so, lo.begin() + so + bytes);
      }
    }
  }
  DFN_CUDA(cudaMalloc(&m->ts_hi, thi.size()));
  DFN_CUDA(cudaMalloc(&m->ts_lo, tlo.size()));
  DFN_CUDA(cudaMemcpyAsync(m->ts_hi, thi.data(), thi.size(), cudaMemcpyHostToDevice, st));
  DFN_CUDA(cudaMemcpyAsync(m->ts_lo, tlo.data(), tlo.size(), cudaMemcpyHostToDevice, st));
  DFN_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int ts_launch(const dfn_model* m, const float* bias_ws, const float* vbias_ws, int64_t R, int S, const float* rays_o,
              const float* rays_d, const float* z_vals, float* raw, int precision, cudaStream_t st) {
  const dfn_model_desc& d = m->desc;
  ts::Params P;
  memset(&P, 0, sizeof(P));
  P.w_hi = m->ts_hi;
  P.w_lo = m->ts_lo;
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
  P.last_pe_layer = 0;
  tc_get_trace(reinterpret_cast<void**>(&P.trace), &P.trace_tiles);
  for (int i = 0; i < m->prog.n_layers; ++i) {
    P.layers[i] = m->prog.layers[i];
    P.layers[i].woff = m->ts_woff[i];
    for (int k = 0; k < P.layers[i].nkb; ++k)
      if (P.layers[i].kb[k] == TC_KB_PE) P.last_pe_layer = i;
  }
  const int grid = P.n_tiles < num_sms() ? P.n_tiles : num_sms();
  if (precision == DFN_PREC_BF16) {
    static bool attr_done = false;
    if (!attr_done) {
      DFN_CUDA(cudaFuncSetAttribute(ts::mlp_ts_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    ts::Cfg<false>::SMEM_TOTAL));
      attr_done = true;
    }
    ts::mlp_ts_kernel<false><<<grid, 512, ts::Cfg<false>::SMEM_TOTAL, st>>>(P);
  } else {
    static bool attr_done = false;
    if (!attr_done) {
      DFN_CUDA(cudaFuncSetAttribute(ts::mlp_ts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    ts::Cfg<true>::SMEM_TOTAL));
      attr_done = true;
    }
    ts::mlp_ts_kernel<true><<<grid, 512, ts::Cfg<true>::SMEM_TOTAL, st>>>(P);
  }
  return 0;
}

}  // namespace dfn
