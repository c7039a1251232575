// tcgen05 path of network_query_fn for the single-pass precisions (bf16 / fp16), CTA-pair generation: the schedule of
// mlp_tc.cu (two 128-point tiles per CTA in flight, activations as 128-byte-swizzled K-major blocks in shared memory,
// accumulators in TMEM, per-slot epilogue warps, positional encoding one tile ahead by helper warps) with the MMAs issued
// as cta_group::2 by the leader CTA of a 2-CTA cluster: one instruction covers M = 256 points (128 per CTA) and takes HALF
// of the weight rows from each CTA's shared memory.
//
// Why (profiles/microbench/): the 1-CTA kernel sits at its shared-memory bandwidth roofline.  Per
// 256x256 layer and 128-point tile an SM moves A reads 64 KB + B reads 128 KB + TMA weight writes 128 KB + epilogue stores
// 64 KB = 384 KB = 3,072 cycles at 128 B/clk against 2,048 tensor-pipe cycles, and the timeline shows exactly that (3,100
// cycles per tile-layer with the MMA issuer never waiting for activations).  With the pair sharing B each SM stores and reads
// only its half of every weight K-block: 64 + 64 + 64 + 64 = 256 KB = 2,048 cycles, and the 64 KB ring holds a whole layer.
// The first CTA-pair experiment (round 1, mlp_tc2.cu) lost to the 1-CTA kernel because its per-slot dependency chain
// (MMA -> accumulator readout + repack -> cross-CTA signal) was longer than the other slot's MMAs; what changed is the
// epilogue: the column-distributed readout (tc_epi.cuh, epilogue_relu_cd) takes half the time of the row-per-thread one,
// whose warp-broadcast bias loads were more than half of it.
//
// Roles per CTA (384 threads): warp 0 weight producer (cp.async.bulk of this CTA's half stage, local `full` barrier);
// warp 1: leader = MMA issuer, peer = relay (forwards its `full` completions to the leader's `pfull` barriers);
// warps 2-3 positional encoding of each slot's NEXT tile (warp 2 also allocates TMEM); warps 4-7 / 8-11 epilogue of
// tile slot 0 / 1.  Cross-CTA signalling: one lane per epilogue warp of both CTAs arrives on the LEADER's `aready`
// barrier (the peer's with a cluster-scope release); the leader's tcgen05.commit multicasts to both CTAs' `empty` and
// `acc` barriers.
//
// The Decoder HEAD field of the live model (DEC:277-349; program built in mlp_dec.cu) runs here too: its two staged inputs -- the
// positional encoding and, for the view layer, the view-direction encoding (TC_KB_DIR) -- take turns in the slot's fifth block (the helper
// warps write the second once the skip layer has released the first), the density is a 16-column layer (TC_EPI_SIGMA).
//
// Reference arithmetic: HELP:21-52 (Embedder), HELP:275-299 (FaceNeRF.forward), HELP:372-396 (NeRF.forward); output
// bit-identical to mlp_tc.cu (same operands, same fp32 bias add and rounding; the accumulation order inside an MMA is
// the hardware's in both).
#include <string.h>

#include "common.cuh"
#include "model.h"
#include "tc_ptx.cuh"
#include "tc_epi.cuh"
#include "pe.cuh"

namespace dfn {
namespace tcp {

using namespace dfn::tc;

static constexpr int ENT_BYTES = 2 * 128 * 128;    // one ring entry: two K-blocks of this CTA's <=128 weight rows x 64 K
static constexpr int N_ENT = 2;
static constexpr int PE_HELPERS = 64;
static constexpr int ARENA_BLOCKS = 2 * TC_KB_PER_TILE;
static constexpr int SMEM_RING = ARENA_BLOCKS * KB_BYTES;
static constexpr int SMEM_BIAS = SMEM_RING + N_ENT * ENT_BYTES;
static constexpr int SMEM_BAR = SMEM_BIAS + 2 * TC_BIAS_STRIDE * 4;
static constexpr int SMEM_DOT = SMEM_BAR + 256;    // Decoder: the folded density head's row as packed 16-bit pairs (512 bytes)
static constexpr int SMEM_TOTAL = SMEM_DOT + 512;
static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget");

struct Params {
  const uint8_t* w;        // cta-pair stage images (tc_pack.h: hi2 / h16_2)
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
  int flags;                  // bit 0: a layer's weights stay in the ring for both slots; bit 1: CTA-scope release on the peer's `aready` arrivals;
                              // bit 2: per-ray bias rows staged in shared memory (column-distributed readout of the view layer)
  int dec;                    // Decoder head program (DEC:277-349): the encoding of DEC:257-275, sigma_out as a 16-column layer, sigmoid colours
  int multires_views;         // Decoder: frequencies of the view-direction encoding
  // Staged inputs: the slot's fifth block is (re)filled n_fills times per tile by the helper warps -- fill k holds staged block
  // fill_kb[k] (TC_KB_PE: positional encoding, computed; TC_KB_IN1: the torso's deformed signal, copied from `scratch`; TC_KB_DIR:
  // view-direction encoding, computed), is first read by layer fill_use[k] and dead once the MMAs of layer fill_rel[k] have completed.
  int n_fills;
  int fill_kb[3], fill_use[3], fill_rel[3];
  const float* dot_w;         // Decoder: density head folded into the last block's epilogue (TC_F_DOT_SIGMA): row [256] + bias, or null
  uint8_t* scratch;           // torso field: [grid][2 slots] byte images of the deformed-signal block (TC_EPI_STAGE writes, fill IN1 reads)
  TcLayer layers[TC_MAX_LAYERS];
};

// kind::f16, M = 256 across the pair, D = f32, both operands K-major; bf16 or fp16 operands
template <bool F16>
__device__ __forceinline__ uint32_t make_idesc_pair(uint32_t n) {
  return (1u << 4) | (F16 ? 0u : ((1u << 7) | (1u << 10))) | ((n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// EW: epilogue warps per tile slot -- 4 (one per TMEM lane quarter, 384 threads) or 8 (two per quarter, each half of the columns; 640
// threads, 96 registers).
// DEC: the Decoder programs (staged fills beyond the encoding, TC_EPI_SIGMA / CONT / STAGE, the folded density head).  The FaceNeRF / NeRF
// instantiation compiles none of that: 3.4 k instead of 5.6 k instructions, which is worth 9 % of the frame (measured on one box: 52.0 ms
// with the Decoder paths merged into the one instantiation, 47.3 ms apart -- the five roles' code then stays in the instruction cache).
template <bool F16, int EW, int MODE>   // MODE 0: FaceNeRF / NeRF; 1: Decoder head; 2: Decoder torso (deformation field in front)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((4 + 2 * EW) * 32, 1) mlp_pair_kernel(const __grid_constant__ Params P) {
  constexpr bool DEC = MODE != 0, TORSO = MODE == 2;
  constexpr int NSLOT = 2;
  constexpr int NH = EW / 4;          // column halves per row
  constexpr int ETH = EW * 32;        // epilogue threads per slot
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t bar_full = sbase + SMEM_BAR;           // [2] ring entry e has landed (leader: in both CTAs; peer: its own half)
  const uint32_t bar_empty = sbase + SMEM_BAR + 32;     // [2] ring entry e consumed by the pair's MMAs
  const uint32_t bar_acc = sbase + SMEM_BAR + 96;       // [2] accumulator of slot s complete
  const uint32_t bar_aready = sbase + SMEM_BAR + 112;   // [2] leader only: both CTAs' activations of slot s written
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SMEM_BAR + 128);
  const uint32_t bar_pefree = sbase + SMEM_BAR + 144;   // [2] the PE block of slot s is no longer read
  const uint32_t bar_peready = sbase + SMEM_BAR + 160;  // [2] the PE block of slot s holds the next tile's encoding
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < N_ENT; ++i) {
      mbar_init(bar_full + 8 * i, leader ? 2 : 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_acc + 8 * s, 1);
      mbar_init(bar_aready + 8 * s, 2 * EW);  // the epilogue warps of both CTAs
      mbar_init(bar_pefree + 8 * s, ETH);
      mbar_init(bar_peready + 8 * s, PE_HELPERS);
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
  if (DEC && P.dot_w != nullptr) {   // (ordered before its first use by the barriers every tile passes)
    for (int i = threadIdx.x; i < 128; i += blockDim.x) {
      const float2 w2 = reinterpret_cast<const float2*>(P.dot_w)[i];
      reinterpret_cast<uint32_t*>(smem + SMEM_DOT)[i] = F16 ? pack_f16(w2.x, w2.y) : pack_bf16(w2.x, w2.y);
    }
    __syncthreads();
  }
  const int C = gridDim.x >> 1, c = (int)blockIdx.x >> 1;
  const int n_groups = (P.n_tiles + 1) >> 1;
  const int n_local = c * NSLOT < n_groups ? (n_groups - c * NSLOT + C * NSLOT - 1) / (C * NSLOT) * NSLOT : 0;
  const int n_iter = n_local / NSLOT;
  auto group_of = [&](int j, int s) { return (j * C + c) * NSLOT + s; };

  // Ring: two entries of two weight K-blocks each (2 x 32 KB) on one `full` / `empty` barrier pair per entry -- the MMA issuer is one
  // thread, an mbarrier poll costs it 130-220 cycles even when the phase is complete and issuing an MMA ~95, so with a poll per K-block
  // (plus one for the peer's half) it could not keep up with the tensor pipe (128 cycles per MMA); per entry of eight MMAs it can.
  // The leader's `full` barrier counts two arrivals: its own producer's expect_tx and the peer relay's forward of the peer's completion.
  // Layers of at most four K-blocks (all but the skip layer) are loaded ONCE per iteration and read by both slots' MMAs (flag bit 0).
  auto n_uses = [&](int live, int nkb) { return ((P.flags & 1) && live == 2 && nkb <= 2 * N_ENT) ? 1 : live; };
  if (warp == 0) {
    // ============================== weight producer: this CTA's half of every K-block ===================
    uint32_t cnt = 0;
    for (int j = 0; j < n_iter; ++j) {
      const int live = (group_of(j, 0) < n_groups ? 1 : 0) + (group_of(j, 1) < n_groups ? 1 : 0);
      for (int l = 0; l < P.n_layers; ++l) {
        const TcLayer& L = P.layers[l];
        const uint32_t bytes = (uint32_t)L.n * 64u;   // one K-block of this CTA: n/2 rows x 128 bytes
        const int uses = n_uses(live, L.nkb);
        for (int u = 0; u < uses; ++u) {
          const uint8_t* src = P.w + L.woff + crank * (uint32_t)L.nkb * bytes;
          for (int kb0 = 0; kb0 < L.nkb; kb0 += 2) {
            const uint32_t nb = L.nkb - kb0 >= 2 ? 2u : 1u;
            const uint32_t e = cnt % N_ENT, par = (cnt / N_ENT) & 1u;
            mbar_wait(bar_empty + 8 * e, par ^ 1u);
            if (elect_one_sync()) {
              mbar_expect_tx(bar_full + 8 * e, nb * bytes);
              tma_bulk_load(sbase + SMEM_RING + e * ENT_BYTES, src, nb * bytes, bar_full + 8 * e);
            }
            __syncwarp();
            src += nb * bytes;
            ++cnt;
          }
        }
      }
    }
    // the leader's last commits still target this CTA's `empty` barriers: wait for the final release of every entry
    for (uint32_t k = 0; k < (uint32_t)N_ENT && k < cnt; ++k) {
      const uint32_t u = cnt - 1u - k;
      mbar_wait(bar_empty + 8 * (u % N_ENT), (u / N_ENT) & 1u);
    }
  } else if (warp == 1 && !leader) {
    // ============================== relay (peer CTA): local `full` -> the leader's `full` ==================
    uint32_t cnt = 0;
    for (int j = 0; j < n_iter; ++j) {
      const int live = (group_of(j, 0) < n_groups ? 1 : 0) + (group_of(j, 1) < n_groups ? 1 : 0);
      for (int l = 0; l < P.n_layers; ++l) {
        const int nkb = P.layers[l].nkb;
        const int n_ent = n_uses(live, nkb) * ((nkb + 1) >> 1);
        for (int u = 0; u < n_ent; ++u) {
          const uint32_t e = cnt % N_ENT, par = (cnt / N_ENT) & 1u;
          mbar_wait(bar_full + 8 * e, par);
          if (elect_one_sync()) mbar_arrive_remote_light(bar_full + 8 * e, 0u);   // forwards a completion it observed: CTA-scope release
          __syncwarp();
          ++cnt;
        }
      }
    }
  } else if (warp == 1) {
    // ================================= MMA issuer (leader CTA) ==================================
    uint32_t cnt = 0;
    uint32_t apar[2] = {0u, 0u};
    for (int j = 0; j < n_iter; ++j) {
      const int live = (group_of(j, 0) < n_groups ? 1 : 0) + (group_of(j, 1) < n_groups ? 1 : 0);
      for (int l = 0; l < P.n_layers; ++l) {
        const TcLayer& L = P.layers[l];
        const uint32_t idesc = make_idesc_pair<F16>(L.n);
        const uint32_t bytes = (uint32_t)L.n * 64u;
        const bool reuse = n_uses(live, L.nkb) == 1 && live == 2;
        const uint32_t base = cnt;
        for (int s = 0; s < NSLOT; ++s) {
          if (group_of(j, s) >= n_groups) continue;
          const bool tr = P.trace != nullptr && blockIdx.x == 0 && j < P.trace_tiles;
          long long t_w0 = 0, t_w1 = 0, t_full = 0;
          if (tr) t_w0 = clock64();
          mbar_wait_cluster(bar_aready + 8 * s, apar[s]);
          apar[s] ^= 1u;
          tcgen05_fence_after();
          if (tr) t_w1 = clock64();
          const uint32_t acc = tmem_base + (uint32_t)s * 256u;
          const bool first_use = !reuse || s == 0, last_use = !reuse || s == 1;
          if (reuse) cnt = base;
          for (int kb0 = 0; kb0 < L.nkb; kb0 += 2) {
            const uint32_t e = cnt % N_ENT, par = (cnt / N_ENT) & 1u;
            if (first_use) {
              long long t_f0 = 0;
              if (tr) t_f0 = clock64();
              mbar_wait_cluster(bar_full + 8 * e, par);
              tcgen05_fence_after();
              if (tr) t_full += clock64() - t_f0;
            }
#pragma unroll 1   // (rolled: the issue loop is a fifth of the kernel's instructions when ptxas clones it per K-block, and one thread issues it)
            for (int b = 0; b < 2; ++b) {
              if (kb0 + b < L.nkb) {
                const uint32_t blk = L.kb[kb0 + b] >= TC_KB_PE ? (uint32_t)TC_KB_PE : (uint32_t)L.kb[kb0 + b];   // staged inputs share block 4
                const uint64_t adesc = make_smem_desc(sbase + ((uint32_t)s * TC_KB_PER_TILE + blk) * KB_BYTES);
                const uint64_t bdesc = make_smem_desc(sbase + SMEM_RING + e * ENT_BYTES + (uint32_t)b * bytes);
#pragma unroll
                for (int q = 0; q < 4; ++q)   // K = 16 per instruction: both operands advance 32 bytes inside the swizzle atom
                  umma_bf16_2cta(acc, adesc + 2 * q, bdesc + 2 * q, idesc, ((kb0 | b | q) != 0 || (TORSO && (L.flags & TC_F_ACCUM))) ? 1u : 0u);
              }
            }
            if (last_use) umma_commit2_mc(bar_empty + 8 * e, (uint16_t)3);
            ++cnt;
          }
          umma_commit2_mc(bar_acc + 8 * s, (uint16_t)3);
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
  } else if (warp >= 4) {
    // ================================ epilogue warps (EW per tile slot) =======================================
    const int ew = warp - 4;
    const int s = ew / EW;                            // tile slot
    const int hf = (ew % EW) >> 2;                    // which part of the columns (NH parts)
    const uint32_t row = (uint32_t)((warp & 3) * 32 + lane);
    const int tid_s = (ew % EW) * 32 + lane;          // 0..ETH-1 within the slot
    uint8_t* arena = smem + (size_t)s * TC_KB_PER_TILE * KB_BYTES;
    float* bias_s = reinterpret_cast<float*>(smem + SMEM_BIAS) + s * TC_BIAS_STRIDE;
    const uint32_t acc = tmem_base + (uint32_t)s * 256u + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc_par = 0u;
    const int nf = DEC ? P.n_fills : 1;   // fills of the slot's staged block per tile
    // fill k >= 1 must be in place before the layer that first reads it is handed to the MMA issuer
    auto wait_fill_for = [&](int next_layer, int j) {
      for (int k = 1; k < nf; ++k)
        if (P.fill_use[k] == next_layer) mbar_wait(bar_peready + 8 * s, (uint32_t)(j * nf + k) & 1u);
    };
    // this warp's writes to the slot are done and its accumulator reads have completed: tell the leader's MMA issuer
    auto signal_ready = [&]() {
      tcgen05_fence_before();
      fence_proxy_async();     // this CTA's tensor core reads this CTA's rows: the cross-CTA part is the barrier below
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(bar_aready + 8 * s);
        else if (P.flags & 2) mbar_arrive_remote_light(bar_aready + 8 * s, 0u);
        else mbar_arrive_remote(bar_aready + 8 * s, 0u);
      }
    };
    auto stage_bias = [&](const float* src) {   // 256 floats
      if (ETH == 256) bias_s[tid_s] = src[tid_s];
      else reinterpret_cast<float2*>(bias_s)[tid_s] = reinterpret_cast<const float2*>(src)[tid_s];
    };

    for (int j = 0; j < n_iter; ++j) {
      if (group_of(j, s) >= n_groups) break;
      const int tile = 2 * group_of(j, s) + (int)crank;
      const long long t_tile0 = clock64();
      int64_t pt = (int64_t)tile * TILE_M + row;
      const bool valid = pt < P.n_points;
      if (!valid) pt = P.n_points - 1;
      const int64_t ray = pt / P.S;

      // Per-ray-bias layer (views_linears.0, HELP:288-292): when the sample count is a multiple of 32 every warp's 32 rows lie on ONE ray,
      // and a 128-point tile touches at most two rays -- their two bias rows [view_w] fit the slot's 256-float bias staging area, so the
      // layer runs the column-distributed readout of every other layer (bias from shared memory, conflict-free) instead of the
      // row-per-thread one with 64 global bias loads per thread, which was the longest epilogue of the program (2.2-2.9k against 1.7k
      // cycles) on the chain of the program's short tail layers.  Same fp32 add, ReLU and rounding: bit-identical.
      // (S % 32 == 0 and S >= 64: at most two rays per tile; the tile's first / last ray follow from this row's own ray without
      // another division)
      const bool vfast = !DEC && (P.flags & 4) && (P.S & 31) == 0 && P.S >= 64 && 2 * P.view_w <= TC_BIAS_STRIDE;
      const int64_t vp0 = (int64_t)tile * TILE_M;
      const int64_t vp1 = vp0 + TILE_M - 1 < P.n_points - 1 ? vp0 + TILE_M - 1 : P.n_points - 1;
      const int64_t vray0 = ray - (ray * P.S > vp0 ? 1 : 0);
      const int64_t vray1 = ray + ((ray + 1) * P.S - 1 < vp1 ? 1 : 0);
      stage_bias(P.bias);   // the first layer's bias
      // the tile's positional encoding was written into the PE K-block by the helper warps (below), one tile ahead
      mbar_wait(bar_peready + 8 * s, (uint32_t)(j * nf) & 1u);
      signal_ready();
      named_bar_sync(1 + s, ETH);  // bias_s visible to the slot's warps

      float alpha = 0.f;
      for (int l = 0; l < P.n_layers; ++l) {
        const bool dot_next_rgb = DEC && P.dot_w != nullptr && NH == 2 && l + 1 < P.n_layers && P.layers[l + 1].epi == TC_EPI_RGB;
        const TcLayer& L = P.layers[l];
        const bool tr = P.trace != nullptr && blockIdx.x == 0 && tid_s == 0 && j < P.trace_tiles;
        long long t_e0 = 0, t_e1 = 0;
        if (tr) t_e0 = clock64();
        if (!DEC && !vfast && L.epi == TC_EPI_VIEW0) prefetch_row_l1(P.view_bias + ray * P.view_w + hf * (P.view_w / NH), P.view_w / NH);   // hidden behind the wait
        mbar_wait(bar_acc + 8 * s, acc_par);
        acc_par ^= 1u;
        tcgen05_fence_after();
        if (tr) t_e1 = clock64();
        for (int k = 0; k < nf; ++k)
          if (P.fill_rel[k] == l) mbar_arrive(bar_pefree + 8 * s);   // its MMAs were the last readers of this fill of the staged block

        if (L.epi == TC_EPI_RGB) {
          if (hf == 0) {
            uint32_t v[16];
            tmem_ld16(acc, v);
            tmem_ld_wait();
            if (valid) {
              float4 o;
              o.x = __uint_as_float(v[0]) + bias_s[0];
              o.y = __uint_as_float(v[1]) + bias_s[1];
              o.z = __uint_as_float(v[2]) + bias_s[2];
              if (DEC) {   // DEC:346-347
                o.x = __fdividef(1.f, 1.f + __expf(-o.x));
                o.y = __fdividef(1.f, 1.f + __expf(-o.y));
                o.z = __fdividef(1.f, 1.f + __expf(-o.z));
              }
              // folded density head: this thread's column part + the other part (left in the bias staging area by its thread) + bias
              o.w = DEC && P.dot_w != nullptr ? alpha + (NH == 2 ? bias_s[128 + row] : 0.f) + __ldg(P.dot_w + TC_BIAS_STRIDE) : alpha;
              reinterpret_cast<float4*>(P.raw)[pt] = o;
            }
          }
          tcgen05_fence_before();
        } else if (DEC && L.epi == TC_EPI_SIGMA) {
          // Decoder density head (DEC:329): column 0 + bias stays in a register until the last layer writes raw
          if (hf == 0) {
            uint32_t v[16];
            tmem_ld16(acc, v);
            tmem_ld_wait();
            alpha = __uint_as_float(v[0]) + bias_s[0];
          }
          wait_fill_for(l + 1, j);
          signal_ready();
        } else if (TORSO && L.epi == TC_EPI_CONT) {
          // partial sums stay in the accumulator (the next layer continues them, TC_F_ACCUM); only the hand-over
          tcgen05_fence_before();
          wait_fill_for(l + 1, j);
          signal_ready();
        } else if (TORSO && L.epi == TC_EPI_STAGE) {
          // deformation output (DEC:299), + bias, no activation: columns 0..63 = PE' replace the encoding in the staged block, columns 64..127
          // = signal' go to hidden block 3 (read by fc_in_torso) and, as a byte image of the block, to the scratch the IN1 fill copies back
          for (int b = hf * (2 / NH); b < (hf + 1) * (2 / NH); ++b) {
            uint8_t* dst = b == 0 ? arena + (size_t)TC_KB_PE * KB_BYTES : arena + (size_t)3 * KB_BYTES;
            uint8_t* gdst = b == 0 ? nullptr : P.scratch + ((size_t)blockIdx.x * 2 + s) * KB_BYTES;
            epilogue_stage_cd<F16>(acc + (uint32_t)b * 64u, smem_u32(bias_s) + (uint32_t)b * 256u, dst, gdst, (uint32_t)((warp & 3) * 32), (uint32_t)lane);
          }
          wait_fill_for(l + 1, j);
          signal_ready();
        } else {
          const bool vf = !DEC && vfast && L.epi == TC_EPI_VIEW0;
          if (!DEC && !vf && L.epi == TC_EPI_VIEW0) {
            const int per = P.view_w / NH;
            epilogue_relu_rows16<true, F16>(acc, hf * per, (hf + 1) * per, P.view_bias + ray * P.view_w, 0u, arena, row);
            if (hf == 0) {
              uint32_t v[16];
              tmem_ld16(acc + P.view_w, v);
              tmem_ld_wait();
              alpha = __uint_as_float(v[0]) + bias_s[P.view_w];
            }
          } else if (DEC && (L.flags & TC_F_DOT_SIGMA)) {
            // last trunk block + sigma_out (DEC:329): the density from the layer's fp32 activations, this warp's columns' part
            const int per = ((int)L.n >> 6) / NH;
            alpha = epilogue_relu_cd_dotpart<F16>(acc, hf * per, (hf + 1) * per, smem_u32(bias_s), sbase + SMEM_DOT, arena,
                                                  (uint32_t)((warp & 3) * 32), (uint32_t)lane);
          } else if (vf || (L.n % (64 * NH)) == 0) {
            // vf: bias_s = [ray vray0's row | ray vray1's row] (staged below instead of the layer's static bias); this warp's rows are one ray's
            const int per = ((vf ? P.view_w : (int)L.n) >> 6) / NH;
            const uint32_t sb = smem_u32(bias_s) + (vf && ray != vray0 ? (uint32_t)P.view_w * 4u : 0u);
            epilogue_relu_cd<F16, EW == 4>(acc, hf * per, (hf + 1) * per, sb, arena, (uint32_t)((warp & 3) * 32), (uint32_t)lane);
            if (vf && hf == 0) {
              uint32_t v[16];
              tmem_ld16(acc + P.view_w, v);
              tmem_ld_wait();
              alpha = __uint_as_float(v[0]) + __ldg(P.bias + l * TC_BIAS_STRIDE + P.view_w);
            }
          } else {
            const int per = (int)L.n / NH;
            epilogue_relu_rows16<false, F16>(acc, hf * per, (hf + 1) * per, nullptr, smem_u32(bias_s), arena, row);
          }
          wait_fill_for(l + 1, j);
          signal_ready();
        }
        if (tr) {
          unsigned long long* r = P.trace + (size_t)P.trace_tiles * P.n_layers * 8 + ((size_t)(j * P.n_layers + l) * 2 + s) * 4;
          r[0] = (unsigned long long)t_e0;
          r[1] = (unsigned long long)t_e1;
          r[2] = (unsigned long long)clock64();
          r[3] = (unsigned long long)t_tile0;
        }
        // swap in the next layer's bias
        if (l + 1 < P.n_layers) {
          named_bar_sync(1 + s, ETH);
          if (dot_next_rgb && hf == 1) bias_s[tid_s] = alpha;     // tid_s = 128 + row: the second column part of the row's density
          else if (!DEC && vfast && P.layers[l + 1].epi == TC_EPI_VIEW0) {
            for (int i = tid_s; i < 2 * P.view_w; i += ETH)
              bias_s[i] = __ldg(P.view_bias + (i < P.view_w ? vray0 : vray1) * P.view_w + (i < P.view_w ? i : i - P.view_w));
          } else stage_bias(P.bias + (l + 1) * TC_BIAS_STRIDE);
          named_bar_sync(1 + s, ETH);
        }
      }
      named_bar_sync(1 + s, ETH);  // everyone done with bias_s before the next tile restages it
    }
  }

  if (warp == 2 || warp == 3) {
    // ============================ positional-encoding helper warps ==============================
    // x = o + d*z -> [x | sin(2^k x) | cos(2^k x)] (HELP:42-52, pe.cuh) for the NEXT tile of each slot, written straight
    // into the slot's PE K-block as soon as the last layer that reads it (the skip layer) has finished its MMAs.
    const int t = (warp - 2) * 32 + lane;
    const int nf = DEC ? P.n_fills : 1;
    for (int j = 0; j < n_iter; ++j) {
      for (int k = 0; k < nf; ++k) {
        for (int s = 0; s < NSLOT; ++s) {
          if (group_of(j, s) >= n_groups) continue;
          const int fill = j * nf + k;         // fills of this slot's staged block so far
          if (fill > 0) mbar_wait(bar_pefree + 8 * s, (uint32_t)(fill - 1) & 1u);
          uint8_t* pe_blk = smem + (size_t)(s * TC_KB_PER_TILE + TC_KB_PE) * KB_BYTES;
          const int tile = 2 * group_of(j, s) + (int)crank;
          if (TORSO && P.fill_kb[k] == TC_KB_IN1) {
            // the deformed signal written by this tile's TC_EPI_STAGE epilogue: a byte image of the block
            const uint4* src = reinterpret_cast<const uint4*>(P.scratch + ((size_t)blockIdx.x * 2 + s) * KB_BYTES);
            uint4* dst = reinterpret_cast<uint4*>(pe_blk);
#pragma unroll 4
            for (int i = t; i < KB_BYTES / 16; i += PE_HELPERS) dst[i] = src[i];
          } else {
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
              const uint32_t row = (uint32_t)(t + 64 * h);
              int64_t pt = (int64_t)tile * TILE_M + row;
              if (pt >= P.n_points) pt = P.n_points - 1;
              const int64_t ray = pt / P.S;
              float pe[64], x[3];
              if (DEC) {   // one copy of the encoding loop for both staged encodings of the Decoder programs
                const bool dir = P.fill_kb[k] == TC_KB_DIR;
                if (dir) view_direction(P.rays_d, ray, x);
                else sample_point(P.rays_o, P.rays_d, ray, P.z_vals[pt], x);
                pe_decoder(x, dir ? P.multires_views : P.multires, pe);
              } else {
                sample_point(P.rays_o, P.rays_d, ray, P.z_vals[pt], x);
                pe_embedder(x, P.multires, pe);
              }
#pragma unroll
              for (int ch = 0; ch < 8; ++ch) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = pe[ch * 8 + e];
                store_chunk<false, F16>(pe_blk, pe_blk, row, (uint32_t)ch, o);
              }
            }
          }
          fence_proxy_async();
          mbar_arrive(bar_peready + 8 * s);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // the pair's MMAs read both CTAs' shared memory and TMEM until the leader is done
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc2(tmem_base, 512);
}

}  // namespace tcp

static int g_pair_ew = 8, g_pair_flags = 7;   // epilogue warps per slot (debug: dfn_debug_set_impl(3) -> 8, (8) -> 4)
void pair_set_epilogue_warps(int ew) { g_pair_ew = ew == 4 ? 4 : 8; }
void pair_set_flags(int flags) { g_pair_flags = flags; }
int pair_get_flags() { return g_pair_flags; }

// prog: layer program; woff2 / w: per-layer offsets into, and the blob of, the CTA-pair stage images (tc_pack.h); decoder: the head
// program of the live model (mlp_dec.cu).
int pair_launch_prog(const TcProgram& prog, const uint32_t* woff2, const uint8_t* w, bool f16, bool decoder, int multires, int multires_views,
                     int view_w, const float* bias_ws, const float* vbias_ws, const float* dot_w, void* scratch, int64_t R, int S,
                     const float* rays_o, const float* rays_d, const float* z_vals, float* raw, cudaStream_t st) {
  tcp::Params P;
  memset(&P, 0, sizeof(P));
  P.w = w;
  P.bias = bias_ws;
  P.view_bias = vbias_ws;
  P.rays_o = rays_o;
  P.rays_d = rays_d;
  P.z_vals = z_vals;
  P.raw = raw;
  P.n_points = R * S;
  P.S = S;
  P.n_tiles = (int)((P.n_points + tc::TILE_M - 1) / tc::TILE_M);
  P.n_layers = prog.n_layers;
  P.multires = multires;
  P.multires_views = multires_views;
  P.view_w = view_w;
  P.dec = decoder ? 1 : 0;
  P.scratch = reinterpret_cast<uint8_t*>(scratch);
  P.dot_w = nullptr;
  tc_get_trace(reinterpret_cast<void**>(&P.trace), &P.trace_tiles);
  P.flags = g_pair_flags;
  bool torso = false;   // the program uses what only the torso instantiation compiles: TC_EPI_CONT / TC_EPI_STAGE / TC_F_ACCUM / TC_KB_IN1
  int use[7], rel[7];
  for (int k = 0; k < 7; ++k) use[k] = rel[k] = -1;
  for (int i = 0; i < prog.n_layers; ++i) {
    P.layers[i] = prog.layers[i];
    P.layers[i].woff = woff2[i];
    const TcLayer& L = prog.layers[i];
    bool ok = L.epi == TC_EPI_RELU || L.epi == TC_EPI_RGB || (!decoder && L.epi == TC_EPI_VIEW0) ||
              (decoder && (L.epi == TC_EPI_SIGMA || L.epi == TC_EPI_CONT || (L.epi == TC_EPI_STAGE && L.n == 128 && scratch != nullptr)));
    if (L.flags & TC_F_DOT_SIGMA) {
      if (!decoder || dot_w == nullptr || L.epi != TC_EPI_RELU || L.n != 256 || P.dot_w != nullptr) ok = false;
      P.dot_w = dot_w;
    }
    if ((L.flags & ~(TC_F_ACCUM | TC_F_DOT_SIGMA)) != 0 || L.nkb > 5) ok = false;
    if (L.epi == TC_EPI_CONT || L.epi == TC_EPI_STAGE || (L.flags & TC_F_ACCUM)) torso = true;
    if ((L.flags & TC_F_ACCUM) && (i == 0 || prog.layers[i - 1].epi != TC_EPI_CONT)) ok = false;
    if (L.epi == TC_EPI_CONT && (i + 1 >= prog.n_layers || !(prog.layers[i + 1].flags & TC_F_ACCUM) || prog.layers[i + 1].n != L.n)) ok = false;
    for (int k = 0; k < L.nkb; ++k)
      if (L.kb[k] >= TC_KB_PE) {
        if (L.kb[k] > TC_KB_DIR || (!decoder && L.kb[k] != TC_KB_PE)) ok = false;
        else {
          if (L.kb[k] == TC_KB_IN1) torso = true;
          if (use[L.kb[k]] < 0) use[L.kb[k]] = i;
          rel[L.kb[k]] = i;
        }
      }
    if (!ok) {
      set_error("pair_launch_prog: layer %d of the program is outside what the CTA-pair kernel runs", i);
      return DFN_E_UNSUPPORTED;
    }
  }
  // fills of the staged block in the order of their first use; each must be dead before the next one is first read
  P.n_fills = 0;
  for (int round = 0; round < 3; ++round) {
    int best = -1;
    for (int kb = TC_KB_PE; kb <= TC_KB_DIR; ++kb)
      if (use[kb] >= 0 && (best < 0 || use[kb] < use[best])) best = kb;
    if (best < 0) break;
    P.fill_kb[P.n_fills] = best;
    P.fill_use[P.n_fills] = use[best];
    P.fill_rel[P.n_fills] = rel[best];
    ++P.n_fills;
    use[best] = -1;
  }
  bool fills_ok = P.n_fills >= 1 && P.fill_kb[0] == TC_KB_PE && P.fill_use[0] == 0;
  for (int k = 1; k < P.n_fills; ++k)
    if (P.fill_rel[k - 1] >= P.fill_use[k] || P.fill_use[k] < 1) fills_ok = false;
  if (!fills_ok) {
    set_error("pair_launch_prog: the program's staged inputs do not take turns in one block");
    return DFN_E_UNSUPPORTED;
  }
  int grid = P.n_tiles < num_sms() ? P.n_tiles : num_sms();
  grid = (grid + 1) & ~1;
  if (grid > num_sms()) grid = num_sms() & ~1;
  const int ew = g_pair_ew;
  auto launch = [&](auto kernel, int threads) -> int {
    DFN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tcp::SMEM_TOTAL));
    kernel<<<grid, threads, tcp::SMEM_TOTAL, st>>>(P);
    return 0;
  };
  if (decoder && torso) {
    if (ew == 8) return f16 ? launch(tcp::mlp_pair_kernel<true, 8, 2>, 640) : launch(tcp::mlp_pair_kernel<false, 8, 2>, 640);
    return f16 ? launch(tcp::mlp_pair_kernel<true, 4, 2>, 384) : launch(tcp::mlp_pair_kernel<false, 4, 2>, 384);
  }
  if (decoder) {
    if (ew == 8) return f16 ? launch(tcp::mlp_pair_kernel<true, 8, 1>, 640) : launch(tcp::mlp_pair_kernel<false, 8, 1>, 640);
    return f16 ? launch(tcp::mlp_pair_kernel<true, 4, 1>, 384) : launch(tcp::mlp_pair_kernel<false, 4, 1>, 384);
  }
  if (ew == 8) return f16 ? launch(tcp::mlp_pair_kernel<true, 8, 0>, 640) : launch(tcp::mlp_pair_kernel<false, 8, 0>, 640);
  return f16 ? launch(tcp::mlp_pair_kernel<true, 4, 0>, 384) : launch(tcp::mlp_pair_kernel<false, 4, 0>, 384);
}

int pair_launch(const dfn_model* m, const float* bias_ws, const float* vbias_ws, int64_t R, int S, const float* rays_o,
                const float* rays_d, const float* z_vals, float* raw, int precision, cudaStream_t st) {
  const bool f16 = precision == DFN_PREC_FP16;
  if ((precision != DFN_PREC_BF16 && !f16) || m->tc2_hi == nullptr || m->tc2_h16 == nullptr) {
    set_error("pair_launch: the cta_group::2 kernel covers DFN_PREC_BF16 / DFN_PREC_FP16");
    return DFN_E_UNSUPPORTED;
  }
  return pair_launch_prog(m->prog, m->tc2_woff, f16 ? m->tc2_h16 : m->tc2_hi, f16, false, m->desc.multires, m->desc.multires_views,
                          m->desc.W / 2, bias_ws, vbias_ws, nullptr, nullptr, R, S, rays_o, rays_d, z_vals, raw, st);
}

}  // namespace dfn
