// tcgen05 path, third generation ("pp"): ping-pong tiles with a cooperative epilogue.
//
// What the timelines of mlp_tc.cu showed (profiles/): with two 128-point tiles in flight the tensor pipe
// waited on (1) the epilogue, 3.2k cycles per 256-wide layer when only the four warps of one tile work on
// it, and (2) the single MMA-issuing warp's barrier round trips.  This kernel keeps the same data layout
// (activations as 128-byte-swizzled K-major blocks in shared memory, accumulators in TMEM, weights
// multicast to a 2-CTA cluster) and changes the schedule:
//   * ALL EIGHT epilogue warps drain whichever tile's accumulator is ready (two warps per TMEM lane
//     quarter, each thread owns half of the row's columns), so a layer's epilogue takes half as long and
//     two warps per scheduler hide each other's TMEM / shared-memory latency;
//   * one ring entry = one K-block of weights on one mbarrier (four MMAs per barrier round trip);
//   * the positional-encoding K-block no longer owns shared memory for the whole tile: four dedicated warps
//     compute PE rows one iteration ahead into an L2-resident scratch buffer, and the epilogue warps copy a
//     tile's PE block into a *ring entry* right before the two layers that consume it (layer 0 and the skip
//     layer).  The 32 KB this frees makes the weight ring three entries (96 KB) deep.
// bf16x3 (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo) runs one tile with hi/lo activation planes in the same arena.
//
// Warp roles (384 threads, 168 registers each): 0 weight producer (TMA multicast), 1 MMA issuer, 2-3 positional
// encoding (warp 2 also allocates TMEM), 4-11 epilogue.
// Reference arithmetic: HELP:21-52 (Embedder), HELP:275-299 (FaceNeRF.forward), HELP:372-396 (NeRF.forward).
//
// DEC = true instantiates the same schedule for the reference's LIVE model, Decoder + DeformationField_ori
// (DEC:77-349; layer programs built in mlp_dec.cu): the positional encoding of DEC:257-275, a density layer that
// keeps its result in a register (TC_EPI_SIGMA), a 256-wide per-ray view bias, a sigmoid on the colours, a second
// staged input block (the deformed per-sample signal of the torso field) whose layers are split in two
// accumulate-chained halves (TC_EPI_CONT / TC_F_ACCUM), and the deformation output written back to the tile's
// scratch blocks (TC_EPI_STAGE).
#include <string.h>

#include <vector>

#include "common.cuh"
#include "model.h"
#include "tc_ptx.cuh"
#include "tc_epi.cuh"
#include "pe.cuh"

namespace dfn {
namespace pp {

using namespace dfn::tc;

static constexpr int SLOT_BYTES = 16384;        // half a ring entry
static constexpr int N_ENTRIES = 3;             // ring entries (2 x 16 KB each)
static constexpr int ARENA_BLOCKS = 8;          // bf16: 2 tiles x 4 blocks; bf16x3: hi 4 + lo 4
static constexpr int SMEM_RING = ARENA_BLOCKS * KB_BYTES;
static constexpr int SMEM_BIAS = SMEM_RING + N_ENTRIES * 2 * SLOT_BYTES;
static constexpr int SMEM_BAR = SMEM_BIAS + 2 * TC_BIAS_STRIDE * 4;
static constexpr int SMEM_DOT = SMEM_BAR + 256;   // density-head row as packed 16-bit pairs (TC_F_DOT_SIGMA), 512 bytes
static constexpr int SMEM_TOTAL = SMEM_DOT + 512;
static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget");
static constexpr int EPI_THREADS = 256;
static constexpr int PE_THREADS = 64;   // warps 2-3: two rows per thread and slot

struct Params {
  const uint8_t* w_hi;
  const uint8_t* w_lo;
  const float* bias;       // [n_layers][256], latent already folded
  const float* view_bias;  // [R][W/2]
  const float* dot_w;      // Decoder density head folded into an epilogue (TC_F_DOT_SIGMA): row [256] + bias, or null
  const float* dot_b;      // TC_F_DOT_ALPHA: alpha_linear's bias (dot_w = its weight row, fp32)
  const float* rays_o;
  const float* rays_d;
  const float* z_vals;
  float* raw;
  uint8_t* pe_scratch;     // [grid][2 buffers][2 slots][1|2 staged blocks][128 rows][row_bytes]
  unsigned long long* trace;
  int trace_tiles;
  int64_t n_points;
  int S;
  int n_tiles;
  int n_layers;
  int multires;
  int multires_views;      // Decoder: frequencies of the view-direction encoding (staged block TC_KB_DIR)
  int view_w;
  int flags;               // bit 0: split schedule stages a layer's input block at the START of the previous layer's epilogue (early_staged); bit 1: weight-stage barrier polled first
  TcLayer layers[TC_MAX_LAYERS];  // woff: offsets into the K=32 / 64-byte-swizzle blobs
};

// staged input block of a layer (0: positional encoding, 1: deformed signal) or -1; at most one per layer
__device__ __forceinline__ int layer_staged(const TcLayer& L) {
  for (int k = 0; k < L.nkb; ++k)
    if (L.kb[k] >= TC_KB_PE) return (int)L.kb[k] - TC_KB_PE;
  return -1;
}
__device__ __forceinline__ bool layer_has_pe(const TcLayer& L) { return layer_staged(L) >= 0; }
// bf16x3 (one tile, cooperative epilogue): the epilogue signals every finished 64-column activation block -- blocks
// before the last on `kready[k]`, the last one (or an epilogue that writes none) on `aready` -- so the next layer's MMAs
// start on block 0 while blocks 1.. are still being drained from the OTHER accumulator buffer (one tile leaves TMEM
// columns 256..511 free for double buffering by layer).  One barrier per block: a single multi-phase barrier would
// run two phases ahead of its waiter and alias parities.  Number of signals of a layer's epilogue:
__device__ __forceinline__ int layer_phases(const TcLayer& L, int view_w) {
  if (L.epi == TC_EPI_RELU) return L.n >= 64 ? (int)L.n >> 6 : 1;
  if (L.epi == TC_EPI_VIEW0) return view_w >> 6;
  return 1;
}
// operand passes of a layer's weights through the ring: hi and lo stages in the split modes, unless the layer is single-pass
template <int NPART>
__device__ __forceinline__ int layer_parts(const TcLayer& L) {
  return (NPART == 2 && (L.flags & TC_F_SINGLE)) ? 1 : NPART;
}
template <int NPART>
__device__ __forceinline__ uint32_t layer_entries(const TcLayer& L) {
  return (uint32_t)L.nkb * (uint32_t)layer_parts<NPART>(L) + (layer_has_pe(L) ? 1u : 0u);
}

// One 64-column row of a staged block -> scratch, bf16 / fp16 (hi [, lo = bf16(v - hi)]) in the [16-byte chunk][row]
// layout (coalesced for these stores and for the epilogue warps' loads).
template <bool X3, bool F16>
__device__ __forceinline__ void scratch_row(uint8_t* blk_base, uint32_t row, const float (&v)[64]) {
  uint4* dst = reinterpret_cast<uint4*>(blk_base) + row;
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    uint4 h;
    if (F16) {
      h.x = pack_f16(v[ch * 8 + 0], v[ch * 8 + 1]);
      h.y = pack_f16(v[ch * 8 + 2], v[ch * 8 + 3]);
      h.z = pack_f16(v[ch * 8 + 4], v[ch * 8 + 5]);
      h.w = pack_f16(v[ch * 8 + 6], v[ch * 8 + 7]);
    } else {
      h.x = pack_bf16(v[ch * 8 + 0], v[ch * 8 + 1]);
      h.y = pack_bf16(v[ch * 8 + 2], v[ch * 8 + 3]);
      h.z = pack_bf16(v[ch * 8 + 4], v[ch * 8 + 5]);
      h.w = pack_bf16(v[ch * 8 + 6], v[ch * 8 + 7]);
    }
    dst[ch * TILE_M] = h;
    if (X3) {
      uint4 l;
      l.x = pack_lo<F16>(v[ch * 8 + 0], v[ch * 8 + 1], h.x);
      l.y = pack_lo<F16>(v[ch * 8 + 2], v[ch * 8 + 3], h.y);
      l.z = pack_lo<F16>(v[ch * 8 + 4], v[ch * 8 + 5], h.z);
      l.w = pack_lo<F16>(v[ch * 8 + 6], v[ch * 8 + 7], h.w);
      dst[(8 + ch) * TILE_M] = l;
    }
  }
}

// TC_EPI_STAGE: + bias (no activation), bf16 (hi[/lo]) of one 32-column chunk -> the tile's staged block in the
// scratch ([16-byte chunk][row] layout, as the PE warps write it).  blk_base = staged block (chunk32 >> 1).
template <bool X3, bool F16>
__device__ __forceinline__ void stage_chunk(const uint32_t (&v)[32], int chunk32, uint32_t sbias, uint8_t* blk_base,
                                            uint32_t row) {
  uint4* dst = reinterpret_cast<uint4*>(blk_base) + row;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float b[8], o[8];
    lds_f32x8(sbias + (uint32_t)(chunk32 * 32 + g * 8) * 4u, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[g * 8 + e]) + b[e];
    const int c16 = (chunk32 & 1) * 4 + g;
    uint4 h;
    if (F16) {
      h.x = pack_f16(o[0], o[1]);
      h.y = pack_f16(o[2], o[3]);
      h.z = pack_f16(o[4], o[5]);
      h.w = pack_f16(o[6], o[7]);
    } else {
      h.x = pack_bf16(o[0], o[1]);
      h.y = pack_bf16(o[2], o[3]);
      h.z = pack_bf16(o[4], o[5]);
      h.w = pack_bf16(o[6], o[7]);
    }
    dst[c16 * TILE_M] = h;
    if (X3) {
      uint4 l;
      l.x = pack_lo<F16>(o[0], o[1], h.x);
      l.y = pack_lo<F16>(o[2], o[3], h.y);
      l.z = pack_lo<F16>(o[4], o[5], h.z);
      l.w = pack_lo<F16>(o[6], o[7], h.w);
      dst[(8 + c16) * TILE_M] = l;
    }
  }
}

__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// X3 = false: bf16, two tile slots.  X3 = true: split-bf16, one slot.  DEC: Decoder programs (see the header).
// F16 (single-pass schedule only): fp16 operands -- P.w_hi then points at the fp16 weight stages (DFN_PREC_FP16).
template <bool X3, bool DEC, bool F16 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1) mlp_pp_kernel(const __grid_constant__ Params P) {
  constexpr int NSLOT = X3 ? 1 : 2;
  constexpr int NPART = X3 ? 2 : 1;
  constexpr int ROWB = X3 ? 256 : 128;  // scratch bytes per row of a staged block
  constexpr int NBLK = DEC ? 3 : 1;     // staged blocks per tile (Decoder: PE | deformed signal | view-direction PE)
  // bf16x3 (one tile): all eight epilogue warps share the tile's columns.  bf16 (two tiles): four warps per tile,
  // the two groups run concurrently (a cooperative, serialised epilogue of two tiles measured slower).
  constexpr bool COOP = X3;
  constexpr int EPI_GROUP = COOP ? EPI_THREADS : EPI_THREADS / 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t bar0 = sbase + SMEM_BAR;
  const uint32_t bar_full = bar0;                 // [3]  ring entry filled (TMA bytes, or the PE copy)
  const uint32_t bar_empty = bar0 + 8 * 4;        // [3]  ring entry consumed by both CTAs' MMAs
  const uint32_t bar_acc = bar0 + 8 * 8;          // [2]  accumulator of slot s complete
  const uint32_t bar_aready = bar0 + 8 * 10;      // [2]  activations of slot s written, accumulator drained
  const uint32_t bar_pe_ready = bar0 + 8 * 12;    // [2]  scratch buffer b holds the PE rows of its iteration
  const uint32_t bar_pe_free = bar0 + 8 * 14;     // [2]  scratch buffer b no longer needed
  const uint32_t bar_kready = bar0 + 8 * 20;      // [3]  bf16x3: hidden block k of the tile written (blocks before the last)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SMEM_BAR + 8 * 16);
  float* bias_s = reinterpret_cast<float*>(smem + SMEM_BIAS);  // [2][256]

  if (threadIdx.x == 0) {
    for (int i = 0; i < N_ENTRIES; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 2);  // MMA commits of both CTAs of the cluster
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_acc + 8 * s, 1);
      mbar_init(bar_aready + 8 * s, EPI_GROUP);
      mbar_init(bar_pe_ready + 8 * s, PE_THREADS);
      mbar_init(bar_pe_free + 8 * s, EPI_GROUP);
    }
    for (int k = 0; k < 3; ++k) mbar_init(bar_kready + 8 * k, EPI_THREADS);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // tile group g = (j*C + c)*NSLOT + s holds tiles 2g (cluster rank 0) and 2g+1 (rank 1); both CTAs of a
  // cluster walk the same (j, layer, slot) sequence because they share every multicast weight entry.
  const uint32_t crank = cluster_ctarank();
  const int C = gridDim.x >> 1, c = (int)blockIdx.x >> 1;
  const int n_groups = (P.n_tiles + 1) >> 1;
  const int n_iter = c * NSLOT < n_groups ? (n_groups - c * NSLOT + C * NSLOT - 1) / (C * NSLOT) : 0;
  auto valid_slot = [&](int j, int s) { return (j * C + c) * NSLOT + s < n_groups; };
  auto tile_of = [&](int j, int s) { return 2 * ((j * C + c) * NSLOT + s) + (int)crank; };
  const int NL = P.n_layers;
  uint8_t* scratch = P.pe_scratch + (size_t)blockIdx.x * 2 * 2 * NBLK * TILE_M * ROWB;
  // staged block `blk` of the tile in slot s of scratch buffer buf
  auto scr = [&](int buf, int s, int blk) { return scratch + (size_t)((buf * 2 + s) * NBLK + blk) * TILE_M * ROWB; };
  // Split schedule (one tile per CTA): the staged input block of layer (j, l) -- the positional encoding of layer 0 and of the skip
  // layer -- does not depend on the previous layer, and its ring entry is free as soon as the previous layer's accumulator is
  // complete, so the epilogue warps copy it in BEFORE they drain that accumulator and the issuer starts layer l on it while the
  // hidden blocks are still being written (the skip layer was 15,000 cycles with the copy at the end of the epilogue and no
  // overlap).  Not for the very first layer (the prologue's copy is ordered by `aready` alone) and not after a TC_EPI_STAGE
  // layer (whose epilogue WRITES the block its successor stages).  Evaluated identically by the issuer and the epilogue warps.
  uint32_t early_mask = 0u;    // bit l: layer l's staged block is copied early (every tile but the very first layer of the kernel)
  if (X3 && (P.flags & 1)) {
    for (int l = 0; l < NL; ++l)
      if (layer_has_pe(P.layers[l]) && !(DEC && P.layers[(l + NL - 1) % NL].epi == TC_EPI_STAGE)) early_mask |= 1u << l;
  }
  auto early_staged = [&](int j, int l) -> bool { return X3 && ((early_mask >> l) & 1u) != 0u && (j | l) != 0; };

  if (warp == 0) {
    // ============================== weight producer (TMA multicast) ==========================
    uint32_t cnt = 0;
    for (int j = 0; j < n_iter; ++j) {
      for (int l = 0; l < NL; ++l) {
        const TcLayer& L = P.layers[l];
        for (int s = 0; s < NSLOT; ++s) {
          if (!valid_slot(j, s)) continue;
          uint32_t off = L.woff;
          const uint32_t bytes = (uint32_t)L.n * 64u;
          for (int kbi = 0; kbi < L.nkb; ++kbi) {
            if (L.kb[kbi] >= TC_KB_PE) ++cnt;  // the entry before this K-block's weights is filled by the epilogue warps
            for (int part = 0; part < layer_parts<NPART>(L); ++part) {
              const uint32_t e = cnt % N_ENTRIES, par = (cnt / N_ENTRIES) & 1u;
              mbar_wait(bar_empty + 8 * e, par ^ 1u);
              if (elect_one_sync()) {
                mbar_expect_tx(bar_full + 8 * e, 2u * bytes);
                if ((cnt & 1u) == crank) {
                  const uint8_t* src = (part == 0 ? P.w_hi : P.w_lo) + off;
                  const uint32_t dst = sbase + SMEM_RING + e * 2u * SLOT_BYTES;
                  tma_bulk_load_mc(dst, src, bytes, bar_full + 8 * e, (uint16_t)3);
                  tma_bulk_load_mc(dst + SLOT_BYTES, src + bytes, bytes, bar_full + 8 * e, (uint16_t)3);
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
    // tail: the peer CTA's last commits still target this CTA's barriers; do not run to the exit (and let the
    // cluster retire) before every entry's final release, which needs both CTAs' arrivals, has landed here
    for (uint32_t k = 0; k < (uint32_t)N_ENTRIES && k < cnt; ++k) {
      const uint32_t u = cnt - 1u - k;
      mbar_wait(bar_empty + 8 * (u % N_ENTRIES), (u / N_ENTRIES) & 1u);
    }
  } else if (warp == 1) {
    // ================================= MMA issuer =========================================
    uint32_t cnt = 0;
    uint32_t apar[2] = {0u, 0u};
    uint32_t abuf = 0u;    // bf16x3: accumulator buffer of the current layer (toggles per layer, not across CONT -> ACCUM)
    uint32_t kpar = 0u;    // bf16x3: phase parities of kready[0..2]
    int nph_prev = 1;      // bf16x3: `aready` phases the previous layer's epilogue signals (first: the initial PE store)
    for (int j = 0; j < n_iter; ++j) {
      for (int l = 0; l < NL; ++l) {
        const TcLayer& L = P.layers[l];
        const uint32_t idesc = F16 ? make_idesc_f16(L.n) : make_idesc(L.n);
        const uint32_t acc0 = DEC && (L.flags & TC_F_ACCUM) ? 1u : 0u;  // continue the previous (TC_EPI_CONT) layer's sums
        for (int s = 0; s < NSLOT; ++s) {
          if (!valid_slot(j, s)) continue;
          const bool tr = P.trace != nullptr && blockIdx.x == 0 && j < P.trace_tiles;
          long long t_w0 = 0, t_w1 = 0, t_full = 0;
          if (tr) t_w0 = clock64();
          const int nk = X3 ? nph_prev - 1 : 0;   // block signals of the previous epilogue that precede its final one
          int kw = 0;                              // ... consumed so far
          bool fin = false;                        // final signal consumed
          auto wait_block = [&](int k) {
            mbar_wait(bar_kready + 8 * k, (kpar >> k) & 1u);
            kpar ^= 1u << k;
            tcgen05_fence_after();
          };
          auto wait_final = [&]() {
            mbar_wait(bar_aready + 8 * s, apar[s]);
            apar[s] ^= 1u;
            tcgen05_fence_after();
            fin = true;
          };
          const bool staged_layer = layer_has_pe(L);
          const bool early = early_staged(j, l);
          if (X3 && !(L.flags & TC_F_ACCUM)) abuf ^= 1u;
          if (!X3 || (staged_layer && !early)) {
            // (split schedule without early staging: the input is copied at the very end of the previous epilogue -- no overlap)
            for (; kw < nk; ++kw) wait_block(kw);
            wait_final();
          }
          if (tr) t_w1 = clock64();
          const uint32_t acc = tmem_base + (X3 ? abuf : (uint32_t)s) * 256u;
          for (int kbi = 0; kbi < L.nkb; ++kbi) {
            uint32_t a_hi, a_lo, pe_entry = 0;
            const bool is_pe = L.kb[kbi] >= TC_KB_PE;
            // split schedule: the K-block's first weight stage does not depend on the activations -- poll its barrier BEFORE the
            // block's (a completed mbarrier wait still costs ~150 cycles, and with one tile per CTA the last block's sits on the
            // layer's critical chain: epilogue -> final signal -> last K-block's MMAs -> accumulator barrier -> epilogue)
            bool prewaited = false;
            if (X3 && (P.flags & 2)) {
              const uint32_t cw = cnt + (is_pe ? 1u : 0u);
              mbar_wait(bar_full + 8 * (cw % N_ENTRIES), (cw / N_ENTRIES) & 1u);
              prewaited = true;
            }
            // bf16x3: hidden block kbi is complete once phase kbi+1 of the previous epilogue has been signalled
            if (X3 && !fin && !early) {
              if (kbi < nk) {
                wait_block(kbi);
                kw = kbi + 1;
              } else {
                wait_final();
              }
            } else if (X3 && !fin && !is_pe) {
              // early-staged layer: the staged block (waited for on its ring entry below) precedes the hidden blocks in K order
              const int hb = (int)L.kb[kbi];
              while (kw < nk && kw <= hb) {
                wait_block(kw);
                ++kw;
              }
              if (hb >= nk) wait_final();
            }
            if (is_pe) {
              pe_entry = cnt % N_ENTRIES;
              mbar_wait(bar_full + 8 * pe_entry, (cnt / N_ENTRIES) & 1u);
              tcgen05_fence_after();
              a_hi = sbase + SMEM_RING + pe_entry * 2u * SLOT_BYTES;
              a_lo = a_hi + SLOT_BYTES;
              ++cnt;
            } else {
              a_hi = sbase + (uint32_t)((X3 ? 0 : s * 4) + L.kb[kbi]) * KB_BYTES;
              a_lo = a_hi + 4u * KB_BYTES;
            }
            const uint64_t adesc_hi = make_smem_desc(a_hi);
            const uint64_t adesc_lo = make_smem_desc(a_lo);
            const int nparts = layer_parts<NPART>(L);
            for (int part = 0; part < nparts; ++part) {
              const uint32_t e = cnt % N_ENTRIES, par = (cnt / N_ENTRIES) & 1u;
              long long t_f0 = 0;
              if (tr) t_f0 = clock64();
              if (!(prewaited && part == 0)) mbar_wait(bar_full + 8 * e, par);
              tcgen05_fence_after();
              if (tr) t_full += clock64() - t_f0;
              const uint64_t bdesc = make_smem_desc_sw64(sbase + SMEM_RING + e * 2u * SLOT_BYTES);
#pragma unroll
              for (int q = 0; q < 4; ++q) {  // q = 2*(K half) + K step
                const uint64_t bd = bdesc + (uint64_t)((q >> 1) * (SLOT_BYTES >> 4) + (q & 1) * 2);
                umma_bf16(acc, adesc_hi + 2 * q, bd, idesc, (kbi | part | q) != 0 ? 1u : acc0);
              }
              if (X3 && part == 0 && nparts == 2) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint64_t bd = bdesc + (uint64_t)((q >> 1) * (SLOT_BYTES >> 4) + (q & 1) * 2);
                  umma_bf16(acc, adesc_lo + 2 * q, bd, idesc, 1u);
                }
              }
              umma_commit_mc(bar_empty + 8 * e, (uint16_t)3);
              ++cnt;
            }
            if (is_pe) umma_commit_mc(bar_empty + 8 * pe_entry, (uint16_t)3);
          }
          if (X3) {
            for (; kw < nk; ++kw) wait_block(kw);   // (a layer that reads fewer blocks than its predecessor wrote)
            if (!fin) wait_final();
            nph_prev = layer_phases(L, P.view_w);
          }
          umma_commit(bar_acc + 8 * s);
          if (tr && lane == 0) {
            unsigned long long* r = P.trace + ((size_t)(j * NL + l) * 2 + s) * 4;
            r[0] = (unsigned long long)t_w0;
            r[1] = (unsigned long long)t_w1;
            r[2] = (unsigned long long)clock64();
            r[3] = (unsigned long long)t_full;
          }
        }
      }
    }
  } else if (!COOP && warp >= 4 && warp < 12) {
    // ===================== epilogue warps, one group of four per tile slot (bf16) =====================
    const int s = (warp - 4) >> 2;        // this group's tile slot
    const int o = 1 - s;                  // the other slot
    const int q = warp & 3;               // TMEM lane quarter
    const int et = (warp & 3) * 32 + lane;  // 0..127 within the group
    const uint32_t row = (uint32_t)et;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t acc = lane_base + (uint32_t)s * 256u;
    uint8_t* arena_hi = smem + (size_t)(s * 4) * KB_BYTES;
    uint8_t* arena_lo = arena_hi;  // unused in bf16
    float* bias_g = bias_s + s * TC_BIAS_STRIDE;
    const uint32_t sbias = smem_u32(bias_g);
    const uint32_t sdot = sbase + SMEM_DOT;
    if (DEC && !X3 && P.dot_w != nullptr) {
      // each slot group writes the whole row (identical values), so its own barrier below orders it before its reads
      const float2 w2 = reinterpret_cast<const float2*>(P.dot_w)[et];
      reinterpret_cast<uint32_t*>(smem + SMEM_DOT)[et] = F16 ? pack_f16(w2.x, w2.y) : pack_bf16(w2.x, w2.y);
    }
    uint32_t acc_par = 0u;
    int pe_waited_j = -1;
    int last_pe_layer = 0;
    for (int l2 = 0; l2 < NL; ++l2)
      if (layer_has_pe(P.layers[l2])) last_pe_layer = l2;
    uint4 vh[8];
    // PE row of tile (jj, s): scratch (L2) -> registers
    auto pe_load = [&](int jj, int blk) {
      if (jj != pe_waited_j) {
        mbar_wait(bar_pe_ready + 8 * (jj & 1), (uint32_t)(jj >> 1) & 1u);
        pe_waited_j = jj;
      }
      const uint4* src = reinterpret_cast<const uint4*>(scr(jj & 1, s, blk)) + row;
#pragma unroll
      for (int k = 0; k < 8; ++k) vh[k] = src[k * TILE_M];
    };
    // registers -> ring entry cnt_e, then hand the slot to the MMA issuer.  The entry is free once every earlier
    // entry has been consumed by this CTA's MMAs, i.e. once the accumulator barrier of the layer-slot that precedes
    // the consumer in MMA order has completed (see the note in the cooperative branch); the caller has waited.
    auto pe_store = [&](int jj, int layer, uint32_t cnt_e) {
      const uint32_t e = cnt_e % N_ENTRIES;
      uint8_t* dst = smem + SMEM_RING + (size_t)e * 2 * SLOT_BYTES;
#pragma unroll
      for (int k = 0; k < 8; ++k) *reinterpret_cast<uint4*>(dst + swz(row, (uint32_t)k)) = vh[k];
      fence_proxy_async();
      if (et == 0) mbar_arrive(bar_full + 8 * e);
      if (layer == last_pe_layer) {   // scratch buffer of iteration jj is dead after its last slot's last copy
        bool last_slot = true;
        for (int s2 = s + 1; s2 < NSLOT; ++s2)
          if (valid_slot(jj, s2)) last_slot = false;
        if (last_slot) mbar_arrive(bar_pe_free + 8 * (jj & 1));
      }
      mbar_arrive(bar_aready + 8 * s);
    };
    // ring position of layer-slot (j, l, s): entries of all earlier layer-slots in MMA order
    uint32_t base = 0;  // entries before (j, l, slot 0)
    if (n_iter > 0 && valid_slot(0, s)) {
      const uint32_t e0 = layer_entries<NPART>(P.layers[0]);
      reinterpret_cast<float2*>(bias_g)[et] = reinterpret_cast<const float2*>(P.bias)[et];
      pe_load(0, layer_staged(P.layers[0]));
      pe_store(0, 0, s == 0 ? 0u : e0);   // nothing precedes the first layer-slots: the ring is empty
      named_bar_sync(1 + s, EPI_GROUP);
    }
    float alpha = 0.f;
    for (int j = 0; j < n_iter; ++j) {
      if (!valid_slot(j, s)) break;
      const uint32_t nv = valid_slot(j, 1) ? 2u : 1u;   // valid slots in this iteration (slot 0 always is)
      const int tile = tile_of(j, s);
      int64_t pt = (int64_t)tile * TILE_M + row;
      const bool valid = pt < P.n_points;
      if (!valid) pt = P.n_points - 1;
      const int64_t ray = pt / P.S;
      for (int l = 0; l < NL; ++l) {
        const TcLayer& L = P.layers[l];
        const uint32_t ents = layer_entries<NPART>(L);
        const int ln = (l + 1) % NL, jn = j + (l + 1 == NL ? 1 : 0);
        const int st_next = layer_staged(P.layers[ln]);
        const bool next_has_pe = st_next >= 0;
        const bool next_valid = jn < n_iter && valid_slot(jn, s);
        // a TC_EPI_STAGE layer writes the block its successor stages: load that one after the epilogue
        const bool late_load = DEC && L.epi == TC_EPI_STAGE;
        float2 nb = make_float2(0.f, 0.f);
        nb = reinterpret_cast<const float2*>(P.bias + ln * TC_BIAS_STRIDE)[et];
        if (next_valid && next_has_pe && !late_load) pe_load(jn, st_next);   // L2 latency hidden behind the accumulator wait
        const bool tr = P.trace != nullptr && blockIdx.x == 0 && et == 0 && j < P.trace_tiles;
        long long t_e0 = 0, t_e1 = 0;
        if (tr) t_e0 = clock64();
        if (L.epi == TC_EPI_VIEW0) prefetch_row_l1(P.view_bias + ray * P.view_w, P.view_w);   // hidden behind the wait
        mbar_wait(bar_acc + 8 * s, acc_par);
        acc_par ^= 1u;
        tcgen05_fence_after();
        if (tr) t_e1 = clock64();

        if (L.epi == TC_EPI_RGB) {
          uint32_t v[16];
          tmem_ld16(acc, v);
          tmem_ld_wait();
          if (valid) {
            float4 ov;
            ov.x = __uint_as_float(v[0]) + bias_g[0];
            ov.y = __uint_as_float(v[1]) + bias_g[1];
            ov.z = __uint_as_float(v[2]) + bias_g[2];
            if (DEC) {  // DEC:346-347
              ov.x = sigmoid_f(ov.x);
              ov.y = sigmoid_f(ov.y);
              ov.z = sigmoid_f(ov.z);
            }
            ov.w = alpha;
            reinterpret_cast<float4*>(P.raw)[pt] = ov;
          }
        } else if (DEC && L.epi == TC_EPI_SIGMA) {
          uint32_t v[16];
          tmem_ld16(acc, v);
          tmem_ld_wait();
          alpha = __uint_as_float(v[0]) + bias_g[0];
        } else if (DEC && L.epi == TC_EPI_CONT) {
          // partial sums stay in the accumulator
        } else if (DEC && L.epi == TC_EPI_STAGE) {
          const int nch = (int)L.n >> 5;
          uint32_t v0[32], v1[32];
          tmem_ld32(acc, v0);
          for (int cc = 0; cc < nch; cc += 2) {
            tmem_ld_wait();
            tmem_ld32(acc + (cc + 1) * 32, v1);
            stage_chunk<X3, F16>(v0, cc, sbias, scr(j & 1, s, cc >> 1), row);
            tmem_ld_wait();
            if (cc + 2 < nch) tmem_ld32(acc + (cc + 2) * 32, v0);
            stage_chunk<X3, F16>(v1, cc + 1, sbias, scr(j & 1, s, (cc + 1) >> 1), row);
          }
        } else if (DEC && !X3 && (L.flags & TC_F_DOT_SIGMA)) {
          // last trunk block + sigma_out (DEC:329): the usual epilogue, and the density from its fp32 activations
          const int nch = (int)L.n >> 5;
          float d4[4] = {0.f, 0.f, 0.f, 0.f};
          uint32_t v0[32], v1[32];
          tmem_ld32(acc, v0);
          for (int cc = 0; cc < nch; cc += 2) {
            tmem_ld_wait();
            tmem_ld32(acc + (cc + 1) * 32, v1);
            epilogue_chunk_dot1<F16>(v0, cc, sbias, sdot, arena_hi, row, d4);
            tmem_ld_wait();
            if (cc + 2 < nch) tmem_ld32(acc + (cc + 2) * 32, v0);
            epilogue_chunk_dot1<F16>(v1, cc + 1, sbias, sdot, arena_hi, row, d4);
          }
          alpha = (d4[0] + d4[1]) + (d4[2] + d4[3]) + P.dot_w[TC_BIAS_STRIDE];
        } else if (!X3 && L.epi == TC_EPI_RELU && (L.n & 63) == 0) {
          // column-distributed readout (tc_epi.cuh): the bias in registers instead of warp-broadcast loads
          epilogue_relu_cd<F16>(acc, 0, (int)L.n >> 6, sbias, arena_hi, (uint32_t)(q * 32), (uint32_t)(threadIdx.x & 31));
        } else {
          const bool per_ray = L.epi == TC_EPI_VIEW0;
          const int nch = (per_ray ? P.view_w : (int)L.n) >> 5;   // 32-column chunks: 8 or 4
          const float* rb = P.view_bias + ray * P.view_w;
          uint32_t v0[32], v1[32];
          tmem_ld32(acc, v0);
          for (int cc = 0; cc < nch; cc += 2) {
            tmem_ld_wait();
            tmem_ld32(acc + (cc + 1) * 32, v1);
            if (per_ray) epilogue_chunk<X3, true, F16>(v0, cc, rb, 0u, arena_hi, arena_lo, row);
            else epilogue_chunk<X3, false, F16>(v0, cc, nullptr, sbias, arena_hi, arena_lo, row);
            tmem_ld_wait();
            if (cc + 2 < nch) tmem_ld32(acc + (cc + 2) * 32, v0);
            if (per_ray) epilogue_chunk<X3, true, F16>(v1, cc + 1, rb, 0u, arena_hi, arena_lo, row);
            else epilogue_chunk<X3, false, F16>(v1, cc + 1, nullptr, sbias, arena_hi, arena_lo, row);
          }
          if (!DEC && per_ray) {
            uint32_t v[16];
            tmem_ld16(acc + P.view_w, v);
            tmem_ld_wait();
            alpha = __uint_as_float(v[0]) + bias_g[P.view_w];
          }
        }
        tcgen05_fence_before();
        fence_proxy_async();

        if (next_valid) {
          if (next_has_pe) {
            if (late_load) pe_load(jn, st_next);
            // consumer (jn, ln, s); the layer-slot issued just before it belongs to the other slot (if that is valid)
            const uint32_t ents_n = layer_entries<NPART>(P.layers[ln]);
            uint32_t cn;
            if (s == 0) {
              cn = base + ents * nv;                       // after both slots' layer l
              if (nv == 2u) mbar_wait(bar_acc + 8 * o, (uint32_t)(j * NL + l) & 1u);            // acc of (j, l, 1)
            } else {
              cn = base + ents * nv + ents_n;              // after slot 0's next layer
              mbar_wait(bar_acc + 8 * o, (uint32_t)(jn * NL + ln) & 1u);                          // acc of (jn, ln, 0)
            }
            pe_store(jn, ln, cn);
          } else {
            mbar_arrive(bar_aready + 8 * s);
          }
        }
        // swap in the next layer's bias
        named_bar_sync(1 + s, EPI_GROUP);
        reinterpret_cast<float2*>(bias_g)[et] = nb;
        named_bar_sync(1 + s, EPI_GROUP);
        base += ents * nv;
        if (tr) {
          unsigned long long* r = P.trace + (size_t)P.trace_tiles * NL * 8 + ((size_t)(j * NL + l) * 2 + s) * 4;
          r[0] = (unsigned long long)t_e0;
          r[1] = (unsigned long long)t_e1;
          r[2] = (unsigned long long)clock64();
          r[3] = 0;
        }
      }
    }
  } else if (COOP && warp >= 4 && warp < 12) {
    // ================================ epilogue warps (cooperative, bf16x3) ============================
    const int q = warp & 3;               // TMEM lane quarter
    const int hf = (warp - 4) >> 2;       // which half of the layer's columns
    const int et = (warp - 4) * 32 + lane;  // 0..255
    const uint32_t row = (uint32_t)(q * 32 + lane);
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t acc_par[2] = {0u, 0u};
    uint32_t abuf = 0u;                   // accumulator buffer of the current layer (mirrors the MMA issuer)
    uint32_t cnt = 0;                     // ring entries before the current (j, l, s) in MMA order
    float alpha[2] = {0.f, 0.f};
    float* alpha_s = reinterpret_cast<float*>(smem + SMEM_DOT);   // TC_F_DOT_ALPHA: the tile's densities [128] (the region is otherwise
    bool dot_alpha_prog = false;                                  // used by the single-pass Decoder kernels only)
    int pe_waited_j = -1;
    int last_pe_layer = 0;
    for (int l2 = 0; l2 < NL; ++l2) {
      if (layer_has_pe(P.layers[l2])) last_pe_layer = l2;
      if (P.layers[l2].flags & (DEC ? TC_F_DOT_SIGMA : TC_F_DOT_ALPHA)) dot_alpha_prog = true;
    }
    // the scratch buffer of iteration jj is dead once its last slot's copy for the last PE-consuming layer is made
    auto maybe_free = [&](int jj, int layer, int s) {
      if (layer != last_pe_layer) return;
      for (int s2 = s + 1; s2 < NSLOT; ++s2)
        if (valid_slot(jj, s2)) return;
      mbar_arrive(bar_pe_free + 8 * (jj & 1));
    };

    // Copy of this thread's part of a PE row, scratch (L2) -> ring entry, in two halves: the loads are issued
    // when the copy is scheduled, the stores when the ring entry is known to be drained.
    // WAR on the ring entry is a LOCAL condition: every earlier entry has been consumed by this CTA's MMAs once
    // the accumulator barrier of the layer-slot that precedes the consumer in MMA order has completed (no
    // multicast write can target the entry before this CTA has released its PE use).  Waiting on the shared
    // `empty` barrier here instead would alias phases when the peer CTA lags by more than one ring turn.
    struct Pending {
      bool on;
      int jj, layer, s;
      uint32_t cnt_e;
      uint4 vh[4], vl[4];
    } pend;
    pend.on = false;
    auto pe_load = [&](int jj, int layer, int s, uint32_t cnt_e) {
      if (jj != pe_waited_j) {
        mbar_wait(bar_pe_ready + 8 * (jj & 1), (uint32_t)(jj >> 1) & 1u);
        pe_waited_j = jj;
      }
      const uint4* src = reinterpret_cast<const uint4*>(scr(jj & 1, s, layer_staged(P.layers[layer]))) + row;
#pragma unroll
      for (int k = 0; k < 4; ++k) pend.vh[k] = src[(4 * hf + k) * TILE_M];
      if (X3) {
#pragma unroll
        for (int k = 0; k < 4; ++k) pend.vl[k] = src[(8 + 4 * hf + k) * TILE_M];
      }
      pend.on = true;
      pend.jj = jj;
      pend.layer = layer;
      pend.s = s;
      pend.cnt_e = cnt_e;
    };
    auto pe_store = [&]() {
      const uint32_t e = pend.cnt_e % N_ENTRIES;
      uint8_t* dst = smem + SMEM_RING + (size_t)e * 2 * SLOT_BYTES;
#pragma unroll
      for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(dst + swz(row, (uint32_t)(4 * hf + k))) = pend.vh[k];
      if (X3) {
#pragma unroll
        for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(dst + SLOT_BYTES + swz(row, (uint32_t)(4 * hf + k))) = pend.vl[k];
      }
      fence_proxy_async();
      if (et == 0) mbar_arrive(bar_full + 8 * e);  // phase bookkeeping; completeness is covered by a_ready
      maybe_free(pend.jj, pend.layer, pend.s);
      mbar_arrive(bar_aready + 8 * pend.s);
      pend.on = false;
    };
    // first layer's bias + the PE blocks of iteration 0
    bias_s[et] = P.bias[et];
    {
      uint32_t c0 = 0;
      for (int s = 0; s < NSLOT; ++s) {
        if (n_iter == 0 || !valid_slot(0, s)) continue;
        pe_load(0, 0, s, c0);
        pe_store();
        c0 += layer_entries<NPART>(P.layers[0]);
      }
    }

    for (int j = 0; j < n_iter; ++j) {
      for (int l = 0; l < NL; ++l) {
        const TcLayer& L = P.layers[l];
        const uint32_t ents = layer_entries<NPART>(L);
        const int gl = j * NL + l;
        // per layer: everyone is done with the other bias buffer -> refill it with the next layer's bias
        named_bar_sync(1, EPI_THREADS);
        bias_s[((gl + 1) & 1) * TC_BIAS_STRIDE + et] = P.bias[((l + 1) % NL) * TC_BIAS_STRIDE + et];
        const uint32_t sbias = smem_u32(bias_s + (gl & 1) * TC_BIAS_STRIDE);
        const float* bl = bias_s + (gl & 1) * TC_BIAS_STRIDE;
        const int ln = (l + 1) % NL, jn = j + (l + 1 == NL ? 1 : 0);
        const bool next_has_pe = layer_has_pe(P.layers[ln]);
        const uint32_t ents_next = layer_entries<NPART>(P.layers[ln]);
        // Per-ray-bias layer (views_linears.0) on the split schedule, as in mlp_pair.cu: with a sample count that is a multiple of 32 a
        // warp's rows lie on one ray and the tile touches at most two, so their two bias rows replace the layer's static bias in the
        // current staging buffer and the layer takes the column-distributed readout below (its row-per-thread readout with 128 global
        // bias loads per thread was 4.3k cycles for 128 columns, on the chain of the program's short tail layers); bit-identical.
        const bool vfast = X3 && !DEC && (P.flags & 8) && L.epi == TC_EPI_VIEW0 && (P.S & 31) == 0 && P.S >= 64 &&
                           2 * P.view_w <= TC_BIAS_STRIDE && valid_slot(j, 0);
        int64_t vray0 = 0;
        if (vfast) {
          const int64_t vp0 = (int64_t)tile_of(j, 0) * TILE_M, last = P.n_points - 1;
          const int64_t vp1 = vp0 + TILE_M - 1 < last ? vp0 + TILE_M - 1 : last;
          const int64_t myray = (vp0 + row < last ? vp0 + row : last) / P.S;
          vray0 = myray - (myray * P.S > vp0 ? 1 : 0);
          const int64_t vray1 = myray + ((myray + 1) * P.S - 1 < vp1 ? 1 : 0);
          if (et < 2 * P.view_w)
            bias_s[(gl & 1) * TC_BIAS_STRIDE + et] = __ldg(P.view_bias + (et < P.view_w ? vray0 : vray1) * P.view_w + (et < P.view_w ? et : et - P.view_w));
          named_bar_sync(2, EPI_THREADS);
        }

        for (int s = 0; s < NSLOT; ++s) {
          if (!valid_slot(j, s)) continue;
          const int tile = tile_of(j, s);
          int64_t pt = (int64_t)tile * TILE_M + row;
          const bool valid = pt < P.n_points;
          if (!valid) pt = P.n_points - 1;
          const int64_t ray = pt / P.S;
          uint8_t* arena_hi = smem + (size_t)(X3 ? 0 : s * 4) * KB_BYTES;
          uint8_t* arena_lo = arena_hi + (size_t)4 * KB_BYTES;
          if (X3 && !(L.flags & TC_F_ACCUM)) abuf ^= 1u;
          const uint32_t acc = lane_base + (X3 ? abuf : (uint32_t)s) * 256u;
          const bool tr = P.trace != nullptr && blockIdx.x == 0 && et == 0 && j < P.trace_tiles;
          long long t_e0 = 0, t_e1 = 0;
          if (tr) t_e0 = clock64();

          if (L.epi == TC_EPI_VIEW0 && !vfast) prefetch_row_l1(P.view_bias + ray * P.view_w, P.view_w);   // hidden behind the wait
          const bool next_valid = jn < n_iter && valid_slot(jn, s);
          const bool early_next = next_valid && early_staged(jn, ln);
          // early staging (see early_staged): the next layer's staged block is pulled from the L2-resident scratch into L1 before the
          // accumulator wait (no registers held across it) ...
          const uint4* esrc = nullptr;
          if (early_next) {
            if (jn != pe_waited_j) {
              mbar_wait(bar_pe_ready + 8 * (jn & 1), (uint32_t)(jn >> 1) & 1u);
              pe_waited_j = jn;
            }
            esrc = reinterpret_cast<const uint4*>(scr(jn & 1, s, layer_staged(P.layers[ln]))) + row;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              asm volatile("prefetch.global.L1 [%0];" ::"l"(esrc + (4 * hf + k) * TILE_M));
              asm volatile("prefetch.global.L1 [%0];" ::"l"(esrc + (8 + 4 * hf + k) * TILE_M));
            }
          }
          mbar_wait(bar_acc + 8 * s, acc_par[s]);
          acc_par[s] ^= 1u;
          tcgen05_fence_after();
          if (tr) t_e1 = clock64();
          if (early_next) {
            // ... and copied into its ring entry (one slot: the consumer's entries follow this layer's, all earlier ones are consumed now), published
            // through the entry's `full` barrier alone: the issuer waits for it before the staged K-block and consumes `aready` after the
            // layer's last hidden block as for any other layer
            const uint32_t e = (cnt + ents) % N_ENTRIES;
            uint8_t* dst = smem + SMEM_RING + (size_t)e * 2 * SLOT_BYTES;
            uint4 eh[4], el[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) eh[k] = esrc[(4 * hf + k) * TILE_M];
#pragma unroll
            for (int k = 0; k < 4; ++k) el[k] = esrc[(8 + 4 * hf + k) * TILE_M];
#pragma unroll
            for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(dst + swz(row, (uint32_t)(4 * hf + k))) = eh[k];
#pragma unroll
            for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(dst + SLOT_BYTES + swz(row, (uint32_t)(4 * hf + k))) = el[k];
            fence_proxy_async();
            named_bar_sync(2, EPI_THREADS);
            if (et == 0) mbar_arrive(bar_full + 8 * e);
            maybe_free(jn, ln, s);
          } else if (pend.on) {
            pe_store();  // every ring entry before the pending PE block has now been consumed
          }

          if (L.epi == TC_EPI_RGB) {
            if (hf == 0) {
              uint32_t v[16];
              tmem_ld16(acc, v);
              tmem_ld_wait();
              if (valid) {
                float4 o;
                o.x = __uint_as_float(v[0]) + bl[0];
                o.y = __uint_as_float(v[1]) + bl[1];
                o.z = __uint_as_float(v[2]) + bl[2];
                if (DEC) {  // DEC:346-347
                  o.x = sigmoid_f(o.x);
                  o.y = sigmoid_f(o.y);
                  o.z = sigmoid_f(o.z);
                }
                o.w = dot_alpha_prog ? alpha_s[row] : alpha[s];
                reinterpret_cast<float4*>(P.raw)[pt] = o;
              }
            }
          } else if (DEC && L.epi == TC_EPI_SIGMA) {
            if (hf == 0) {
              uint32_t v[16];
              tmem_ld16(acc, v);
              tmem_ld_wait();
              alpha[s] = __uint_as_float(v[0]) + bl[0];
            }
          } else if (DEC && L.epi == TC_EPI_CONT) {
            // partial sums stay in the accumulator
          } else if (DEC && L.epi == TC_EPI_STAGE) {
            // this thread's half of the columns is one whole staged block (N = 128: PE' | signal')
            const int nch = (int)L.n >> 6;
            const int ch0 = hf * nch;
            uint32_t v0[32], v1[32];
            tmem_ld32(acc + ch0 * 32, v0);
            for (int cc = 0; cc < nch; cc += 2) {
              tmem_ld_wait();
              tmem_ld32(acc + (ch0 + cc + 1) * 32, v1);
              stage_chunk<X3, F16>(v0, ch0 + cc, sbias, scr(j & 1, s, (ch0 + cc) >> 1), row);
              tmem_ld_wait();
              if (cc + 2 < nch) tmem_ld32(acc + (ch0 + cc + 2) * 32, v0);
              stage_chunk<X3, F16>(v1, ch0 + cc + 1, sbias, scr(j & 1, s, (ch0 + cc + 1) >> 1), row);
            }
          } else if ((L.epi == TC_EPI_RELU && (L.n & 63) == 0) || vfast) {
            // column-distributed readout (tc_epi.cuh): the two warps of a lane quarter take its two 16-lane halves of every
            // 64-column block, so block k is complete after everybody's k-th piece and is signalled as below; the lo plane is
            // skipped when the consuming layer is single-pass
            const int nkb_out = (vfast ? P.view_w : (int)L.n) >> 6;
            const uint32_t sbias_l = sbias + (vfast && ray != vray0 ? (uint32_t)P.view_w * 4u : 0u);   // vfast: this warp's ray's row
            const bool need_lo = !(P.layers[ln].flags & TC_F_SINGLE);
            const uint32_t accp = acc + ((uint32_t)(16 * hf) << 16);
            const uint32_t r0 = (uint32_t)(q * 32 + 16 * hf) + ((uint32_t)lane >> 2);
            uint32_t v0[32], v1[32];
            auto bias16 = [&](int kb, float (&b)[16]) {
              const uint32_t ba = sbias_l + (uint32_t)(kb * 64 + 2 * (lane & 3)) * 4u;
#pragma unroll
              for (int g = 0; g < 8; ++g)
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(b[2 * g]), "=f"(b[2 * g + 1]) : "r"(ba + (uint32_t)g * 32u));
            };
            auto done = [&](int k) {
              if (k + 1 < nkb_out) {
                tcgen05_fence_before();
                fence_proxy_async();
                mbar_arrive(bar_kready + 8 * k);
              }
            };
            // TC_F_DOT_ALPHA: density = alpha_linear.weight . relu(out) + bias from the fp32 activations (rows r0 and r0 + 8 of this thread,
            // its 16 columns of every block; the four lanes of a quad hold a row's other columns)
            const bool dot_alpha = DEC ? (L.flags & TC_F_DOT_SIGMA) != 0 : (L.flags & TC_F_DOT_ALPHA) != 0;   // (Decoder: sigma_out, DEC:329)
            float dsum[2] = {0.f, 0.f};
            auto dot16 = [&](const uint32_t (&v)[32], const float (&b)[16], int kb) {
              const float2* wp = reinterpret_cast<const float2*>(P.dot_w + kb * 64 + 2 * (lane & 3));
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float2 w = __ldg(wp + 4 * g);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const float x0 = fmaxf(__uint_as_float(v[4 * g + 2 * h]) + b[2 * g], 0.f);
                  const float x1 = fmaxf(__uint_as_float(v[4 * g + 2 * h + 1]) + b[2 * g + 1], 0.f);
                  dsum[h] = fmaf(x1, w.y, fmaf(x0, w.x, dsum[h]));
                }
              }
            };
            tmem_ld_16x256b_x8(accp, v0);
            for (int kb = 0; kb < nkb_out; kb += 2) {
              float b[16];
              bias16(kb, b);
              tmem_ld_wait();
              if (kb + 1 < nkb_out) tmem_ld_16x256b_x8(accp + (uint32_t)(kb + 1) * 64u, v1);
              epilogue_piece_cd_split<F16>(v0, b, arena_hi + (size_t)kb * KB_BYTES, arena_lo + (size_t)kb * KB_BYTES, need_lo, r0, (uint32_t)lane);
              done(kb);
              if (dot_alpha) dot16(v0, b, kb);
              if (kb + 1 < nkb_out) {
                bias16(kb + 1, b);
                tmem_ld_wait();
                if (kb + 2 < nkb_out) tmem_ld_16x256b_x8(accp + (uint32_t)(kb + 2) * 64u, v0);
                epilogue_piece_cd_split<F16>(v1, b, arena_hi + (size_t)(kb + 1) * KB_BYTES, arena_lo + (size_t)(kb + 1) * KB_BYTES, need_lo, r0,
                                             (uint32_t)lane);
                done(kb + 1);
                if (dot_alpha) dot16(v1, b, kb + 1);
              }
            }
            if (dot_alpha) {
              const float ab = __ldg(P.dot_b);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                float t = dsum[h];
                t += __shfl_xor_sync(0xffffffffu, t, 1);
                t += __shfl_xor_sync(0xffffffffu, t, 2);
                if ((lane & 3) == 0) alpha_s[r0 + 8u * (uint32_t)h] = t + ab;     // read by the row's owner in the TC_EPI_RGB epilogue
              }
            }
            if (vfast && hf == 0) {  // density head riding as accumulator column view_w (its bias is no longer in the staging buffer)
              uint32_t v[16];
              tmem_ld16(acc + P.view_w, v);
              tmem_ld_wait();
              alpha[s] = __uint_as_float(v[0]) + __ldg(P.bias + l * TC_BIAS_STRIDE + P.view_w);
            }
          } else {
            const bool per_ray = L.epi == TC_EPI_VIEW0;
            const int n_relu = per_ray ? P.view_w : (int)L.n;
            const int nch = n_relu >> 6;            // 32-column chunks per thread (4 or 2)
            // this thread's k-th chunk is 2k + hf: the two column halves of 64-column block k belong to the two warps of
            // a lane quarter, so block k is complete -- and signalled to the MMA issuer, which may start the next
            // layer on it -- after everybody's k-th chunk (the last block is signalled by the hand-back below)
            const float* rb = P.view_bias + ray * P.view_w;
            uint32_t v0[32], v1[32];
            auto block_done = [&](int k) {
              if (k + 1 < nch) {
                tcgen05_fence_before();
                fence_proxy_async();
                mbar_arrive(bar_kready + 8 * k);
              }
            };
            tmem_ld32(acc + hf * 32, v0);
            for (int cc = 0; cc < nch; cc += 2) {
              const int c0 = 2 * cc + hf, c1 = c0 + 2;
              tmem_ld_wait();
              tmem_ld32(acc + c1 * 32, v1);
              if (per_ray) epilogue_chunk<X3, true, F16>(v0, c0, rb, 0u, arena_hi, arena_lo, row);
              else epilogue_chunk<X3, false, F16>(v0, c0, nullptr, sbias, arena_hi, arena_lo, row);
              block_done(cc);
              tmem_ld_wait();
              if (cc + 2 < nch) tmem_ld32(acc + (c1 + 2) * 32, v0);
              if (per_ray) epilogue_chunk<X3, true, F16>(v1, c1, rb, 0u, arena_hi, arena_lo, row);
              else epilogue_chunk<X3, false, F16>(v1, c1, nullptr, sbias, arena_hi, arena_lo, row);
              block_done(cc + 1);
            }
            if (!DEC && per_ray && hf == 0) {  // density head: accumulator column view_w, no activation
              uint32_t v[16];
              tmem_ld16(acc + P.view_w, v);
              tmem_ld_wait();
              alpha[s] = __uint_as_float(v[0]) + bl[P.view_w];
            }
          }
          tcgen05_fence_before();
          fence_proxy_async();

          // hand the slot back to the MMA issuer; if its next layer consumes the PE block, stage that first
          if (next_valid) {
            if (next_has_pe && !early_next) {
              // ring position of the next layer of this slot: the rest of this layer, then the earlier slots of the next
              uint32_t cn = cnt + ents;
              bool between = false;  // another layer-slot is issued between this one and the consumer
              for (int s2 = s + 1; s2 < NSLOT; ++s2)
                if (valid_slot(j, s2)) {
                  cn += ents;
                  between = true;
                }
              for (int s2 = 0; s2 < s; ++s2)
                if (valid_slot(jn, s2)) {
                  cn += ents_next;
                  between = true;
                }
              // the staged block was written by other threads of the group in this layer's epilogue
              if (DEC && L.epi == TC_EPI_STAGE) named_bar_sync(2, EPI_THREADS);
              pe_load(jn, ln, s, cn);
              if (!between) pe_store();  // else: stored (and a_ready signalled) after that layer-slot's accumulator wait
            } else {
              mbar_arrive(bar_aready + 8 * s);
            }
          }
          cnt += ents;
          if (tr) {
            unsigned long long* r = P.trace + (size_t)P.trace_tiles * NL * 8 + ((size_t)(j * NL + l) * 2 + s) * 4;
            r[0] = (unsigned long long)t_e0;
            r[1] = (unsigned long long)t_e1;
            r[2] = (unsigned long long)clock64();
            r[3] = 0;
          }
        }
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ============================ positional-encoding warps ===================================
    // One iteration ahead of the MLP: PE rows (bf16 hi [, lo]) of both slots' tiles -> L2-resident scratch,
    // laid out [16-byte chunk][row] so that both these stores and the epilogue warps' loads are coalesced.
    // sin/cos: the argument x*2^k is exact; two-constant Cody-Waite reduction to [-pi, pi] (error ~1e-8 for
    // |x*2^k| < 1e3) and the MUFU sin/cos (abs error 2^-21.4 on that range) -- well below the bf16 (hi) and
    // hi+lo (2^-17) resolution the MMA operands keep.
    for (int j = 0; j < n_iter; ++j) {
      const int buf = j & 1;
      mbar_wait(bar_pe_free + 8 * buf, ((uint32_t)(j >> 1) & 1u) ^ 1u);
      for (int sr = 0; sr < 2 * NSLOT; ++sr) {
        const int s = sr >> 1;
        const uint32_t row = (uint32_t)((warp - 2) * 32 + lane + (sr & 1) * 64);
        if (!valid_slot(j, s)) continue;
        int64_t pt = (int64_t)tile_of(j, s) * TILE_M + row;
        if (pt >= P.n_points) pt = P.n_points - 1;
        const int64_t ray = pt / P.S;
        const float z = P.z_vals[pt];
        float pe[64], x[3];
        sample_point(P.rays_o, P.rays_d, ray, z, x);
        if (!DEC) pe_embedder(x, P.multires, pe);
        else pe_decoder(x, P.multires, pe);
        scratch_row<X3, F16>(scr(buf, s, 0), row, pe);
        if (DEC) {
          // view-direction encoding (DEC:337-338): d / |d|, halved, [sin(2^k pi d) | cos(2^k pi d)]_k -- constant along a
          // ray, but a K-block of the view layer's MMA is cheaper than a per-ray bias row read per eight columns in its
          // epilogue (global broadcast loads made that epilogue 2.5x longer than any other)
          pe_decoder_viewdir(P.rays_d, ray, P.multires_views, pe);
          scratch_row<X3, F16>(scr(buf, s, 2), row, pe);
        }
      }
      mbar_arrive(bar_pe_ready + 8 * buf);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace pp

static int g_pp_flags = 15;   // bit 0: early staging in the split schedule, bit 1: weight barrier polled before the activation block's,
                             // bit 2: the density head (alpha_linear in fp16x3m, the Decoder's sigma_out in bf16x3) is evaluated in fp32 inside the last
                             // trunk layer's epilogue, bit 3: the per-ray-bias layer's two bias rows staged in shared memory, column-distributed readout
                             // (debug: dfn_debug_set_pp_flags)
void pp_set_flags(int flags) { g_pp_flags = flags; }
int pp_get_flags() { return g_pp_flags; }

int64_t pp_scratch_bytes() { return (int64_t)(num_sms() + 1) * 2 * 2 * tc::TILE_M * 256; }
int64_t pp_dec_scratch_bytes() { return 3 * pp_scratch_bytes(); }  // three staged blocks per tile

template <bool X3, bool DEC, bool F16 = false>
static int pp_launch_t(const pp::Params& P, int grid, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    DFN_CUDA(cudaFuncSetAttribute(pp::mlp_pp_kernel<X3, DEC, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, pp::SMEM_TOTAL));
    attr_done = true;
  }
  pp::mlp_pp_kernel<X3, DEC, F16><<<grid, 384, pp::SMEM_TOTAL, st>>>(P);
  return 0;
}

// prog: layer program (woff32: per-layer offsets into the K=32 stage blobs w_hi / w_lo); decoder: Decoder programs
// (DEC kernel instantiation, 256-wide per-ray view bias, PE of DEC:257-275 with n_freq = multires).
int pp_launch_prog(const TcProgram& prog, const uint32_t* woff32, const uint8_t* w_hi, const uint8_t* w_lo, const float* dot_w,
                   bool decoder, int multires, int multires_views, int view_w, const float* bias_ws, const float* vbias_ws, void* scratch, int64_t R, int S,
                   const float* rays_o, const float* rays_d, const float* z_vals, float* raw, int precision,
                   cudaStream_t st, const float* dot_b) {
  pp::Params P;
  memset(&P, 0, sizeof(P));
  P.w_hi = w_hi;
  P.w_lo = w_lo;
  P.bias = bias_ws;
  P.view_bias = vbias_ws;
  P.dot_w = dot_w;
  P.dot_b = dot_b;
  P.rays_o = rays_o;
  P.rays_d = rays_d;
  P.z_vals = z_vals;
  P.raw = raw;
  P.pe_scratch = reinterpret_cast<uint8_t*>(scratch);
  tc_get_trace(reinterpret_cast<void**>(&P.trace), &P.trace_tiles);
  P.n_points = R * S;
  P.S = S;
  P.n_tiles = (int)((P.n_points + tc::TILE_M - 1) / tc::TILE_M);
  P.n_layers = prog.n_layers;
  P.multires = multires;
  P.multires_views = multires_views;
  P.view_w = view_w;
  P.flags = g_pp_flags;
  for (int i = 0; i < prog.n_layers; ++i) {
    P.layers[i] = prog.layers[i];
    P.layers[i].woff = woff32[i];
  }
  int grid = P.n_tiles < num_sms() ? P.n_tiles : num_sms();
  grid = (grid + 1) & ~1;
  if (grid > num_sms()) grid = num_sms() & ~1;
  const bool x3 = precision == DFN_PREC_BF16X3 || precision == DFN_PREC_FP16X3M;
  // folded density head: packed 16-bit row in shared memory on the single-pass Decoder kernels; fp32 row + bias (dot_w, dot_b) read
  // through L1 by the split kernels' column-distributed epilogue
  for (int i = 0; i < prog.n_layers; ++i)
    if ((prog.layers[i].flags & TC_F_DOT_SIGMA) && (!decoder || dot_w == nullptr || (x3 && dot_b == nullptr))) {
      set_error("pp_launch_prog: a program with a folded density head needs the Decoder kernels and its head row");
      return DFN_E_STATE;
    }
  if (precision == DFN_PREC_FP16X3M) {   // w_hi / w_lo: the fp16 stages and their fp16 residuals; TC_F_SINGLE set by the caller
    if (decoder) {
      set_error("pp_launch_prog: DFN_PREC_FP16X3M is wired for FaceNeRF / NeRF only");
      return DFN_E_UNSUPPORTED;
    }
    return pp_launch_t<true, false, true>(P, grid, st);
  }
  if (precision == DFN_PREC_FP16) {   // w_hi: the fp16 stages (Decoder programs; FaceNeRF / NeRF use mlp_pair.cu)
    if (!decoder) {
      set_error("pp_launch_prog: DFN_PREC_FP16 is wired for the Decoder programs only");
      return DFN_E_UNSUPPORTED;
    }
    return pp_launch_t<false, true, true>(P, grid, st);
  }
  if (decoder) return x3 ? pp_launch_t<true, true>(P, grid, st) : pp_launch_t<false, true>(P, grid, st);
  return x3 ? pp_launch_t<true, false>(P, grid, st) : pp_launch_t<false, false>(P, grid, st);
}

int pp_launch(const dfn_model* m, const float* bias_ws, const float* vbias_ws, void* scratch, int64_t R, int S,
              const float* rays_o, const float* rays_d, const float* z_vals, float* raw, int precision,
              cudaStream_t st) {
  if (precision == DFN_PREC_FP16X3M) {
    // fp16 operands; three products (hi hi + lo hi + hi lo) only where the density is formed after the skip connection -- the trunk
    // layers from the one that consumes the skip on, and views_linears.0 whose MMA carries alpha_linear as an output row -- one
    // product everywhere else (profiles/precision_emulation.py --sweep: those four layers carry the whole sensitivity).
    TcProgram prog = m->prog;
    const int D = m->desc.D, skip = m->desc.skip;
    // bit 2 of the schedule switches: alpha_linear in fp32 inside the epilogue of the last trunk layer (TC_F_DOT_ALPHA), which leaves
    // views_linears.0 single-pass as well (its density row is computed and ignored)
    const bool dot_alpha = (g_pp_flags & 4) != 0 && m->alpha.w != nullptr && m->alpha.b != nullptr && m->desc.W == 256 && D >= 2 &&
                           prog.layers[D - 1].epi == TC_EPI_RELU && prog.layers[D].epi == TC_EPI_VIEW0;
    for (int i = 0; i < prog.n_layers; ++i)
      if (!(i > skip && i <= (dot_alpha ? D - 1 : D))) prog.layers[i].flags |= TC_F_SINGLE;
    if (dot_alpha) prog.layers[D - 1].flags |= TC_F_DOT_ALPHA;
    return pp_launch_prog(prog, m->tc32_woff, m->tc_h16, m->tc_l16, dot_alpha ? m->alpha.w : nullptr, false, m->desc.multires,
                          m->desc.multires_views, m->desc.W / 2, bias_ws, vbias_ws, scratch, R, S, rays_o, rays_d, z_vals, raw, precision, st,
                          dot_alpha ? m->alpha.b : nullptr);
  }
  return pp_launch_prog(m->prog, m->tc32_woff, m->tc_hi, m->tc_lo, nullptr, false, m->desc.multires, m->desc.multires_views, m->desc.W / 2,
                        bias_ws, vbias_ws,
                        scratch, R, S, rays_o, rays_d, z_vals, raw, precision, st);
}

}  // namespace dfn
