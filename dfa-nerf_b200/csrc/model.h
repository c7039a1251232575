// Model handle shared by the fp32 FFMA path, the tcgen05 path and the C ABI.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#include <vector>

#include "../../include/dfn.h"

namespace dfn {

struct Fp32Layer {
  const float* w = nullptr;  // [out, in] row-major (nn.Linear layout), device
  const float* b = nullptr;  // [out], device
  int in = 0, out = 0;
};

// ---- tcgen05 path: per-layer program -------------------------------------------------------
// A operand blocks ("K-blocks") are [128 rows x 64 bf16] tiles in the 128-byte-swizzled K-major
// layout.  Per tile slot: blocks 0..3 hold the 256-wide hidden state h, block 4 the positional
// encoding of xyz (63 columns + one zero column).  Block ids >= TC_KB_PE are "staged" inputs: in mlp_pp.cu
// they live in an L2-resident scratch and pass through a weight-ring entry right before the layer that
// consumes them (TC_KB_PE: the positional encoding; TC_KB_IN1: the deformed per-sample signal of the torso
// field, DEC:297-299; TC_KB_DIR: the Decoder's view-direction encoding, DEC:337-338).  A layer has at most one staged
// input block.
enum { TC_KB_H0 = 0, TC_KB_PE = 4, TC_KB_IN1 = 5, TC_KB_DIR = 6, TC_KB_PER_TILE = 5 };
enum {
  TC_EPI_RELU = 0,   // + bias, relu -> hidden blocks H0..
  TC_EPI_VIEW0 = 1,  // + per-ray bias, relu -> hidden blocks (FaceNeRF/NeRF: density rides as column view_w)
  TC_EPI_RGB = 2,    // last layer: (rgb, density) -> raw[pt]
  TC_EPI_SIGMA = 3,  // Decoder: column 0 + bias -> density register, nothing stored (DEC:329)
  TC_EPI_CONT = 4,   // no epilogue: the next layer accumulates onto this one's partial sums (second staged input)
  TC_EPI_STAGE = 5   // + bias, no activation -> staged blocks PE / IN1 of the tile (deformation output, DEC:299)
};
enum {
  TC_F_ACCUM = 1,      // TcLayer.flags: the first MMA accumulates (layer continues a TC_EPI_CONT layer)
  // Decoder, single-pass precisions: sigma_out (DEC:329) is evaluated on the CUDA cores inside the epilogue of the block
  // that produces its input -- an M=128 x K=16 MMA costs the same for N=16 as for N=256, and every layer is one more
  // MMA -> epilogue -> barrier step on the tile's dependency chain.  dot_w = the head row [256] + its bias.
  TC_F_DOT_SIGMA = 2,  // TC_EPI_RELU layer: density register = dot_w[0..255] . relu(out) + dot_w[256]
  // split modes (mlp_pp.cu, X3): this layer issues the hi x hi product only (DFN_PREC_FP16X3M: everything but the layers that form the
  // density after the skip connection)
  TC_F_SINGLE = 4,
  // split modes, FaceNeRF / NeRF (DFN_PREC_FP16X3M): alpha_linear (HELP:286 / 383) is evaluated in fp32 on the CUDA cores inside the
  // epilogue of the last trunk layer, from its fp32 activations -- so views_linears.0, whose MMA otherwise carries the density as an
  // output row and needs the three split products for it, runs single-pass.  dot_w = alpha_linear.weight [W], dot_b = its bias.
  TC_F_DOT_ALPHA = 8
};
static constexpr int TC_DOT_FLOATS = 256 + 4;
static constexpr int TC_MAX_LAYERS = 20;
static constexpr int TC_BIAS_STRIDE = 256;  // floats per layer in the bias blob

struct TcLayer {
  uint32_t woff;     // byte offset of this layer's first weight stage in the (hi) blob
  uint16_t n;        // output columns computed by the MMAs (multiple of 16)
  uint8_t nkb;       // number of input K-blocks
  uint8_t kb[6];     // their block ids, in weight order
  uint8_t epi;       // TC_EPI_*
  uint8_t flags;     // TC_F_*
  uint8_t pad;
};
static_assert(sizeof(TcLayer) == 16, "TcLayer is passed by value in kernel parameters");

struct TcProgram {
  int n_layers = 0;
  TcLayer layers[TC_MAX_LAYERS];
  int fold_layer[2] = {-1, -1};  // layers whose bias absorbs W[:, latent cols] @ latent
};

}  // namespace dfn

struct dfn_model {
  dfn_model_desc desc;
  bool loaded = false;
  int n_views = 0;  // number of views_linears
  // fp32 path
  float* fp32_blob = nullptr;  // one device allocation holding all weights and biases
  dfn::Fp32Layer pts[16];
  dfn::Fp32Layer views[8];
  dfn::Fp32Layer feature, alpha, rgb;
  // tcgen05 path
  dfn::TcProgram prog;
  uint32_t tc32_woff[dfn::TC_MAX_LAYERS] = {};  // per-layer offsets into tc_hi / tc_lo
  uint8_t* tc_hi = nullptr;     // packed bf16 (hi) weight stages
  uint8_t* tc_lo = nullptr;     // packed bf16 (lo = bf16(w - hi)) weight stages, same offsets
  uint8_t* tc_h16 = nullptr;    // packed fp16 weight stages (DFN_PREC_FP16), same offsets
  uint8_t* tc_l16 = nullptr;    // fp16 residuals fp16(w - fp16(w)) (DFN_PREC_FP16X3M), same offsets
  int64_t tc_blob_bytes = 0;
  float* tc_bias = nullptr;     // [n_layers][256] static biases
  float* tc_fold_w = nullptr;   // [2][W][dim_aud] latent columns of the two folding layers
  float* tc_view_w = nullptr;   // [W/2][input_ch_views] view-direction columns of views_linears.0
  float* tc_view_b = nullptr;   // [W/2] its (composed) bias
  // cta_group::2 kernel (mlp_pair.cu): per CTA of the pair, per K-block, [n/2 rows x 64 K] stages
  uint8_t* tc2_hi = nullptr;
  uint8_t* tc2_h16 = nullptr;   // fp16 images, same offsets
  uint32_t tc2_woff[dfn::TC_MAX_LAYERS] = {};
};

namespace dfn {
int64_t mlp_fp32_workspace_bytes(const dfn_model* m, int64_t P);
int mlp_fp32_forward(const dfn_model* m, int64_t P, const float* x, float* out, void* workspace,
                     int64_t workspace_bytes, cudaStream_t st);

void tc_set_trace(void* dev_ptr, int tiles);
void tc_get_trace(void** dev_ptr, int* tiles);
void tc_set_impl(int impl);  // debug: 1 mlp_tc.cu, 2 mlp_pp.cu, 3 mlp_pair.cu (see tc_query_points)
int64_t pp_scratch_bytes();
int64_t pp_dec_scratch_bytes();
int pp_launch_prog(const TcProgram& prog, const uint32_t* woff32, const uint8_t* w_hi, const uint8_t* w_lo, const float* dot_w, bool decoder,
                   int multires, int multires_views, int view_w, const float* bias_ws, const float* vbias_ws, void* scratch, int64_t R, int S,
                   const float* rays_o, const float* rays_d, const float* z_vals, float* raw, int precision,
                   cudaStream_t st, const float* dot_b = nullptr);
int pp_launch(const dfn_model* m, const float* bias_ws, const float* vbias_ws, void* scratch, int64_t R, int S,
              const float* rays_o, const float* rays_d, const float* z_vals, float* raw, int precision,
              cudaStream_t st);
int pair_launch(const dfn_model* m, const float* bias_ws, const float* vbias_ws, int64_t R, int S, const float* rays_o,
                const float* rays_d, const float* z_vals, float* raw, int precision, cudaStream_t st);
int pair_launch_prog(const TcProgram& prog, const uint32_t* woff2, const uint8_t* w, bool f16, bool decoder, int multires, int multires_views,
                     int view_w, const float* bias_ws, const float* vbias_ws, const float* dot_w, void* scratch, int64_t R, int S,
                     const float* rays_o, const float* rays_d, const float* z_vals, float* raw, cudaStream_t st);
void pair_set_epilogue_warps(int ew);
int pair_get_flags();
void pair_set_flags(int flags);
void pp_set_flags(int flags);   // mlp_pp.cu schedule switches (dfn_debug_set_pp_flags)
int pp_get_flags();
struct TcHostDump {   // host-only view of a packed model (dfn_model_program_host)
  std::vector<float> dense;    // [layer][256][6][64]
  std::vector<float> bias;     // [TC_MAX_LAYERS][256]
  std::vector<float> fold_w;   // [2][W][dim_aud]
  std::vector<float> view_w;   // [W/2][input_ch_views]
  std::vector<float> view_b;   // [W/2]
};
int tc_pack_model(dfn_model* m, const float* const* t, cudaStream_t st, TcHostDump* dump = nullptr);
void tc_free_model(dfn_model* m);
int64_t tc_query_workspace_bytes(const dfn_model* m, int64_t R, int S);
int tc_query_points(const dfn_model* m, int64_t R, int S, const float* rays_o, const float* rays_d,
                    const float* viewdirs, const float* z_vals, const float* latent, float* raw,
                    int precision, void* workspace, int64_t workspace_bytes, cudaStream_t st, const void* pre = nullptr);
int64_t tc_prep_bytes(const dfn_model* m, int64_t R);
int tc_prep_launch(const dfn_model* a, const dfn_model* b, int64_t R, int Nc, const float* viewdirs, const float* latent,
                   const float* t_vals, const float* near, const float* far, const float* rnd, float* z0, void* pre_a, void* pre_b,
                   cudaStream_t st);
}  // namespace dfn
