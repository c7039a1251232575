/*
 * dfn.h -- C ABI of libdfn.so, the B200 (sm_100a) volume-rendering hot path for DFA-NeRF.
 *
 * The reference (ShunyuYao/DFA-NeRF) has no plugin / FFI layer for this path: it is plain
 * PyTorch called from NeRFs/DFANeRF/run_nerf_com_trainExpLater.py.
 * Each entry point below therefore cites the reference *function* it replaces; the
 * Python shim in dfa-nerf_b200/ binds these with ctypes and re-exports the reference's
 * own names (INTEGRATION.md shows the stub a maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in _host;
 *   - tensors are dense row-major fp32 (indices int64) with the shapes given;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); nothing
 *     synchronises, nothing allocates: the caller owns outputs and workspaces;
 *   - return value: 0 = ok, >0 = cudaError_t, <0 = DFN_E_* below; dfn_last_error()
 *     returns a thread-local message for the last non-zero return;
 *   - HELP = NeRFs/DFANeRF/run_nerf_helpers.py, MAIN = NeRFs/DFANeRF/run_nerf_com_trainExpLater.py,
 *     DEC = NeRFs/DFANeRF/decoder.py (paths relative to the reference root).
 */
#ifndef DFN_H_
#define DFN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFN_ABI_VERSION 1

enum {
  DFN_E_ARG = -1,         /* bad argument (null pointer, size, unsupported shape) */
  DFN_E_STATE = -2,       /* model not loaded / wrong kind */
  DFN_E_WORKSPACE = -3,   /* workspace too small */
  DFN_E_UNSUPPORTED = -4  /* configuration the kernels do not cover */
};

/* Arithmetic used for the MLP layers. */
enum {
  DFN_PREC_FP32 = 0,   /* fp32 FFMA kernels (no tensor cores), reference-exact up to summation order */
  DFN_PREC_BF16 = 1,   /* tcgen05 kind::f16 bf16 operands, fp32 accumulate in TMEM (throughput mode) */
  DFN_PREC_FP16 = 2,   /* same kernel and speed with fp16 operands (11-bit significands, saturating): ~10x closer to
                          fp32 than bf16 */
  DFN_PREC_BF16X3 = 3, /* tcgen05 split-bf16: A_hi*W_hi + A_lo*W_hi + A_hi*W_lo (fp32-parity mode) */
  DFN_PREC_FP16X3M = 4 /* FaceNeRF / NeRF: fp16 operands, the three split products only on the layers that form the density after the
                          skip connection (trunk layers skip+1 .. D-1 and views_linears.0 + alpha_linear), one product elsewhere: the
                          fast fp32-parity mode (teacher-forced RGB <= 1e-4; ~1.8x the single-pass MMA work instead of 3x) */
};

enum {
  DFN_MODEL_FACENERF = 0, /* HELP:242-299 FaceNeRF(use_viewdirs=True) */
  DFN_MODEL_NERF = 1      /* HELP:342-396 NeRF(use_viewdirs=True)     */
};

int dfn_abi_version(void);
const char* dfn_last_error(void);

/* ---- a1  get_rays  (HELP:449-465) -------------------------------------------------------
 * xs[n_cols], ys[n_rows]: the torch.linspace pixel-centre tables (W//stride, H//stride entries);
 * c2w_host: 12 floats, row-major [3,4], HOST memory.  Outputs [n_rows*n_cols,3]; rays_o is the
 * broadcast origin; viewdirs (nullable) = rays_d/||rays_d|| (upstream render(); DEC:337). */
int dfn_get_rays(int n_rows, int n_cols, const float* xs, const float* ys, float focal, float cx,
                 float cy, const float* c2w_host, float* rays_o, float* rays_d, float* viewdirs,
                 void* stream);

/* ---- a2  depth sampling  (MAIN:617-619; upstream stratified block) ------------------------
 * z[r,s] = near[r]*(1-t[s]) + far[r]*t[s];  rand (nullable, [R,S]) applies the stratified jitter. */
int dfn_z_vals(int R, int S, const float* t_vals, const float* near, const float* far,
               const float* rand, float* z_out, void* stream);

/* ---- a3  sample points  (MAIN:638-641) -----------------------------------------------------------------
 * pts[r,s,:] = rays_o[r] + rays_d[r]*z_vals[r,s]; dirs[r,s,:] = rays_d[r].  Either output may be null.  Only the
 * explicit-points interface of the Decoder needs these tensors; dfn_query_points never materialises them. */
int dfn_make_points(int R, int S, const float* rays_o, const float* rays_d, const float* z_vals, float* pts,
                    float* dirs, void* stream);

/* ---- a4 / a4'  positional encodings  (HELP:21-70 Embedder; DEC:257-275 transform_points) ---
 * kind 0: [x, sin(2^k x), cos(2^k x)]_k  -> out [P, 3+6L];
 * kind 1: p/=2; [sin(2^k pi p), cos(2^k pi p)]_k -> out [P, 6L];
 * kind 2: kind 1 of x/||x|| (view directions, DEC:337-338). */
int dfn_embed(int64_t P, const float* x, int L, int kind, float* out, void* stream);

/* ---- a7  composite_function  (MAIN:146-166) -------------------------------------------------
 * sigma [n_box,n], feat [n_box,n,3] -> sigma_sum [n], feat_w [n,3]. */
int dfn_composite_fields(int n_box, int64_t n, const float* sigma, const float* feat,
                         float* sigma_sum, float* feat_w, void* stream);

/* ---- a8  calc_volume_weights  (MAIN:169-179) ------------------------------------------------ */
int dfn_calc_volume_weights(int R, int S, const float* z_vals, const float* ray_vector,
                            const float* sigma, float last_dist, float* weights, void* stream);

/* ---- a8+a9  raw2outputs  (upstream name; core = MAIN:169-179 + MAIN:706) ----------------------
 * raw [R,S,4] = (rgb pre-sigmoid, sigma).  bc_rgb (nullable, [R,3]) replaces the last sample's
 * colour (MAIN:669-671).  raw_is_feat != 0: channels 0..2 are already colours (Decoder head, DEC:346).
 * Outputs (each nullable): rgb_map [R,3], disp_map [R], acc_map [R], weights [R,S], depth_map [R]. */
int dfn_raw2outputs(int R, int S, const float* raw, const float* z_vals, const float* rays_d,
                    const float* bc_rgb, int raw_is_feat, int white_bkgd, float last_dist,
                    float* rgb_map, float* disp_map, float* acc_map, float* weights,
                    float* depth_map, void* stream);

/* ---- a6+a7+a8+a9  live two-field compositing of one chunk  (MAIN:669-708, concate_bg) ------------------
 * feat_* [R,S,3] (sigmoid already applied, DEC:346), sigma_* [R,S] (raw, relu applied here, MAIN:688-689),
 * bc_rgb [R,3], z_vals [R,S], rays_d_* [R,3] -> rgb_head [R,3] (head field alone), rgb_person [R,3]
 * (density-weighted head+torso mix).  Either output may be null. */
int dfn_composite_head_torso(int R, int S, const float* feat_head, const float* sigma_head,
                             const float* feat_torso, const float* sigma_torso, const float* bc_rgb,
                             const float* z_vals, const float* rays_d_head, const float* rays_d_torso,
                             float last_dist, float* rgb_head, float* rgb_person, void* stream);

/* ---- a10  sample_pdf  (HELP:537-581) --------------------------------------------------------
 * bins [R,nb], weights [R,nb-1] (row stride w_stride floats, so a [R,S] weights tensor can be
 * passed as weights+1 with w_stride=S for the weights[...,1:-1] slice), u [N] (u_per_ray=0) or
 * [R,N] (u_per_ray=1).  samples [R,N]; inds (nullable) int64 [R,N] = searchsorted(cdf,u,right). */
int dfn_sample_pdf(int R, int nb, const float* bins, const float* weights, int64_t w_stride,
                   int N, const float* u, int u_per_ray, float* samples, int64_t* inds,
                   void* stream);

/* Inversion only, with the cdf [R,nb] given (HELP:563-579): used to check indices bit-for-bit. */
int dfn_invert_cdf(int R, int nb, const float* bins, const float* cdf, int N, const float* u,
                   int u_per_ray, float* samples, int64_t* inds, void* stream);

/* ---- a11  merge  (upstream: z_vals,_ = sort(cat([z_vals, z_samples]))) ------------------------ */
int dfn_sort_merge(int R, int na, const float* a, int nb, const float* b, float* out, void* stream);

/* ---- a8+a9+a10+a11 fused: coarse pass -> fine depths  (upstream render_rays) ------------
 * raw2outputs(raw0) -> z_mid = .5 (z[1:] + z[:-1]) -> sample_pdf(z_mid, weights[..., 1:-1], N_importance, u) ->
 * z_all = sort(cat(z_vals, z_samples)) in ONE launch; weights, midpoints, cdf and samples stay in shared memory.  The
 * arithmetic is that of dfn_raw2outputs / dfn_sample_pdf / dfn_sort_merge, operation for operation: z_all is bit-identical
 * to the chain of the three.  rgb0 [R,3] (coarse image) and z_samples_out [R,N_importance] are optional outputs;
 * z_samples_in (nullable) injects the new depths instead of sampling them (teacher forcing).  u: [N_importance]
 * (u_per_ray 0) or [R,N_importance]. */
int dfn_coarse_to_fine(int R, int N_samples, int N_importance, const float* raw0, const float* z_vals, const float* rays_d,
                       const float* bc_rgb, int white_bkgd, float last_dist, const float* u, int u_per_ray,
                       const float* z_samples_in, float* rgb0, float* z_samples_out, float* z_all, void* stream);

/* ---- output side  to8b  (HELP:17; MAIN:714-715) ---------------------------------------------
 * out[i] = uint8(255 * clip(x[i], 0, 1)) (fp32 product, truncation) -- the frame leaves the device as 3 bytes/pixel. */
int dfn_to8b(int64_t n, const float* x, uint8_t* out, void* stream);

/* ---- a5 / a5'  the 8x256 skip-MLP  (HELP:242-299 FaceNeRF, HELP:342-396 NeRF) ------------------ */
typedef struct dfn_model dfn_model;

typedef struct {
  int kind;           /* DFN_MODEL_* */
  int D;              /* 8 */
  int W;              /* 256 */
  int input_ch;       /* 63 = 3+6*multires */
  int input_ch_views; /* 27 = 3+6*multires_views */
  int dim_aud;        /* 64 (FaceNeRF) / 0 (NeRF) */
  int skip;           /* 4 : layer index after which [input_pts, h] is concatenated */
  int multires;       /* 10 */
  int multires_views; /* 4 */
} dfn_model_desc;

int dfn_model_create(const dfn_model_desc* desc, dfn_model** out);
void dfn_model_destroy(dfn_model* m);

/* Number of tensors dfn_model_load expects, in this order (state_dict order of the reference):
 * pts_linears.{0..D-1}.{weight,bias}, views_linears.{0..nv-1}.{weight,bias},
 * feature_linear.{weight,bias}, alpha_linear.{weight,bias}, rgb_linear.{weight,bias}
 * with nv = 1 + D/4 for FaceNeRF (HELP:265-266) and 1 for NeRF (HELP:358-359). */
int dfn_model_num_tensors(const dfn_model* m);

/* tensors_host[i]: HOST fp32 pointers in the order above; repacks to the kernel layouts
 * (fp32 row-major for the FFMA path; bf16 hi/lo planes, K padded to 64, 128-byte swizzled
 * K-major tiles for the tcgen05 path) and uploads on `stream`. */
int dfn_model_load(dfn_model* m, const float* const* tensors_host, int n_tensors, void* stream);

/* One fused nn.Linear of the fp32 path: Y[p,n] = act(sum_k X(p,k) W[n,k] + bias[n]) (+ addend[p*ld_add + n]);
 * X(p,k) = k < K1 ? X1[p*ld1 + k] : X2[p*ld2 + k-K1] (concatenated inputs are never materialised; a row
 * stride of 0 broadcasts one row, e.g. the per-frame signal of DEC:293-295).  act & 3: 0 none, 1 relu,
 * 2 sigmoid, 3 LeakyReLU(0.02); act & 4: addend is added before the activation instead of after it.
 * bias / X2 / addend may be null.  Building block of the Decoder / DeformationField_ori shim (DEC:109-349). */
int dfn_linear(int64_t P, int N, int K1, const float* X1, int64_t ld1, int K2, const float* X2, int64_t ld2,
               const float* W, const float* bias, int act, const float* addend, int64_t ld_add, float* Y,
               int64_t ldy, void* stream);

/* Module forward on explicit embedded inputs: x [P, input_ch+dim_aud+input_ch_views] -> out [P,4]
 * (HELP:275-299 / HELP:372-396).  fp32 FFMA path.  workspace: dfn_mlp_workspace_bytes(P) bytes. */
int64_t dfn_mlp_workspace_bytes(const dfn_model* m, int64_t P);
int dfn_mlp_forward(const dfn_model* m, int64_t P, const float* x, float* out, void* workspace,
                    int64_t workspace_bytes, void* stream);

/* network_query_fn (upstream run_network): pts = rays_o + rays_d*z, PE(pts) | latent | PE(viewdir)
 * -> model, fused.  rays_o/rays_d/viewdirs [R,3], z_vals [R,S], latent [dim_aud] (nullable for
 * NeRF) -> raw [R,S,4].  precision: DFN_PREC_*.  workspace: dfn_query_workspace_bytes(R,S). */
int64_t dfn_query_workspace_bytes(const dfn_model* m, int64_t R, int S, int precision);
int dfn_query_points(const dfn_model* m, int64_t R, int S, const float* rays_o, const float* rays_d,
                     const float* viewdirs, const float* z_vals, const float* latent, float* raw,
                     int precision, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- a5''  Decoder + DeformationField_ori on the tensor cores  (DEC:77-134, DEC:137-349) ---------------
 * The reference's LIVE model (MAIN:518: Decoder(z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field)).
 * dfn_decoder_query is the fused network query of one field for pts = rays_o + rays_d*z: positional encoding of
 * DEC:257-275, per-frame latents (signal, z_shape, z_app) folded into biases, the view-direction encoding as an input block, the torso's
 * deformation field in front of its trunk; raw [R,S,4] = (feat after the sigmoid of DEC:346-347, sigma before the relu
 * of MAIN:688).  rays_d is NOT normalised by the caller (DEC:337 normalises the view direction itself). */
typedef struct dfn_decoder dfn_decoder;

typedef struct {
  int hidden;        /* 256 */
  int z_dim;         /* 256 */
  int dim_signal;    /* 96  head signal (scripts/test_obama.sh:8) */
  int dim_et_embed;  /* 42  torso signal */
  int n_freq;        /* 10  n_freq_posenc */
  int n_freq_views;  /* 4   n_freq_posenc_views */
  int n_blocks;      /* 8 */
  int skip;          /* 4   skips=[4] */
} dfn_decoder_desc;

int dfn_decoder_create(const dfn_decoder_desc* desc, dfn_decoder** out);
void dfn_decoder_destroy(dfn_decoder* m);

/* 64 HOST fp32 tensors, {weight, bias} of the reference's modules in this order:
 * deform_net.blocks_embed.{0..4}, deform_net.out_embed, deform_net.blocks_signal.{0..4}, deform_net.out_signal,
 * deform_net.fc_embed_skips.0, deform_net.fc_signal_skips.0, fc_in, fc_in_torso, fc_z, blocks.{0..6}, fc_z_skips.0,
 * fc_p_skips.0, fc_p_skips_torso.0, sigma_out, fc_z_view, feat_view, fc_view, feat_out. */
int dfn_decoder_num_tensors(const dfn_decoder* m);
int dfn_decoder_load(dfn_decoder* m, const float* const* tensors_host, int n_tensors, void* stream);

/* field: 0 head (signal [dim_signal]), 1 torso (signal [dim_et_embed]); z_shape, z_app [z_dim];
 * precision: DFN_PREC_BF16, DFN_PREC_FP16 or DFN_PREC_BF16X3. */
int64_t dfn_decoder_query_workspace_bytes(const dfn_decoder* m, int64_t R, int S);
int dfn_decoder_query(const dfn_decoder* m, int field, int64_t R, int S, const float* rays_o, const float* rays_d,
                      const float* z_vals, const float* z_shape, const float* z_app, const float* signal, float* raw,
                      int precision, void* workspace, int64_t workspace_bytes, void* stream);
/* The same with the head's expression term (use_expression, DEC:279-281, 333-334): view_term [hidden] = expnet(expression),
 * a per-frame row the caller forms with dfn_linear, added to the input of the view layer's relu; null = dfn_decoder_query. */
int dfn_decoder_query_ex(const dfn_decoder* m, int field, int64_t R, int S, const float* rays_o, const float* rays_d,
                         const float* z_vals, const float* z_shape, const float* z_app, const float* signal,
                         const float* view_term, float* raw, int precision, void* workspace, int64_t workspace_bytes,
                         void* stream);
/* algorithmic MACs per sample of a field with the per-frame and per-ray terms folded (roofline accounting) */
double dfn_decoder_macs_per_sample(const dfn_decoder* m, int field);

/* Host-only introspection (no CUDA calls; used by the CPU tests): the layer program dfn_decoder_load compiles for one
 * field.  layers[l]: output columns n, input K-blocks kb[0..nkb) (0..3 hidden blocks, 4 = positional encoding / deformed
 * encoding, 5 = deformed signal, 6 = view-direction encoding), epilogue (0 relu, 1 per-ray-bias relu, 2 rgb+sigmoid, 3 sigma, 4 continue, 5 write
 * staged blocks) and flags (1 = accumulates onto the previous layer; 2 = folded-head program: the epilogue also forms the
 * density dot_w[0..255] . relu(out) + dot_w[256]).  folded_heads = 0: the program of DFN_PREC_BF16X3 (sigma_out as a
 * 16-column layer); 1: the program of DFN_PREC_BF16 / DFN_PREC_FP16, with dot_w [260] filled.  weights: dense fp32 [max_layers][256][6*64]
 * (row n, input block slot i, position k), bias [max_layers][256], fold_w [n_fold][dimL][256] with
 * bias[fold_layer[i]][n] += sum_j fold_w[i][j][n] * latent[j], latent = [signal | z_shape | z_app].
 * max_layers >= 20; fold_layer has 8 entries; fold_w holds 8*dimL*256 floats (dimL <= 1024). */
typedef struct {
  int n, nkb, epi, flags;
  int kb[6];
} dfn_layer_info;
int dfn_decoder_program_host(const dfn_decoder_desc* desc, const float* const* tensors_host, int n_tensors, int field,
                             int folded_heads, int max_layers, dfn_layer_info* layers, int* n_layers, float* weights,
                             float* bias, int* n_fold, int* fold_layer, float* fold_w, int* dimL, int* view_layer,
                             float* dot_w);

/* The same for a FaceNeRF / NeRF model (m: created, weights not needed): layers as above (epilogue 1 = views_linears.0
 * with the density head as column W/2 and the per-ray view term as bias); fold_layer[2] / fold_w [2][W][dim_aud]: the
 * latent columns of the two layers that read the input (bias[fold_layer[i]][n] += fold_w[i][n][:] . latent); view_w
 * [W/2][input_ch_views], view_b [W/2]: the per-ray term of views_linears.0 (for NeRF with feature_linear composed in). */
int dfn_model_program_host(const dfn_model* m, const float* const* tensors_host, int n_tensors, int max_layers,
                           dfn_layer_info* layers, int* n_layers, float* weights, float* bias, int* fold_layer,
                           float* fold_w, float* view_w, float* view_b);

/* ---- one chunk of the live render loop  (MAIN:617-619, MAIN:633-708) ----------------------------------
 * z sampling -> head field on the head-pose rays, torso field (with deformation) on the body-pose rays ->
 * background splice, two-field density mix, weights, colour sums.  z_shape / z_app: [2, z_dim] (row 0 head,
 * row 1 torso, MAIN:664-674).  Outputs (each nullable): rgb_head [R,3], rgb_person [R,3]. */
typedef struct {
  const float* rays_o_head;
  const float* rays_d_head;
  const float* rays_o_torso;
  const float* rays_d_torso;
  const float* near;
  const float* far;
  const float* t_vals;
  const float* bc_rgb;
  const float* z_shape;
  const float* z_app;
  const float* signal;
  const float* signal_torso;
  float* rgb_head;
  float* rgb_person;
  float last_dist; /* <= 0: 1e10 (MAIN:169) */
  const float* expression_term; /* head field: expnet(expression) [hidden] (DEC:279-281), or null */
} dfn_head_torso_io;

int64_t dfn_render_head_torso_workspace_bytes(const dfn_decoder* m, int64_t R, int S);
int dfn_render_head_torso(const dfn_decoder* m, int64_t R, int S, const dfn_head_torso_io* io, int precision,
                          void* workspace, int64_t workspace_bytes, void* stream);

/* ---- latent encoders: the callers of the path  (HELP:109-240, MAIN:28-111) -----------
 * The MLP encoders (AudioNet_W2L HELP:165, ExpressionEnc HELP:182) are chains of dfn_linear with act = 3.
 * A conv stack is `n` Conv1d(kernel 3, padding 1, one common stride) layers, each followed by LeakyReLU(0.02);
 * w[i] is [ch[i+1], ch[i], 3], b[i] is [ch[i+1]]; all DEVICE pointers. */
typedef struct {
  int n;             /* number of conv layers (<= 6) */
  int ch[7];         /* channels: ch[0] input ... ch[n] output */
  int stride;        /* 2 (AudioNet) or 1 (AudioAttNet) */
  const float* w[6];
  const float* b[6];
} dfn_conv_stack;

/* AudioNet.forward (HELP:133-141): x [N,16,29] (DeepSpeech windows) -> out [N,dim_aud]; four stride-2 convs
 * (29->32->32->64->64 over 16->8->4->2->1 steps), Linear(64,64)+LeakyReLU, Linear(64,dim_aud).  win_size must be 16. */
int dfn_audionet_forward(int N, int dim_aud, const float* x, const dfn_conv_stack* convs, const float* fc1_w,
                         const float* fc1_b, const float* fc2_w, const float* fc2_b, float* out, void* stream);

/* AudioAttNet.forward over a whole sequence (HELP:232-240 with the window logic of MAIN:35-61 / MAIN:85-101):
 * feats [N,D] per-frame features; for frame i the window is rows [i-seq_len/2, i+seq_len/2), rows outside [0,N)
 * replaced by pad_row [D] (the features of an all-zero input, as the reference pads before encoding); attention
 * weights from the first dim_att columns (five stride-1 convs, Linear(seq,seq), softmax) applied to all D columns.
 * out [N,D].  seq_len even, <= 16; D <= 128. */
int dfn_att_smooth(int N, int D, int dim_att, int seq_len, const float* feats, const float* pad_row,
                   const dfn_conv_stack* convs, const float* lin_w, const float* lin_b, float* out, void* stream);

/* pose_to_euler_trans + Embedder (MAIN:182-205, MAIN:106-109): poses [N, pose_stride floats] row-major 3x4 or 4x4
 * (pose_stride 12 or 16) -> out [N, 2*(3+6L)] = [embed_L(euler) | embed_L(trans)]; et_out (nullable) [N,6] = euler|trans. */
int dfn_pose_signal(int N, const float* poses, int pose_stride, int L, float* out, float* et_out, void* stream);

/* ---- render_rays  (upstream name; MAIN:114 is the reference's dead stub) -----------------------
 * coarse pass (N_samples) -> raw2outputs -> sample_pdf(z_mid, w[1:-1], N_importance) -> sort-merge
 * -> fine pass (N_samples+N_importance) with `fine` (or `coarse` when null) -> raw2outputs.
 * N_importance = 0 stops after the coarse pass.
 * Inputs: rays_o, rays_d, viewdirs [R,3]; near, far [R]; bc_rgb [R,3] (nullable); latent [dim_aud];
 *         t_vals [N_samples] and u_vals [N_importance] = torch.linspace(0,1,.) tables;
 *         perturb_rand (nullable, [R,N_samples]); z_samples_in (nullable, [R,N_importance]):
 *         teacher-forced fine depths that replace the sample_pdf output.
 * Outputs (each nullable): rgb_map [R,3], disp_map [R], acc_map [R], last_weight [R],
 *         rgb0 [R,3], z_samples_out [R,N_importance], z_vals_out [R,N_samples+N_importance]. */
typedef struct {
  const float* rays_o;
  const float* rays_d;
  const float* viewdirs;
  const float* near;
  const float* far;
  const float* bc_rgb;
  const float* latent;
  const float* t_vals;
  const float* u_vals;
  const float* perturb_rand;
  const float* z_samples_in;
  float* rgb_map;
  float* disp_map;
  float* acc_map;
  float* last_weight;
  float* rgb0;
  float* z_samples_out;
  float* z_vals_out;
  int64_t u_per_ray; /* 0: u_vals is [N_importance]; 1: u_vals is [R,N_importance] (perturb > 0) */
} dfn_render_io;

int64_t dfn_render_workspace_bytes(const dfn_model* coarse, int64_t R, int N_samples,
                                   int N_importance, int precision);
/* The same for a coarse / fine pair whose networks differ in shape (width, latent size): the query scratch and the
 * prepared-bias regions are sized for the larger of the two.  fine == NULL: the coarse network runs both passes.  This is
 * the size dfn_render_rays checks its workspace against. */
int64_t dfn_render_workspace_bytes2(const dfn_model* coarse, const dfn_model* fine, int64_t R, int N_samples,
                                    int N_importance, int precision);
int dfn_render_rays(const dfn_model* coarse, const dfn_model* fine, int64_t R, int N_samples,
                    int N_importance, const dfn_render_io* io, int white_bkgd, int precision,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* Number of kernels the last dfn_render_rays / dfn_query_points call on this thread launched. */
int dfn_last_launch_count(void);

/* Measurement hook for bench.py: while enabled, every launch of the tcgen05 MLP kernel is bracketed
 * by CUDA events on its own stream (up to 4096 launches).  dfn_profile_collect synchronises those
 * events, returns the summed kernel time (ms), the number of launches and the algorithmic MACs they
 * covered (constant-folded count), and resets the counters. */
int dfn_profile_enable(int on);

/* Debug timeline of the tcgen05 kernel: while dev_buffer is non-null, CTA 0 records clock64 stamps for its
 * first `tiles` tile iterations: [MMA issuer | epilogue] x [tile][layer][slot] x {wait begin, wait end,
 * done, aux} (uint64 each; 2*tiles*20*2*4 entries at most).  Pass null to switch it off. */
int dfn_debug_trace(void* dev_buffer, int tiles);

/* Selects the tcgen05 kernel variant (A/B measurements): -1 (default) the fastest measured per precision (bf16 / fp16 -> 3,
 * bf16x3 / fp16x3m -> 2).  Low 4 bits: 1 the 1-CTA generation, two tiles in flight with per-tile epilogue warps (mlp_tc.cu);
 * 2 cooperative epilogue + PE through the weight ring, the split-precision schedule (mlp_pp.cu); 3 CTA pairs, cta_group::2
 * MMAs (mlp_pair.cu); 8 the same with four epilogue warps per slot.  Higher bits: (flags + 1) << 4 for the pair kernel --
 * flag 1 a layer's weights are loaded once for both slots, 2 CTA-scope release on the peer's arrivals, 4 the per-ray-bias
 * layer's two bias rows staged in shared memory and the column-distributed readout for it (default 7), 8 / 16 keep the Decoder
 * head / torso programs on mlp_pp.cu. */
int dfn_debug_set_impl(int impl);
/* Schedule switches of the split-precision kernel (mlp_pp.cu; A/B measurements).  Bit 0 (default on): a layer's staged input
 * block (the positional encoding of layer 0 and of the skip layer) is copied into its ring entry at the START of the previous
 * layer's epilogue, so that layer's MMAs overlap the epilogue; 0: copied at its end (the round-1 schedule; bit-identical output).
 * Bit 1 (default on): the issuer polls a K-block's weight-stage barrier before the activation block's.  Bit 2 (default on): the
 * density head (alpha_linear in DFN_PREC_FP16X3M, the Decoder's sigma_out in DFN_PREC_BF16X3) is evaluated in fp32 inside the last
 * trunk layer's epilogue.  Bit 3 (default on): the per-ray-bias layer reads its two bias rows from shared memory (bit-identical). */
int dfn_debug_set_pp_flags(int flags);
int dfn_profile_collect(double* kernel_ms, int64_t* launches, double* algorithmic_macs);

/* ---- f-3  training step  (MAIN:855-931: two-field forward on N_rand rays, img2mse x2, loss.backward(), Adam) ---------
 * The three GEMMs of every nn.Linear -- forward, data gradient, weight gradient (what torch.autograd runs inside
 * MAIN:923's loss.backward()) -- are ONE strided tcgen05 kernel: both operands carry a (row, k) stride pair, so no
 * transposed copy of an activation, gradient or weight is ever made; operands are converted fp32 -> split bf16 in the
 * kernel's loaders; the activation derivative is applied while the gradient operand is loaded.
 *   C[m,n] = act( sum_k A(m,k) B(n,k) + bias[n] (+ addend[m,n] if act & 4) ) (+ addend[m,n] otherwise) (+ C[m,n] if beta)
 *   A(m,k) = A[m*a_ld_r + k*a_ld_k] * mask'(A_mask[same index]);   B(n,k) = B[n*b_ld_r + k*b_ld_k];  N <= 256 per call.
 * act & 3 as dfn_linear (0 none, 1 relu, 2 sigmoid, 3 LeakyReLU(0.02)).  k_splits > 1 splits the contraction over CTAs
 * and accumulates with fp32 atomics into C (zeroed or running sums; needs act = 0, beta = 1).
 * precision: DFN_PREC_BF16X3 (A_hi B_hi + A_lo B_hi + A_hi B_lo: products good to ~2^-18), DFN_PREC_BF16, or DFN_PREC_FP32 (FFMA on
 * the CUDA cores: the reference-exact mode -- the tensor core's fp32 accumulation truncates, see gemm_tc.cu). */
enum { DFN_MASK_NONE = 0, DFN_MASK_RELU = 1, DFN_MASK_LEAKY = 2, DFN_MASK_SIGMOID = 3 };  /* act'(y) from the OUTPUT y */
typedef struct {
  const float* A; int64_t a_ld_r, a_ld_k;
  const float* A_mask; int a_mask_mode;       /* nullable; same strides as A */
  const float* B; int64_t b_ld_r, b_ld_k;
  float* C; int64_t c_ld_r, c_ld_c;
  const float* bias;                          /* [N], nullable */
  const float* addend; int64_t add_ld_r, add_ld_c;   /* nullable; a row stride of 0 broadcasts one row */
  int act;
  int M, N, K;
  int beta;
  int k_splits;
  int precision;
} dfn_gemm_desc;
int dfn_gemm(const dfn_gemm_desc* d, void* stream);

/* Bias gradient: out[n] += sum_m X[m*ld + n] * act'(Y[m*ld + n]) (Y nullable / mask_mode 0: plain column sums). */
int dfn_colsum(int64_t M, int N, const float* X, int64_t ld, const float* Y, int mask_mode, float* out, void* stream);

/* MAIN:884-907 and its backward: the live two-field compositing of dfn_composite_head_torso, the two image losses
 * loss2[0] += img2mse(rgb_head, target_head), loss2[1] += img2mse(rgb_person, target_person) (HELP:11; loss2 is
 * accumulated into, zero it first), and the gradient of their sum with respect to the fields' outputs:
 * dpre_* [R,S,3] = d loss / d (colour BEFORE the final sigmoid of DEC:346-347), dsigma_* [R,S] = d loss / d (raw density
 * before MAIN:688's relu).  rgb_head / rgb_person (nullable) receive the rendered pixels.  S <= 128. */
int dfn_head_torso_loss_bwd(int R, int S, const float* feat_head, const float* sigma_head, const float* feat_torso,
                            const float* sigma_torso, const float* bc_rgb, const float* z_vals, const float* rays_d_head,
                            const float* rays_d_torso, float last_dist, const float* target_head, const float* target_person,
                            float* loss2, float* rgb_head, float* rgb_person, float* dpre_head, float* dsigma_head,
                            float* dpre_torso, float* dsigma_torso, void* stream);

/* torch.optim.Adam(lr, betas, eps) step number `step` (1-based) over one flat parameter group (MAIN:522-535, 924-931):
 * parameters whose gradient is zero and whose moments are zero do not move, like parameters torch skips for grad=None. */
int dfn_adam_step(int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float lr, float beta1,
                  float beta2, float eps, int step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DFN_H_ */
