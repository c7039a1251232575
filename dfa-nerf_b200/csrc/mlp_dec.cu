// The reference's LIVE radiance model on the tcgen05 kernel: Decoder (DEC:137-349) + DeformationField_ori
// (DEC:77-134) as layer programs for mlp_pp_kernel<., DEC = true>, plus the per-call prep kernels and the fused
// head + torso chunk of MAIN:633-708.
//
// What the programs fold away (all exact in real arithmetic; products formed on the host in fp64):
//   * per-frame inputs -- the head's audio/expression signal (DEC:293-295), z_shape (fc_z, fc_z_skips) and z_app
//     (fc_z_view) -- become fp32 bias terms (fold_kernel);
//   * fc_view(PE(ray_d / |ray_d|)) (DEC:337-339) is constant along a ray, but stays a K-block of the view layer's MMA
//     (staged block TC_KB_DIR): a per-ray bias row read in that layer's epilogue measured 2.5x the cost of any other
//     epilogue (global broadcast loads), one extra K-block is 4 of ~180 MMA instructions per tile;
//   * the additive skips (DEC:317-325, DEC:118-121) are applied AFTER the relu and feed a linear layer, so
//     blocks[4](relu + skip(p)) = blocks[4](relu) + (W4 Wskip) p + W4 b_skip: the skip becomes extra input K-blocks of
//     blocks[4] with composed weights, exactly the [input | h] form the FaceNeRF skip layer already has;
//   * the two 64-wide branches of the deformation field run as one 128-wide block-diagonal network; its residual
//     output `deform_net(p) + p` (DEC:299) is an identity block in the output layer's weights.
// The torso's deformed signal is per-sample, so it is a real (staged) input block: layers that read both staged
// blocks are split in two accumulate-chained halves (TC_EPI_CONT, TC_F_ACCUM).
// Each field is compiled twice: the plain program (bf16x3 kernel) keeps sigma_out as a 16-column MMA layer; the
// "folded-head" program of the single-pass kernels evaluates it on the CUDA cores inside the epilogue of blocks[6]
// (TC_F_DOT_SIGMA) -- one layer fewer on every tile's dependency chain.
#include <math.h>
#include <string.h>

#include <functional>
#include <new>
#include <vector>

#include "common.cuh"
#include "model.h"
#include "tc_pack.h"

namespace dfn {

static inline int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }

struct DecField {
  TcProgram prog;
  uint32_t woff32[TC_MAX_LAYERS] = {};
  uint8_t* w_hi = nullptr;
  uint8_t* w_lo = nullptr;
  uint8_t* w_h16 = nullptr;  // fp16 stages (DFN_PREC_FP16), same offsets
  uint32_t woff2[TC_MAX_LAYERS] = {};   // CTA-pair stage images (mlp_pair.cu: the head field's single-pass path), or w2_hi == nullptr
  uint8_t* w2_hi = nullptr;
  uint8_t* w2_h16 = nullptr;
  float* bias = nullptr;     // [n_layers][256] static part
  int n_fold = 0;
  int fold_layer[8] = {};
  float* fold_w = nullptr;   // [n_fold][dimL][256]: bias[layer][n] += sum_j fold_w[f][j][n] * latent[j]
  int dimL = 0;              // latent = [signal | z_shape | z_app]
  int view_layer = -1;
  float* dot_w = nullptr;    // folded-head program: sigma_out row [256] + its bias (TC_DOT_FLOATS)
  double macs_pt = 0.0;      // algorithmic MACs per sample with the per-frame / per-ray terms folded
};

}  // namespace dfn

struct dfn_decoder {
  dfn_decoder_desc desc;
  bool loaded = false;
  dfn::DecField f[2];        // 0 head, 1 torso: plain programs (bf16x3)
  dfn::DecField g[2];        // the same fields with sigma_out folded into an epilogue (bf16 / fp16)
  dfn::DecField tp;          // the torso field's folded-head program in the layout of the CTA-pair kernel (fc_in_torso as one layer)
};

namespace dfn {

// ------------------------------------------------------------------------------- prep kernels
struct FoldArgs {
  int n_fold;
  int layer[8];
  int seg[3];
  const float* lat[3];
  const float* extra;   // [hidden] added to the biases of layer extra_layer (expnet(expression), DEC:279-281,333-334), or null
  int extra_layer;
};

// bias_out[l][n] = bias[l][n] + sum_j fold_w[f(l)][j][n] * latent[j]; one block per layer, fp32, sequential in j.
__global__ void dec_fold_kernel(int n_layers, int dimL, const float* __restrict__ bias, const float* __restrict__ fold_w,
                                FoldArgs a, float* __restrict__ bias_out) {
  __shared__ float lat[1024];
  const int l = blockIdx.x, n = threadIdx.x;
  int f = -1;
  for (int i = 0; i < a.n_fold; ++i)
    if (a.layer[i] == l) f = i;
  if (f >= 0) {
    int o = 0;
    for (int sgi = 0; sgi < 3; ++sgi) {
      for (int j = threadIdx.x; j < a.seg[sgi]; j += blockDim.x) lat[o + j] = a.lat[sgi][j];
      o += a.seg[sgi];
    }
  }
  __syncthreads();
  float v = bias[l * TC_BIAS_STRIDE + n];
  if (f >= 0) {
    const float* w = fold_w + (size_t)f * dimL * TC_BIAS_STRIDE + n;
    float acc = 0.f;
    for (int j = 0; j < dimL; ++j) acc = fmaf(w[(size_t)j * TC_BIAS_STRIDE], lat[j], acc);
    v += acc;
  }
  if (a.extra != nullptr && l == a.extra_layer) v += a.extra[n];
  bias_out[l * TC_BIAS_STRIDE + n] = v;
}

// ------------------------------------------------------------------------------- host: programs
namespace {

struct Lin {
  const float* w;  // [out][in]
  const float* b;
  int in, out;
  float W(int n, int k) const { return w[(size_t)n * in + k]; }
};

// C[n][k] = sum_q A[n][q] * B[q][k0 + k]  (fp64), A: [rows x inner] row-major with leading dimension lda
std::vector<double> compose(const float* A, int rows, int inner, int lda, const Lin& B, int k0, int kn) {
  std::vector<double> C((size_t)rows * kn, 0.0);
  for (int n = 0; n < rows; ++n)
    for (int q = 0; q < inner; ++q) {
      const double a = A[(size_t)n * lda + q];
      if (a == 0.0) continue;
      for (int k = 0; k < kn; ++k) C[(size_t)n * kn + k] += a * (double)B.W(q, k0 + k);
    }
  return C;
}
// y[n] = sum_q A[n][q] * b[q]
std::vector<double> compose_vec(const float* A, int rows, int inner, int lda, const float* b) {
  std::vector<double> y(rows, 0.0);
  for (int n = 0; n < rows; ++n)
    for (int q = 0; q < inner; ++q) y[n] += (double)A[(size_t)n * lda + q] * (double)b[q];
  return y;
}

struct Builder {
  tc::Packer pk;
  DecField* F;
  std::vector<float> bias;                 // [TC_MAX_LAYERS][256]
  std::vector<std::vector<float>> folds;   // each [dimL][256]
  std::vector<float> dot;                  // folded heads (TC_DOT_FLOATS) or empty
  int nl = 0;
  bool pair_layout = false;   // program for mlp_pair.cu: the torso's fc_in as ONE layer over [PE' | signal' in hidden block 3]
  explicit Builder(DecField* f, int dimL) : F(f), bias((size_t)TC_MAX_LAYERS * TC_BIAS_STRIDE, 0.f) {
    pk.want2 = false;   // set by the caller for the program that also runs on mlp_pair.cu
    F->prog = TcProgram();
    F->dimL = dimL;
    F->n_fold = 0;
    F->macs_pt = 0.0;
  }
  // wfun(n, kbi, k): weight of output row n for position k of the layer's kbi-th input block
  int layer(int n, int epi, int flags, std::initializer_list<int> kbs, const std::function<float(int, int, int)>& wfun) {
    TcLayer L;
    memset(&L, 0, sizeof(L));
    L.n = (uint16_t)n;
    L.epi = (uint8_t)epi;
    L.flags = (uint8_t)flags;
    for (int kb : kbs) L.kb[L.nkb++] = (uint8_t)kb;
    pk.add_layer(n, L.nkb, wfun);
    F->woff32[nl] = pk.last32;
    F->woff2[nl] = pk.last2;
    F->prog.layers[nl] = L;
    return nl++;
  }
  float& b(int l, int n) { return bias[(size_t)l * TC_BIAS_STRIDE + n]; }
  // fold matrix of layer l (created on first use): [dimL][256]
  float* fold(int l) {
    for (int i = 0; i < F->n_fold; ++i)
      if (F->fold_layer[i] == l) return folds[i].data();
    F->fold_layer[F->n_fold++] = l;
    folds.emplace_back((size_t)F->dimL * TC_BIAS_STRIDE, 0.f);
    return folds.back().data();
  }
  int upload(cudaStream_t st) {
    F->prog.n_layers = nl;
    DFN_CUDA(cudaMalloc(&F->w_hi, pk.hi32.size()));
    DFN_CUDA(cudaMalloc(&F->w_lo, pk.lo32.size()));
    DFN_CUDA(cudaMalloc(&F->w_h16, pk.h16.size()));
    DFN_CUDA(cudaMemcpyAsync(F->w_h16, pk.h16.data(), pk.h16.size(), cudaMemcpyHostToDevice, st));
    if (pk.want2 && !pk.hi2.empty()) {
      DFN_CUDA(cudaMalloc(&F->w2_hi, pk.hi2.size()));
      DFN_CUDA(cudaMalloc(&F->w2_h16, pk.h16_2.size()));
      DFN_CUDA(cudaMemcpyAsync(F->w2_hi, pk.hi2.data(), pk.hi2.size(), cudaMemcpyHostToDevice, st));
      DFN_CUDA(cudaMemcpyAsync(F->w2_h16, pk.h16_2.data(), pk.h16_2.size(), cudaMemcpyHostToDevice, st));
    }
    DFN_CUDA(cudaMalloc(&F->bias, bias.size() * 4));
    std::vector<float> fw;
    for (auto& f : folds) fw.insert(fw.end(), f.begin(), f.end());
    if (fw.empty()) fw.resize(1, 0.f);
    DFN_CUDA(cudaMalloc(&F->fold_w, fw.size() * 4));
    DFN_CUDA(cudaMemcpyAsync(F->w_hi, pk.hi32.data(), pk.hi32.size(), cudaMemcpyHostToDevice, st));
    DFN_CUDA(cudaMemcpyAsync(F->w_lo, pk.lo32.data(), pk.lo32.size(), cudaMemcpyHostToDevice, st));
    DFN_CUDA(cudaMemcpyAsync(F->bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice, st));
    DFN_CUDA(cudaMemcpyAsync(F->fold_w, fw.data(), fw.size() * 4, cudaMemcpyHostToDevice, st));
    if (!dot.empty()) {
      DFN_CUDA(cudaMalloc(&F->dot_w, dot.size() * 4));
      DFN_CUDA(cudaMemcpyAsync(F->dot_w, dot.data(), dot.size() * 4, cudaMemcpyHostToDevice, st));
    }
    DFN_CUDA(cudaStreamSynchronize(st));
    return 0;
  }
};

enum {  // load order (dfn.h)
  T_DE0 = 0, T_DEOUT = 5, T_DS0 = 6, T_DSOUT = 11, T_DESKIP = 12, T_DSSKIP = 13, T_FCIN = 14, T_FCIN_TORSO = 15, T_FCZ = 16,
  T_BLOCK0 = 17, T_FCZSKIP = 24, T_FCPSKIP = 25, T_FCPSKIP_TORSO = 26, T_SIGMA = 27, T_FCZVIEW = 28, T_FEATVIEW = 29,
  T_FCVIEW = 30, T_FEATOUT = 31, T_COUNT = 32
};

}  // namespace

static void free_field(DecField& F) {
  cudaFree(F.w_hi);
  cudaFree(F.w_lo);
  cudaFree(F.w_h16);
  cudaFree(F.w2_hi);
  cudaFree(F.w2_h16);
  F.w_h16 = F.w2_hi = F.w2_h16 = nullptr;
  cudaFree(F.bias);
  cudaFree(F.fold_w);
  cudaFree(F.dot_w);
  F.w_hi = F.w_lo = nullptr;
  F.bias = F.fold_w = F.dot_w = nullptr;
}

// Layers shared by both fields from fc_in on.  The field's point input occupies the staged blocks
// (head: PE; torso: PE', signal').  lat offsets: z_shape at zs0, z_app at za0 inside the latent vector.
static void build_trunk(Builder& B, const std::vector<Lin>& T, const dfn_decoder_desc& d, bool torso, int zs0, int za0,
                        bool fold_heads) {
  const int H = d.hidden, de = 6 * d.n_freq;
  const Lin& fcin = T[torso ? T_FCIN_TORSO : T_FCIN];
  const Lin& pskip = T[torso ? T_FCPSKIP_TORSO : T_FCPSKIP];
  const Lin& fcz = T[T_FCZ];
  const int dsig = torso ? d.dim_et_embed : d.dim_signal;
  // ---- fc_in (+ fc_z(z_shape)), relu (DEC:303-312)
  int l;
  if (!torso) {
    l = B.layer(H, TC_EPI_RELU, 0, {TC_KB_PE}, [&](int n, int, int k) { return k < de ? fcin.W(n, k) : 0.f; });
    float* fw = B.fold(l);
    for (int n = 0; n < H; ++n)
      for (int j = 0; j < dsig; ++j) fw[(size_t)j * TC_BIAS_STRIDE + n] = fcin.W(n, de + j);   // per-frame signal columns
  } else if (B.pair_layout) {
    // mlp_pair.cu: the deformation output's epilogue leaves PE' in the staged block and signal' in hidden block 3 (free under the 128-wide
    // deformation layers), so fc_in_torso reads both at once
    l = B.layer(H, TC_EPI_RELU, 0, {TC_KB_PE, 3}, [&](int n, int kbi, int k) {
      return kbi == 0 ? (k < de ? fcin.W(n, k) : 0.f) : (k < dsig ? fcin.W(n, de + k) : 0.f);
    });
  } else {
    B.layer(H, TC_EPI_CONT, 0, {TC_KB_PE}, [&](int n, int, int k) { return k < de ? fcin.W(n, k) : 0.f; });
    l = B.layer(H, TC_EPI_RELU, TC_F_ACCUM, {TC_KB_IN1}, [&](int n, int, int k) { return k < dsig ? fcin.W(n, de + k) : 0.f; });
  }
  {
    float* fw = B.fold(l);
    for (int n = 0; n < H; ++n) {
      B.b(l, n) = fcin.b[n] + fcz.b[n];
      for (int j = 0; j < d.z_dim; ++j) fw[(size_t)(zs0 + j) * TC_BIAS_STRIDE + n] = fcz.W(n, j);
    }
  }
  B.F->macs_pt += (double)H * (de + (torso ? dsig : 0));
  // ---- blocks[0..6]; the additive skip after blocks[skip-1] is composed into blocks[skip] (DEC:314-325)
  const int nb = d.n_blocks - 1;
  for (int i = 0; i < nb; ++i) {
    const Lin& blk = T[T_BLOCK0 + i];
    if (i != d.skip) {
      // folded head: the last block's epilogue also forms sigma_out . relu(out)
      l = B.layer(H, TC_EPI_RELU, (fold_heads && i == nb - 1) ? TC_F_DOT_SIGMA : 0, {0, 1, 2, 3},
                  [&](int n, int kbi, int k) { return blk.W(n, kbi * 64 + k); });
      for (int n = 0; n < H; ++n) B.b(l, n) = blk.b[n];
      B.F->macs_pt += (double)H * H;
      continue;
    }
    const Lin& zskip = T[T_FCZSKIP];
    const std::vector<double> Wp = compose(blk.w, H, H, H, pskip, 0, de + dsig);     // W4 . fc_p_skips
    const std::vector<double> Wz = compose(blk.w, H, H, H, zskip, 0, d.z_dim);       // W4 . fc_z_skips
    std::vector<float> bsum(H);
    for (int q = 0; q < H; ++q) bsum[q] = zskip.b[q] + pskip.b[q];
    const std::vector<double> bb = compose_vec(blk.w, H, H, H, bsum.data());
    const int ldp = de + dsig;
    if (!torso) {
      l = B.layer(H, TC_EPI_RELU, 0, {TC_KB_PE, 0, 1, 2, 3}, [&](int n, int kbi, int k) {
        if (kbi == 0) return k < de ? (float)Wp[(size_t)n * ldp + k] : 0.f;
        return blk.W(n, (kbi - 1) * 64 + k);
      });
      float* fw = B.fold(l);
      for (int n = 0; n < H; ++n)
        for (int j = 0; j < dsig; ++j) fw[(size_t)j * TC_BIAS_STRIDE + n] = (float)Wp[(size_t)n * ldp + de + j];
    } else {
      B.layer(H, TC_EPI_CONT, 0, {TC_KB_PE}, [&](int n, int, int k) { return k < de ? (float)Wp[(size_t)n * ldp + k] : 0.f; });
      l = B.layer(H, TC_EPI_RELU, TC_F_ACCUM, {TC_KB_IN1, 0, 1, 2, 3}, [&](int n, int kbi, int k) {
        if (kbi == 0) return k < dsig ? (float)Wp[(size_t)n * ldp + de + k] : 0.f;
        return blk.W(n, (kbi - 1) * 64 + k);
      });
    }
    float* fw = B.fold(l);
    for (int n = 0; n < H; ++n) {
      B.b(l, n) = (float)((double)blk.b[n] + bb[n]);
      for (int j = 0; j < d.z_dim; ++j) fw[(size_t)(zs0 + j) * TC_BIAS_STRIDE + n] = (float)Wz[(size_t)n * d.z_dim + j];
    }
    B.F->macs_pt += (double)H * H + (double)H * (de + (torso ? dsig : 0));
  }
  if (fold_heads) {   // sigma_out (DEC:329) rides the epilogue of the last block: its row and bias
    const Lin& sg = T[T_SIGMA];
    B.dot.assign(TC_DOT_FLOATS, 0.f);
    for (int k = 0; k < H; ++k) B.dot[k] = sg.W(0, k);
    B.dot[TC_BIAS_STRIDE] = sg.b[0];
    B.F->macs_pt += H;
  } else {
    // ---- sigma_out (DEC:329): 16-column layer, column 0 kept in a register
    const Lin& sg = T[T_SIGMA];
    l = B.layer(16, TC_EPI_SIGMA, 0, {0, 1, 2, 3}, [&](int n, int kbi, int k) { return n == 0 ? sg.W(0, kbi * 64 + k) : 0.f; });
    B.b(l, 0) = sg.b[0];
    B.F->macs_pt += H;
  }
  // ---- feat_view + fc_z_view(z_app) + fc_view(PE(dir)) -> relu (DEC:331-340)
  {
    const Lin& fv = T[T_FEATVIEW];
    const Lin& zv = T[T_FCZVIEW];
    const Lin& vw = T[T_FCVIEW];
    const int dv = 6 * d.n_freq_views;
    // the view-direction encoding is a staged input block of this layer (TC_KB_DIR), fc_view its weights
    l = B.layer(H, TC_EPI_RELU, 0, {TC_KB_DIR, 0, 1, 2, 3}, [&](int n, int kbi, int k) {
      if (kbi == 0) return k < dv ? vw.W(n, k) : 0.f;
      return fv.W(n, (kbi - 1) * 64 + k);
    });
    float* fw = B.fold(l);
    for (int n = 0; n < H; ++n) {
      B.b(l, n) = fv.b[n] + zv.b[n] + vw.b[n];
      for (int j = 0; j < d.z_dim; ++j) fw[(size_t)(za0 + j) * TC_BIAS_STRIDE + n] = zv.W(n, j);
    }
    B.F->view_layer = l;
    B.F->macs_pt += (double)H * H;
  }
  // ---- feat_out -> sigmoid (DEC:344-347)
  {
    const Lin& fo = T[T_FEATOUT];
    l = B.layer(16, TC_EPI_RGB, 0, {0, 1, 2, 3}, [&](int n, int kbi, int k) { return n < 3 ? fo.W(n, kbi * 64 + k) : 0.f; });
    for (int n = 0; n < 3; ++n) B.b(l, n) = fo.b[n];
    B.F->macs_pt += 3.0 * H;
  }
}

// DeformationField_ori (DEC:109-134) as a 128-wide block-diagonal network: rows 0..63 the embed branch, rows 64..127
// the signal branch.  The (per-frame) torso signal enters through biases; the output layer writes
// [PE + d_embed | signal + d_signal] to the tile's staged blocks (DEC:299).
static void build_deform(Builder& B, const std::vector<Lin>& T, const dfn_decoder_desc& d) {
  const int de = 6 * d.n_freq, dt = d.dim_et_embed, HD = 64;
  auto E = [&](int i) -> const Lin& { return T[T_DE0 + i]; };
  auto S = [&](int i) -> const Lin& { return T[T_DS0 + i]; };
  int l = B.layer(2 * HD, TC_EPI_RELU, 0, {TC_KB_PE}, [&](int n, int, int k) {
    if (k >= de) return 0.f;
    return n < HD ? E(0).W(n, k) : S(0).W(n - HD, k);
  });
  {
    float* fw = B.fold(l);
    for (int n = 0; n < 2 * HD; ++n) {
      const Lin& L0 = n < HD ? E(0) : S(0);
      const int r = n < HD ? n : n - HD;
      B.b(l, n) = L0.b[r];
      for (int j = 0; j < dt; ++j) fw[(size_t)j * TC_BIAS_STRIDE + n] = L0.W(r, de + j);
    }
  }
  B.F->macs_pt += 2.0 * HD * de;
  const int n_lay = 5, skip_after = 3;  // layers per branch; (idx+1) in skips=[4] and idx < n-1 (DEC:118, DEC:128)
  for (int i = 1; i < n_lay; ++i) {
    if (i != skip_after + 1) {
      l = B.layer(2 * HD, TC_EPI_RELU, 0, {0, 1}, [&](int n, int kbi, int k) {
        if (n < HD) return kbi == 0 ? E(i).W(n, k) : 0.f;
        return kbi == 1 ? S(i).W(n - HD, k) : 0.f;
      });
      for (int n = 0; n < 2 * HD; ++n) B.b(l, n) = n < HD ? E(i).b[n] : S(i).b[n - HD];
      B.F->macs_pt += 2.0 * HD * HD;
      continue;
    }
    const Lin& es = T[T_DESKIP];
    const Lin& ss = T[T_DSSKIP];
    const std::vector<double> We = compose(E(i).w, HD, HD, HD, es, 0, de);   // W_e4 . fc_embed_skips  (on PE)
    const std::vector<double> Ws = compose(S(i).w, HD, HD, HD, ss, 0, dt);   // W_s4 . fc_signal_skips (on the signal)
    const std::vector<double> be = compose_vec(E(i).w, HD, HD, HD, es.b);
    const std::vector<double> bs = compose_vec(S(i).w, HD, HD, HD, ss.b);
    l = B.layer(2 * HD, TC_EPI_RELU, 0, {TC_KB_PE, 0, 1}, [&](int n, int kbi, int k) {
      if (kbi == 0) return (n < HD && k < de) ? (float)We[(size_t)n * de + k] : 0.f;
      if (n < HD) return kbi == 1 ? E(i).W(n, k) : 0.f;
      return kbi == 2 ? S(i).W(n - HD, k) : 0.f;
    });
    float* fw = B.fold(l);
    for (int n = 0; n < 2 * HD; ++n) {
      if (n < HD) {
        B.b(l, n) = (float)((double)E(i).b[n] + be[n]);
      } else {
        B.b(l, n) = (float)((double)S(i).b[n - HD] + bs[n - HD]);
        for (int j = 0; j < dt; ++j) fw[(size_t)j * TC_BIAS_STRIDE + n] = (float)Ws[(size_t)(n - HD) * dt + j];
      }
    }
    B.F->macs_pt += 2.0 * HD * HD + (double)HD * de;
  }
  // output: rows 0..de-1 = PE + out_embed(h_e); rows 64..64+dt-1 = signal + out_signal(h_s)
  const Lin& oe = T[T_DEOUT];
  const Lin& os = T[T_DSOUT];
  l = B.layer(2 * HD, TC_EPI_STAGE, 0, {TC_KB_PE, 0, 1}, [&](int n, int kbi, int k) {
    if (kbi == 0) return (n < de && k == n) ? 1.f : 0.f;
    if (n < de) return kbi == 1 ? oe.W(n, k) : 0.f;
    if (n >= HD && n < HD + dt) return kbi == 2 ? os.W(n - HD, k) : 0.f;
    return 0.f;
  });
  float* fw = B.fold(l);
  for (int n = 0; n < de; ++n) B.b(l, n) = oe.b[n];
  for (int n = 0; n < dt; ++n) {
    B.b(l, HD + n) = os.b[n];
    fw[(size_t)n * TC_BIAS_STRIDE + HD + n] = 1.f;   // + signal (residual)
  }
  B.F->macs_pt += (double)HD * (de + dt);
}

// {weight, bias} pointers + shapes of the reference's modules in load order
static std::vector<Lin> tensor_table(const dfn_decoder_desc& d, const float* const* t) {
  const int H = d.hidden, de = 6 * d.n_freq, dv = 6 * d.n_freq_views, dt = d.dim_et_embed;
  std::vector<Lin> T(T_COUNT);
  auto set = [&](int i, int in, int out) { T[i] = Lin{t[2 * i], t[2 * i + 1], in, out}; };
  set(T_DE0, de + dt, 64);
  set(T_DS0, de + dt, 64);
  for (int i = 1; i < 5; ++i) {
    set(T_DE0 + i, 64, 64);
    set(T_DS0 + i, 64, 64);
  }
  set(T_DEOUT, 64, de);
  set(T_DSOUT, 64, dt);
  set(T_DESKIP, de, 64);
  set(T_DSSKIP, dt, 64);
  set(T_FCIN, de + d.dim_signal, H);
  set(T_FCIN_TORSO, de + dt, H);
  set(T_FCZ, d.z_dim, H);
  for (int i = 0; i < 7; ++i) set(T_BLOCK0 + i, H, H);
  set(T_FCZSKIP, d.z_dim, H);
  set(T_FCPSKIP, de + d.dim_signal, H);
  set(T_FCPSKIP_TORSO, de + dt, H);
  set(T_SIGMA, H, 1);
  set(T_FCZVIEW, d.z_dim, H);
  set(T_FEATVIEW, H, H);
  set(T_FCVIEW, dv, H);
  set(T_FEATOUT, H, 3);

  return T;
}

static int64_t dec_workspace_bytes(const dfn_decoder* m, int64_t R) {
  (void)m;
  (void)R;
  return align256((int64_t)TC_MAX_LAYERS * TC_BIAS_STRIDE * 4) + align256(pp_dec_scratch_bytes());
}

static int dec_query(const dfn_decoder* m, int field, int64_t R, int S, const float* rays_o, const float* rays_d,
                     const float* z_vals, const float* z_shape, const float* z_app, const float* signal, float* raw,
                     int precision, void* workspace, int64_t workspace_bytes, cudaStream_t st, const float* view_term = nullptr) {
  DFN_CHECK_ARG(m && (field == 0 || field == 1) && R > 0 && S > 0 && rays_o && rays_d && z_vals && z_shape && z_app && signal &&
                    raw && workspace,
                "dfn_decoder_query: bad argument");
  if (!m->loaded) {
    set_error("dfn_decoder_query: decoder has no weights");
    return DFN_E_STATE;
  }
  DFN_CHECK_ARG(precision == DFN_PREC_BF16 || precision == DFN_PREC_FP16 || precision == DFN_PREC_BF16X3,
                "dfn_decoder_query: precision must be DFN_PREC_BF16, DFN_PREC_FP16 or DFN_PREC_BF16X3 (the fp32 path is dfn_linear)");
  DFN_CHECK_ARG((reinterpret_cast<uintptr_t>(raw) & 15) == 0, "dfn_decoder_query: raw must be 16-byte aligned");
  if (workspace_bytes < dec_workspace_bytes(m, R)) {
    set_error("dfn_decoder_query: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)dec_workspace_bytes(m, R));
    return DFN_E_WORKSPACE;
  }
  // split precision: the plain programs on mlp_pp.cu; single-pass: the folded-head programs on the CTA-pair kernel (mlp_pair.cu; the
  // torso's in its own layout), or -- debug flags 8 / 16, hidden != 256 -- on mlp_pp.cu
  const dfn_decoder_desc& d0 = m->desc;
  const DecField& P2 = field == 0 ? m->g[0] : m->tp;
  const bool on_pair = precision != DFN_PREC_BF16X3 && P2.w2_hi != nullptr && !(pair_get_flags() & (field == 0 ? 8 : 16));
  // (split precision with schedule switch 4: the folded-head program as well, sigma_out in fp32 from the fp32 activations of blocks[6])
  const bool x3_folded = precision == DFN_PREC_BF16X3 && (pp_get_flags() & 4) != 0 && m->g[field].dot_w != nullptr && d0.hidden == 256;
  const DecField& F = on_pair ? P2 : (precision == DFN_PREC_BF16X3 && !x3_folded ? m->f[field] : m->g[field]);
  const dfn_decoder_desc& d = m->desc;
  float* bias_ws = reinterpret_cast<float*>(workspace);
  void* scratch = reinterpret_cast<char*>(workspace) + align256((int64_t)TC_MAX_LAYERS * TC_BIAS_STRIDE * 4);

  FoldArgs fa;
  memset(&fa, 0, sizeof(fa));
  fa.n_fold = F.n_fold;
  for (int i = 0; i < F.n_fold; ++i) fa.layer[i] = F.fold_layer[i];
  fa.seg[0] = field == 0 ? d.dim_signal : d.dim_et_embed;
  fa.seg[1] = fa.seg[2] = d.z_dim;
  fa.lat[0] = signal;
  fa.lat[1] = z_shape;
  fa.lat[2] = z_app;
  fa.extra = view_term;
  fa.extra_layer = F.view_layer;
  dec_fold_kernel<<<F.prog.n_layers, TC_BIAS_STRIDE, 0, st>>>(F.prog.n_layers, F.dimL, F.bias, F.fold_w, fa, bias_ws);
  DFN_LAUNCH_CHECK();
  const bool prof = profile_begin(st, F.macs_pt * (double)R * S);
  int rc = on_pair ? pair_launch_prog(F.prog, F.woff2, precision == DFN_PREC_FP16 ? F.w2_h16 : F.w2_hi, precision == DFN_PREC_FP16, true, d.n_freq,
                                      d.n_freq_views, d.hidden, bias_ws, nullptr, F.dot_w, scratch, R, S, rays_o, rays_d, z_vals, raw, st)
                   : pp_launch_prog(F.prog, F.woff32, precision == DFN_PREC_FP16 ? F.w_h16 : F.w_hi, F.w_lo, F.dot_w, true, d.n_freq, d.n_freq_views,
                                    d.hidden, bias_ws, nullptr, scratch, R, S, rays_o, rays_d, z_vals, raw, precision, st,
                                    F.dot_w != nullptr ? F.dot_w + TC_BIAS_STRIDE : nullptr);
  if (prof) profile_end(st);
  if (rc) return rc;
  DFN_LAUNCH_CHECK();
  return 0;
}

}  // namespace dfn

using namespace dfn;

extern "C" int dfn_decoder_create(const dfn_decoder_desc* desc, dfn_decoder** out) {
  DFN_CHECK_ARG(desc && out, "dfn_decoder_create: null argument");
  DFN_CHECK_ARG(desc->hidden == 256 && desc->n_blocks == 8 && desc->skip == 4,
                "dfn_decoder_create: the tcgen05 path covers hidden_size=256, n_blocks=8, skips=[4] (MAIN:518)");
  DFN_CHECK_ARG(desc->n_freq >= 1 && desc->n_freq <= 10 && desc->n_freq_views >= 1 && desc->n_freq_views <= 10,
                "dfn_decoder_create: n_freq_posenc / n_freq_posenc_views must be in 1..10");
  DFN_CHECK_ARG(desc->dim_et_embed >= 1 && desc->dim_et_embed <= 64 && desc->dim_signal >= 1 && desc->z_dim >= 1 &&
                    desc->dim_signal + 2 * desc->z_dim <= 1024,
                "dfn_decoder_create: dim_et_embed <= 64, dim_signal + 2*z_dim <= 1024");
  dfn_decoder* m = new (std::nothrow) dfn_decoder();
  if (!m) {
    set_error("dfn_decoder_create: out of memory");
    return DFN_E_STATE;
  }
  m->desc = *desc;
  *out = m;
  return 0;
}

extern "C" void dfn_decoder_destroy(dfn_decoder* m) {
  if (!m) return;
  for (int i = 0; i < 2; ++i) {
    free_field(m->f[i]);
    free_field(m->g[i]);
  }
  free_field(m->tp);
  delete m;
}

extern "C" int dfn_decoder_num_tensors(const dfn_decoder* m) { return m ? 2 * T_COUNT : 0; }

extern "C" int dfn_decoder_load(dfn_decoder* m, const float* const* t, int n_tensors, void* stream) {
  DFN_CHECK_ARG(m && t, "dfn_decoder_load: null argument");
  DFN_CHECK_ARG(n_tensors == 2 * T_COUNT, "dfn_decoder_load: expected %d tensors, got %d", 2 * T_COUNT, n_tensors);
  for (int i = 0; i < n_tensors; ++i) DFN_CHECK_ARG(t[i] != nullptr, "dfn_decoder_load: tensor %d is null", i);
  cudaStream_t st = (cudaStream_t)stream;
  const dfn_decoder_desc& d = m->desc;
  const int dt = d.dim_et_embed;
  const std::vector<Lin> T = tensor_table(d, t);

  for (int i = 0; i < 2; ++i) {
    free_field(m->f[i]);
    free_field(m->g[i]);
  }
  free_field(m->tp);
  m->loaded = false;
  if (d.hidden == 256) {
    Builder B(&m->tp, dt + 2 * d.z_dim);
    B.pk.want2 = true;
    B.pair_layout = true;
    build_deform(B, T, d);
    build_trunk(B, T, d, true, dt, dt + d.z_dim, true);
    int rc = B.upload(st);
    if (rc) return rc;
  }
  for (int folded = 0; folded < 2; ++folded) {
    DecField* F = folded ? m->g : m->f;
    {
      Builder B(&F[0], d.dim_signal + 2 * d.z_dim);
      B.pk.want2 = folded == 1 && d.hidden == 256;   // the head's folded-head program runs on the CTA-pair kernel (single-pass precisions)
      build_trunk(B, T, d, false, d.dim_signal, d.dim_signal + d.z_dim, folded != 0);
      int rc = B.upload(st);
      if (rc) return rc;
    }
    {
      Builder B(&F[1], dt + 2 * d.z_dim);
      build_deform(B, T, d);
      build_trunk(B, T, d, true, dt, dt + d.z_dim, folded != 0);
      int rc = B.upload(st);
      if (rc) return rc;
    }
  }
  m->loaded = true;
  return 0;
}

// Host-only: the layer program dfn_decoder_load builds for one field, as dense fp32 (no CUDA calls).
extern "C" int dfn_decoder_program_host(const dfn_decoder_desc* desc, const float* const* t, int n_tensors, int field,
                                        int folded_heads, int max_layers, dfn_layer_info* layers, int* n_layers, float* weights,
                                        float* bias, int* n_fold, int* fold_layer, float* fold_w, int* dimL, int* view_layer,
                                        float* dot_w) {
  DFN_CHECK_ARG(desc && t && layers && n_layers && weights && bias && n_fold && fold_layer && fold_w && dimL && view_layer,
                "dfn_decoder_program_host: null argument");
  DFN_CHECK_ARG(n_tensors == 2 * T_COUNT && (field == 0 || field == 1) && max_layers >= TC_MAX_LAYERS,
                "dfn_decoder_program_host: expected %d tensors, field 0|1, max_layers >= %d", 2 * T_COUNT, TC_MAX_LAYERS);
  DFN_CHECK_ARG((folded_heads != 1 && folded_heads != 3) || dot_w, "dfn_decoder_program_host: dot_w is required for the folded-head program");
  DFN_CHECK_ARG(desc->hidden == 256 && desc->n_blocks == 8 && desc->skip == 4 && desc->n_freq >= 1 && desc->n_freq <= 10 &&
                    desc->dim_et_embed >= 1 && desc->dim_et_embed <= 64,
                "dfn_decoder_program_host: unsupported decoder shape");
  const std::vector<Lin> T = tensor_table(*desc, t);
  DecField F;
  const int dsig = field == 0 ? desc->dim_signal : desc->dim_et_embed;
  Builder B(&F, dsig + 2 * desc->z_dim);
  std::vector<float> dense;
  B.pk.dense = &dense;
  B.pair_layout = folded_heads >= 2;   // 2 / 3: the plain / folded-head program in the layout mlp_pair.cu runs (torso: fc_in_torso as one layer)
  const bool fold = folded_heads == 1 || folded_heads == 3;
  if (field == 1) build_deform(B, T, *desc);
  build_trunk(B, T, *desc, field == 1, dsig, dsig + desc->z_dim, fold);
  if (fold) memcpy(dot_w, B.dot.data(), B.dot.size() * 4);
  *n_layers = B.nl;
  for (int l = 0; l < B.nl; ++l) {
    const TcLayer& L = F.prog.layers[l];
    layers[l].n = L.n;
    layers[l].nkb = L.nkb;
    layers[l].epi = L.epi;
    layers[l].flags = L.flags;
    for (int k = 0; k < 6; ++k) layers[l].kb[k] = L.kb[k];
  }
  memset(weights, 0, (size_t)max_layers * tc::Packer::kDenseLayer * 4);
  memcpy(weights, dense.data(), dense.size() * 4);
  memcpy(bias, B.bias.data(), B.bias.size() * 4);
  *n_fold = F.n_fold;
  *dimL = F.dimL;
  *view_layer = F.view_layer;
  for (int i = 0; i < F.n_fold; ++i) {
    fold_layer[i] = F.fold_layer[i];
    memcpy(fold_w + (size_t)i * F.dimL * TC_BIAS_STRIDE, B.folds[i].data(), (size_t)F.dimL * TC_BIAS_STRIDE * 4);
  }
  return 0;
}

extern "C" int64_t dfn_decoder_query_workspace_bytes(const dfn_decoder* m, int64_t R, int S) {
  (void)S;
  return m && R > 0 ? dec_workspace_bytes(m, R) : 0;
}

extern "C" int dfn_decoder_query(const dfn_decoder* m, int field, int64_t R, int S, const float* rays_o, const float* rays_d,
                                 const float* z_vals, const float* z_shape, const float* z_app, const float* signal,
                                 float* raw, int precision, void* workspace, int64_t workspace_bytes, void* stream) {
  reset_launch_count();
  return dec_query(m, field, R, S, rays_o, rays_d, z_vals, z_shape, z_app, signal, raw, precision, workspace, workspace_bytes,
                   (cudaStream_t)stream);
}

extern "C" int dfn_decoder_query_ex(const dfn_decoder* m, int field, int64_t R, int S, const float* rays_o, const float* rays_d,
                                    const float* z_vals, const float* z_shape, const float* z_app, const float* signal,
                                    const float* view_term, float* raw, int precision, void* workspace, int64_t workspace_bytes,
                                    void* stream) {
  reset_launch_count();
  return dec_query(m, field, R, S, rays_o, rays_d, z_vals, z_shape, z_app, signal, raw, precision, workspace, workspace_bytes,
                   (cudaStream_t)stream, view_term);
}

extern "C" double dfn_decoder_macs_per_sample(const dfn_decoder* m, int field) {
  return m && m->loaded && (field == 0 || field == 1) ? m->f[field].macs_pt : 0.0;
}

// ------------------------------------------------------------------ fused live chunk (MAIN:633-708)
extern "C" int64_t dfn_render_head_torso_workspace_bytes(const dfn_decoder* m, int64_t R, int S) {
  if (!m || R <= 0 || S <= 0) return 0;
  return align256(R * S * 4) + 2 * align256(R * S * 16) + dec_workspace_bytes(m, R);
}

extern "C" int dfn_render_head_torso(const dfn_decoder* m, int64_t R, int S, const dfn_head_torso_io* io, int precision,
                                     void* workspace, int64_t workspace_bytes, void* stream) {
  reset_launch_count();
  DFN_CHECK_ARG(m && io && workspace && R > 0 && R < (1ll << 31) && S >= 2 && S <= 256, "dfn_render_head_torso: bad argument");
  DFN_CHECK_ARG(io->rays_o_head && io->rays_d_head && io->rays_o_torso && io->rays_d_torso && io->near && io->far &&
                    io->t_vals && io->bc_rgb && io->z_shape && io->z_app && io->signal && io->signal_torso,
                "dfn_render_head_torso: a required input pointer is null");
  if (workspace_bytes < dfn_render_head_torso_workspace_bytes(m, R, S)) {
    set_error("dfn_render_head_torso: workspace %lld < %lld bytes", (long long)workspace_bytes,
              (long long)dfn_render_head_torso_workspace_bytes(m, R, S));
    return DFN_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = reinterpret_cast<char*>(workspace);
  float* z = reinterpret_cast<float*>(ws);
  float* raw_h = reinterpret_cast<float*>(ws + align256(R * S * 4));
  float* raw_t = reinterpret_cast<float*>(ws + align256(R * S * 4) + align256(R * S * 16));
  void* qws = ws + align256(R * S * 4) + 2 * align256(R * S * 16);
  const int64_t qbytes = workspace_bytes - (align256(R * S * 4) + 2 * align256(R * S * 16));
  const int zd = m->desc.z_dim;
  int rc = dfn_z_vals((int)R, S, io->t_vals, io->near, io->far, nullptr, z, st);   // MAIN:617-619
  if (rc) return rc;
  rc = dec_query(m, 0, R, S, io->rays_o_head, io->rays_d_head, z, io->z_shape, io->z_app, io->signal, raw_h, precision, qws,
                 qbytes, st, io->expression_term);
  if (rc) return rc;
  rc = dec_query(m, 1, R, S, io->rays_o_torso, io->rays_d_torso, z, io->z_shape + zd, io->z_app + zd, io->signal_torso, raw_t,
                 precision, qws, qbytes, st);
  if (rc) return rc;
  return launch_head_torso((int)R, S, raw_h, 4, raw_h + 3, 4, raw_t, 4, raw_t + 3, 4, io->bc_rgb, z, io->rays_d_head,
                           io->rays_d_torso, io->last_dist > 0.f ? io->last_dist : 1e10f, io->rgb_head, io->rgb_person, st);
}
