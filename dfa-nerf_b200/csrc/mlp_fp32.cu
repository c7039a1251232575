// fp32 FFMA path of the skip-MLP (HELP:275-299 FaceNeRF.forward, HELP:372-396 NeRF.forward):
// one tiled SGEMM-with-epilogue launch per nn.Linear.  It takes arbitrary embedded inputs, so it
// backs the drop-in module forward and serves as the on-device fp32 cross-check for the tcgen05
// kernel at sizes the CPU oracle cannot reach.  Concatenations ([input_pts, h], [h, views]) are
// never materialised: a layer reads up to two sources and walks the weight columns across them.
#include "common.cuh"
#include "model.h"

namespace dfn {

static constexpr int BM = 64, BN = 64, BK = 16;

// Y[p, n] = act(sum_k X(p,k) * W[n,k] + b[n]) (+ A[p*ld_add + n]),  X(p,k) = k<K1 ? X1[p*ld1+k] : X2[p*ld2+k-K1]
// act & 3: 0 none, 1 relu, 2 sigmoid, 3 LeakyReLU(0.02) (HELP:171); act & 4: the addend is added BEFORE the activation (DEC:331-340) instead
// of after it (DEC:316-325).  A row stride of 0 broadcasts one row (per-frame latent terms, DEC:311,319).
__global__ void __launch_bounds__(256)
linear_fp32_kernel(int64_t P, int N, int K1, int K2, const float* __restrict__ X1, int64_t ld1,
                   const float* __restrict__ X2, int64_t ld2, const float* __restrict__ Wt,
                   const float* __restrict__ bias, int relu, float* __restrict__ Y, int64_t ldy,
                   const float* __restrict__ addend = nullptr, int64_t ld_add = 0) {
  __shared__ float Xs[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int K = K1 + K2;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t p0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // 64 rows x 16 k = 1024 elements per operand, 4 per thread; k fastest for coalescing
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = threadIdx.x + e * 256;
      int r = idx >> 4, kk = idx & 15;
      int k = k0 + kk;
      int64_t p = p0 + r;
      float xv = 0.f, wv = 0.f;
      if (k < K) {
        if (p < P) xv = k < K1 ? X1[p * ld1 + k] : X2[p * ld2 + (k - K1)];
        int n = n0 + r;
        if (n < N) wv = Wt[(int64_t)n * K + k];
      }
      Xs[kk][r] = xv;
      Ws[kk][r] = wv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Xs[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t p = p0 + ty * 4 + i;
    if (p >= P) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (addend && (relu & 4)) v += addend[p * ld_add + n];
      if ((relu & 3) == 1) v = fmaxf(v, 0.f);
      else if ((relu & 3) == 2) v = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v)));
      else if ((relu & 3) == 3) v = v > 0.f ? v : __fmul_rn(0.02f, v);   // nn.LeakyReLU(0.02) of the latent encoders
      if (addend && !(relu & 4)) v += addend[p * ld_add + n];
      Y[p * ldy + n] = v;
    }
  }
}

static int launch_linear(int64_t P, const Fp32Layer& L, const float* X1, int64_t ld1, int K1,
                         const float* X2, int64_t ld2, int K2, int relu, float* Y, int64_t ldy,
                         cudaStream_t st) {
  if (K1 + K2 != L.in) {
    set_error("fp32 mlp: layer expects K=%d, got %d+%d", L.in, K1, K2);
    return DFN_E_STATE;
  }
  dim3 grid(ceil_div(P, BM), ceil_div(L.out, BN));
  linear_fp32_kernel<<<grid, 256, 0, st>>>(P, L.out, K1, K2, X1, ld1, X2, ld2, L.w, L.b, relu, Y, ldy);
  DFN_LAUNCH_CHECK();
  return 0;
}

}  // namespace dfn

// Generic fused linear layer of the fp32 path (also the building block of the Decoder shim, DEC:277-349).
extern "C" int dfn_linear(int64_t P, int N, int K1, const float* X1, int64_t ld1, int K2, const float* X2, int64_t ld2,
                          const float* W, const float* bias, int act, const float* addend, int64_t ld_add, float* Y,
                          int64_t ldy, void* stream) {
  using namespace dfn;
  DFN_CHECK_ARG(P > 0 && N > 0 && K1 > 0 && K2 >= 0 && X1 && W && Y && (K2 == 0 || X2) && act >= 0 && act <= 7 && ldy >= N,
                "dfn_linear: bad argument");
  dim3 grid(ceil_div(P, BM), ceil_div(N, BN));
  linear_fp32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P, N, K1, K2, X1, ld1, X2, ld2, W, bias, act, Y, ldy, addend,
                                                            ld_add);
  DFN_LAUNCH_CHECK();
  return 0;
}

namespace dfn {

int64_t mlp_fp32_workspace_bytes(const dfn_model* m, int64_t P) {
  // two ping-pong activation buffers of width W, 256-byte aligned
  int64_t one = ((P * m->desc.W * (int64_t)sizeof(float) + 255) / 256) * 256;
  return 2 * one;
}

// x [P, in_dim] -> out [P,4].  Layer order follows HELP:275-299 / HELP:372-396.
int mlp_fp32_forward(const dfn_model* m, int64_t P, const float* x, float* out, void* workspace,
                     int64_t workspace_bytes, cudaStream_t st) {
  const dfn_model_desc& d = m->desc;
  if (workspace_bytes < mlp_fp32_workspace_bytes(m, P)) {
    set_error("dfn_mlp_forward: workspace %lld < %lld bytes", (long long)workspace_bytes,
              (long long)mlp_fp32_workspace_bytes(m, P));
    return DFN_E_WORKSPACE;
  }
  const int n_pts = d.input_ch + d.dim_aud;
  const int64_t ldx = n_pts + d.input_ch_views;
  int64_t one = ((P * d.W * (int64_t)sizeof(float) + 255) / 256) * 256;
  float* bufA = reinterpret_cast<float*>(workspace);
  float* bufB = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + one);
  const float* h = nullptr;
  int rc;
  for (int i = 0; i < d.D; ++i) {
    float* dst = (i & 1) ? bufB : bufA;
    if (i == 0) {
      rc = launch_linear(P, m->pts[i], x, ldx, n_pts, nullptr, 0, 0, 1, dst, d.W, st);
    } else if (i - 1 == d.skip) {
      rc = launch_linear(P, m->pts[i], x, ldx, n_pts, h, d.W, d.W, 1, dst, d.W, st);
    } else {
      rc = launch_linear(P, m->pts[i], h, d.W, d.W, nullptr, 0, 0, 1, dst, d.W, st);
    }
    if (rc) return rc;
    h = dst;
  }
  if (d.skip == d.D - 1) {
    set_error("fp32 mlp: skip at the last trunk layer is not supported");
    return DFN_E_UNSUPPORTED;
  }
  // alpha head (no activation) -> out[:,3]
  rc = launch_linear(P, m->alpha, h, d.W, d.W, nullptr, 0, 0, 0, out + 3, 4, st);
  if (rc) return rc;
  float* other = (h == bufA) ? bufB : bufA;
  const float* feat = h;
  if (d.kind == DFN_MODEL_NERF) {  // HELP:384 feature_linear applied; FaceNeRF bypasses it (HELP:287)
    rc = launch_linear(P, m->feature, h, d.W, d.W, nullptr, 0, 0, 0, other, d.W, st);
    if (rc) return rc;
    feat = other;
    other = (float*)h;
  }
  // views_linears.0 on [feature, views]; the rest W/2 -> W/2
  const int Wh = d.W / 2;
  rc = launch_linear(P, m->views[0], feat, d.W, d.W, x + n_pts, ldx, d.input_ch_views, 1, other, Wh, st);
  if (rc) return rc;
  const float* hv = other;
  float* spare = (float*)feat;
  for (int i = 1; i < m->n_views; ++i) {
    rc = launch_linear(P, m->views[i], hv, Wh, Wh, nullptr, 0, 0, 1, spare, Wh, st);
    if (rc) return rc;
    float* t = (float*)hv;
    hv = spare;
    spare = t;
  }
  return launch_linear(P, m->rgb, hv, Wh, Wh, nullptr, 0, 0, 0, out, 4, st);
}

}  // namespace dfn
