// Latent encoders -- the callers immediately before the hot path: AudioNet (HELP:109-141),
// AudioAttNet temporal smoothing (HELP:210-240 with the window logic of MAIN:35-61 / MAIN:85-101) and the torso's pose
// signal (MAIN:182-205, MAIN:106-109).  Tiny per-frame networks: one thread block per frame, everything in shared
// memory, a whole sequence per launch instead of dozens of micro-kernels per frame.  The MLP encoders (AudioNet_W2L,
// ExpressionEnc) are dfn_linear chains (mlp_fp32.cu, act = 3).
#include <math.h>

#include "common.cuh"

namespace dfn {

struct ConvStackDev {
  int n, stride;
  int ch[7];
  const float* w[6];
  const float* b[6];
};

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : __fmul_rn(0.02f, v); }

// Conv1d(kernel 3, padding 1) + LeakyReLU(0.02), block-cooperative: in [cin][Tin] -> out [cout][Tout] (shared memory)
__device__ void conv1d_leaky(const float* in, int cin, int Tin, const float* __restrict__ w, const float* __restrict__ b,
                             int cout, int stride, float* out, int Tout) {
  for (int idx = threadIdx.x; idx < cout * Tout; idx += blockDim.x) {
    const int co = idx / Tout, t = idx % Tout;
    float acc = b[co];
    for (int ci = 0; ci < cin; ++ci) {
      const float* wk = w + ((size_t)co * cin + ci) * 3;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int ti = t * stride + k - 1;
        if (ti >= 0 && ti < Tin) acc = fmaf(wk[k], in[ci * Tin + ti], acc);
      }
    }
    out[idx] = leaky(acc);
  }
}

static constexpr int kBuf = 64 * 16;   // largest activation: 64 channels x 16 steps

__global__ void audionet_kernel(int N, int dim_aud, const float* __restrict__ x, ConvStackDev cs, const float* __restrict__ fc1_w,
                                const float* __restrict__ fc1_b, const float* __restrict__ fc2_w,
                                const float* __restrict__ fc2_b, float* __restrict__ out) {
  __shared__ float bufA[kBuf], bufB[kBuf];
  const int n = blockIdx.x;
  // x[n, t, c] -> in[c][t]  (HELP:135: permute(0, 2, 1))
  for (int i = threadIdx.x; i < 16 * 29; i += blockDim.x) {
    const int t = i / 29, c = i % 29;
    bufA[c * 16 + t] = x[(size_t)n * 16 * 29 + i];
  }
  __syncthreads();
  float* a = bufA;
  float* bq = bufB;
  int T = 16;
  for (int l = 0; l < cs.n; ++l) {
    const int Tout = (T - 1) / cs.stride + 1;
    conv1d_leaky(a, cs.ch[l], T, cs.w[l], cs.b[l], cs.ch[l + 1], cs.stride, bq, Tout);
    __syncthreads();
    float* tmp = a;
    a = bq;
    bq = tmp;
    T = Tout;
  }
  // a = [64][1]; Linear(64,64) + LeakyReLU; Linear(64, dim_aud)   (HELP:127-131)
  const int H = cs.ch[cs.n];
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    float acc = fc1_b[j];
    for (int k = 0; k < H; ++k) acc = fmaf(fc1_w[j * H + k], a[k], acc);
    bq[j] = leaky(acc);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < dim_aud; j += blockDim.x) {
    float acc = fc2_b[j];
    for (int k = 0; k < H; ++k) acc = fmaf(fc2_w[j * H + k], bq[k], acc);
    out[(size_t)n * dim_aud + j] = acc;
  }
}

__global__ void att_smooth_kernel(int N, int D, int dim_att, int seq, const float* __restrict__ feats,
                                  const float* __restrict__ pad_row, ConvStackDev cs, const float* __restrict__ lin_w,
                                  const float* __restrict__ lin_b, float* __restrict__ out) {
  __shared__ float X[16 * 128];
  __shared__ float bufA[128 * 16], bufB[16 * 16];
  __shared__ float att[16];
  const int i = blockIdx.x, half = seq / 2;
  for (int idx = threadIdx.x; idx < seq * D; idx += blockDim.x) {
    const int r = idx / D, d = idx % D;
    const int f = i - half + r;
    X[idx] = (f >= 0 && f < N) ? feats[(size_t)f * D + d] : pad_row[d];
  }
  __syncthreads();
  // y = x[..., :dim_att].permute(1, 0): [dim_att][seq]   (HELP:233)
  for (int idx = threadIdx.x; idx < dim_att * seq; idx += blockDim.x) bufA[idx] = X[(idx % seq) * D + idx / seq];
  __syncthreads();
  float* a = bufA;
  float* bq = bufB;
  for (int l = 0; l < cs.n; ++l) {
    conv1d_leaky(a, cs.ch[l], seq, cs.w[l], cs.b[l], cs.ch[l + 1], 1, bq, seq);
    __syncthreads();
    float* tmp = a;
    a = bq;
    bq = tmp;
  }
  // attentionNet: Linear(seq, seq) + softmax   (HELP:226-230, HELP:237)
  if (threadIdx.x == 0) {
    float z[16], m = -INFINITY;
    for (int j = 0; j < seq; ++j) {
      float acc = lin_b[j];
      for (int t = 0; t < seq; ++t) acc = fmaf(lin_w[j * seq + t], a[t], acc);
      z[j] = acc;
      m = fmaxf(m, acc);
    }
    float ssum = 0.f;
    for (int j = 0; j < seq; ++j) {
      z[j] = expf(z[j] - m);
      ssum += z[j];
    }
    for (int j = 0; j < seq; ++j) att[j] = __fdiv_rn(z[j], ssum);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {      // torch.sum(y * x, dim=0)
    float acc = 0.f;
    for (int t = 0; t < seq; ++t) acc = __fadd_rn(acc, __fmul_rn(att[t], X[t * D + d]));
    out[(size_t)i * D + d] = acc;
  }
}

__global__ void pose_signal_kernel(int N, const float* __restrict__ poses, int stride, int L, float* __restrict__ out,
                                   float* __restrict__ et_out) {
  const int per = 3 + 6 * L;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < N; f += gridDim.x * blockDim.x) {
    const float* R = poses + (size_t)f * stride;   // row-major [.,4]
    float et[6];
    et[2] = atan2f(R[0], -R[1]);     // atan2(R00, -R01)   (MAIN:196)
    et[1] = asinf(-R[2]);            // asin(-R02)
    et[0] = atan2f(R[10], R[6]);     // atan2(R22, R12)
    et[3] = R[3];
    et[4] = R[7];
    et[5] = R[11];
    if (et_out)
      for (int k = 0; k < 6; ++k) et_out[(size_t)f * 6 + k] = et[k];
    float* o = out + (size_t)f * 2 * per;
    for (int h = 0; h < 2; ++h) {
      for (int j = 0; j < per; ++j) {
        const float v = et[h * 3 + (j < 3 ? j : (j - 3) % 3)];
        float r;
        if (j < 3) {
          r = v;
        } else {
          const int k = (j - 3) / 6, q = (j - 3) % 6;
          const float a = __fmul_rn(v, pow2i(k));
          r = q < 3 ? sinf(a) : cosf(a);
        }
        o[h * per + j] = r;
      }
    }
  }
}

static int to_dev(const dfn_conv_stack* c, ConvStackDev* d, const char* what) {
  if (!c || c->n < 1 || c->n > 6 || c->stride < 1 || c->stride > 2) {
    set_error("%s: bad conv stack", what);
    return DFN_E_ARG;
  }
  d->n = c->n;
  d->stride = c->stride;
  for (int i = 0; i <= c->n; ++i) {
    if (c->ch[i] < 1 || c->ch[i] > 128) {
      set_error("%s: conv channels must be in 1..128", what);
      return DFN_E_ARG;
    }
    d->ch[i] = c->ch[i];
  }
  for (int i = 0; i < c->n; ++i) {
    if (!c->w[i] || !c->b[i]) {
      set_error("%s: null conv weights", what);
      return DFN_E_ARG;
    }
    d->w[i] = c->w[i];
    d->b[i] = c->b[i];
  }
  return 0;
}

}  // namespace dfn

using namespace dfn;

extern "C" int dfn_audionet_forward(int N, int dim_aud, const float* x, const dfn_conv_stack* convs, const float* fc1_w,
                                    const float* fc1_b, const float* fc2_w, const float* fc2_b, float* out, void* stream) {
  DFN_CHECK_ARG(N > 0 && dim_aud > 0 && x && fc1_w && fc1_b && fc2_w && fc2_b && out, "dfn_audionet_forward: bad argument");
  ConvStackDev cs;
  int rc = to_dev(convs, &cs, "dfn_audionet_forward");
  if (rc) return rc;
  DFN_CHECK_ARG(cs.ch[0] == 29 && cs.stride == 2 && cs.n == 4 && cs.ch[cs.n] <= 64,
                "dfn_audionet_forward: expects the four stride-2 convs of HELP:113-125 on 29 input channels");
  for (int i = 0; i <= cs.n; ++i) DFN_CHECK_ARG(cs.ch[i] <= 64, "dfn_audionet_forward: at most 64 channels");
  audionet_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(N, dim_aud, x, cs, fc1_w, fc1_b, fc2_w, fc2_b, out);
  DFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dfn_att_smooth(int N, int D, int dim_att, int seq_len, const float* feats, const float* pad_row,
                              const dfn_conv_stack* convs, const float* lin_w, const float* lin_b, float* out, void* stream) {
  DFN_CHECK_ARG(N > 0 && D > 0 && D <= 128 && dim_att > 0 && dim_att <= D && seq_len >= 2 && seq_len <= 16 && seq_len % 2 == 0 &&
                    feats && pad_row && lin_w && lin_b && out,
                "dfn_att_smooth: bad argument (D <= 128, seq_len even and <= 16)");
  ConvStackDev cs;
  int rc = to_dev(convs, &cs, "dfn_att_smooth");
  if (rc) return rc;
  DFN_CHECK_ARG(cs.ch[0] == dim_att && cs.ch[cs.n] == 1 && cs.stride == 1, "dfn_att_smooth: conv stack must map dim_att -> 1, stride 1");
  for (int i = 1; i <= cs.n; ++i) DFN_CHECK_ARG(cs.ch[i] <= 16, "dfn_att_smooth: hidden conv channels <= 16 (HELP:216-224)");
  att_smooth_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(N, D, dim_att, seq_len, feats, pad_row, cs, lin_w, lin_b, out);
  DFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dfn_pose_signal(int N, const float* poses, int pose_stride, int L, float* out, float* et_out, void* stream) {
  DFN_CHECK_ARG(N > 0 && poses && out && (pose_stride == 12 || pose_stride == 16) && L >= 0 && L <= 16,
                "dfn_pose_signal: bad argument (pose_stride 12 or 16)");
  pose_signal_kernel<<<ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(N, poses, pose_stride, L, out, et_out);
  DFN_LAUNCH_CHECK();
  return 0;
}
