// Training-step kernels around the GEMMs (MAIN:855-931): bias gradients, the backward of the live two-field
// compositing + the two image losses, and the Adam update.
//   dfn_colsum                  db = sum over the batch of dH * act'(Y)                      (autograd of nn.Linear's bias)
//   dfn_head_torso_loss_bwd     MAIN:884-907 forward (as head_torso_kernel) + img2mse x2 (HELP:11) and their gradients with
//                               respect to both fields' colours (pre-sigmoid) and densities: a reverse per-ray scan
//   dfn_adam_step               torch.optim.Adam(betas=(0.9, 0.999), eps=1e-8) over a flat parameter group (MAIN:522-535, 924-931)
#include <math_constants.h>

#include "common.cuh"

namespace dfn {

__device__ __forceinline__ float act_grad(int mode, float y) {
  if (mode == DFN_MASK_RELU) return y > 0.f ? 1.f : 0.f;
  if (mode == DFN_MASK_LEAKY) return y > 0.f ? 1.f : 0.02f;
  if (mode == DFN_MASK_SIGMOID) return y * (1.f - y);
  return 1.f;
}

// out[n] += sum_{m in this block's rows} X[m*ld + n] * act'(Y[m*ld + n]); thread = column (coalesced rows), fp32 atomics
__global__ void colsum_kernel(int64_t M, int N, const float* __restrict__ X, int64_t ld, const float* __restrict__ Y, int mode,
                              int rows_per_block, float* __restrict__ out) {
  const int n = threadIdx.x;
  if (n >= N) return;
  const int64_t m0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t m1 = m0 + rows_per_block < M ? m0 + rows_per_block : M;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int64_t m = m0;
  for (; m + 4 <= m1; m += 4) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x = X[(m + j) * ld + n];
      if (Y) x *= act_grad(mode, Y[(m + j) * ld + n]);
      acc[j] += x;
    }
  }
  for (; m < m1; ++m) {
    float x = X[m * ld + n];
    if (Y) x *= act_grad(mode, Y[m * ld + n]);
    acc[0] += x;
  }
  atomicAdd(out + n, (acc[0] + acc[1]) + (acc[2] + acc[3]));
}

// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double shfl_up_d(double v, int delta) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_up_sync(0xffffffffu, lo, delta);
  hi = __shfl_up_sync(0xffffffffu, hi, delta);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_down_d(double v, int delta) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_down_sync(0xffffffffu, lo, delta);
  hi = __shfl_down_sync(0xffffffffu, hi, delta);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// One warp per ray; lane l owns the contiguous samples [l*seg, (l+1)*seg).  Forward exactly as head_torso_kernel (stages.cu:
// MAIN:669-708); then, with g_f = d loss / d rgb_f = 2 (rgb_f - target_f) / (3R) for the two images f,
//   G_s = c_s . g,   d c_s = w_s g,   d alpha_s = T_s G_s - (sum_{j>s} w_j G_j) / (1 - alpha_s + 1e-10),
//   d t_s = d alpha_s * dist_s * exp(-t_s dist_s)        (t = relu(sigma) + 1e-6, MAIN:174)
// and the chain rule of composite_function (MAIN:146-166: w_b = sigma_b / den with the den == 0 -> 1e-4 entries detached by
// the in-place masked assignment), of the relu's (MAIN:688-689, MAIN:174) and of the final sigmoid (DEC:346-347).
template <int SEG>
__global__ void head_torso_bwd_kernel(int R, int S, const float* __restrict__ feat_h, const float* __restrict__ sig_h,
                                      const float* __restrict__ feat_t, const float* __restrict__ sig_t,
                                      const float* __restrict__ bc_rgb, const float* __restrict__ z_vals,
                                      const float* __restrict__ rays_d_h, const float* __restrict__ rays_d_t, float last_dist,
                                      const float* __restrict__ target_head, const float* __restrict__ target_person,
                                      float inv_n, float* __restrict__ loss2, float* __restrict__ rgb_head,
                                      float* __restrict__ rgb_person, float* __restrict__ dpre_h, float* __restrict__ dsig_h,
                                      float* __restrict__ dpre_t, float* __restrict__ dsig_t) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int seg = (S + 31) / 32;
  float loss_acc[2] = {0.f, 0.f};
  for (int ray = blockIdx.x * warps_per_block + (threadIdx.x >> 5); ray < R; ray += gridDim.x * warps_per_block) {
    float nrm[2];
    {
      const float* d0 = rays_d_h + ray * 3;
      const float* d1 = rays_d_t + ray * 3;
      nrm[0] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d0[0], d0[0]), __fmul_rn(d0[1], d0[1])), __fmul_rn(d0[2], d0[2])));
      nrm[1] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d1[0], d1[0]), __fmul_rn(d1[1], d1[1])), __fmul_rn(d1[2], d1[2])));
    }
    const float* zr = z_vals + (int64_t)ray * S;
    float alpha[2][SEG], ex[2][SEG], dist[2][SEG], col[2][SEG][3];
    float fh[SEG][3], ft[SEG][3], sh[SEG], st[SEG], rawh[SEG], rawt[SEG], wh[SEG], wt[SEG], den[SEG], ssum[SEG];
    double local[2] = {1.0, 1.0};
    const int s0 = lane * seg;
#pragma unroll
    for (int k = 0; k < SEG; ++k) {
      const int s = s0 + k;
      alpha[0][k] = alpha[1][k] = 0.f;
      if (k < seg && s < S) {
        const int64_t i = (int64_t)ray * S + s;
        const bool last = s == S - 1;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          fh[k][c] = last ? bc_rgb[ray * 3 + c] : feat_h[i * 3 + c];
          ft[k][c] = feat_t[i * 3 + c];
        }
        rawh[k] = sig_h[i];
        rawt[k] = sig_t[i];
        sh[k] = fmaxf(rawh[k], 0.f);
        st[k] = last ? 0.f : fmaxf(rawt[k], 0.f);
        const float sh1 = last ? __fadd_rn(sh[k], 1e-6f) : sh[k];
        if (last) st[k] = __fadd_rn(st[k], 1e-6f);
        float dn = __fadd_rn(sh[k], st[k]);
        ssum[k] = dn;
        if (dn == 0.f) dn = 1e-4f;
        den[k] = dn;
        wh[k] = __fdiv_rn(sh[k], dn);
        wt[k] = __fdiv_rn(st[k], dn);
        const float dz = last ? last_dist : __fsub_rn(zr[s + 1], zr[s]);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          col[0][k][c] = fh[k][c];
          col[1][k][c] = __fadd_rn(__fmul_rn(fh[k][c], wh[k]), __fmul_rn(ft[k][c], wt[k]));
        }
        const float sg[2] = {sh1, ssum[k]};
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          dist[f][k] = __fmul_rn(dz, nrm[f]);
          const float e = expf(-__fmul_rn(__fadd_rn(fmaxf(sg[f], 0.f), 1e-6f), dist[f][k]));
          ex[f][k] = e;
          const float a = __fsub_rn(1.0f, e);
          alpha[f][k] = a;
          local[f] *= (double)__fadd_rn(__fsub_rn(1.0f, a), 1e-10f);
        }
      }
    }
    float dsh[SEG], dst_[SEG], dfh[SEG][3], dft[SEG][3];
#pragma unroll
    for (int k = 0; k < SEG; ++k) {
      dsh[k] = dst_[k] = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) dfh[k][c] = dft[k][c] = 0.f;
    }
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      // ---- forward: transmittance (exclusive product), weights, colour sums
      double incl = local[f];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        double up = shfl_up_d(incl, d);
        if (lane >= d) incl *= up;
      }
      double run = shfl_up_d(incl, 1);
      if (lane == 0) run = 1.0;
      float T[SEG], w[SEG];
      float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < SEG; ++k) {
        const int s = s0 + k;
        T[k] = w[k] = 0.f;
        if (k < seg && s < S) {
          T[k] = (float)run;
          w[k] = __fmul_rn(alpha[f][k], T[k]);
          run *= (double)__fadd_rn(__fsub_rn(1.0f, alpha[f][k]), 1e-10f);
          acc[0] += w[k] * col[f][k][0];
          acc[1] += w[k] * col[f][k][1];
          acc[2] += w[k] * col[f][k][2];
        }
      }
      float g[3];
      const float* tgt = (f == 0 ? target_head : target_person) + ray * 3;
      float* out = f == 0 ? rgb_head : rgb_person;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        acc[c] = warp_sum_f(acc[c]);
        const float diff = acc[c] - tgt[c];
        g[c] = 2.f * diff * inv_n;
        if (lane == 0) {
          loss_acc[f] += diff * diff * inv_n;
          if (out) out[ray * 3 + c] = acc[c];
        }
      }
      // ---- backward: exclusive SUFFIX sums of w_j G_j
      float G[SEG];
      double lsum = 0.0;
#pragma unroll
      for (int k = 0; k < SEG; ++k) {
        G[k] = col[f][k][0] * g[0] + col[f][k][1] * g[1] + col[f][k][2] * g[2];
        lsum += (double)w[k] * (double)G[k];
      }
      double suf = lsum;     // inclusive suffix over lanes
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        double dn = shfl_down_d(suf, d);
        if (lane + d < 32) suf += dn;
      }
      double after = shfl_down_d(suf, 1);   // sum over the lanes after this one
      if (lane == 31) after = 0.0;
#pragma unroll
      for (int k = SEG - 1; k >= 0; --k) {
        const int s = s0 + k;
        if (k < seg && s < S) {
          const bool last = s == S - 1;
          const float u = __fadd_rn(__fsub_rn(1.0f, alpha[f][k]), 1e-10f);
          const float dalpha = T[k] * G[k] - (float)(after / (double)u);
          after += (double)w[k] * (double)G[k];
          const float dt = dalpha * dist[f][k] * ex[f][k];
          float dc[3] = {w[k] * g[0], w[k] * g[1], w[k] * g[2]};
          if (f == 0) {
            // head-only stack: t = relu(relu(sig_h) [+1e-6 at the last sample]) + 1e-6; colour = feat_h (background at the last sample)
            dsh[k] += dt;      // masked by [sig_h > 0] below
            if (!last) {
#pragma unroll
              for (int c = 0; c < 3; ++c) dfh[k][c] += dc[c];
            }
          } else {
            const float dwh = fh[k][0] * dc[0] + fh[k][1] * dc[1] + fh[k][2] * dc[2];
            const float dwt = ft[k][0] * dc[0] + ft[k][1] * dc[1] + ft[k][2] * dc[2];
            const float inv = 1.0f / den[k];
            if (ssum[k] != 0.f) {
              const float q = (dwh * sh[k] + dwt * st[k]) * inv * inv;
              dsh[k] += dwh * inv - q;
              dst_[k] += dwt * inv - q;
            } else {            // den was replaced by the constant 1e-4 (MAIN:159): no gradient through it
              dsh[k] += dwh * inv;
              dst_[k] += dwt * inv;
            }
            const float dss = ssum[k] > 0.f ? dt : 0.f;     // relu inside calc_volume_weights (MAIN:174)
            dsh[k] += dss;
            dst_[k] += dss;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              if (!last) dfh[k][c] += wh[k] * dc[c];
              dft[k][c] += wt[k] * dc[c];
            }
          }
        }
      }
    }
    // ---- write the gradients of this ray's samples
#pragma unroll
    for (int k = 0; k < SEG; ++k) {
      const int s = s0 + k;
      if (k < seg && s < S) {
        const int64_t i = (int64_t)ray * S + s;
        const bool last = s == S - 1;
        dsig_h[i] = rawh[k] > 0.f ? dsh[k] : 0.f;
        dsig_t[i] = (!last && rawt[k] > 0.f) ? dst_[k] : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float yh = feat_h[i * 3 + c];                 // (the last sample's own colour gets no gradient: dfh = 0 there)
          dpre_h[i * 3 + c] = dfh[k][c] * yh * (1.f - yh);
          dpre_t[i * 3 + c] = dft[k][c] * ft[k][c] * (1.f - ft[k][c]);
        }
      }
    }
  }
  if (lane == 0 && loss2) {
    if (loss_acc[0] != 0.f) atomicAdd(loss2 + 0, loss_acc[0]);
    if (loss_acc[1] != 0.f) atomicAdd(loss2 + 1, loss_acc[1]);
  }
}

// torch.optim.Adam single-tensor update (no weight decay, no amsgrad) on a flat group:
//   m <- m + (g - m)(1 - b1);  v <- b2 v + (1 - b2) g^2;  p <- p - step_size * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void adam_kernel(int64_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, float step_size, float b1, float b2, float inv_bc2_sqrt, float eps) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * (1.f - b1);
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * (mi / (sqrtf(vi) * inv_bc2_sqrt + eps));
  }
}

}  // namespace dfn

using namespace dfn;

extern "C" int dfn_colsum(int64_t M, int N, const float* X, int64_t ld, const float* Y, int mask_mode, float* out, void* stream) {
  DFN_CHECK_ARG(M > 0 && N > 0 && N <= 1024 && X && out && ld >= N && mask_mode >= 0 && mask_mode <= 3, "dfn_colsum: bad argument");
  const int threads = (N + 31) / 32 * 32;
  int64_t blocks = (M + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  const int rows_per_block = (int)((M + blocks - 1) / blocks);
  blocks = (M + rows_per_block - 1) / rows_per_block;
  colsum_kernel<<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(M, N, X, ld, mask_mode ? Y : nullptr, mask_mode, rows_per_block, out);
  DFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dfn_head_torso_loss_bwd(int R, int S, const float* feat_head, const float* sigma_head, const float* feat_torso,
                                       const float* sigma_torso, const float* bc_rgb, const float* z_vals, const float* rays_d_head,
                                       const float* rays_d_torso, float last_dist, const float* target_head,
                                       const float* target_person, float* loss2, float* rgb_head, float* rgb_person,
                                       float* dpre_head, float* dsigma_head, float* dpre_torso, float* dsigma_torso, void* stream) {
  DFN_CHECK_ARG(R > 0 && S > 0 && S <= 128 && feat_head && sigma_head && feat_torso && sigma_torso && bc_rgb && z_vals &&
                    rays_d_head && rays_d_torso && target_head && target_person && loss2 && dpre_head && dsigma_head &&
                    dpre_torso && dsigma_torso,
                "dfn_head_torso_loss_bwd: bad argument (S <= 128)");
  const float inv_n = 1.0f / (3.0f * (float)R);
  int64_t blocks = ((int64_t)R + 3) / 4;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  const int seg = (S + 31) / 32;
#define CALL_BWD(SEG)                                                                                                       \
  head_torso_bwd_kernel<SEG><<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(                                                \
      R, S, feat_head, sigma_head, feat_torso, sigma_torso, bc_rgb, z_vals, rays_d_head, rays_d_torso, last_dist, target_head, \
      target_person, inv_n, loss2, rgb_head, rgb_person, dpre_head, dsigma_head, dpre_torso, dsigma_torso)
  if (seg <= 1) CALL_BWD(1);
  else if (seg <= 2) CALL_BWD(2);
  else CALL_BWD(4);
#undef CALL_BWD
  DFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dfn_adam_step(int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float lr, float beta1,
                             float beta2, float eps, int step, void* stream) {
  DFN_CHECK_ARG(n > 0 && params && grads && exp_avg && exp_avg_sq && step >= 1, "dfn_adam_step: bad argument");
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  adam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(n, params, grads, exp_avg, exp_avg_sq, step_size, beta1, beta2, inv_bc2_sqrt, eps);
  DFN_LAUNCH_CHECK();
  return 0;
}
