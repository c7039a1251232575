// Stage kernels either side of the MLP: rays, depth sampling, positional encoding, field
// compositing, sigma->alpha->weights, raw2outputs, inverse-CDF resampling, sort-merge.
// All are HBM-bound elementwise / per-ray work: coalesced loads, one warp per ray for the
// scans, warp-shuffle prefix products and sums.  Reference lines are cited per kernel.
#include <math_constants.h>

#include "common.cuh"

namespace dfn {

static constexpr int kThreads = 256;

static inline int grid_for(int64_t n, int threads = kThreads, int max_waves = 8) {
  int64_t blocks = (n + threads - 1) / threads;
  int64_t cap = (int64_t)num_sms() * max_waves * (2048 / threads);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ------------------------------------------------------------------------------- a1 get_rays
struct Pose {
  float m[12];
};

// HELP:449-465.  dirs = ((x-cx)/f, -(y-cy)/f, -1); rays_d[a] = (dx*R[a][0] + dy*R[a][1]) + dz*R[a][2]
// with every product and sum rounded separately in that order (what torch.sum over the size-3
// axis does), so rays_d is bit-identical to the reference.
__global__ void get_rays_kernel(int n_rows, int n_cols, const float* __restrict__ xs,
                                const float* __restrict__ ys, float focal, float cx, float cy,
                                Pose c2w, float* __restrict__ rays_o, float* __restrict__ rays_d,
                                float* __restrict__ viewdirs) {
  int64_t n = (int64_t)n_rows * n_cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int row = (int)(i / n_cols), col = (int)(i % n_cols);
    float dx = __fdiv_rn(__fsub_rn(xs[col], cx), focal);
    float dy = -__fdiv_rn(__fsub_rn(ys[row], cy), focal);
    float dz = -1.0f;
    float d[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float p0 = __fmul_rn(dx, c2w.m[a * 4 + 0]);
      float p1 = __fmul_rn(dy, c2w.m[a * 4 + 1]);
      float p2 = __fmul_rn(dz, c2w.m[a * 4 + 2]);
      d[a] = __fadd_rn(__fadd_rn(p0, p1), p2);
    }
    if (rays_o) {
      rays_o[i * 3 + 0] = c2w.m[3];
      rays_o[i * 3 + 1] = c2w.m[7];
      rays_o[i * 3 + 2] = c2w.m[11];
    }
    rays_d[i * 3 + 0] = d[0];
    rays_d[i * 3 + 1] = d[1];
    rays_d[i * 3 + 2] = d[2];
    if (viewdirs) {
      float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])),
                                  __fmul_rn(d[2], d[2])));
      viewdirs[i * 3 + 0] = __fdiv_rn(d[0], nrm);
      viewdirs[i * 3 + 1] = __fdiv_rn(d[1], nrm);
      viewdirs[i * 3 + 2] = __fdiv_rn(d[2], nrm);
    }
  }
}

// ------------------------------------------------------------------------------- a2 z sampling
// MAIN:617-619: z = near*(1-t) + far*t (separately rounded).  Optional stratified jitter
// (upstream render_rays): mids=.5(z[1:]+z[:-1]); lower+(upper-lower)*rand.
__global__ void z_vals_kernel(int R, int S, const float* __restrict__ t_vals,
                              const float* __restrict__ near, const float* __restrict__ far,
                              const float* __restrict__ rnd, float* __restrict__ z_out) {
  int64_t n = (int64_t)R * S;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int r = (int)(i / S), s = (int)(i % S);
    float nr = near[r], fr = far[r];
    auto zf = [&](int k) {
      float t = t_vals[k];
      return __fadd_rn(__fmul_rn(nr, __fsub_rn(1.0f, t)), __fmul_rn(fr, t));
    };
    float z = zf(s);
    if (rnd) {
      float lower = s == 0 ? z : __fmul_rn(0.5f, __fadd_rn(z, zf(s - 1)));
      float upper = s == S - 1 ? z : __fmul_rn(0.5f, __fadd_rn(zf(s + 1), z));
      z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), rnd[i]));
    }
    z_out[i] = z;
  }
}

// ------------------------------------------------------------------------------- a3 points
// MAIN:638-641: p = o + d*z (product and sum rounded separately) and the ray direction repeated per sample.
__global__ void make_points_kernel(int R, int S, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                   const float* __restrict__ z_vals, float* __restrict__ pts, float* __restrict__ dirs) {
  const int64_t n = (int64_t)R * S * 3;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pt = i / 3;
    const int c = (int)(i % 3);
    const int64_t ray = pt / S;
    const float d = rays_d[ray * 3 + c];
    if (pts) pts[i] = __fadd_rn(rays_o[ray * 3 + c], __fmul_rn(d, z_vals[pt]));
    if (dirs) dirs[i] = d;
  }
}

// ------------------------------------------------------------------------- a4 / a4' encodings
// One thread per output element so the [P, D] store is coalesced.
// kind 0 (HELP:21-52): [x | sin(2^k x) | cos(2^k x)]_k, exact power-of-two scaling, no pi.
// kind 1 (DEC:257-275): p/2 then [sin(2^k pi p) | cos(2^k pi p)]_k with fl32(2^k*pi) as in torch.
// kind 2: kind 1 applied to x/||x|| (the view-direction branch, DEC:337-338).
__global__ void embed_kernel(int64_t P, const float* __restrict__ x, int L, int kind,
                             float* __restrict__ out) {
  int D = kind == 0 ? 3 + 6 * L : 6 * L;
  int64_t n = P * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / D;
    int j = (int)(i % D);
    float v;
    if (kind == 0) {
      if (j < 3) {
        v = x[p * 3 + j];
      } else {
        int k = (j - 3) / 6, r = (j - 3) % 6;
        float a = __fmul_rn(x[p * 3 + (r % 3)], pow2i(k));
        v = r < 3 ? sinf(a) : cosf(a);
      }
    } else {
      int k = j / 6, r = j % 6;
      float xv = x[p * 3 + (r % 3)];
      if (kind == 2) {
        const float a0 = x[p * 3], a1 = x[p * 3 + 1], a2 = x[p * 3 + 2];
        xv = __fdiv_rn(xv, sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1)), __fmul_rn(a2, a2))));
      }
      float ph = __fmul_rn(xv, 0.5f);
      float a = __fmul_rn(__fmul_rn(pow2i(k), CUDART_PI_F), ph);
      v = r < 3 ? sinf(a) : cosf(a);
    }
    out[i] = v;
  }
}

// --------------------------------------------------------------------- a7 composite_function
// MAIN:146-166 (sum composition): den = sum_b sigma_b (0 -> 1e-4); feat = sum_b feat_b*sigma_b/den.
__global__ void composite_fields_kernel(int n_box, int64_t n, const float* __restrict__ sigma,
                                        const float* __restrict__ feat,
                                        float* __restrict__ sigma_sum, float* __restrict__ feat_w) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (n_box == 1) {
      sigma_sum[i] = sigma[i];
      feat_w[i * 3 + 0] = feat[i * 3 + 0];
      feat_w[i * 3 + 1] = feat[i * 3 + 1];
      feat_w[i * 3 + 2] = feat[i * 3 + 2];
      continue;
    }
    float den = 0.f;
    for (int b = 0; b < n_box; ++b) den = b == 0 ? sigma[i] : __fadd_rn(den, sigma[b * n + i]);
    float ssum = den;
    if (den == 0.f) den = 1e-4f;
    float acc[3] = {0.f, 0.f, 0.f};
    for (int b = 0; b < n_box; ++b) {
      float w = __fdiv_rn(sigma[b * n + i], den);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float t = __fmul_rn(feat[(b * n + i) * 3 + c], w);
        acc[c] = b == 0 ? t : __fadd_rn(acc[c], t);
      }
    }
    sigma_sum[i] = ssum;
    feat_w[i * 3 + 0] = acc[0];
    feat_w[i * 3 + 1] = acc[1];
    feat_w[i * 3 + 2] = acc[2];
  }
}

// ------------------------------------------------------ a8 / a9 weights and raw2outputs
__device__ __forceinline__ double shfl_up_f64(double v, int delta) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_up_sync(0xffffffffu, lo, delta);
  hi = __shfl_up_sync(0xffffffffu, hi, delta);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_xor_f64(double v, int m) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(0xffffffffu, lo, m);
  hi = __shfl_xor_sync(0xffffffffu, hi, m);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

static constexpr int kMaxSeg = 8;  // samples per lane: S <= 256

// One warp per ray.  Lane l owns the contiguous samples [l*seg, (l+1)*seg).
// MAIN:169-179: alpha = 1-exp(-(relu(sigma)+1e-6)*dist*|d|); w = alpha * prod_{j<s}(1-alpha_j+1e-10).
// The running product is carried in fp64 and rounded per element, as torch's CPU cumprod does;
// a warp-shuffle exclusive scan combines the per-lane segment products.
// MODE 0: sigma [R,S] -> weights.  MODE 1: raw [R,S,4] -> maps (raw2outputs).
// SEG = samples per lane (compile-time, >= ceil(S/32)): the per-lane arrays live in registers, so a 64-sample ray (SEG 2)
// does not pay the register footprint -- and the occupancy -- of a 256-sample one.
template <int MODE, int SEG>
__global__ void __launch_bounds__(256, 4) volume_weights_kernel(int R, int S, const float* __restrict__ src,
                                      const float* __restrict__ z_vals,
                                      const float* __restrict__ rays_d,
                                      const float* __restrict__ bc_rgb, int raw_is_feat,
                                      int white_bkgd, float last_dist, float* __restrict__ rgb_map,
                                      float* __restrict__ disp_map, float* __restrict__ acc_map,
                                      float* __restrict__ weights, float* __restrict__ depth_map,
                                      float* __restrict__ last_weight) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int seg = (S + 31) / 32;
  for (int ray = blockIdx.x * warps_per_block + (threadIdx.x >> 5); ray < R;
       ray += gridDim.x * warps_per_block) {
    const float dx = rays_d[ray * 3 + 0], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    const float* zr = z_vals + (int64_t)ray * S;
    float alpha[SEG], zv[SEG], col[SEG][3];
    double local = 1.0;
    const int s0 = lane * seg;
#pragma unroll
    for (int k = 0; k < SEG; ++k) {
      int s = s0 + k;
      alpha[k] = 0.f;
      zv[k] = 0.f;
      if (k < seg && s < S) {
        float sig;
        if (MODE == 0) {
          sig = src[(int64_t)ray * S + s];
        } else {
          float4 rv = reinterpret_cast<const float4*>(src)[(int64_t)ray * S + s];
          sig = rv.w;
          if (raw_is_feat) {
            col[k][0] = rv.x; col[k][1] = rv.y; col[k][2] = rv.z;
          } else {
            col[k][0] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-rv.x)));
            col[k][1] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-rv.y)));
            col[k][2] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-rv.z)));
          }
          if (s == S - 1 && bc_rgb) {
            col[k][0] = bc_rgb[ray * 3 + 0]; col[k][1] = bc_rgb[ray * 3 + 1]; col[k][2] = bc_rgb[ray * 3 + 2];
          }
        }
        float z = zr[s];
        zv[k] = z;
        float dist = s == S - 1 ? last_dist : __fsub_rn(zr[s + 1], z);
        dist = __fmul_rn(dist, nrm);
        float t = __fadd_rn(fmaxf(sig, 0.f), 1e-6f);
        float a = __fsub_rn(1.0f, expf(-__fmul_rn(t, dist)));
        alpha[k] = a;
        local *= (double)__fadd_rn(__fsub_rn(1.0f, a), 1e-10f);
      }
    }
    // exclusive scan of the lane products
    double incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      double up = shfl_up_f64(incl, d);
      if (lane >= d) incl *= up;
    }
    double run = shfl_up_f64(incl, 1);
    if (lane == 0) run = 1.0;
    float acc_rgb[3] = {0.f, 0.f, 0.f}, acc_w = 0.f, acc_d = 0.f;
#pragma unroll
    for (int k = 0; k < SEG; ++k) {
      int s = s0 + k;
      if (k < seg && s < S) {
        float T = (float)run;
        float w = __fmul_rn(alpha[k], T);
        run *= (double)__fadd_rn(__fsub_rn(1.0f, alpha[k]), 1e-10f);
        if (weights) weights[(int64_t)ray * S + s] = w;
        if (last_weight && s == S - 1) last_weight[ray] = w;
        if (MODE == 1) {
          acc_rgb[0] += w * col[k][0];
          acc_rgb[1] += w * col[k][1];
          acc_rgb[2] += w * col[k][2];
          acc_w += w;
          acc_d += w * zv[k];
        }
      }
    }
    if (MODE == 1) {
      acc_rgb[0] = warp_sum(acc_rgb[0]);
      acc_rgb[1] = warp_sum(acc_rgb[1]);
      acc_rgb[2] = warp_sum(acc_rgb[2]);
      acc_w = warp_sum(acc_w);
      acc_d = warp_sum(acc_d);
      if (lane == 0) {
        if (white_bkgd) {
          float bgw = 1.0f - acc_w;
          acc_rgb[0] += bgw; acc_rgb[1] += bgw; acc_rgb[2] += bgw;
        }
        if (rgb_map) {
          rgb_map[ray * 3 + 0] = acc_rgb[0]; rgb_map[ray * 3 + 1] = acc_rgb[1]; rgb_map[ray * 3 + 2] = acc_rgb[2];
        }
        if (acc_map) acc_map[ray] = acc_w;
        if (depth_map) depth_map[ray] = acc_d;
        if (disp_map) disp_map[ray] = __fdiv_rn(1.0f, fmaxf(1e-10f, __fdiv_rn(acc_d, acc_w)));
      }
    }
  }
}

// ------------------------------------------------ a6 + a7 + a8 + a9: live head + torso compositing
// One warp per ray, MAIN:669-708 with concate_bg: head colour of the last sample <- background pixel; torso
// sigma of the last sample <- 0; relu; +1e-6 on the last sample of the LAST field of each stack (head for the
// head-only image, torso for the person image); two-field mix den = s_h + s_t (0 -> 1e-4),
// feat = f_h*(s_h/den) + f_t*(s_t/den), sigma = s_h + s_t (MAIN:146-166); weights with the head rays' norm for
// the head image and the torso rays' norm for the person image (MAIN:704-705); rgb = sum w*feat.
// feat_* / sig_* are addressed with element strides (3 and 1 for separate tensors; 4 and 4 for the fused kernels'
// interleaved raw [R,S,4] = (feat, sigma)).
template <int SEG>
__global__ void __launch_bounds__(256, 4) head_torso_kernel(int R, int S, const float* __restrict__ feat_h, int fs_h, const float* __restrict__ sig_h,
                                  int ss_h, const float* __restrict__ feat_t, int fs_t, const float* __restrict__ sig_t,
                                  int ss_t, const float* __restrict__ bc_rgb, const float* __restrict__ z_vals,
                                  const float* __restrict__ rays_d_h, const float* __restrict__ rays_d_t,
                                  float last_dist, float* __restrict__ rgb_head, float* __restrict__ rgb_person) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int seg = (S + 31) / 32;
  for (int ray = blockIdx.x * warps_per_block + (threadIdx.x >> 5); ray < R; ray += gridDim.x * warps_per_block) {
    float nrm[2];
    {
      const float* d0 = rays_d_h + ray * 3;
      const float* d1 = rays_d_t + ray * 3;
      nrm[0] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d0[0], d0[0]), __fmul_rn(d0[1], d0[1])), __fmul_rn(d0[2], d0[2])));
      nrm[1] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d1[0], d1[0]), __fmul_rn(d1[1], d1[1])), __fmul_rn(d1[2], d1[2])));
    }
    const float* zr = z_vals + (int64_t)ray * S;
    float alpha[2][SEG], col[2][SEG][3];
    double local[2] = {1.0, 1.0};
    const int s0 = lane * seg;
#pragma unroll
    for (int k = 0; k < SEG; ++k) {
      const int s = s0 + k;
      alpha[0][k] = alpha[1][k] = 0.f;
      if (k < seg && s < S) {
        const int64_t i = (int64_t)ray * S + s;
        const bool last = s == S - 1;
        float fh[3] = {feat_h[i * fs_h], feat_h[i * fs_h + 1], feat_h[i * fs_h + 2]};
        if (last) {
          fh[0] = bc_rgb[ray * 3];
          fh[1] = bc_rgb[ray * 3 + 1];
          fh[2] = bc_rgb[ray * 3 + 2];
        }
        const float ft[3] = {feat_t[i * fs_t], feat_t[i * fs_t + 1], feat_t[i * fs_t + 2]};
        const float sh = fmaxf(sig_h[i * ss_h], 0.f);
        float st = last ? 0.f : fmaxf(sig_t[i * ss_t], 0.f);
        const float sh1 = last ? __fadd_rn(sh, 1e-6f) : sh;  // head-only stack: the head is the last field
        if (last) st = __fadd_rn(st, 1e-6f);                  // two-field stack: the torso is the last field
        float den = __fadd_rn(sh, st);
        const float ssum = den;
        if (den == 0.f) den = 1e-4f;
        const float wh = __fdiv_rn(sh, den), wt = __fdiv_rn(st, den);
        const float dz = last ? last_dist : __fsub_rn(zr[s + 1], zr[s]);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          col[0][k][c] = fh[c];
          col[1][k][c] = __fadd_rn(__fmul_rn(fh[c], wh), __fmul_rn(ft[c], wt));
        }
        const float sg[2] = {sh1, ssum};
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          const float dist = __fmul_rn(dz, nrm[f]);
          const float a = __fsub_rn(1.0f, expf(-__fmul_rn(__fadd_rn(fmaxf(sg[f], 0.f), 1e-6f), dist)));
          alpha[f][k] = a;
          local[f] *= (double)__fadd_rn(__fsub_rn(1.0f, a), 1e-10f);
        }
      }
    }
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      double incl = local[f];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        double up = shfl_up_f64(incl, d);
        if (lane >= d) incl *= up;
      }
      double run = shfl_up_f64(incl, 1);
      if (lane == 0) run = 1.0;
      float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < SEG; ++k) {
        const int s = s0 + k;
        if (k < seg && s < S) {
          const float w = __fmul_rn(alpha[f][k], (float)run);
          run *= (double)__fadd_rn(__fsub_rn(1.0f, alpha[f][k]), 1e-10f);
          acc[0] += w * col[f][k][0];
          acc[1] += w * col[f][k][1];
          acc[2] += w * col[f][k][2];
        }
      }
      acc[0] = warp_sum(acc[0]);
      acc[1] = warp_sum(acc[1]);
      acc[2] = warp_sum(acc[2]);
      float* out = f == 0 ? rgb_head : rgb_person;
      if (lane == 0 && out) {
        out[ray * 3 + 0] = acc[0];
        out[ray * 3 + 1] = acc[1];
        out[ray * 3 + 2] = acc[2];
      }
    }
  }
}

// ---------------------------------------------------------------------------- a10 sample_pdf
static constexpr int kMaxBins = 256;

// One warp per ray (HELP:537-581).  cdf: w+=1e-5; pdf = w/sum(w); running fp64 sum rounded per
// element (torch CPU cumsum).  The row sum is taken in fp64 (order-independent to fp32 rounding).
// Inversion: inds = #(cdf <= u) (searchsorted right=True); below=max(0,inds-1); above=min(nb-1,inds);
// denom<1e-5 -> 1; sample = bins_b + (u-cdf_b)/denom*(bins_a-bins_b), each op rounded separately.
template <bool HAVE_CDF>
__global__ void sample_pdf_kernel(int R, int nb, const float* __restrict__ bins,
                                  const float* __restrict__ win, int64_t w_stride, int N,
                                  const float* __restrict__ u, int u_per_ray,
                                  float* __restrict__ samples, int64_t* __restrict__ inds) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  float* cdf = sm + (size_t)wib * 2 * kMaxBins;
  float* bn = cdf + kMaxBins;
  for (int ray = blockIdx.x * warps_per_block + wib; ray < R; ray += gridDim.x * warps_per_block) {
    __syncwarp();
    for (int i = lane; i < nb; i += 32) bn[i] = bins[(int64_t)ray * nb + i];
    if (HAVE_CDF) {
      for (int i = lane; i < nb; i += 32) cdf[i] = win[(int64_t)ray * nb + i];
    } else {
      const int nw = nb - 1;
      const int seg = (nw + 31) / 32;
      const float* wr = win + (int64_t)ray * w_stride;
      float wv[kMaxBins / 32];
      double part = 0.0;
#pragma unroll
      for (int k = 0; k < kMaxBins / 32; ++k) {
        int i = lane * seg + k;
        wv[k] = 0.f;
        if (k < seg && i < nw) {
          wv[k] = __fadd_rn(wr[i], 1e-5f);
          part += (double)wv[k];
        }
      }
      double tot = part;
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) tot += shfl_xor_f64(tot, m);
      const float total = (float)tot;
      double lsum = 0.0;
#pragma unroll
      for (int k = 0; k < kMaxBins / 32; ++k) {
        int i = lane * seg + k;
        if (k < seg && i < nw) {
          wv[k] = __fdiv_rn(wv[k], total);
          lsum += (double)wv[k];
        }
      }
      double incl = lsum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        double up = shfl_up_f64(incl, d);
        if (lane >= d) incl += up;
      }
      double run = shfl_up_f64(incl, 1);
      if (lane == 0) {
        run = 0.0;
        cdf[0] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < kMaxBins / 32; ++k) {
        int i = lane * seg + k;
        if (k < seg && i < nw) {
          run += (double)wv[k];
          cdf[i + 1] = (float)run;
        }
      }
    }
    __syncwarp();
    for (int j = lane; j < N; j += 32) {
      float uu = u_per_ray ? u[(int64_t)ray * N + j] : u[j];
      int lo = 0, hi = nb;  // first index with cdf > uu
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (cdf[mid] <= uu) lo = mid + 1; else hi = mid;
      }
      int below = max(0, lo - 1), above = min(nb - 1, lo);
      float cb = cdf[below], ca = cdf[above];
      float den = __fsub_rn(ca, cb);
      if (den < 1e-5f) den = 1.0f;
      float t = __fdiv_rn(__fsub_rn(uu, cb), den);
      float bb = bn[below], ba = bn[above];
      samples[(int64_t)ray * N + j] = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
      if (inds) inds[(int64_t)ray * N + j] = lo;
    }
  }
}

// ----------------------------------------------------------------------------- a11 sort-merge
// out = sort(cat(a, b)) per ray (upstream render_rays); values only, so the result equals torch.sort's.
// Fast path: at render time both runs are already ascending (uniform z_vals; sample_pdf of a sorted u is monotone), so
// every element's output position is its own index plus its rank in the other run (two binary searches in shared
// memory) -- checked per ray, warp-uniformly.  Otherwise (stratified / random u, NaNs): bitonic network over the next
// power of two (padding +inf).
__global__ void sort_merge_kernel(int R, int na, const float* __restrict__ a, int nb,
                                  const float* __restrict__ b, float* __restrict__ out, int npow2) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  float* v = sm + (size_t)wib * 2 * npow2;
  const int n = na + nb;
  for (int ray = blockIdx.x * warps_per_block + wib; ray < R; ray += gridDim.x * warps_per_block) {
    __syncwarp();
    for (int i = lane; i < npow2; i += 32) {
      float x = CUDART_INF_F;
      if (i < na) x = a[(int64_t)ray * na + i];
      else if (i < n) x = b[(int64_t)ray * nb + (i - na)];
      v[i] = x;
    }
    __syncwarp();
    {
      bool sorted = true;
      for (int i = lane; i < n; i += 32)
        if (i != 0 && i != na) sorted = sorted && (v[i - 1] <= v[i]);
      if (__all_sync(0xFFFFFFFFu, sorted)) {
        float* w = v + npow2;   // second buffer: merged run, stored coalesced afterwards
        for (int i = lane; i < n; i += 32) {
          const float x = v[i];
          const bool from_a = i < na;
          const float* other = from_a ? v + na : v;
          int lo = 0, hi = from_a ? nb : na;
          while (lo < hi) {   // from a: #{b < x} (lower bound); from b: #{a <= x} (upper bound) -> distinct positions on ties
            const int mid = (lo + hi) >> 1;
            const float y = other[mid];
            if (from_a ? (y < x) : (y <= x)) lo = mid + 1;
            else hi = mid;
          }
          w[(from_a ? i : i - na) + lo] = x;
        }
        __syncwarp();
        for (int i = lane; i < n; i += 32) out[(int64_t)ray * n + i] = w[i];
        continue;
      }
    }
    for (int k = 2; k <= npow2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = lane; t < (npow2 >> 1); t += 32) {
          int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j clear
          int p = i | j;
          bool up = (i & k) == 0;
          float x = v[i], y = v[p];
          if ((x > y) == up) {
            v[i] = y;
            v[p] = x;
          }
        }
        __syncwarp();
      }
    }
    for (int i = lane; i < n; i += 32) out[(int64_t)ray * n + i] = v[i];
  }
}

// ------------------------------------------------- coarse pass -> fine depths, one kernel (a8 + a9 + a10 + a11)
// raw2outputs(coarse) -> z_mid -> sample_pdf -> sort(cat(z, z_samples)) of upstream render_rays for one
// ray per warp, with the arithmetic of volume_weights_kernel<1>, sample_pdf_kernel<false> and sort_merge_kernel above -- the
// same operations in the same order, so the merged depths are bit-identical to the chain of separate launches -- but the
// weights, the bin midpoints, the cdf and the new samples live in shared memory instead of making five round trips
// through HBM (w0, z_mid, z_samples written and re-read), and the frame needs one launch instead of four.
// Per warp: the ray's raw row [Nc,4] and depth row [Nc] are fetched with coalesced loads into shared memory (lane l then
// owns the contiguous samples [l*seg, (l+1)*seg) for the fp64 transmittance scan); rgb0 is formed only when requested.
struct C2FSmem {   // floats per warp
  static __host__ __device__ int raw(int Nc) { return 4 * Nc; }
  static __host__ __device__ int total(int Nc, int npow2) { return 4 * Nc + 4 * Nc + 2 * npow2; }   // raw | w, zmid, cdf, pad | v, merged
};

template <int SEG>
__global__ void __launch_bounds__(256, 4) coarse_to_fine_kernel(int R, int Nc, int Nf, const float* __restrict__ raw0, const float* __restrict__ z0,
                                      const float* __restrict__ rays_d, const float* __restrict__ bc_rgb, int white_bkgd,
                                      float last_dist, const float* __restrict__ u, int u_per_ray,
                                      const float* __restrict__ zs_in, float* __restrict__ rgb0, float* __restrict__ zs_out,
                                      float* __restrict__ zall, int npow2) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  float* base = sm + (size_t)wib * C2FSmem::total(Nc, npow2);
  float4* rawS = reinterpret_cast<float4*>(base);
  float* wgt = base + 4 * Nc;          // [Nc] weights
  float* zmid = wgt + Nc;              // [Nc-1]
  float* cdf = zmid + Nc;              // [Nc-1]
  float* v = cdf + 2 * Nc;             // [npow2]: z (Nc) | samples (Nf) | +inf
  float* mrg = v + npow2;              // [npow2] merged run
  const int seg = (Nc + 31) / 32;
  const int n = Nc + Nf;
  const int nb = Nc - 1, nw = Nc - 2;  // bins = z_mid, weights[..., 1:-1]
  for (int ray = blockIdx.x * warps_per_block + wib; ray < R; ray += gridDim.x * warps_per_block) {
    __syncwarp();
    // ---- coalesced fetch of the ray's rows
    const float4* rsrc = reinterpret_cast<const float4*>(raw0) + (int64_t)ray * Nc;
    for (int i = lane; i < Nc; i += 32) {
      rawS[i] = rsrc[i];
      v[i] = z0[(int64_t)ray * Nc + i];
    }
    const float dx = rays_d[ray * 3 + 0], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    __syncwarp();
    // ---- weights (MAIN:169-179) [+ colour sum, MAIN:706] exactly as volume_weights_kernel<1, SEG>
    float alpha[SEG], col[SEG][3];
    double local = 1.0;
    const int s0 = lane * seg;
#pragma unroll
    for (int k = 0; k < SEG; ++k) {
      const int s = s0 + k;
      alpha[k] = 0.f;
      if (k < seg && s < Nc) {
        const float4 rv = rawS[s];
        if (rgb0) {
          col[k][0] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-rv.x)));
          col[k][1] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-rv.y)));
          col[k][2] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-rv.z)));
          if (s == Nc - 1 && bc_rgb) {
            col[k][0] = bc_rgb[ray * 3 + 0]; col[k][1] = bc_rgb[ray * 3 + 1]; col[k][2] = bc_rgb[ray * 3 + 2];
          }
        }
        const float z = v[s];
        float dist = s == Nc - 1 ? last_dist : __fsub_rn(v[s + 1], z);
        dist = __fmul_rn(dist, nrm);
        const float t = __fadd_rn(fmaxf(rv.w, 0.f), 1e-6f);
        const float a = __fsub_rn(1.0f, expf(-__fmul_rn(t, dist)));
        alpha[k] = a;
        local *= (double)__fadd_rn(__fsub_rn(1.0f, a), 1e-10f);
      }
    }
    double incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      double up = shfl_up_f64(incl, d);
      if (lane >= d) incl *= up;
    }
    double run = shfl_up_f64(incl, 1);
    if (lane == 0) run = 1.0;
    float acc_rgb[3] = {0.f, 0.f, 0.f}, acc_w = 0.f;
#pragma unroll
    for (int k = 0; k < SEG; ++k) {
      const int s = s0 + k;
      if (k < seg && s < Nc) {
        const float T = (float)run;
        const float w = __fmul_rn(alpha[k], T);
        run *= (double)__fadd_rn(__fsub_rn(1.0f, alpha[k]), 1e-10f);
        wgt[s] = w;
        if (rgb0) {
          acc_rgb[0] += w * col[k][0];
          acc_rgb[1] += w * col[k][1];
          acc_rgb[2] += w * col[k][2];
          acc_w += w;
        }
      }
    }
    if (rgb0) {
      acc_rgb[0] = warp_sum(acc_rgb[0]);
      acc_rgb[1] = warp_sum(acc_rgb[1]);
      acc_rgb[2] = warp_sum(acc_rgb[2]);
      acc_w = warp_sum(acc_w);
      if (lane == 0) {
        if (white_bkgd) {
          const float bgw = 1.0f - acc_w;
          acc_rgb[0] += bgw; acc_rgb[1] += bgw; acc_rgb[2] += bgw;
        }
        rgb0[ray * 3 + 0] = acc_rgb[0]; rgb0[ray * 3 + 1] = acc_rgb[1]; rgb0[ray * 3 + 2] = acc_rgb[2];
      }
    }
    __syncwarp();
    if (zs_in == nullptr) {
      // ---- z_mid and the cdf of weights[..., 1:-1] + 1e-5 (HELP:539-544), as sample_pdf_kernel<false>
      for (int i = lane; i < nb; i += 32) zmid[i] = __fmul_rn(0.5f, __fadd_rn(v[i + 1], v[i]));
      {
        const int sg = (nw + 31) / 32;
        float wv[SEG];   // (Nc - 2 + 31) / 32 <= SEG
        double part = 0.0;
#pragma unroll
        for (int k = 0; k < SEG; ++k) {
          const int i = lane * sg + k;
          wv[k] = 0.f;
          if (k < sg && i < nw) {
            wv[k] = __fadd_rn(wgt[1 + i], 1e-5f);
            part += (double)wv[k];
          }
        }
        double tot = part;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) tot += shfl_xor_f64(tot, m);
        const float total = (float)tot;
        double lsum = 0.0;
#pragma unroll
        for (int k = 0; k < SEG; ++k) {
          const int i = lane * sg + k;
          if (k < sg && i < nw) {
            wv[k] = __fdiv_rn(wv[k], total);
            lsum += (double)wv[k];
          }
        }
        double inc2 = lsum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          double up = shfl_up_f64(inc2, d);
          if (lane >= d) inc2 += up;
        }
        double r2 = shfl_up_f64(inc2, 1);
        if (lane == 0) {
          r2 = 0.0;
          cdf[0] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < SEG; ++k) {
          const int i = lane * sg + k;
          if (k < sg && i < nw) {
            r2 += (double)wv[k];
            cdf[i + 1] = (float)r2;
          }
        }
      }
      __syncwarp();
      // ---- inverse-cdf samples (HELP:563-579)
      for (int j = lane; j < Nf; j += 32) {
        const float uu = u_per_ray ? u[(int64_t)ray * Nf + j] : u[j];
        int lo = 0, hi = nb;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (cdf[mid] <= uu) lo = mid + 1; else hi = mid;
        }
        const int below = max(0, lo - 1), above = min(nb - 1, lo);
        const float cb = cdf[below], ca = cdf[above];
        float den = __fsub_rn(ca, cb);
        if (den < 1e-5f) den = 1.0f;
        const float t = __fdiv_rn(__fsub_rn(uu, cb), den);
        const float bb = zmid[below], ba = zmid[above];
        const float smp = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
        v[Nc + j] = smp;
        if (zs_out) zs_out[(int64_t)ray * Nf + j] = smp;
      }
    } else {
      for (int j = lane; j < Nf; j += 32) {
        const float smp = zs_in[(int64_t)ray * Nf + j];
        v[Nc + j] = smp;
        if (zs_out && zs_out != zs_in) zs_out[(int64_t)ray * Nf + j] = smp;
      }
    }
    for (int i = n + lane; i < npow2; i += 32) v[i] = CUDART_INF_F;
    __syncwarp();
    // ---- sort(cat(z, z_samples)) as sort_merge_kernel: merge by rank when both runs ascend, bitonic network otherwise
    {
      bool sorted = true;
      for (int i = lane; i < n; i += 32)
        if (i != 0 && i != Nc) sorted = sorted && (v[i - 1] <= v[i]);
      if (__all_sync(0xFFFFFFFFu, sorted)) {
        // merge path: lane l produces the outputs [l*per, (l+1)*per) -- ONE binary search along its diagonal for the split between the
        // two runs (ties take the coarse depth first, as the by-rank merge of sort_merge_kernel does), then `per` sequential steps;
        // the by-rank form cost a binary search per element (a sixth of this kernel's instructions).  Values only: the same output.
        const float* a = v;
        const float* b = v + Nc;
        const int per = (n + 31) >> 5;
        const int d = min(lane * per, n);
        int lo = max(0, d - Nf), hi = min(d, Nc);
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (a[mid] <= b[d - mid - 1]) lo = mid + 1;
          else hi = mid;
        }
        int ia = lo, ib = d - lo;
        for (int t = 0; t < per && d + t < n; ++t) {
          const bool take_a = ia < Nc && (ib >= Nf || a[ia] <= b[ib]);
          mrg[d + t] = take_a ? a[ia] : b[ib];
          ia += take_a ? 1 : 0;
          ib += take_a ? 0 : 1;
        }
        __syncwarp();
        for (int i = lane; i < n; i += 32) zall[(int64_t)ray * n + i] = mrg[i];
        continue;
      }
    }
    for (int k = 2; k <= npow2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = lane; t < (npow2 >> 1); t += 32) {
          const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          const int p = i | j;
          const bool up = (i & k) == 0;
          const float x = v[i], y = v[p];
          if ((x > y) == up) {
            v[i] = y;
            v[p] = x;
          }
        }
        __syncwarp();
      }
    }
    for (int i = lane; i < n; i += 32) zall[(int64_t)ray * n + i] = v[i];
  }
}

// ---------------------------------------------------------------------------- to8b (HELP:17)
// (255 * clip(x, 0, 1)).astype(uint8): fp32 product, truncation.  Four values per thread (one 32-bit store).
__global__ void to8b_kernel(int64_t n, const float* __restrict__ x, uint8_t* __restrict__ out) {
  const int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const uint32_t b0 = (uint32_t)__fmul_rn(255.f, fminf(fmaxf(v.x, 0.f), 1.f));
    const uint32_t b1 = (uint32_t)__fmul_rn(255.f, fminf(fmaxf(v.y, 0.f), 1.f));
    const uint32_t b2 = (uint32_t)__fmul_rn(255.f, fminf(fmaxf(v.z, 0.f), 1.f));
    const uint32_t b3 = (uint32_t)__fmul_rn(255.f, fminf(fmaxf(v.w, 0.f), 1.f));
    reinterpret_cast<uint32_t*>(out)[i] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    out[i] = (uint8_t)__fmul_rn(255.f, fminf(fmaxf(x[i], 0.f), 1.f));
  }
}

}  // namespace dfn

using namespace dfn;

extern "C" int dfn_to8b(int64_t n, const float* x, uint8_t* out, void* stream) {
  DFN_CHECK_ARG(n > 0 && x && out, "dfn_to8b: bad argument");
  DFN_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0,
                "dfn_to8b: x must be 16-byte and out 4-byte aligned");
  to8b_kernel<<<grid_for((n + 3) / 4), kThreads, 0, (cudaStream_t)stream>>>(n, x, out);
  DFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dfn_get_rays(int n_rows, int n_cols, const float* xs, const float* ys, float focal,
                            float cx, float cy, const float* c2w_host, float* rays_o, float* rays_d,
                            float* viewdirs, void* stream) {
  DFN_CHECK_ARG(n_rows > 0 && n_cols > 0 && xs && ys && c2w_host && rays_d, "dfn_get_rays: bad argument");
  Pose p;
  for (int i = 0; i < 12; ++i) p.m[i] = c2w_host[i];
  int64_t n = (int64_t)n_rows * n_cols;
  get_rays_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(n_rows, n_cols, xs, ys, focal, cx, cy, p,
                                                                     rays_o, rays_d, viewdirs);
  DFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dfn_z_vals(int R, int S, const float* t_vals, const float* near, const float* far,
                          const float* rnd, float* z_out, void* stream) {
  DFN_CHECK_ARG(R > 0 && S > 0 && t_vals && near && far && z_out, "dfn_z_vals: bad argument");
  z_vals_kernel<<<grid_for((int64_t)R * S), kThreads, 0, (cudaStream_t)stream>>>(R, S, t_vals, near, far, rnd, z_out);
  DFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dfn_make_points(int R, int S, const float* rays_o, const float* rays_d, const float* z_vals, float* pts,
                               float* dirs, void* stream) {
  DFN_CHECK_ARG(R > 0 && S > 0 && rays_o && rays_d && z_vals && (pts || dirs), "dfn_make_points: bad argument");
  make_points_kernel<<<grid_for((int64_t)R * S * 3), kThreads, 0, (cudaStream_t)stream>>>(R, S, rays_o, rays_d, z_vals, pts, dirs);
  DFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dfn_embed(int64_t P, const float* x, int L, int kind, float* out, void* stream) {
  DFN_CHECK_ARG(P > 0 && x && out && L > 0 && L <= 16 && kind >= 0 && kind <= 2, "dfn_embed: bad argument");
  int D = kind == 0 ? 3 + 6 * L : 6 * L;
  embed_kernel<<<grid_for(P * D), kThreads, 0, (cudaStream_t)stream>>>(P, x, L, kind, out);
  DFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dfn_composite_fields(int n_box, int64_t n, const float* sigma, const float* feat,
                                    float* sigma_sum, float* feat_w, void* stream) {
  DFN_CHECK_ARG(n_box >= 1 && n > 0 && sigma && feat && sigma_sum && feat_w, "dfn_composite_fields: bad argument");
  composite_fields_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(n_box, n, sigma, feat, sigma_sum, feat_w);
  DFN_LAUNCH_CHECK();
  return 0;
}

// instantiate KERNEL<..., SEG> for the smallest SEG in {1, 2, 4, 6, 8} that covers ceil(S/32)
#define DFN_SEG_DISPATCH(S, CALL)          \
  do {                                     \
    const int seg__ = ((S) + 31) / 32;     \
    if (seg__ <= 1) { CALL(1); }           \
    else if (seg__ <= 2) { CALL(2); }      \
    else if (seg__ <= 4) { CALL(4); }      \
    else if (seg__ <= 6) { CALL(6); }      \
    else { CALL(8); }                      \
  } while (0)

static int rays_grid(int R, int warps_per_block) {
  int64_t blocks = ((int64_t)R + warps_per_block - 1) / warps_per_block;
  int64_t cap = (int64_t)num_sms() * 16;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

extern "C" int dfn_calc_volume_weights(int R, int S, const float* z_vals, const float* ray_vector,
                                       const float* sigma, float last_dist, float* weights, void* stream) {
  DFN_CHECK_ARG(R > 0 && S > 0 && S <= 32 * kMaxSeg && z_vals && ray_vector && sigma && weights,
                "dfn_calc_volume_weights: bad argument (S <= 256)");
#define CALL_VW0(SEG)                                                                                     \
  volume_weights_kernel<0, SEG><<<rays_grid(R, 8), 256, 0, (cudaStream_t)stream>>>(                       \
      R, S, sigma, z_vals, ray_vector, nullptr, 0, 0, last_dist, nullptr, nullptr, nullptr, weights, nullptr, nullptr)
  DFN_SEG_DISPATCH(S, CALL_VW0);
#undef CALL_VW0
  DFN_LAUNCH_CHECK();
  return 0;
}

int dfn::launch_raw2outputs(int R, int S, const float* raw, const float* z_vals, const float* rays_d,
                            const float* bc_rgb, int raw_is_feat, int white_bkgd, float last_dist, float* rgb_map,
                            float* disp_map, float* acc_map, float* weights, float* depth_map, float* last_weight,
                            cudaStream_t st) {
  DFN_CHECK_ARG(R > 0 && S > 0 && S <= 32 * kMaxSeg && raw && z_vals && rays_d,
                "dfn_raw2outputs: bad argument (S <= 256)");
  DFN_CHECK_ARG((reinterpret_cast<uintptr_t>(raw) & 15) == 0, "dfn_raw2outputs: raw must be 16-byte aligned");
#define CALL_VW1(SEG)                                                                                          \
  volume_weights_kernel<1, SEG><<<rays_grid(R, 8), 256, 0, st>>>(R, S, raw, z_vals, rays_d, bc_rgb, raw_is_feat,       \
                                                                white_bkgd, last_dist, rgb_map, disp_map, acc_map,    \
                                                                weights, depth_map, last_weight)
  DFN_SEG_DISPATCH(S, CALL_VW1);
#undef CALL_VW1
  DFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dfn_raw2outputs(int R, int S, const float* raw, const float* z_vals, const float* rays_d,
                               const float* bc_rgb, int raw_is_feat, int white_bkgd, float last_dist,
                               float* rgb_map, float* disp_map, float* acc_map, float* weights,
                               float* depth_map, void* stream) {
  return launch_raw2outputs(R, S, raw, z_vals, rays_d, bc_rgb, raw_is_feat, white_bkgd, last_dist, rgb_map, disp_map,
                            acc_map, weights, depth_map, nullptr, (cudaStream_t)stream);
}

extern "C" int dfn_composite_head_torso(int R, int S, const float* feat_head, const float* sigma_head,
                                        const float* feat_torso, const float* sigma_torso, const float* bc_rgb,
                                        const float* z_vals, const float* rays_d_head, const float* rays_d_torso,
                                        float last_dist, float* rgb_head, float* rgb_person, void* stream) {
  DFN_CHECK_ARG(R > 0 && S > 0 && S <= 32 * kMaxSeg && feat_head && sigma_head && feat_torso && sigma_torso && bc_rgb &&
                    z_vals && rays_d_head && rays_d_torso,
                "dfn_composite_head_torso: bad argument (S <= 256)");
  return launch_head_torso(R, S, feat_head, 3, sigma_head, 1, feat_torso, 3, sigma_torso, 1, bc_rgb, z_vals, rays_d_head,
                           rays_d_torso, last_dist, rgb_head, rgb_person, (cudaStream_t)stream);
}

int dfn::launch_head_torso(int R, int S, const float* feat_h, int fstride_h, const float* sig_h, int sstride_h,
                           const float* feat_t, int fstride_t, const float* sig_t, int sstride_t, const float* bc_rgb,
                           const float* z_vals, const float* rays_d_h, const float* rays_d_t, float last_dist,
                           float* rgb_head, float* rgb_person, cudaStream_t st) {
#define CALL_HT(SEG)                                                                                             \
  head_torso_kernel<SEG><<<rays_grid(R, 8), 256, 0, st>>>(R, S, feat_h, fstride_h, sig_h, sstride_h, feat_t, fstride_t, \
                                                         sig_t, sstride_t, bc_rgb, z_vals, rays_d_h, rays_d_t,         \
                                                         last_dist, rgb_head, rgb_person)
  DFN_SEG_DISPATCH(S, CALL_HT);
#undef CALL_HT
  DFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dfn_sample_pdf(int R, int nb, const float* bins, const float* weights, int64_t w_stride,
                              int N, const float* u, int u_per_ray, float* samples, int64_t* inds,
                              void* stream) {
  DFN_CHECK_ARG(R > 0 && nb >= 2 && nb <= kMaxBins && bins && weights && N > 0 && u && samples && w_stride >= nb - 1,
                "dfn_sample_pdf: bad argument (nb <= 256)");
  const int wpb = 8;
  size_t smem = (size_t)wpb * 2 * kMaxBins * sizeof(float);
  sample_pdf_kernel<false><<<rays_grid(R, wpb), wpb * 32, smem, (cudaStream_t)stream>>>(
      R, nb, bins, weights, w_stride, N, u, u_per_ray, samples, inds);
  DFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dfn_invert_cdf(int R, int nb, const float* bins, const float* cdf, int N, const float* u,
                              int u_per_ray, float* samples, int64_t* inds, void* stream) {
  DFN_CHECK_ARG(R > 0 && nb >= 2 && nb <= kMaxBins && bins && cdf && N > 0 && u && samples,
                "dfn_invert_cdf: bad argument (nb <= 256)");
  const int wpb = 8;
  size_t smem = (size_t)wpb * 2 * kMaxBins * sizeof(float);
  sample_pdf_kernel<true><<<rays_grid(R, wpb), wpb * 32, smem, (cudaStream_t)stream>>>(
      R, nb, bins, cdf, nb, N, u, u_per_ray, samples, inds);
  DFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dfn_coarse_to_fine(int R, int N_samples, int N_importance, const float* raw0, const float* z_vals, const float* rays_d,
                                  const float* bc_rgb, int white_bkgd, float last_dist, const float* u, int u_per_ray,
                                  const float* z_samples_in, float* rgb0, float* z_samples_out, float* z_all, void* stream) {
  DFN_CHECK_ARG(R > 0 && N_samples >= 3 && N_samples <= 32 * 4 && N_importance > 0 && N_samples + N_importance <= 512 && raw0 &&
                    z_vals && rays_d && z_all && (u || z_samples_in),
                "dfn_coarse_to_fine: bad argument (3 <= N_samples <= 128, N_samples + N_importance <= 512)");
  DFN_CHECK_ARG((reinterpret_cast<uintptr_t>(raw0) & 15) == 0, "dfn_coarse_to_fine: raw0 must be 16-byte aligned");
  int npow2 = 2;
  while (npow2 < N_samples + N_importance) npow2 <<= 1;
  const int per_warp = C2FSmem::total(N_samples, npow2);
  int wpb = 8;
  while (wpb > 1 && (size_t)wpb * per_warp * sizeof(float) > 48 * 1024) wpb >>= 1;
  const size_t smem = (size_t)wpb * per_warp * sizeof(float);
  const int seg = (N_samples + 31) / 32;
#define CALL_C2F(SEG)                                                                                                     \
  coarse_to_fine_kernel<SEG><<<rays_grid(R, wpb), wpb * 32, smem, (cudaStream_t)stream>>>(                                \
      R, N_samples, N_importance, raw0, z_vals, rays_d, bc_rgb, white_bkgd, last_dist, u, u_per_ray, z_samples_in, rgb0,  \
      z_samples_out, z_all, npow2)
  if (seg <= 1) CALL_C2F(1);
  else if (seg <= 2) CALL_C2F(2);
  else CALL_C2F(4);
#undef CALL_C2F
  DFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dfn_sort_merge(int R, int na, const float* a, int nb, const float* b, float* out, void* stream) {
  DFN_CHECK_ARG(R > 0 && na >= 0 && nb >= 0 && na + nb > 0 && na + nb <= 1024 && out && (na == 0 || a) && (nb == 0 || b),
                "dfn_sort_merge: bad argument (na+nb <= 1024)");
  int npow2 = 2;
  while (npow2 < na + nb) npow2 <<= 1;
  int wpb = 8;
  while (wpb > 1 && (size_t)wpb * 2 * npow2 * sizeof(float) > 48 * 1024) wpb >>= 1;
  size_t smem = (size_t)wpb * 2 * npow2 * sizeof(float);
  sort_merge_kernel<<<rays_grid(R, wpb), wpb * 32, smem, (cudaStream_t)stream>>>(R, na, a, nb, b, out, npow2);
  DFN_LAUNCH_CHECK();
  return 0;
}
