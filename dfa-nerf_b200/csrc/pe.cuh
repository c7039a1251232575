// Positional encodings evaluated inside the tcgen05 kernels, one 64-wide K-block row per call.
//
// sin/cos: the argument is formed exactly as the reference forms it (exact 2^k scaling of an fp32 value), reduced to
// [-pi, pi] with a two-constant Cody-Waite step (error ~1e-8 for |arg| < 1e3) and evaluated with the MUFU sin/cos
// (abs error 2^-21.4 on that range): below the bf16 / fp16 / hi+lo resolution the MMA operands keep, and four times
// shorter than sincosf on a tile's critical path.
#pragma once
#include "common.cuh"

namespace dfn {
namespace tc {

__device__ __forceinline__ void sincos_reduced(float t, float& sv, float& cv) {
  const float n = rintf(t * 0.15915494309189535f);
  float r = fmaf(-n, 6.28125f, t);
  r = fmaf(-n, 1.9353071795864769e-3f, r);
  sv = __sinf(r);
  cv = __cosf(r);
}

// sample point x = o + d*z (MAIN:638-641), rounded as the reference's broadcast fma-free expression
__device__ __forceinline__ void sample_point(const float* __restrict__ rays_o, const float* __restrict__ rays_d, int64_t ray,
                                             float z, float (&x)[3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(rays_o[ray * 3 + c], __fmul_rn(rays_d[ray * 3 + c], z));
}

// HELP:42-52 Embedder: [x | sin(2^k x) | cos(2^k x)]_k, 3 + 6L <= 63 columns, the rest zero.
__device__ __forceinline__ void pe_embedder(const float (&x)[3], int L, float (&pe)[64]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) pe[c] = x[c];
#pragma unroll
  for (int k = 0; k < 10; ++k) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float sv = 0.f, cv = 0.f;
      if (k < L) sincos_reduced(__fmul_rn(x[c], pow2i(k)), sv, cv);
      pe[3 + 6 * k + c] = sv;
      pe[6 + 6 * k + c] = cv;
    }
  }
  pe[63] = 0.f;
}

// DEC:257-275 Decoder.transform_points: p /= 2; [sin(2^k pi p) | cos(2^k pi p)]_k, no identity term, 6L <= 60 columns.
// torch multiplies the fp32 point by fl32(2^k pi) = 2^k fl32(pi), so the argument is exactly 2^k * fl32(fl32(pi) * p).
__device__ __forceinline__ void pe_decoder(const float (&p)[3], int L, float (&pe)[64]) {
  float a0[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) a0[c] = __fmul_rn(3.14159274101257324f, __fmul_rn(p[c], 0.5f));
#pragma unroll
  for (int k = 0; k < 10; ++k) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float sv = 0.f, cv = 0.f;
      if (k < L) sincos_reduced(__fmul_rn(a0[c], pow2i(k)), sv, cv);
      pe[6 * k + c] = sv;
      pe[6 * k + 3 + c] = cv;
    }
  }
  pe[60] = pe[61] = pe[62] = pe[63] = 0.f;
}

// DEC:337-338: the view direction d / |d| (then through transform_points(views=True) = pe_decoder)
__device__ __forceinline__ void view_direction(const float* __restrict__ rays_d, int64_t ray, float (&dn)[3]) {
  const float dx = rays_d[ray * 3], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  dn[0] = __fdiv_rn(dx, nrm);
  dn[1] = __fdiv_rn(dy, nrm);
  dn[2] = __fdiv_rn(dz, nrm);
}
__device__ __forceinline__ void pe_decoder_viewdir(const float* __restrict__ rays_d, int64_t ray, int L, float (&pe)[64]) {
  float dn[3];
  view_direction(rays_d, ray, dn);
  pe_decoder(dn, L, pe);
}

}  // namespace tc
}  // namespace dfn
