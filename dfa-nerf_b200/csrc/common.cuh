// Shared helpers for libdfn (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dfn.h"

namespace dfn {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
void reset_launch_count();

#define DFN_CHECK_ARG(cond, ...)                \
  do {                                          \
    if (!(cond)) {                              \
      ::dfn::set_error(__VA_ARGS__);            \
      return DFN_E_ARG;                         \
    }                                           \
  } while (0)

#define DFN_CUDA(expr)                                                                    \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      ::dfn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                       __LINE__);                                                         \
      return (int)e__;                                                                    \
    }                                                                                     \
  } while (0)

#define DFN_LAUNCH_CHECK()                        \
  do {                                            \
    ::dfn::count_launch();                        \
    DFN_CUDA(cudaGetLastError());                 \
  } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

int num_sms();
// profiling hook (api.cu): returns false when disabled
bool profile_begin(cudaStream_t st, double macs);
void profile_end(cudaStream_t st);

// exact 2^k (exp2f is not guaranteed exact); the reference's frequency bands are exact powers of two
__device__ __forceinline__ float pow2i(int k) { return __int_as_float((127 + k) << 23); }

int launch_raw2outputs(int R, int S, const float* raw, const float* z_vals, const float* rays_d,
                       const float* bc_rgb, int raw_is_feat, int white_bkgd, float last_dist, float* rgb_map,
                       float* disp_map, float* acc_map, float* weights, float* depth_map, float* last_weight,
                       cudaStream_t st);

int launch_head_torso(int R, int S, const float* feat_h, int fstride_h, const float* sig_h, int sstride_h,
                      const float* feat_t, int fstride_t, const float* sig_t, int sstride_t, const float* bc_rgb,
                      const float* z_vals, const float* rays_d_h, const float* rays_d_t, float last_dist, float* rgb_head,
                      float* rgb_person, cudaStream_t st);

}  // namespace dfn
