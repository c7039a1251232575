// Micro-benchmark (sm_100a): the single-pass epilogue of mlp_tc.cu in isolation -- TMEM accumulator [128 x 256] fp32 -> (+ bias) -> ReLU ->
// bf16 -> swizzled K-major activation blocks in shared memory -- with no MMAs running, for different tcgen05.ld widths, bias sources and
// warps per TMEM lane quarter.  Cycles per 128 x 256 tile-layer.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../dfa-nerf_b200/csrc -o epilogue epilogue.cu && ./epilogue
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
#include "tc_epi.cuh"

using namespace dfn::tc;

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}

// BIAS: 0 none, 1 LDS.128 broadcast (product), 2 registers preloaded per 16 columns from a per-lane LDS + shuffles (not built)
template <int BIAS>
__device__ __forceinline__ void chunk8(const uint32_t* v, int col, uint32_t sbias, uint8_t* arena, uint32_t row) {
  float b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (BIAS == 1) lds_f32x8(sbias + (uint32_t)col * 4u, b);
  uint4 h;
  h.x = add_relu_pack(v[0], v[1], b[0], b[1]);
  h.y = add_relu_pack(v[2], v[3], b[2], b[3]);
  h.z = add_relu_pack(v[4], v[5], b[4], b[5]);
  h.w = add_relu_pack(v[6], v[7], b[6], b[7]);
  uint8_t* dst = arena + (size_t)(col >> 6) * KB_BYTES;
  *reinterpret_cast<uint4*>(dst + swz(row, (uint32_t)((col & 63) >> 3))) = h;
}

template <int LDW, int BIAS>
__device__ __forceinline__ void epi_cols(uint32_t acc, int c_begin, int c_end, uint32_t sbias, uint8_t* arena, uint32_t row) {
  if (LDW == 32) {
    uint32_t v0[32], v1[32];
    tmem_ld32(acc + c_begin, v0);
    for (int c = c_begin; c < c_end; c += 64) {
      tmem_ld_wait();
      tmem_ld32(acc + c + 32, v1);
#pragma unroll
      for (int g = 0; g < 4; ++g) chunk8<BIAS>(v0 + g * 8, c + g * 8, sbias, arena, row);
      tmem_ld_wait();
      if (c + 64 < c_end) tmem_ld32(acc + c + 64, v0);
#pragma unroll
      for (int g = 0; g < 4; ++g) chunk8<BIAS>(v1 + g * 8, c + 32 + g * 8, sbias, arena, row);
    }
  } else if (LDW == 16) {
    uint32_t v0[16], v1[16];
    tmem_ld16(acc + c_begin, v0);
    for (int c = c_begin; c < c_end; c += 32) {
      tmem_ld_wait();
      tmem_ld16(acc + c + 16, v1);
#pragma unroll
      for (int g = 0; g < 2; ++g) chunk8<BIAS>(v0 + g * 8, c + g * 8, sbias, arena, row);
      tmem_ld_wait();
      if (c + 32 < c_end) tmem_ld16(acc + c + 32, v0);
#pragma unroll
      for (int g = 0; g < 2; ++g) chunk8<BIAS>(v1 + g * 8, c + 16 + g * 8, sbias, arena, row);
    }
  } else if (LDW == 160) {   // four x16 loads in flight (64 registers), one wait per 64 columns
    uint32_t v[4][16];
    for (int c = c_begin; c < c_end; c += 64) {
#pragma unroll
      for (int q = 0; q < 4; ++q) tmem_ld16(acc + c + q * 16, v[q]);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int g = 0; g < 2; ++g) chunk8<BIAS>(v[q] + g * 8, c + q * 16 + g * 8, sbias, arena, row);
    }
  } else {   // 8
    uint32_t v0[8], v1[8];
    tmem_ld8(acc + c_begin, v0);
    for (int c = c_begin; c < c_end; c += 16) {
      tmem_ld_wait();
      tmem_ld8(acc + c + 8, v1);
      chunk8<BIAS>(v0, c, sbias, arena, row);
      tmem_ld_wait();
      if (c + 16 < c_end) tmem_ld8(acc + c + 16, v0);
      chunk8<BIAS>(v1, c + 8, sbias, arena, row);
    }
  }
}

// blockDim = NSLOT * WPQ * 128 threads.  Slot s owns accumulator columns [256 s, 256 s + 256) and arena blocks [4 s, 4 s + 4).
template <int LDW, int BIAS, int WPQ>
__global__ void __launch_bounds__(256 * WPQ, 1) k_epi(int iters, int nslot, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(smem_u32(&tptr), 512);
  float* bias_all = reinterpret_cast<float*>(smem + 8 * KB_BYTES);
  for (int i = threadIdx.x; i < 512; i += blockDim.x) bias_all[i] = 0.01f * (float)i;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const int s = warp / (4 * WPQ), w = warp % (4 * WPQ);
  const int part = w >> 2;
  const uint32_t row = (uint32_t)((w & 3) * 32 + lane);
  const uint32_t acc = tptr + (uint32_t)s * 256u + ((uint32_t)((w & 3) * 32) << 16);
  uint8_t* arena = smem + (size_t)s * 4 * KB_BYTES;
  const uint32_t sbias = smem_u32(bias_all + s * 256);
  const int per = 256 / WPQ;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    epi_cols<LDW, BIAS>(acc, part * per, (part + 1) * per, sbias, arena, row);
    tcgen05_fence_before();
    fence_proxy_async();
    named_bar_sync(1 + s, 128 * WPQ);
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (warp == 0) tmem_dealloc(tptr, 512);
}

template <int LDW, int BIAS, int WPQ>
static void bench(const char* name, int nslot, unsigned long long* d_out) {
  const int iters = 512;
  const size_t sm = 8 * KB_BYTES + 2048;
  cudaFuncSetAttribute(k_epi<LDW, BIAS, WPQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  for (int rep = 0; rep < 2; ++rep) {
    k_epi<LDW, BIAS, WPQ><<<148, nslot * WPQ * 128, sm>>>(iters, nslot, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return; }
  }
  unsigned long long h[148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < 148; ++i) s += (double)h[i];
  const double per_iter = s / 148 / iters;
  printf("%-44s slots %d warps/slot %d: %7.0f cyc per epilogue pass -> %7.0f cyc per tile-layer of SM time\n", name, nslot, 4 * WPQ, per_iter,
         per_iter / nslot);
}

int main() {
  unsigned long long* d_out;
  cudaMalloc(&d_out, 148 * 8);
  for (int nslot = 1; nslot <= 2; ++nslot) {
    bench<32, 1, 1>("x32 loads, LDS.128 bias (product)", nslot, d_out);
    bench<32, 0, 1>("x32 loads, no bias", nslot, d_out);
    bench<16, 1, 1>("x16 loads, LDS.128 bias", nslot, d_out);
    bench<16, 0, 1>("x16 loads, no bias", nslot, d_out);
    bench<160, 1, 1>("4 x x16 loads in flight, LDS.128 bias", nslot, d_out);
    bench<160, 0, 1>("4 x x16 loads in flight, no bias", nslot, d_out);
    bench<8, 1, 1>("x8 loads, LDS.128 bias", nslot, d_out);
    bench<8, 0, 1>("x8 loads, no bias", nslot, d_out);
    bench<32, 1, 2>("x32 loads, LDS.128 bias", nslot, d_out);
    bench<32, 0, 2>("x32 loads, no bias", nslot, d_out);
    bench<16, 1, 2>("x16 loads, LDS.128 bias", nslot, d_out);
    bench<16, 0, 2>("x16 loads, no bias", nslot, d_out);
    bench<160, 1, 2>("4 x x16 loads in flight, LDS.128 bias", nslot, d_out);
    bench<160, 0, 2>("4 x x16 loads in flight, no bias", nslot, d_out);
  }
  return 0;
}
