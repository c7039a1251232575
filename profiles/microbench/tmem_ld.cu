// Micro-benchmark (sm_100a): how fast can the epilogue warps read a TMEM accumulator?  One CTA per SM, W warps, each warp reads its
// lane quarter (warp % 4) with tcgen05.ld in a loop.  Prints cycles per 128x256 fp32 accumulator (128 KB) and bytes/clk/SM.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_ld tmem_ld.cu && ./tmem_ld
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int SHAPE>   // 0: 32x32b.x32   1: 32x32b.x64   2: 16x256b.x8 (two per 32 lanes)   3: 32x32b.x16
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t& sink) {
  if (SHAPE == 0) {
    uint32_t v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) sink ^= v[i];
  } else if (SHAPE == 3) {
    uint32_t v[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) sink ^= v[i];
  } else if (SHAPE == 2) {
    uint32_t v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) sink ^= v[i];
  }
}

// bytes one call of ld<SHAPE> moves per warp
template <int SHAPE> __host__ __device__ constexpr int ld_bytes() { return SHAPE == 0 ? 4096 : (SHAPE == 3 ? 2048 : (SHAPE == 2 ? 4096 : 8192)); }
template <int SHAPE> __host__ __device__ constexpr int ld_cols() { return SHAPE == 0 ? 32 : (SHAPE == 3 ? 16 : 64); }

template <int SHAPE>
__global__ void __launch_bounds__(1024, 1) k_ldtm(int iters, unsigned long long* out, uint32_t* sink_out) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t sink = 0;
  // warps sharing a lane quarter read disjoint column ranges
  const int per = 512 / ((nw + 3) / 4);
  const int c0 = (warp >> 2) * per;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (SHAPE == 2) {
      // 16 lanes x 64 columns per instruction; second half of the quarter at lane offset 16
      const int c = c0 + (it * 64) % per;
      ld<2>(base + c, sink);
      ld<2>(base + (16u << 16) + c, sink);
    } else {
      const int c = c0 + (it * ld_cols<SHAPE>()) % per;
      ld<SHAPE>(base + c, sink);
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (sink == 0x12345678u) sink_out[0] = sink;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "r"(512) : "memory");
}

// Two loads in flight per warp (the product epilogue's schedule): wait::ld after issuing the next one.
__global__ void __launch_bounds__(1024, 1) k_ldtm_pipelined(int iters, unsigned long long* out, uint32_t* sink_out) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t sink = 0;
  const int per = 512 / ((nw + 3) / 4);
  const int c0 = (warp >> 2) * per;
  __syncthreads();
  const long long t0 = clock64();
#define LD32(V, ADDR) asm volatile( \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 " \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
        : "=r"(V[0]), "=r"(V[1]), "=r"(V[2]), "=r"(V[3]), "=r"(V[4]), "=r"(V[5]), "=r"(V[6]), "=r"(V[7]), \
          "=r"(V[8]), "=r"(V[9]), "=r"(V[10]), "=r"(V[11]), "=r"(V[12]), "=r"(V[13]), "=r"(V[14]), "=r"(V[15]), \
          "=r"(V[16]), "=r"(V[17]), "=r"(V[18]), "=r"(V[19]), "=r"(V[20]), "=r"(V[21]), "=r"(V[22]), "=r"(V[23]), \
          "=r"(V[24]), "=r"(V[25]), "=r"(V[26]), "=r"(V[27]), "=r"(V[28]), "=r"(V[29]), "=r"(V[30]), "=r"(V[31]) \
        : "r"(ADDR) : "memory")
  uint32_t a[32], b[32];
  LD32(a, base + c0);
  for (int it = 0; it < iters; it += 2) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    LD32(b, base + c0 + ((it + 1) * 32) % per);
#pragma unroll
    for (int i = 0; i < 32; ++i) sink ^= a[i];
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    LD32(a, base + c0 + ((it + 2) * 32) % per);
#pragma unroll
    for (int i = 0; i < 32; ++i) sink ^= b[i];
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (sink == 0x12345678u) sink_out[0] = sink ^ a[0];
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "r"(512) : "memory");
}

// Shared-memory side of the epilogue: warp-broadcast LDS.128 (the bias loads), and the swizzled 16-byte row stores.
__global__ void __launch_bounds__(1024, 1) k_lds(int iters, int mode, unsigned long long* out, float* sink_out) {
  __shared__ __align__(16) float buf[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) buf[i] = (float)i;
  __syncthreads();
  const uint32_t sb = smem_u32(buf);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float x, y, z, w;
    uint32_t addr;
    if (mode == 0) addr = sb + ((it * 16 + warp * 64) & 0x3ff0);                       // broadcast 16 B
    else if (mode == 1) addr = sb + (((it * 512 + lane * 16) + warp * 64) & 0x7ff0);   // 32 distinct consecutive 16 B
    else addr = sb + ((it * 4 + warp * 64) & 0x3ffc);
    if (mode == 2) {
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(addr));
      y = z = w = 0.f;
    } else if (mode == 3) {
      addr = sb + ((it * 8 + warp * 64) & 0x3ff8);
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(addr));
      z = w = 0.f;
    } else {
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(addr));
    }
    acc += x + y + z + w;
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 1.2345f) sink_out[0] = acc;
}


// Throughput form: eight independent loads per iteration.
__global__ void __launch_bounds__(1024, 1) k_lds_tp(int iters, int mode, unsigned long long* out, float* sink_out) {
  __shared__ __align__(16) float buf[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) buf[i] = (float)i;
  __syncthreads();
  const uint32_t sb = smem_u32(buf);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it += 8) {
    float x[8][4];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      x[q][1] = x[q][2] = x[q][3] = 0.f;
      if (mode == 0) {
        const uint32_t addr = sb + (((it + q) * 16 + warp * 64) & 0x3ff0);
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[q][0]), "=f"(x[q][1]), "=f"(x[q][2]), "=f"(x[q][3]) : "r"(addr));
      } else if (mode == 1) {
        const uint32_t addr = sb + ((((it + q) * 512 + lane * 16) + warp * 64) & 0x7ff0);
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[q][0]), "=f"(x[q][1]), "=f"(x[q][2]), "=f"(x[q][3]) : "r"(addr));
      } else if (mode == 2) {
        const uint32_t addr = sb + (((it + q) * 4 + warp * 64) & 0x3ffc);
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[q][0]) : "r"(addr));
      } else if (mode == 3) {
        const uint32_t addr = sb + (((it + q) * 8 + warp * 64) & 0x3ff8);
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x[q][0]), "=f"(x[q][1]) : "r"(addr));
      } else {   // 4: each lane of a quad-group its own 16 bytes: lanes 0..7 distinct, replicated 4x (128 distinct bytes per instruction)
        const uint32_t addr = sb + ((((it + q) * 128 + (lane & 7) * 16) + warp * 64) & 0x7ff0);
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[q][0]), "=f"(x[q][1]), "=f"(x[q][2]), "=f"(x[q][3]) : "r"(addr));
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) acc += x[q][0] + x[q][1] + x[q][2] + x[q][3];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 1.2345f) sink_out[0] = acc;
}

template <class F>
static double run(F launch, unsigned long long* d_out) {
  launch();
  cudaDeviceSynchronize();
  launch();
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return -1; }
  unsigned long long h[148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < 148; ++i) s += (double)h[i];
  return s / 148;
}

int main() {
  unsigned long long* d_out;
  uint32_t* d_sink;
  cudaMalloc(&d_out, 148 * 8);
  cudaMalloc(&d_sink, 64);
  const int iters = 2048;
  const char* names[4] = {"32x32b.x32", "32x32b.x64(n/a)", "16x256b.x8 (x2)", "32x32b.x16"};
  for (int nw : {1, 4, 8, 16}) {
    {
      double c = run([&] { k_ldtm<0><<<148, nw * 32, 0>>>(iters, d_out, d_sink); }, d_out);
      double bytes = (double)iters * 4096 * nw;
      printf("LDTM %-16s warps %2d  serial    : %8.0f cyc, %6.1f B/clk/SM, %6.0f cyc per 128 KB accumulator\n", names[0], nw, c, bytes / c, c / bytes * 131072);
    }
    {
      double c = run([&] { k_ldtm<3><<<148, nw * 32, 0>>>(iters, d_out, d_sink); }, d_out);
      double bytes = (double)iters * 2048 * nw;
      printf("LDTM %-16s warps %2d  serial    : %8.0f cyc, %6.1f B/clk/SM, %6.0f cyc per 128 KB accumulator\n", names[3], nw, c, bytes / c, c / bytes * 131072);
    }
    {
      double c = run([&] { k_ldtm<2><<<148, nw * 32, 0>>>(iters, d_out, d_sink); }, d_out);
      double bytes = (double)iters * 8192 * nw;
      printf("LDTM %-16s warps %2d  serial    : %8.0f cyc, %6.1f B/clk/SM, %6.0f cyc per 128 KB accumulator\n", names[2], nw, c, bytes / c, c / bytes * 131072);
    }
    {
      double c = run([&] { k_ldtm_pipelined<<<148, nw * 32, 0>>>(iters, d_out, d_sink); }, d_out);
      double bytes = (double)iters * 4096 * nw;
      printf("LDTM %-16s warps %2d  2 in flight: %8.0f cyc, %6.1f B/clk/SM, %6.0f cyc per 128 KB accumulator\n", names[0], nw, c, bytes / c, c / bytes * 131072);
    }
  }
  const char* lm[4] = {"LDS.128 broadcast", "LDS.128 distinct", "LDS.32 broadcast", "LDS.64 broadcast"};
  for (int nw : {4, 8, 16})
    for (int mode = 0; mode < 4; ++mode) {
      double c = run([&] { k_lds<<<148, nw * 32, 0>>>(iters, mode, d_out, (float*)d_sink); }, d_out);
      printf("%-18s warps %2d: %6.2f cyc per warp-instruction (all warps: %.2f instr/clk/SM)\n", lm[mode], nw, c / iters, (double)iters * nw / c);
    }
  const char* lt[5] = {"LDS.128 broadcast", "LDS.128 distinct", "LDS.32 broadcast", "LDS.64 broadcast", "LDS.128 8 distinct x4"};
  for (int nw : {4, 8, 16})
    for (int mode = 0; mode < 5; ++mode) {
      double c = run([&] { k_lds_tp<<<148, nw * 32, 0>>>(iters, mode, d_out, (float*)d_sink); }, d_out);
      printf("throughput %-22s warps %2d: %6.2f cyc per warp-instruction per warp; SM-wide %.2f cyc per instruction\n", lt[mode], nw, c / iters, c / iters / nw);
    }
  return 0;
}
