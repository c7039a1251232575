// Micro-benchmark (sm_100a): how fast can one SM pull L2-resident weight stages into shared memory with cp.async.bulk, as a function of the
// bytes in flight (ring depth) -- unicast and multicast to a CTA pair.  All 148 SMs stream the same 1.1 MB blob (the FaceNeRF weights) in
// 16 KB stages, no consumer.  Reports B/clk/SM and the implied round-trip latency of one stage.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../dfa-nerf_b200/csrc -o tma_ingest tma_ingest.cu && ./tma_ingest
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"

using namespace dfn::tc;

template <bool MC>
__global__ void __launch_bounds__(32, 1) k_tma(const uint8_t* blob, int blob_stages, int stage_bytes, int nst, int iters, unsigned long long* out, int stagger) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + (uint32_t)nst * (uint32_t)stage_bytes;
  const uint32_t crank = MC ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nst; ++i) mbar_init(bars + 8 * i, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (MC) cluster_sync_all();
  const int soff = stagger ? (int)(blockIdx.x >> (MC ? 1 : 0)) * stagger : 0;
  const long long t0 = clock64();
  // prologue: fill the ring
  for (int i = 0; i < nst; ++i) {
    if (elect_one_sync()) {
      mbar_expect_tx(bars + 8 * i, (uint32_t)stage_bytes);
      const uint8_t* src = blob + (size_t)((i + soff) % blob_stages) * stage_bytes;
      if (MC) {
        if ((i & 1) == (int)crank) tma_bulk_load_mc(sbase + i * stage_bytes, src, (uint32_t)stage_bytes, bars + 8 * i, (uint16_t)3);
      } else {
        tma_bulk_load(sbase + i * stage_bytes, src, (uint32_t)stage_bytes, bars + 8 * i);
      }
    }
    __syncwarp();
  }
  long long t_wait = 0, t_issue = 0;
  const int lg = 31 - __clz(nst);
  const uint32_t bmask = (uint32_t)blob_stages - 1u;
  for (int it = 0; it < iters; ++it) {
    const int e = it & (nst - 1);
    const uint32_t par = (uint32_t)(it >> lg) & 1u;
    mbar_wait(bars + 8 * e, par);
    if (MC) {
      // the peer must have seen the stage too before it is overwritten: in the product the release needs both CTAs' MMAs; here a
      // cluster barrier per stage would dominate, so the copy is simply re-armed (both CTAs wait on their own barrier first)
    }
    if (it + nst < iters + nst) {
      if (elect_one_sync()) {
        mbar_expect_tx(bars + 8 * e, (uint32_t)stage_bytes);
        const uint8_t* src = blob + (((uint32_t)(it + nst + soff) & bmask) * (uint32_t)stage_bytes);
        if (MC) {
          if (((it + nst) & 1) == (int)crank) tma_bulk_load_mc(sbase + e * stage_bytes, src, (uint32_t)stage_bytes, bars + 8 * e, (uint16_t)3);
        } else {
          tma_bulk_load(sbase + e * stage_bytes, src, (uint32_t)stage_bytes, bars + 8 * e);
        }
      }
      __syncwarp();
    }
  }
  // drain
  for (int i = 0; i < nst; ++i) {
    const int it = iters + i;
    mbar_wait(bars + 8 * (it % nst), (uint32_t)(it / nst) & 1u);
  }
  const long long t1 = clock64();
  if (MC) cluster_sync_all();
  if (threadIdx.x == 0) { out[blockIdx.x] = (unsigned long long)(t1 - t0); out[148 + blockIdx.x] = (unsigned long long)t_wait; out[296 + blockIdx.x] = (unsigned long long)t_issue; }
}

int main() {
  const int blob_bytes = 1024 * 1024;
  uint8_t* blob;
  unsigned long long* d_out;
  cudaMalloc(&blob, blob_bytes);
  cudaMemset(blob, 1, blob_bytes);
  cudaMalloc(&d_out, 3 * 148 * 8);
  const int iters = 4096;
  for (int grid : {148})
  for (int stagger : {0})
  for (int mc = 0; mc < 2; ++mc)
    for (int stage_kb : {8, 16, 32})
      for (int nst : {1, 2, 4, 8}) {
        const int stage_bytes = stage_kb * 1024;
        if ((size_t)nst * stage_bytes > 200 * 1024) continue;
        const size_t sm = (size_t)nst * stage_bytes + 128;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(32);
        cfg.dynamicSmemBytes = sm;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = mc ? 2 : 1;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        const uint8_t* cb = blob;
        int bs = blob_bytes / stage_bytes;
        for (int rep = 0; rep < 2; ++rep) {
          if (mc) {
            cudaFuncSetAttribute(k_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            cudaLaunchKernelEx(&cfg, k_tma<true>, cb, bs, stage_bytes, nst, iters, d_out, stagger);
          } else {
            cudaFuncSetAttribute(k_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            cudaLaunchKernelEx(&cfg, k_tma<false>, cb, bs, stage_bytes, nst, iters, d_out, stagger);
          }
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        }
        unsigned long long h[3 * 148];
        cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        double s = 0;
        for (int i = 0; i < grid; ++i) s += (double)h[i];
        const double cyc = s / grid;
        const double bytes = (double)(iters + nst) * stage_bytes;
        printf("grid %3d stagger %d %s stage %2d KB x %2d in flight (%3d KB): %6.1f B/clk/SM ingest, %6.0f cyc per stage, implied round trip %6.0f cyc; per stage: wait %5.0f issue %5.0f\n",
               grid, stagger, mc ? "multicast x2" : "unicast     ", stage_kb, nst, nst * stage_kb, bytes / cyc, cyc / (iters + nst), cyc / (iters + nst) * nst, (double)h[148] / iters, (double)h[296] / iters);
      }
  return 0;
}
